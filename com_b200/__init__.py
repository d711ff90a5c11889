"""com_b200 — B200-native (sm_100a) implementation of COM's voxel-detector hot path.

Layers:  include/comb200.h (C-ABI)  <-  com_b200/csrc/*.cu (kernels)  <-  com_b200._lib (ctypes)
         <-  com_b200.ops (torch plumbing)  <-  reference-facing mirrors:
             com_b200.sparse (spconv.pytorch), com_b200.voxel (spconv.utils / VoxelGeneratorWrapper),
             com_b200.pcdet_ops (iou3d_nms, roiaware_pool3d), com_b200.models (MeanVFE,
             VoxelResBackBone8x, HeightCompression), com_b200.pipeline (fused frame pipeline).
"""
import os
import sys

__version__ = "0.1.0"


def install_dropins():
    """Make `import spconv`, `import cumm` and the two pcdet pybind modules resolve to com_b200.

    Call before importing pcdet (or put com_b200/dropin on PYTHONPATH for spconv/cumm)."""
    here = os.path.dirname(os.path.abspath(__file__))
    dropin = os.path.join(here, "dropin")
    if dropin not in sys.path:
        sys.path.insert(0, dropin)
    from .pcdet_ops import iou3d_nms_cuda, roiaware_pool3d_cuda
    sys.modules.setdefault("pcdet.ops.iou3d_nms.iou3d_nms_cuda", iou3d_nms_cuda)
    sys.modules.setdefault("pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda", roiaware_pool3d_cuda)
