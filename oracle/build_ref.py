"""Compile the REFERENCE's own iou3d_nms and roiaware_pool3d extensions from the sources where they
lie under /root/reference into oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).

Used to pin oracle.c's box-op restatement (CPU entry points) and, on the GPU box, as the bit-exact
oracle for the device flavour (nms_gpu, boxes_iou_bev_gpu, points_in_boxes_gpu).  No reference
source is copied into the repo.  `-O2` is REQUIRED: `check_rect_cross` is `inline` in iou3d_cpu.cpp
but a non-inline __device__ function in iou3d_nms_kernel.cu, and without inlining nvcc's host stub
(exit(1)) wins at link time (SURVEY.md §7 hard part 9).
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("COM_REFERENCE", "/root/reference")
MODS = {
    "ref_iou3d_nms_cuda": ["iou3d_nms/src/" + f for f in
                           ("iou3d_cpu.cpp", "iou3d_nms_api.cpp", "iou3d_nms.cpp", "iou3d_nms_kernel.cu")],
    "ref_roiaware_pool3d_cuda": ["roiaware_pool3d/src/" + f for f in
                                 ("roiaware_pool3d.cpp", "roiaware_pool3d_kernel.cu")],
}


def build(verbose=False):
    """Build both modules (no-op when /root/reference is absent, e.g. on the GPU box)."""
    if not os.path.isdir(os.path.join(REF, "pcdet", "ops")):
        return False
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    for name, srcs in MODS.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name, sources=[os.path.join(REF, "pcdet", "ops", s) for s in srcs], extra_cflags=["-O2"],
             extra_cuda_cflags=["-O2"], build_directory=bdir, verbose=verbose)
    from . import ref_py
    ref_py.build_pyc()          # byte-code of the reference's hot-path Python modules (see ref_py.py)
    return True


def available():
    from . import ref_py
    return all(os.path.exists(os.path.join(OUT, n, n + ".so")) for n in MODS) and ref_py.pyc_available()


def load_ref(name):
    """Import a prebuilt reference module from oracle/_ref (works without /root/reference)."""
    import torch  # noqa: F401  (the extension links against libtorch)
    path = os.path.join(OUT, name, name + ".so")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    __package__ = "oracle"
    import oracle  # noqa: F401
    ok = build(verbose="-v" in sys.argv)
    print("built" if ok else "reference tree not found; nothing built", OUT)
