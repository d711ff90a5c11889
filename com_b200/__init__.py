"""com_b200 — B200-native (sm_100a) implementation of COM's voxel-detector hot path.

Layers:  include/comb200.h (C-ABI)  <-  com_b200/csrc/*.cu (kernels)  <-  com_b200._lib (ctypes)
         <-  com_b200.ops (torch plumbing)  <-  reference-facing mirrors:
             com_b200.sparse (spconv.pytorch), com_b200.voxel (spconv.utils / VoxelGeneratorWrapper),
             com_b200.pcdet_ops (iou3d_nms, roiaware_pool3d), com_b200.models (MeanVFE,
             VoxelResBackBone8x, HeightCompression), com_b200.pipeline (fused frame pipeline).
"""
import importlib.abc
import importlib.util
import os
import sys

__version__ = "0.2.0"


# ---------------------------------------------------------------------------------------------------------------------
# Post-import hooks: reference modules that are patched right after the reference's own import machinery has executed
# them (zero edits in the reference tree).
def _hook_spconv_backbone(mod):
    """pcdet/models/backbones_3d/spconv_backbone.py: the registry's VoxelResBackBone8x gets the fused eval path."""
    from .models import patch_reference_backbone
    if hasattr(mod, "VoxelResBackBone8x"):
        patch_reference_backbone(mod.VoxelResBackBone8x)


def _hook_box_utils(mod):
    """pcdet/utils/box_utils.py:117-131: remove_points_in_boxes3d through the any-box kernel (same result, P bytes
    back to the host instead of the (Nb,P) int32 mask)."""
    from .pcdet_ops import box_ops
    reference_fn = getattr(mod, "remove_points_in_boxes3d", None)
    if reference_fn is None or getattr(reference_fn, "_comb", False):
        return

    def remove_points_in_boxes3d(points, boxes3d):
        return box_ops.remove_points_in_boxes3d(points, boxes3d)

    remove_points_in_boxes3d.__doc__ = reference_fn.__doc__
    remove_points_in_boxes3d._comb = True
    remove_points_in_boxes3d.reference = reference_fn
    mod.remove_points_in_boxes3d = remove_points_in_boxes3d


def _hook_center_head(mod):
    """pcdet/models/dense_heads/center_head.py:266-317 (and the COM head in curriculum_center_head.py): post-processing
    through comb_centerhead_decode_nms when the configuration is covered, the reference method otherwise."""
    from .pcdet_ops import center_decode
    for name in ("CenterHead", "CurriculumCenterHead"):
        cls = getattr(mod, name, None)
        if cls is None or getattr(cls.generate_predicted_boxes, "_comb", False):
            continue
        reference_method = cls.generate_predicted_boxes

        def generate_predicted_boxes(self, batch_size, pred_dicts, _ref=reference_method):
            import os
            if os.environ.get("COMB_FUSED_DECODE", "1") != "0" and center_decode.supported(self) \
                    and pred_dicts[0]["hm"].is_cuda:
                return center_decode.generate_predicted_boxes(self, batch_size, pred_dicts)
            return _ref(self, batch_size, pred_dicts)

        generate_predicted_boxes._comb = True
        generate_predicted_boxes.reference = reference_method
        cls.generate_predicted_boxes = generate_predicted_boxes
    _hook_center_targets(mod)


def _hook_center_targets(mod):
    """curriculum_center_head.py:203-296, 431-473: target assignment and the curriculum groups on the device
    (comb_centerhead_assign_targets / comb_centerhead_cluster_groups), the in-place relabelling of gt_boxes included."""
    from .pcdet_ops import center_targets
    plain = getattr(mod, "CenterHead", None)
    if plain is not None and not getattr(plain.assign_targets, "_comb", False):
        ref_plain = plain.assign_targets

        def assign_targets_plain(self, gt_boxes, feature_map_size=None, _ref=ref_plain, **kwargs):
            if center_targets.supported_head(self, gt_boxes):
                return center_targets.assign_targets_plain(self, gt_boxes, feature_map_size=feature_map_size, **kwargs)
            return _ref(self, gt_boxes, feature_map_size=feature_map_size, **kwargs)

        assign_targets_plain._comb = True
        assign_targets_plain.reference = ref_plain
        plain.assign_targets = assign_targets_plain
    cls = getattr(mod, "CurriculumCenterHead", None)
    if cls is None or getattr(cls.assign_targets, "_comb", False):
        return
    ref_assign, ref_cluster = cls.assign_targets, cls.cluster

    def assign_targets(self, gt_boxes, feature_map_size=None, npgt=None, true_object=None, _ref=ref_assign, **kwargs):
        if center_targets.supported_head(self, gt_boxes):
            return center_targets.assign_targets(self, gt_boxes, feature_map_size=feature_map_size, npgt=npgt,
                                                 true_object=true_object, **kwargs)
        return _ref(self, gt_boxes, feature_map_size=feature_map_size, npgt=npgt, true_object=true_object, **kwargs)

    def cluster(self, gt_boxes, true_object, occupancy_ratio, facade_type, _ref=ref_cluster):
        if center_targets.supported_head(self, gt_boxes) and true_object is not None:
            return center_targets.cluster(self, gt_boxes, true_object, occupancy_ratio, facade_type)
        return _ref(self, gt_boxes, true_object, occupancy_ratio, facade_type)

    for new, ref in ((assign_targets, ref_assign), (cluster, ref_cluster)):
        new._comb = True
        new.reference = ref
    cls.assign_targets = assign_targets
    cls.cluster = cluster


def _hook_loss_utils(mod):
    """pcdet/utils/loss_utils.py:1180-1309: the object loop and the group confidences of the COM focal loss on the
    device (comb_comloss_reweight / comb_comloss_group_confidence)."""
    from .pcdet_ops import center_targets
    cls = getattr(mod, "FocalLossCenterCurriculum", None)
    if cls is None or getattr(cls.neg_loss, "_comb", False):
        return
    reference_method = cls.neg_loss

    def neg_loss(self, pred, gt, radius_map, box_mask, mask=None, epoch=None, _ref=reference_method):
        if center_targets.supported_loss(self, pred, radius_map, mask):
            return center_targets.neg_loss(self, pred, gt, radius_map, box_mask, mask=mask, epoch=epoch)
        return _ref(self, pred, gt, radius_map, box_mask, mask=mask, epoch=epoch)

    neg_loss._comb = True
    neg_loss.reference = reference_method
    cls.neg_loss = neg_loss


def _hook_height_compression(mod):
    """pcdet/models/backbones_2d/map_to_bev/height_compression.py:9-26 — with com_b200.sparse.config.bev == "bf16" the BEV
    image is produced directly as channels-last bf16 (f4); otherwise the reference method runs unchanged."""
    cls = getattr(mod, "HeightCompression", None)
    if cls is None or getattr(cls.forward, "_comb", False):
        return
    reference_method = cls.forward

    def forward(self, batch_dict, _ref=reference_method):
        from .sparse import config
        t = batch_dict["encoded_spconv_tensor"]
        if config.bev != "bf16" or not hasattr(t, "dense_bev_bf16"):
            return _ref(self, batch_dict)
        batch_dict["spatial_features"] = t.dense_bev_bf16()
        batch_dict["spatial_features_stride"] = batch_dict["encoded_spconv_tensor_stride"]
        return batch_dict

    forward._comb = True
    forward.reference = reference_method
    cls.forward = forward


def _hook_bev_backbone(mod):
    """pcdet/models/backbones_2d/base_bev_backbone.py:81-112 — a bf16 channels-last input (f4) runs the unchanged forward
    under bf16 autocast; `spatial_features_2d` is handed on as fp32 (the dense head keeps its fp32 weights)."""
    import torch
    cls = getattr(mod, "BaseBEVBackbone", None)
    if cls is None or getattr(cls.forward, "_comb", False):
        return
    reference_method = cls.forward

    def forward(self, data_dict, _ref=reference_method):
        x = data_dict["spatial_features"]
        if x.dtype != torch.bfloat16:
            return _ref(self, data_dict)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            data_dict = _ref(self, data_dict)
        data_dict["spatial_features_2d"] = data_dict["spatial_features_2d"].float()
        return data_dict

    forward._comb = True
    forward.reference = reference_method
    cls.forward = forward


POST_IMPORT_HOOKS = {
    "pcdet.models.backbones_3d.spconv_backbone": _hook_spconv_backbone,
    "pcdet.utils.box_utils": _hook_box_utils,
    "pcdet.models.dense_heads.center_head": _hook_center_head,
    "pcdet.models.dense_heads.curriculum_center_head": _hook_center_head,
    "pcdet.utils.loss_utils": _hook_loss_utils,
    "pcdet.models.backbones_2d.map_to_bev.height_compression": _hook_height_compression,
    "pcdet.models.backbones_2d.base_bev_backbone": _hook_bev_backbone,
}


class _HookLoader(importlib.abc.Loader):
    def __init__(self, loader, hook):
        self._loader, self._hook = loader, hook

    def create_module(self, spec):
        return self._loader.create_module(spec)

    def exec_module(self, module):
        self._loader.exec_module(module)
        self._hook(module)

    def __getattr__(self, name):           # get_code / get_source / is_package ... of the wrapped loader
        return getattr(self._loader, name)


class _PostImportFinder(importlib.abc.MetaPathFinder):
    """Finds the hooked modules with the REMAINING finders and wraps their loader so that the hook runs right after
    the module body."""

    def __init__(self):
        self._busy = False

    def find_spec(self, name, path=None, target=None):
        if self._busy or name not in POST_IMPORT_HOOKS:
            return None
        self._busy = True
        try:
            spec = None
            for finder in sys.meta_path:
                if finder is self or not hasattr(finder, "find_spec"):
                    continue
                spec = finder.find_spec(name, path, target)
                if spec is not None:
                    break
        finally:
            self._busy = False
        if spec is None or spec.loader is None:
            return None
        spec.loader = _HookLoader(spec.loader, POST_IMPORT_HOOKS[name])
        return spec


_finder = None


def install_dropins(accelerate=True):
    """Make `import spconv`, `import cumm` and the two pcdet pybind modules resolve to com_b200.

    Call before importing pcdet (or put com_b200/dropin on PYTHONPATH for spconv/cumm).  With `accelerate` (default)
    the reference modules listed in POST_IMPORT_HOOKS are additionally patched right after THEY are imported by the
    reference: the registry's VoxelResBackBone8x takes the fused bf16 tensor-core path in eval mode and
    box_utils.remove_points_in_boxes3d uses the any-box kernel.  Nothing in the reference tree is edited."""
    global _finder
    here = os.path.dirname(os.path.abspath(__file__))
    dropin = os.path.join(here, "dropin")
    if dropin not in sys.path:
        sys.path.insert(0, dropin)
    from .pcdet_ops import iou3d_nms_cuda, roiaware_pool3d_cuda
    sys.modules.setdefault("pcdet.ops.iou3d_nms.iou3d_nms_cuda", iou3d_nms_cuda)
    sys.modules.setdefault("pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda", roiaware_pool3d_cuda)
    if accelerate:
        if _finder is None:
            _finder = _PostImportFinder()
            sys.meta_path.insert(0, _finder)
        for name, hook in POST_IMPORT_HOOKS.items():      # already imported: patch now
            if name in sys.modules:
                hook(sys.modules[name])
