"""Host-side mirrors of the two pcdet op packages on the COM hot path:
pcdet/ops/iou3d_nms and pcdet/ops/roiaware_pool3d (points_in_boxes_* only)."""
from . import box_ops, center_decode, iou3d_nms_cuda, iou3d_nms_utils, roiaware_pool3d_cuda, roiaware_pool3d_utils  # noqa: F401
