"""Alias module: same import path tail as pcdet/ops/iou3d_nms/iou3d_nms_utils.py."""
from .box_ops import boxes_bev_iou_cpu, boxes_iou3d_gpu, boxes_iou_bev, nms_gpu, nms_normal_gpu  # noqa: F401
