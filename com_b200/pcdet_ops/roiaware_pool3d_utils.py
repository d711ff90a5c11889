"""Alias module: same import path tail as pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py."""
from .box_ops import points_in_boxes_cpu, points_in_boxes_gpu  # noqa: F401
