"""Reference-facing Python API of the box ops (same names/arguments as the reference wrappers).

Mirrors, by name:
  pcdet/ops/iou3d_nms/iou3d_nms_utils.py:12-116   boxes_bev_iou_cpu, boxes_iou_bev, boxes_iou3d_gpu,
                                                   nms_gpu, nms_normal_gpu
  pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-41   points_in_boxes_cpu, points_in_boxes_gpu
  pcdet/utils/box_utils.py:117-131,187-200        remove_points_in_boxes3d, enlarge_box3d
All arithmetic runs in libcomb200 kernels; numpy in -> numpy out like the reference.
"""
import numpy as np
import torch

from .. import ops as _ops
from . import iou3d_nms_cuda as _iou
from . import roiaware_pool3d_cuda as _roi


def _to_torch(x):
    """(tensor, was_numpy) — the reference's check_numpy_to_torch contract."""
    return (torch.from_numpy(x).float(), True) if isinstance(x, np.ndarray) else (x, False)


def _ret(t, was_numpy):
    return t.numpy() if was_numpy else t


def _seven(*boxes):
    for b in boxes:
        if b.shape[1] != 7:
            raise AssertionError("boxes must be (N, 7) [x, y, z, dx, dy, dz, heading]")


# ------------------------------------------------------------------ rotated BEV IoU
def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """(N,7),(M,7) CPU tensors or numpy -> (N,M) rotated BEV IoU."""
    a, np_in = _to_torch(boxes_a)
    b, _ = _to_torch(boxes_b)
    if a.is_cuda or b.is_cuda:
        raise AssertionError("Only support CPU tensors")
    _seven(a, b)
    iou = a.new_zeros((a.shape[0], b.shape[0]))
    _iou.boxes_iou_bev_cpu(a.contiguous(), b.contiguous(), iou)
    return _ret(iou, np_in)


def boxes_iou_bev(boxes_a, boxes_b):
    """CUDA (N,7),(M,7) -> (N,M) rotated BEV IoU (device arithmetic of the reference kernel)."""
    _seven(boxes_a, boxes_b)
    iou = boxes_a.new_zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32)
    _iou.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), iou)
    return iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """3D IoU = BEV overlap x height overlap / union volume."""
    _seven(boxes_a, boxes_b)
    bev = boxes_a.new_zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32)
    _iou.boxes_overlap_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), bev)

    def z_span(b):
        return b[:, 2] - b[:, 5] / 2, b[:, 2] + b[:, 5] / 2

    (a_lo, a_hi), (b_lo, b_hi) = z_span(boxes_a), z_span(boxes_b)
    h = (torch.min(a_hi[:, None], b_hi[None, :]) - torch.max(a_lo[:, None], b_lo[None, :])).clamp(min=0)
    inter = bev * h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5])[:, None]
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5])[None, :]
    return inter / (vol_a + vol_b - inter).clamp(min=1e-6)


# ------------------------------------------------------------------ NMS
def _nms_common(fn, boxes, scores, thresh, pre_maxsize):
    _seven(boxes)
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    sorted_boxes = boxes[order].contiguous()
    keep = torch.empty(sorted_boxes.size(0), dtype=torch.int64)      # CPU out-param like the reference
    n = fn(sorted_boxes, keep, thresh)
    return order[keep[:n].to(boxes.device)].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """Rotated-IoU NMS; returns (indices into `boxes` kept, None)."""
    return _nms_common(_iou.nms_gpu, boxes, scores, thresh, pre_maxsize)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """Axis-aligned NMS; returns (indices into `boxes` kept, None)."""
    return _nms_common(_iou.nms_normal_gpu, boxes, scores, thresh, None)


# ------------------------------------------------------------------ points in boxes
def points_in_boxes_cpu(points, boxes):
    """points (P,3), boxes (N,7) (CPU tensors or numpy) -> (N,P) int32 0/1 mask, MARGIN 1e-2."""
    if boxes.shape[1] != 7 or points.shape[1] != 3:
        raise AssertionError("points must be (P,3) and boxes (N,7)")
    pts, np_in = _to_torch(points)
    bxs, _ = _to_torch(boxes)
    mask = pts.new_zeros((bxs.shape[0], pts.shape[0]), dtype=torch.int)
    _roi.points_in_boxes_cpu(bxs.float().contiguous(), pts.float().contiguous(), mask)
    return _ret(mask, np_in)


def points_in_boxes_gpu(points, boxes):
    """points (B,M,3), boxes (B,T,7) CUDA -> (B,M) int32 index of the first containing box, -1 = none."""
    if boxes.shape[0] != points.shape[0] or boxes.shape[2] != 7 or points.shape[2] != 3:
        raise AssertionError("points must be (B,M,3) and boxes (B,T,7)")
    idx = torch.full(points.shape[:2], -1, dtype=torch.int, device=points.device)
    _roi.points_in_boxes_gpu(boxes.contiguous(), points.contiguous(), idx)
    return idx


def enlarge_box3d(boxes3d, extra_width=(0, 0, 0)):
    b, _ = _to_torch(boxes3d)
    out = b.clone()
    out[:, 3:6] += b.new_tensor(extra_width)[None, :]
    return out


def remove_points_in_boxes3d(points, boxes3d):
    """Drop every point that lies in any box (pcdet/utils/box_utils.py:117-131; COMAug,
    database_sampler_v2.py:535-539).  Same result as the reference's `points_in_boxes_cpu(...).sum(0) == 0` filter;
    the (Nb,P) mask is never materialised (comb_points_in_any_box returns one byte per point)."""
    bxs, _ = _to_torch(boxes3d)
    pts, np_in = _to_torch(points)
    inside_any = _roi.points_in_any_box_cpu(bxs, pts[:, 0:3])
    return _ret(pts[~inside_any], np_in)


# ---- f3: the device side of a COMAug sampler step ----------------------------------------------------------------------
def comaug_place_sampled_boxes(sampled_boxes, existed_boxes):
    """The placement test of DataBaseSampler.__call__ (pcdet/datasets/augmentor/database_sampler_v2.py:600-611): which
    of the `sampled_boxes` (S,7+) drawn from the ground-truth database may be pasted into a scene that already holds
    `existed_boxes` (E,7+).  -> (valid_idx int64 (V,), existed_boxes with the valid sampled boxes appended), exactly
    what the reference derives from its two boxes_bev_iou_cpu matrices; here one byte per sampled box crosses PCIe."""
    sampled_boxes = np.asarray(sampled_boxes)
    existed_boxes = np.asarray(existed_boxes)
    valid_idx = _ops.comaug_valid_mask(sampled_boxes, existed_boxes).nonzero()[0]
    valid = sampled_boxes[valid_idx]
    return valid_idx, np.concatenate((existed_boxes, valid[:, :existed_boxes.shape[-1]]), axis=0)


def comaug_add_to_scene(points, sampled_gt_boxes, obj_points, extra_width=(0.0, 0.0, 0.0)):
    """The point-cloud side of DataBaseSampler.add_sampled_boxes_to_scene (database_sampler_v2.py:535-539): scene points
    inside the sampled boxes enlarged by REMOVE_EXTRA_WIDTH are dropped (any-box kernel, P bytes back), the database
    objects' points are put in front.  numpy in, numpy out."""
    large = np.array(sampled_gt_boxes[:, 0:7], dtype=np.float32, copy=True)
    large[:, 3:6] += np.asarray(extra_width, dtype=np.float32)[None, :]
    kept = remove_points_in_boxes3d(points, large)
    return np.concatenate([obj_points[:, :kept.shape[-1]], kept], axis=0)
