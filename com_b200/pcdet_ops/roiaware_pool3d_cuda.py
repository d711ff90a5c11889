"""Drop-in for the pybind module `pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda`
(pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:172-177): points_in_boxes_cpu / _gpu.
`forward`/`backward` (RoI-aware pooling) are not on the COM hot path and raise."""
import torch

from .. import ops


def points_in_boxes_cpu(boxes_tensor, pts_tensor, pts_indices_tensor):
    """(boxes (N,7), pts (P,3), out (N,P) int32), all CPU — note the (boxes, pts) argument order
    (roiaware_pool3d.cpp:143-168).  Evaluated on the GPU with the CPU build's exact arithmetic; worker-process
    policy in ops.host_op_device."""
    dev = ops.host_op_device()
    boxes = boxes_tensor.float().contiguous()
    pts = pts_tensor.float().contiguous()
    with torch.cuda.device(dev):
        trig = torch.from_numpy(ops.box_trig_host(boxes.numpy())).to(dev)
        mask = ops.points_in_boxes_mask(pts.to(dev), boxes.to(dev), trig)
        pts_indices_tensor.copy_(mask)
    return 1


def points_in_any_box_cpu(boxes_tensor, pts_tensor):
    """(boxes (N,7), pts (P,>=3)) CPU -> (P,) bool CPU: the point lies in at least one box.  Equals
    `points_in_boxes_cpu(...).sum(0) != 0` (what remove_points_in_boxes3d needs, box_utils.py:128-129) with P bytes
    of device->host traffic instead of N*P*4."""
    dev = ops.host_op_device()
    boxes = boxes_tensor.float().contiguous()
    pts = pts_tensor.float().contiguous()
    with torch.cuda.device(dev):
        trig = torch.from_numpy(ops.box_trig_host(boxes.numpy())).to(dev)
        any_ = ops.points_in_any_box(pts.to(dev), boxes.to(dev), trig)
        return any_.cpu().bool()


def points_in_boxes_gpu(boxes_tensor, pts_tensor, box_idx_of_points_tensor):
    """(boxes (B,T,7), pts (B,P,3), out (B,P) int32 prefilled with -1), all CUDA
    (roiaware_pool3d.cpp:96-117, kernel roiaware_pool3d_kernel.cu:313-336)."""
    ops.points_in_boxes_index(pts_tensor.contiguous(), boxes_tensor.contiguous(), out=box_idx_of_points_tensor)
    return 1


def forward(*args, **kwargs):
    raise NotImplementedError("roiaware_pool3d forward is outside the COM hot path (SURVEY.md §8)")


def backward(*args, **kwargs):
    raise NotImplementedError("roiaware_pool3d backward is outside the COM hot path (SURVEY.md §8)")
