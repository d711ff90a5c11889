"""Drop-in for the pybind module `pcdet.ops.iou3d_nms.iou3d_nms_cuda`
(pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17).  Same function names, argument order and
out-parameter convention; argument errors raise RuntimeError instead of exit(-1)."""
import numpy as np
import torch

from .. import ops


def _check_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous tensor" % name)


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """iou3d_nms.cpp:49-68"""
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_overlap, "ans_overlap")):
        _check_cuda(t, n)
    ops.boxes_bev(boxes_a, boxes_b, flavour="gpu", what="overlap", out=ans_overlap)
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    """iou3d_nms.cpp:70-88"""
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        _check_cuda(t, n)
    ops.boxes_bev(boxes_a, boxes_b, flavour="gpu", what="iou", out=ans_iou)
    return 1


def _nms(boxes, keep, thresh, rotated):
    _check_cuda(boxes, "boxes")
    if not keep.is_contiguous():
        raise RuntimeError("keep must be contiguous tensor")
    keep_dev, num_dev = ops.nms(boxes, thresh, rotated=rotated, flavour="gpu")
    num = int(num_dev.item())
    keep[:num] = keep_dev[:num].to(keep.device)   # keep is a CPU LongTensor in the reference wrapper
    return num


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:90-136 — boxes sorted by score, `keep` (N,) int64 out-param, returns the count."""
    return _nms(boxes, keep, nms_overlap_thresh, True)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:139-188"""
    return _nms(boxes, keep, nms_overlap_thresh, False)


def boxes_iou_bev_cpu(boxes_a_tensor, boxes_b_tensor, ans_iou_tensor):
    """iou3d_cpu.cpp:232-252 — CPU tensors in/out; computed on the GPU with the CPU build's exact
    arithmetic (host-libm trig tables, no FMA).  Called from DataLoader workers by COMAug
    (database_sampler_v2.py:600-604): see ops.host_op_device for the worker-process policy."""
    if not (boxes_a_tensor.is_contiguous() and boxes_b_tensor.is_contiguous()):
        raise RuntimeError("boxes must be contiguous tensor")
    dev = ops.host_op_device()
    a = boxes_a_tensor.float()
    b = boxes_b_tensor.float()
    with torch.cuda.device(dev):
        ta = torch.from_numpy(ops.box_trig4_host(a.numpy())).to(dev)
        tb = torch.from_numpy(ops.box_trig4_host(b.numpy())).to(dev)
        out = ops.boxes_bev(a.to(dev).contiguous(), b.to(dev).contiguous(), flavour="cpu", what="iou", trig_a=ta,
                            trig_b=tb)
        ans_iou_tensor.copy_(out)
    return 1
