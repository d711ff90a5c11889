"""Drop-in bodies for the training-side Python loops of the COM head (f1):

  CurriculumCenterHead.cluster / .assign_targets      pcdet/models/dense_heads/curriculum_center_head.py:431-473, 203-296
  FocalLossCenterCurriculum.neg_loss                   pcdet/utils/loss_utils.py:1180-1309

The reference walks the ground-truth boxes one by one on the host (the targets are built on CPU tensors and copied
back; the loss loop reads `pred[...]` with `.item()` per object and calls 288 tiny `torch.where` chains for the group
confidences): 33 + 35 ms of a 86 ms training step at batch 2 (profiles/r2_full_step_centerpoint_com.json).  Here each of
them is one kernel (csrc/targets.cu); everything that is a whole-tensor torch expression in the reference stays a torch
expression, evaluated in the same order."""
import os

import torch

from .. import ops


def enabled():
    return os.environ.get("COMB_FUSED_TARGETS", "1") != "0"


def supported_head(head, gt_boxes):
    """CUDA fp32 contiguous gt_boxes (the reference relabels `gt_boxes[..., -1]` IN PLACE while it walks a head,
    curriculum_center_head.py:252-254, and later heads see it: the kernel writes into the caller's tensor to reproduce
    that, so it must not be handed a copy), NUM_MAX_OBJS <= 1024."""
    try:
        return (enabled() and gt_boxes.is_cuda and gt_boxes.dtype == torch.float32 and gt_boxes.is_contiguous()
                and int(head.model_cfg.TARGET_ASSIGNER_CONFIG.NUM_MAX_OBJS) <= 1024 and gt_boxes.shape[-1] >= 8)
    except AttributeError:
        return False


def cluster(head, gt_boxes, true_object, occupancy_ratio, facade_type):
    """Same arguments and return value as CurriculumCenterHead.cluster: group (B, M) int64."""
    return ops.centerhead_cluster_groups(gt_boxes.contiguous(), true_object, occupancy_ratio, facade_type)


def assign_targets(head, gt_boxes, feature_map_size=None, npgt=None, true_object=None, **kwargs):
    """Same arguments and return value as CurriculumCenterHead.assign_targets (true_object carries the group)."""
    feature_map_size = feature_map_size[::-1]          # [H, W] -> [x, y]
    cfg = head.model_cfg.TARGET_ASSIGNER_CONFIG
    assert gt_boxes.shape[:-1] == npgt.shape
    ret = {"heatmaps": [], "target_boxes": [], "inds": [], "masks": [], "heatmap_masks": [], "radius_map": [],
           "heatmap_mask": []}
    all_names = ["bg", *head.class_names]
    gt = gt_boxes                                      # contiguous (supported_head): relabelled in place like the reference
    for cur_class_names in head.class_names_each_head:
        cls_map = torch.tensor([cur_class_names.index(n) if n in cur_class_names else -1 for n in all_names],
                               dtype=torch.int32, device=gt.device)
        heatmap, ret_boxes, inds, mask, radius_map = ops.centerhead_assign_targets(
            gt, npgt, true_object, cls_map, len(cur_class_names), feature_map_size, cfg.FEATURE_MAP_STRIDE,
            head.point_cloud_range, head.voxel_size, num_max_objs=cfg.NUM_MAX_OBJS, gaussian_overlap=cfg.GAUSSIAN_OVERLAP,
            min_radius=cfg.MIN_RADIUS, filter_points=head.epoch <= head.epoch_thredhold, min_points=head.min_points,
            relabel_in_place=True)
        ret["heatmaps"].append(heatmap)
        ret["target_boxes"].append(ret_boxes)
        ret["inds"].append(inds)
        ret["masks"].append(mask)
        ret["radius_map"].append(radius_map)
        ret["heatmap_mask"].append(torch.ones_like(heatmap))
    return ret


def assign_targets_plain(head, gt_boxes, feature_map_size=None, **kwargs):
    """Same arguments and return value as CenterHead.assign_targets (pcdet/models/dense_heads/center_head.py:161-220):
    the assignment without the point-count filter and the group column; masks are int64 there."""
    feature_map_size = feature_map_size[::-1]          # [H, W] -> [x, y]
    cfg = head.model_cfg.TARGET_ASSIGNER_CONFIG
    ret = {"heatmaps": [], "target_boxes": [], "inds": [], "masks": [], "heatmap_masks": []}
    all_names = ["bg", *head.class_names]
    npgt = torch.zeros(gt_boxes.shape[:-1], dtype=torch.float32, device=gt_boxes.device)
    for cur_class_names in head.class_names_each_head:
        cls_map = torch.tensor([cur_class_names.index(n) if n in cur_class_names else -1 for n in all_names],
                               dtype=torch.int32, device=gt_boxes.device)
        heatmap, ret_boxes, inds, mask, _ = ops.centerhead_assign_targets(
            gt_boxes, npgt, None, cls_map, len(cur_class_names), feature_map_size, cfg.FEATURE_MAP_STRIDE,
            head.point_cloud_range, head.voxel_size, num_max_objs=cfg.NUM_MAX_OBJS, gaussian_overlap=cfg.GAUSSIAN_OVERLAP,
            min_radius=cfg.MIN_RADIUS, filter_points=False, relabel_in_place=True)
        ret["heatmaps"].append(heatmap)
        ret["target_boxes"].append(ret_boxes)
        ret["inds"].append(inds)
        ret["masks"].append(mask.long())
    return ret


def supported_loss(mod, pred, radius_map, mask):
    return (enabled() and pred.is_cuda and pred.dtype == torch.float32 and mask is not None and radius_map.shape[-1] >= 5
            and radius_map.dtype == torch.int64)


def neg_loss(mod, pred, gt, radius_map, box_mask, mask=None, epoch=None):
    """Same arguments, return value and side effects (mod.confidence_all, mod.avg_confidence, in-place box_mask / mask)
    as FocalLossCenterCurriculum.neg_loss."""
    predc = pred.detach().contiguous()
    radius_map = radius_map.contiguous()
    if mod.conf_shape is not None:
        mod.confidence_all = list(ops.comloss_group_confidence(predc, radius_map, mod.conf_shape[0], mod.conf_shape[1]))
    confidence_true, confidence_aug = 1, 2

    pos_inds = gt.eq(1).float()
    neg_inds = gt.lt(1).float()
    neg_weights = torch.pow(1 - gt, 4)
    loss = 0
    pos_loss = torch.log(pred) * torch.pow(1 - pred, 2) * pos_inds
    neg_loss_ = torch.log(1 - pred) * torch.pow(pred, 2) * neg_weights * neg_inds
    num_obj = pos_inds.float().sum()
    avg_confidence = (pred * pos_inds).sum() / num_obj
    avg_value = avg_confidence.item()
    mod.avg_confidence = mod.alpha * avg_value + (1 - mod.alpha) * mod.avg_confidence

    if mod.use_curriculum_loss:
        threshold = mod.threshold if mod.fix_threshold else mod.avg_confidence * mod.threshold
        assert box_mask.is_contiguous() and mask.is_contiguous()
        ops.comloss_reweight(predc, radius_map, box_mask, mask, threshold, mod.elongation, mod.height, K=mod.K,
                             mode=1 if mod.straight else 2 if mod.tuning else 0, fixed_radius=mod.radius,
                             add_radius=mod.add, only_center=mod.only_center,
                             active=mod.start_epoch <= epoch <= mod.end_epoch)

    if mask is not None:
        mask = mask[:, None, :, :].float()
        pos_loss = pos_loss * mask
        neg_loss_ = neg_loss_ * mask
        num_pos = (pos_inds.float() * mask).sum()
    else:
        num_pos = pos_inds.float().sum()
    pos_loss = pos_loss.sum()
    neg_loss_ = neg_loss_.sum()
    if num_pos == 0:
        loss = loss - neg_loss_
    else:
        loss = loss - (pos_loss + neg_loss_) / num_pos
    return loss, box_mask, avg_value, confidence_true, confidence_aug
