"""Drop-in body for `CenterHead.generate_predicted_boxes` (pcdet/models/dense_heads/center_head.py:266-317; the COM
head `CurriculumCenterHead` carries the same method, curriculum_center_head.py): heat-map top-K, gathers, box decode,
range / score mask and rotated NMS run as ONE launch sequence per separate head (comb_centerhead_decode_nms) and the
host reads the per-frame detection counts of all heads once, at the end.

The reference runs, per head and per frame: two torch.topk, ~25 gather / elementwise kernels, boolean-mask indexing
(a host synchronisation per frame), a third topk inside class_agnostic_nms (model_nms_utils.py:6-25) and the NMS with
its blocking copies."""
import torch

from .. import ops

MAX_K = 1024


def supported(head):
    """Configurations the fused kernel covers; anything else keeps the reference method."""
    try:
        cfg = head.model_cfg.POST_PROCESSING
        return cfg.NMS_CONFIG.NMS_TYPE == "nms_gpu" and int(cfg.MAX_OBJ_PER_SAMPLE) <= MAX_K
    except AttributeError:
        return False


def generate_predicted_boxes(head, batch_size, pred_dicts):
    """Same arguments and return value as the reference method: a list (one dict per frame) of pred_boxes (n,7) — (n,9)
    for heads with a 'vel' branch (center_head.py:280) —, pred_scores (n,), pred_labels (n,) int64 1-based."""
    cfg = head.model_cfg.POST_PROCESSING
    nms = cfg.NMS_CONFIG
    K = int(cfg.MAX_OBJ_PER_SAMPLE)
    outs = []
    for idx, pd in enumerate(pred_dicts):
        hm = pd["hm"].detach().float().contiguous()
        label_map = head.class_id_mapping_each_head[idx].to(device=hm.device, dtype=torch.int32).contiguous()
        outs.append(ops.centerhead_decode_nms(
            hm, pd["center"].detach().float().contiguous(), pd["center_z"].detach().float().contiguous(),
            pd["dim"].detach().float().contiguous(), pd["rot"].detach().float().contiguous(), K,
            head.feature_map_stride, head.voxel_size, head.point_cloud_range, cfg.POST_CENTER_LIMIT_RANGE,
            cfg.SCORE_THRESH, nms.NMS_THRESH, nms.NMS_PRE_MAXSIZE, nms.NMS_POST_MAXSIZE, label_map=label_map,
            vel=pd["vel"].detach().float().contiguous() if "vel" in head.separate_head_cfg.HEAD_ORDER else None))
    counts = torch.stack([o[3] for o in outs]).tolist()                  # the one host read: [head][frame]
    ret = []
    for k in range(batch_size):
        bx = [o[0][k, : counts[h][k]] for h, o in enumerate(outs)]
        sc = [o[1][k, : counts[h][k]] for h, o in enumerate(outs)]
        lb = [o[2][k, : counts[h][k]].long() for h, o in enumerate(outs)]
        ret.append({"pred_boxes": torch.cat(bx, dim=0), "pred_scores": torch.cat(sc, dim=0),
                    "pred_labels": torch.cat(lb, dim=0)})
    return ret
