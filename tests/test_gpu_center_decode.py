"""f2 — CenterHead post-processing fused on the device (comb_centerhead_decode_nms) against the reference's own
Python: `CenterHead.generate_predicted_boxes` (pcdet/models/dense_heads/center_head.py:266-317) calling
`centernet_utils.decode_bbox_from_heatmap` (centernet_utils.py:199-279) and `model_nms_utils.class_agnostic_nms`
(model_nms_utils.py:6-25), loaded unmodified by oracle/ref_py.py.

Bars: the selected heat-map cells (order included), labels and detection counts are bit-exact; boxes and scores 1e-6
(the decode of the reference runs on CPU tensors: different libm for exp / atan2 / sigmoid)."""
import types

import numpy as np
import pytest
import torch

from com_b200 import ops
from com_b200.pcdet_ops import center_decode
from oracle import ref_py

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_py.available(), reason="reference Python not available")]

RANGE = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]
VSIZE = [0.1, 0.1, 0.15]


def head_outputs(B, C, H, W, seed, peaks=900):
    g = torch.Generator().manual_seed(seed)
    hm = torch.randn((B, C, H, W), generator=g) * 1.2 - 5.0
    for b in range(B):                                     # object-like peaks, some in clusters (NMS has work to do)
        ys, xs = torch.randint(2, H - 2, (peaks,), generator=g), torch.randint(2, W - 2, (peaks,), generator=g)
        cs = torch.randint(0, C, (peaks,), generator=g)
        hm[b, cs, ys, xs] = torch.randn((peaks,), generator=g) * 2.0 + 1.0
        hm[b, cs, ys, (xs + 1).clamp(max=W - 1)] = torch.randn((peaks,), generator=g) * 2.0 + 0.5
    rot = torch.randn((B, 2, H, W), generator=g)
    return {"hm": hm, "center": torch.rand((B, 2, H, W), generator=g), "center_z": torch.randn((B, 1, H, W), generator=g),
            "dim": torch.randn((B, 3, H, W), generator=g) * 0.3 + torch.tensor([1.5, 0.7, 0.5]).view(1, 3, 1, 1), "rot": rot}


def fake_head(E, K=500, score_thresh=0.1, nms_thresh=0.7, classes=3, vel=False):
    """What generate_predicted_boxes reads from `self` (center_head.py:55-70,266-317)."""
    cfg = E(POST_PROCESSING=E(SCORE_THRESH=score_thresh, POST_CENTER_LIMIT_RANGE=[-75.2, -75.2, -2, 75.2, 75.2, 4],
                              MAX_OBJ_PER_SAMPLE=K,
                              NMS_CONFIG=E(NMS_TYPE="nms_gpu", NMS_THRESH=nms_thresh, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500)))
    return types.SimpleNamespace(model_cfg=cfg, point_cloud_range=RANGE, voxel_size=VSIZE, feature_map_stride=8,
                                 class_id_mapping_each_head=[torch.arange(classes).cuda()],
                                 separate_head_cfg=E(HEAD_ORDER=["center", "center_z", "dim", "rot"] + (["vel"] if vel else [])))


def same_detections(g, w, vel=False):
    """Bit-exact labels / counts, 1e-6 boxes and scores, in the reference's order — except inside runs of EQUAL scores:
    the reference ranks sigmoid(hm) (center_head.py:276), where neighbouring logits collapse onto one fp32 score and
    torch.topk leaves their order unspecified; the fused kernel ranks the logits (ties by flat index).  Positions that
    differ must therefore sit in such a run, and the two lists must agree once both are put in a canonical order."""
    assert g["pred_boxes"].shape == w["pred_boxes"].shape and g["pred_boxes"].shape[0] > 20
    assert g["pred_labels"].dtype == torch.int64
    assert torch.allclose(g["pred_scores"], w["pred_scores"], rtol=0, atol=1e-6)
    ws = w["pred_scores"]
    tied = torch.zeros_like(ws, dtype=torch.bool)
    eq = (ws[1:] - ws[:-1]).abs() <= 1e-6
    tied[1:] |= eq
    tied[:-1] |= eq
    differ = (g["pred_labels"] != w["pred_labels"]) | ((g["pred_boxes"] - w["pred_boxes"]).abs() > 1e-5).any(dim=1)
    assert not bool((differ & ~tied).any()), "order differs outside runs of equal scores"
    assert int(differ.sum()) <= 8                      # a handful of swapped pairs at most

    def canon(d):
        b = d["pred_boxes"].double().cpu().numpy()
        order = np.lexsort((b[:, 1], b[:, 0], -np.round(d["pred_scores"].double().cpu().numpy(), 5)))
        return d["pred_boxes"].cpu()[order], d["pred_labels"].cpu()[order]
    (gb, gl), (wb, wl) = canon(g), canon(w)
    assert torch.equal(gl, wl)
    assert torch.allclose(gb[:, :7], wb[:, :7], rtol=1e-6, atol=1e-6)
    if vel:
        assert gb.shape[1] == 9 and torch.equal(gb[:, 7:], wb[:, 7:])


@pytest.mark.parametrize("B,H,W,K,seed", [(2, 188, 188, 500, 0), (4, 188, 188, 500, 1), (1, 64, 80, 100, 2), (2, 188, 188, 1000, 3)])
def test_fused_postprocessing_vs_reference_method(B, H, W, K, seed):
    E = ref_py.EasyDict
    ch = ref_py.load("pcdet.models.dense_heads.center_head")
    assert ch.CenterHead.generate_predicted_boxes._comb                    # the post-import hook is in place
    reference_method = ch.CenterHead.generate_predicted_boxes.reference     # the reference's own function object
    head = fake_head(E, K=K)
    pd = {k: v.cuda() for k, v in head_outputs(B, 3, H, W, seed).items()}
    want = reference_method(head, B, [dict(pd)])
    assert center_decode.supported(head)
    got = ch.CenterHead.generate_predicted_boxes(head, B, [dict(pd)])
    assert len(got) == len(want) == B
    for g, w in zip(got, want):
        same_detections(g, w)


@pytest.mark.parametrize("B,H,W,K,seed", [(2, 128, 128, 500, 11), (1, 180, 180, 83, 12)])
def test_velocity_head_vs_reference_method(B, H, W, K, seed):
    """Heads with a 'vel' branch (nuscenes_models/cbgs_*_centerpoint.yaml, waymo_models/centerpoint_4frames.yaml): boxes
    carry nine columns, the velocity pair is copied untouched (bit-exact) and the NMS looks at the first seven."""
    E = ref_py.EasyDict
    ch = ref_py.load("pcdet.models.dense_heads.center_head")
    reference_method = ch.CenterHead.generate_predicted_boxes.reference
    head = fake_head(E, K=K, vel=True)
    pd = head_outputs(B, 3, H, W, seed)
    pd["vel"] = torch.randn((B, 2, H, W), generator=torch.Generator().manual_seed(seed + 100)) * 3.0
    pd = {k: v.cuda() for k, v in pd.items()}
    want = reference_method(head, B, [dict(pd)])
    assert center_decode.supported(head)
    got = ch.CenterHead.generate_predicted_boxes(head, B, [dict(pd)])
    for g, w in zip(got, want):
        same_detections(g, w, vel=True)


def test_decode_stage_vs_reference_on_cpu_tensors():
    """Top-K selection + decode + mask alone (NMS threshold 2.0 keeps everything) against decode_bbox_from_heatmap run
    by the reference on CPU tensors: same cells in the same order, boxes 1e-6."""
    cu = ref_py.load("pcdet.models.model_utils.centernet_utils")
    B, C, H, W, K = 2, 3, 188, 188, 500
    pd = head_outputs(B, C, H, W, 7)
    want = cu.decode_bbox_from_heatmap(
        heatmap=pd["hm"].sigmoid(), rot_cos=pd["rot"][:, 0].unsqueeze(1), rot_sin=pd["rot"][:, 1].unsqueeze(1),
        center=pd["center"], center_z=pd["center_z"], dim=pd["dim"].exp(), point_cloud_range=RANGE, voxel_size=VSIZE,
        feature_map_stride=8, K=K, circle_nms=False, score_thresh=0.1,
        post_center_limit_range=torch.tensor([-75.2, -75.2, -2, 75.2, 75.2, 4]).float())
    d = {k: v.cuda() for k, v in pd.items()}
    boxes, scores, labels, counts = ops.centerhead_decode_nms(
        d["hm"], d["center"], d["center_z"], d["dim"], d["rot"], K, 8, VSIZE, RANGE, [-75.2, -75.2, -2, 75.2, 75.2, 4], 0.1,
        nms_thresh=2.0, nms_pre_max=4096, nms_post_max=K)
    cnt = counts.tolist()
    for b in range(B):
        w = want[b]
        assert cnt[b] == w["pred_boxes"].shape[0] > 50
        assert torch.equal(labels[b, : cnt[b]].cpu().long() - 1, w["pred_labels"].long())
        assert torch.allclose(scores[b, : cnt[b]].cpu(), w["pred_scores"], rtol=0, atol=1e-6)
        assert torch.allclose(boxes[b, : cnt[b]].cpu(), w["pred_boxes"], rtol=1e-6, atol=2e-6)
        assert bool((scores[b, 1: cnt[b]] <= scores[b, : cnt[b] - 1]).all())          # sorted by score


def test_ties_and_small_maps():
    """Equal logits: the smaller flat index wins (deterministic); K larger than the map: everything is returned."""
    B, C, H, W = 1, 2, 4, 5
    hm = torch.full((B, C, H, W), -1.0)
    hm[0, 1, 2, 3] = 3.0
    d = {"hm": hm.cuda(), "center": torch.zeros((B, 2, H, W)).cuda(), "center_z": torch.zeros((B, 1, H, W)).cuda(),
         "dim": torch.zeros((B, 3, H, W)).cuda(), "rot": torch.ones((B, 2, H, W)).cuda()}
    boxes, scores, labels, counts = ops.centerhead_decode_nms(
        d["hm"], d["center"], d["center_z"], d["dim"], d["rot"], 8, 8, VSIZE, [0, 0, -2, 75.2, 75.2, 4], [-80, -80, -10, 80, 80, 10],
        0.0, nms_thresh=2.0, nms_pre_max=4096, nms_post_max=500)
    assert int(counts[0]) == 8
    assert int(labels[0, 0]) == 2 and abs(float(scores[0, 0]) - float(torch.sigmoid(torch.tensor(3.0)))) < 1e-6
    assert labels[0, 1:8].tolist() == [1] * 7                                            # flat indices 0..6 of class 0
    assert torch.allclose(boxes[0, 1:8, 0].cpu(), torch.tensor([0.0, 0.8, 1.6, 2.4, 3.2, 0.0, 0.8]), atol=1e-6)
    b2, s2, l2, c2 = ops.centerhead_decode_nms(
        d["hm"], d["center"], d["center_z"], d["dim"], d["rot"], 100, 8, VSIZE, [0, 0, -2, 75.2, 75.2, 4],
        [-80, -80, -10, 80, 80, 10], 0.0, nms_thresh=2.0, nms_pre_max=4096, nms_post_max=500)
    assert int(c2[0]) == C * H * W
