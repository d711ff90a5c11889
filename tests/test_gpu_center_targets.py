"""f1 — CenterHead target assignment and the COM loss re-weighting on the device (csrc/targets.cu) against the
reference's own Python, loaded unmodified by oracle/ref_py.py:
`CurriculumCenterHead.cluster / assign_targets` (pcdet/models/dense_heads/curriculum_center_head.py:431-473, 120-296)
and `FocalLossCenterCurriculum.neg_loss` (pcdet/utils/loss_utils.py:1180-1309) with conf_shape (3, 96), the setting of
CurriculumCenterHead_x5.

Bars: groups, heat maps, inds, masks and radius_map are bit-exact; the regression targets 1e-6 (device logf / cosf /
sinf against the CPU's); box_mask, the weight mask, the group confidences and the loss 1e-5."""
import types

import numpy as np
import pytest
import torch

from com_b200 import ops
from com_b200.pcdet_ops import center_targets
from oracle import ref_py

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_py.available(), reason="reference Python not available")]

RANGE = np.array([-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], dtype=np.float32)
VSIZE = [0.1, 0.1, 0.15]
NAMES = ["Vehicle", "Pedestrian", "Cyclist"]


def scene(B, M, seed, extra=0, pad=7):
    """gt_boxes (B, M, 8 + extra) with a padded tail (class 0), point counts and the COMAug attributes."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.zeros((B, M, 8 + extra))
    n = M - pad
    gt[:, :n, 0] = (torch.rand((B, n), generator=g) - 0.5) * 160.0        # some centres leave the range: clamped
    gt[:, :n, 1] = (torch.rand((B, n), generator=g) - 0.5) * 160.0
    gt[:, :n, 2] = torch.randn((B, n), generator=g)
    cls = torch.randint(1, 4, (B, n), generator=g)
    size = torch.tensor([[4.7, 2.1, 1.7], [0.9, 0.85, 1.75], [1.8, 0.85, 1.75]])[cls - 1]
    gt[:, :n, 3:6] = size * (0.6 + 0.8 * torch.rand((B, n, 3), generator=g))
    gt[:, : n // 20, 3] = 11.0                                             # long vehicles: radius > min, length > 6
    gt[:, 3, 3] = 0.0                                                      # degenerate box: skipped
    gt[:, :n, 6] = (torch.rand((B, n), generator=g) - 0.5) * 6.2
    if extra:
        gt[:, :n, 7:7 + extra] = torch.randn((B, n, extra), generator=g)
    gt[:, :n, -1] = cls.float()
    gt[0, 5, 0:2] = gt[0, 6, 0:2]                                          # two objects on one cell
    npgt = torch.randint(0, 40, (B, M), generator=g).float()
    true_object = torch.randint(1, 3, (B, M), generator=g)
    occ = torch.rand((B, M), generator=g)
    facade = torch.randint(0, 4, (B, M), generator=g)
    return gt, npgt, true_object, occ, facade


def fake_head(ch, E, epoch=0, min_points=5, max_objs=500):
    head = object.__new__(ch.CurriculumCenterHead)
    cfg = E(TARGET_ASSIGNER_CONFIG=E(FEATURE_MAP_STRIDE=8, NUM_MAX_OBJS=max_objs, GAUSSIAN_OVERLAP=0.1, MIN_RADIUS=2))
    head.__dict__.update(model_cfg=cfg, class_names=NAMES, class_names_each_head=[list(NAMES)], point_cloud_range=RANGE,
                         voxel_size=VSIZE, epoch=epoch, epoch_thredhold=100, min_points=min_points)
    return head


@pytest.mark.parametrize("B,M,extra,max_objs,seed", [(2, 120, 0, 500, 0), (4, 300, 0, 500, 1), (2, 90, 2, 500, 2), (1, 260, 0, 64, 3)])
def test_cluster_and_assign_targets_vs_reference_methods(B, M, extra, max_objs, seed):
    E = ref_py.EasyDict
    ch = ref_py.load("pcdet.models.dense_heads.curriculum_center_head")
    assert ch.CurriculumCenterHead.assign_targets._comb and ch.CurriculumCenterHead.cluster._comb   # hooks in place
    ref_assign, ref_cluster = ch.CurriculumCenterHead.assign_targets.reference, ch.CurriculumCenterHead.cluster.reference
    head = fake_head(ch, E, max_objs=max_objs)
    gt, npgt, to, occ, fac = (t.cuda() for t in scene(B, M, seed, extra=extra))
    want_group = ref_cluster(head, gt.clone(), to, occ, fac)
    got_group = ch.CurriculumCenterHead.cluster(head, gt.clone(), to, occ, fac)
    assert got_group.dtype == torch.int64 and torch.equal(got_group, want_group) and int(want_group.max()) > 40
    fm = (188, 188)
    want = ref_assign(head, gt.clone(), feature_map_size=fm, npgt=npgt, true_object=want_group)
    assert center_targets.supported_head(head, gt)
    got = ch.CurriculumCenterHead.assign_targets(head, gt.clone(), feature_map_size=fm, npgt=npgt, true_object=want_group)
    assert set(got) == set(want)
    for key in ("heatmaps", "inds", "masks", "radius_map", "heatmap_mask"):
        assert len(got[key]) == len(want[key]) == 1
        g, w = got[key][0], want[key][0]
        assert g.shape == w.shape and g.dtype == w.dtype and g.device == w.device, key
        assert torch.equal(g, w), key
    assert float(want["masks"][0].sum()) > 10 and float(want["heatmaps"][0].max()) == 1.0
    g, w = got["target_boxes"][0], want["target_boxes"][0]
    assert g.shape == w.shape and torch.allclose(g, w, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("heads", [[["Vehicle"], ["Pedestrian", "Cyclist"]], [["Pedestrian", "Cyclist"], ["Vehicle"]]])
def test_assign_targets_multi_head_relabels_gt_in_place_like_the_reference(heads):
    """Several separate heads: the reference overwrites the class column of gt_boxes with the head-local id while it
    walks a head (curriculum_center_head.py:252-254), so the NEXT head — and the caller — see relabelled boxes (with
    the second head order pedestrians re-enter the Vehicle head as class 1).  Same targets, same mutated gt_boxes."""
    E = ref_py.EasyDict
    ch = ref_py.load("pcdet.models.dense_heads.curriculum_center_head")
    head = fake_head(ch, E)
    head.__dict__["class_names_each_head"] = heads
    gt, npgt, to, occ, fac = (t.cuda() for t in scene(2, 130, 21))
    grp = ops.centerhead_cluster_groups(gt, to, occ, fac)
    gt_ref, gt_got = gt.clone(), gt.clone()
    want = ch.CurriculumCenterHead.assign_targets.reference(head, gt_ref, feature_map_size=(188, 188), npgt=npgt, true_object=grp)
    assert center_targets.supported_head(head, gt_got)
    got = ch.CurriculumCenterHead.assign_targets(head, gt_got, feature_map_size=(188, 188), npgt=npgt, true_object=grp)
    assert torch.equal(gt_got, gt_ref) and not torch.equal(gt_ref, gt)        # the side effect, reproduced
    for key in ("heatmaps", "inds", "masks", "radius_map", "heatmap_mask"):
        assert len(got[key]) == len(want[key]) == 2
        for h in range(2):
            assert got[key][h].shape == want[key][h].shape and torch.equal(got[key][h], want[key][h]), (key, h)
    for h in range(2):
        assert torch.allclose(got["target_boxes"][h], want["target_boxes"][h], rtol=1e-6, atol=1e-6)
        assert float(want["masks"][h].sum()) > 5


@pytest.mark.parametrize("heads", [[list(NAMES)], [["Vehicle"], ["Pedestrian", "Cyclist"]]])
def test_plain_center_head_assign_targets_vs_reference_method(heads):
    """CenterHead.assign_targets (center_head.py:161-220, the head of the plain CenterPoint configurations): same
    kernel without the point filter and the group column; masks are int64 there."""
    E = ref_py.EasyDict
    chm = ref_py.load("pcdet.models.dense_heads.center_head")
    assert chm.CenterHead.assign_targets._comb
    head = object.__new__(chm.CenterHead)
    cfg = E(TARGET_ASSIGNER_CONFIG=E(FEATURE_MAP_STRIDE=8, NUM_MAX_OBJS=500, GAUSSIAN_OVERLAP=0.1, MIN_RADIUS=2))
    head.__dict__.update(model_cfg=cfg, class_names=NAMES, class_names_each_head=heads, point_cloud_range=RANGE, voxel_size=VSIZE)
    gt = scene(3, 160, 31, extra=2)[0].cuda()
    gt_ref, gt_got = gt.clone(), gt.clone()
    want = chm.CenterHead.assign_targets.reference(head, gt_ref, feature_map_size=(188, 188))
    got = chm.CenterHead.assign_targets(head, gt_got, feature_map_size=(188, 188))
    assert set(got) == set(want) and torch.equal(gt_got, gt_ref)
    for h in range(len(heads)):
        for key in ("heatmaps", "inds", "masks"):
            assert got[key][h].dtype == want[key][h].dtype and torch.equal(got[key][h], want[key][h]), (key, h)
        assert torch.allclose(got["target_boxes"][h], want["target_boxes"][h], rtol=1e-6, atol=1e-6)
    assert got["masks"][0].dtype == torch.int64 and int(want["masks"][0].sum()) > 10


def test_assign_targets_point_filter_is_epoch_gated():
    """MIN_POINTS drops sparse boxes only while epoch <= EPOCH_THRED (curriculum_center_head.py:167-168)."""
    E = ref_py.EasyDict
    ch = ref_py.load("pcdet.models.dense_heads.curriculum_center_head")
    gt, npgt, to, occ, fac = (t.cuda() for t in scene(2, 150, 7))
    for epoch in (0, 101):
        head = fake_head(ch, E, epoch=epoch, min_points=20)
        want = ch.CurriculumCenterHead.assign_targets.reference(head, gt.clone(), feature_map_size=(188, 188), npgt=npgt,
                                                                true_object=to.long())
        got = center_targets.assign_targets(head, gt.clone(), feature_map_size=(188, 188), npgt=npgt, true_object=to.long())
        for key in ("heatmaps", "inds", "masks", "radius_map"):
            assert torch.equal(got[key][0], want[key][0]), (epoch, key)


def test_empty_and_all_padding_ground_truth():
    """Frames without objects (M = 0) and frames whose gt rows are all padding (class 0): zero targets, as the
    reference produces; the loss runs on them (no positives: `loss = -neg_loss`)."""
    E = ref_py.EasyDict
    ch = ref_py.load("pcdet.models.dense_heads.curriculum_center_head")
    lu = ref_py.load("pcdet.utils.loss_utils")
    head = fake_head(ch, E)
    for M in (0, 6):
        gt = torch.zeros((2, M, 8), device="cuda")
        npgt = torch.zeros((2, M), device="cuda")
        grp = torch.zeros((2, M), dtype=torch.int64, device="cuda")
        want = ch.CurriculumCenterHead.assign_targets.reference(head, gt.clone(), feature_map_size=(188, 188), npgt=npgt, true_object=grp)
        got = center_targets.assign_targets(head, gt, feature_map_size=(188, 188), npgt=npgt, true_object=grp)
        for key in ("heatmaps", "inds", "masks", "radius_map", "heatmap_mask", "target_boxes"):
            assert got[key][0].shape == want[key][0].shape and torch.equal(got[key][0], want[key][0]), (M, key)
        assert ops.centerhead_cluster_groups(gt, npgt, npgt, npgt).shape == (2, M)
    pred = torch.rand((2, 3, 188, 188), device="cuda").clamp(1e-4, 1 - 1e-4)
    outs = []
    for fn in (lu.FocalLossCenterCurriculum.neg_loss.reference, lu.FocalLossCenterCurriculum.neg_loss):
        mod = loss_module(lu, E)
        outs.append(fn(mod, pred, got["heatmaps"][0], got["radius_map"][0], got["masks"][0].clone(),
                       mask=got["heatmap_mask"][0].clone(), epoch=1))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-6) and torch.equal(outs[0][1], outs[1][1])
    # the device side of a COMAug step with nothing to place
    from com_b200.pcdet_ops import box_ops
    idx, ex = box_ops.comaug_place_sampled_boxes(np.zeros((0, 7), np.float32), np.zeros((3, 7), np.float32))
    assert idx.shape == (0,) and ex.shape == (3, 7)


def loss_module(lu, E, **curriculum):
    cfg = E(LOSS_CURRICULUM=E(**curriculum))
    mod = lu.FocalLossCenterCurriculum(cfg, conf_shape=(3, 96))
    mod.avg_confidence = 0.05
    return mod


@pytest.mark.parametrize("curriculum,epoch", [({}, 3), ({"HEIGHT": 0.8, "ELONGATION": -6, "ADD": 1}, 3), ({"STRAIGHT": True, "K": 0.7}, 2),
                                              ({"CENTER": True}, 1), ({"RADIUS": 3, "FIX": True}, 5), ({"START": 10}, 3)])
def test_com_neg_loss_vs_reference_method(curriculum, epoch):
    E = ref_py.EasyDict
    lu = ref_py.load("pcdet.utils.loss_utils")
    ch = ref_py.load("pcdet.models.dense_heads.curriculum_center_head")
    assert lu.FocalLossCenterCurriculum.neg_loss._comb
    reference_method = lu.FocalLossCenterCurriculum.neg_loss.reference
    B, M = 2, 140
    head = fake_head(ch, E)
    gt, npgt, to, occ, fac = (t.cuda() for t in scene(B, M, 11))
    group = ops.centerhead_cluster_groups(gt, to, occ, fac)
    tg = center_targets.assign_targets(head, gt, feature_map_size=(188, 188), npgt=npgt, true_object=group)
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn((B, 3, 188, 188), generator=g) * 1.5 - 3.0).cuda().requires_grad_(True)

    def run(fn, mod):
        pred = torch.clamp(logits.sigmoid(), min=1e-4, max=1 - 1e-4)
        box_mask, mask = tg["masks"][0].clone(), tg["heatmap_mask"][0].clone()
        loss, bm, avg, ct, ca = fn(mod, pred, tg["heatmaps"][0], tg["radius_map"][0], box_mask, mask=mask, epoch=epoch)
        grad, = torch.autograd.grad(loss, logits)
        return loss.detach(), bm, mask, avg, ct, ca, grad, mod

    want = run(reference_method, loss_module(lu, E, **curriculum))
    got = run(lu.FocalLossCenterCurriculum.neg_loss, loss_module(lu, E, **curriculum))
    assert torch.allclose(got[0], want[0], rtol=1e-5, atol=1e-6)
    assert torch.allclose(got[1], want[1], rtol=1e-5, atol=1e-6) and torch.allclose(got[2], want[2], rtol=1e-5, atol=1e-6)
    if curriculum.get("START", 0) <= epoch:
        assert not torch.equal(want[1], tg["masks"][0])                    # the curriculum did re-weight
    assert abs(got[3] - want[3]) < 1e-7 and got[4:6] == want[4:6]
    assert torch.allclose(got[6], want[6], rtol=1e-4, atol=1e-7)
    assert abs(got[7].avg_confidence - want[7].avg_confidence) < 1e-9
    for a, b in zip(got[7].confidence_all, want[7].confidence_all):
        assert a.shape == b.shape == (3, 96) and torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert float(want[7].confidence_all[1].sum()) > 10
