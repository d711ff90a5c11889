"""Pin oracle/oracle.c's box-op restatement: bit-exact against (a) golden vectors generated from the
reference's compiled CPU entry points (tests/golden/box_ops_ref.npz, made by make_golden.py) and
(b) the reference itself when oracle/_ref is present (iou3d_cpu.cpp:232-252, roiaware_pool3d.cpp:143-168)."""
import os

import numpy as np
import pytest

import oracle
from com_b200 import synth
from oracle import build_ref
from util import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "box_ops_ref.npz"))


@pytest.mark.parametrize("name", ["uc", "cc", "self", "known"])
def test_iou_golden_bit_exact(gold, name):
    got = oracle.boxes_bev_cpu(gold["iou_%s_a" % name], gold["iou_%s_b" % name])
    want = gold["iou_%s" % name]
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_iou_known_answers(gold):
    """IoU(self)=1.0000, IoU(box, box shifted (1,0.5))=0.4906, disjoint=0."""
    got = oracle.boxes_bev_cpu(gold["iou_known_a"], gold["iou_known_b"])[0]
    assert abs(got[0] - 1.0) < 1e-4
    assert abs(got[1] - 0.4906) < 1e-4
    assert got[2] == 0.0


def test_iou_golden_has_overlaps(gold):
    # the fixture must exercise the polygon code, not only the disjoint early path
    assert (gold["iou_cc"] > 0).sum() > 200
    assert (gold["iou_self"].diagonal() > 0.99).all()


def test_points_in_boxes_golden_bit_exact(gold):
    got = oracle.points_in_boxes_cpu(gold["pib_points"], gold["pib_boxes"])
    want = np.unpackbits(gold["pib_mask_packed"], axis=1)[:, : got.shape[1]].astype(np.int32)
    assert int(gold["pib_hits"][0]) == want.sum() > 500
    assert np.array_equal(got, want)


def test_points_in_boxes_margin_edge_cases():
    """MARGIN 1e-2 in x/y (strict <), none in z (reject only when |dz| > dz/2): roiaware_pool3d.cpp:131-138."""
    box = np.array([[0, 0, 0, 4, 2, 2, 0]], dtype=np.float32)
    pts = np.array([[2.005, 0, 0], [2.02, 0, 0], [0, 1.005, 0], [0, 1.02, 0], [0, 0, 1.0], [0, 0, 1.0001],
                    [0, 0, -1.0]], dtype=np.float32)
    assert oracle.points_in_boxes_cpu(pts, box)[0].tolist() == [1, 0, 1, 0, 1, 0, 1]


def test_nms_sweep_matches_bruteforce():
    boxes = synth.make_clustered_boxes(200, seed=5)
    iou = oracle.boxes_bev_cpu(boxes, boxes)
    keep, sup = [], np.zeros(200, dtype=bool)
    for i in range(200):
        if sup[i]:
            continue
        keep.append(i)
        sup[i + 1:] |= iou[i, i + 1:] > 0.3
    got = oracle.nms_cpu(boxes, 0.3)
    assert got.tolist() == keep
    assert 0 < len(keep) < 200
    assert oracle.nms_cpu(boxes[:0], 0.3).tolist() == []


@pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not built (no /root/reference)")
def test_against_compiled_reference():
    import torch
    iou = build_ref.load_ref("ref_iou3d_nms_cuda")
    roi = build_ref.load_ref("ref_roiaware_pool3d_cuda")
    a, b = synth.make_clustered_boxes(150, seed=11), synth.make_clustered_boxes(170, seed=12)
    want = torch.zeros((150, 170))
    iou.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), want)
    got = oracle.boxes_bev_cpu(a, b)
    assert np.array_equal(got.view(np.uint32), want.numpy().view(np.uint32))
    rng = np.random.default_rng(13)
    pts = (a[rng.integers(0, 150, 20000), :3] + rng.normal(0, 1.5, size=(20000, 3))).astype(np.float32)
    m = torch.zeros((150, 20000), dtype=torch.int32)
    roi.points_in_boxes_cpu(torch.from_numpy(a), torch.from_numpy(pts), m)
    assert np.array_equal(oracle.points_in_boxes_cpu(pts, a), m.numpy())
