"""a10-a16 parity: dense() scatter, points_in_boxes (both semantics), rotated BEV IoU and NMS through
the C-ABI vs the CPU oracle / golden vectors of the reference / the reference's own CUDA build."""
import os

import numpy as np
import pytest
import torch

import oracle
from com_b200 import ops, synth
from com_b200.pcdet_ops import box_ops
from oracle import build_ref
from util import GOLDEN, random_coords

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "box_ops_ref.npz"))


@pytest.mark.parametrize("shape", [[2, 47, 45], [2, 188, 188], [3, 20, 36]])    # DHW % 4 != 0 and == 0 (vector path)
@pytest.mark.parametrize("C,dtype", [(128, torch.float32), (128, torch.bfloat16), (5, torch.float32), (33, torch.float32)])
def test_dense(C, dtype, shape):
    rng = np.random.default_rng(C)
    batch = 3
    coords = random_coords(rng, 2500, batch, shape)
    feats = rng.normal(size=(len(coords), C)).astype(np.float32)
    f = cuda(feats).to(dtype)
    got = ops.dense(f, cuda(coords), batch, shape).cpu().numpy()
    want = oracle.dense(f.float().cpu().numpy(), coords, batch, shape)
    assert got.shape == (batch, C, *shape) and np.array_equal(got, want)
    empty = ops.dense(f[:0], cuda(coords[:0]), batch, shape)
    assert float(empty.abs().sum()) == 0.0


def test_points_in_boxes_cpu_semantics_golden(gold):
    boxes, pts = gold["pib_boxes"], gold["pib_points"]
    want = np.unpackbits(gold["pib_mask_packed"], axis=1)[:, : len(pts)].astype(np.int32)
    got = box_ops.points_in_boxes_cpu(pts, boxes)           # numpy in -> numpy out like the reference wrapper
    assert isinstance(got, np.ndarray) and got.dtype == np.int32 and np.array_equal(got, want)


def test_points_in_boxes_cpu_semantics_waymo_size():
    pts = synth.make_frame(seed=1002)[:, :3].copy()
    boxes = synth.make_boxes(500, seed=0)
    boxes[:, 2] = -1.0
    want = oracle.points_in_boxes_cpu(pts, boxes)
    got = box_ops.points_in_boxes_cpu(torch.from_numpy(pts), torch.from_numpy(boxes))
    assert want.sum() > 1000 and np.array_equal(got.numpy(), want)
    kept = box_ops.remove_points_in_boxes3d(pts, boxes[:35])
    assert len(kept) == len(pts) - int((want[:35].sum(0) != 0).sum())


def test_bev_iou_cpu_semantics_golden(gold):
    for name in ("uc", "cc", "self", "known"):
        got = box_ops.boxes_bev_iou_cpu(gold["iou_%s_a" % name], gold["iou_%s_b" % name])
        want = gold["iou_%s" % name]
        bad = int((got.view(np.uint32) != want.view(np.uint32)).sum())
        assert np.array_equal(got == 0, want == 0), name        # what COMAug consumes (== 0 test) must be exact
        assert bad == 0, "%s: %d of %d IoU values differ in bits (max abs %.3g)" % (
            name, bad, want.size, np.abs(got - want).max())


def test_bev_iou_cpu_semantics_500x500():
    a = synth.make_clustered_boxes(500, seed=21)
    want = oracle.boxes_bev_cpu(a, a)
    got = box_ops.boxes_bev_iou_cpu(a, a)
    assert (want > 0).sum() > 2000
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_nms_cpu_flavour_keep_list():
    boxes = synth.make_clustered_boxes(500, seed=22)
    for thresh in (0.1, 0.7):
        want = oracle.nms_cpu(boxes, thresh)
        trig = cuda(ops.box_trig4_host(boxes))
        keep, num = ops.nms(cuda(boxes), thresh, rotated=True, flavour="cpu", trig=trig)
        got = keep[: int(num)].cpu().numpy()
        assert np.array_equal(got, want) and 0 < len(want) < 500


ref_cuda = pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref (reference CUDA build) not present")


@ref_cuda
def test_gpu_flavour_vs_reference_cuda_iou_and_nms():
    """Device flavour vs the reference's own kernels (iou3d_nms_kernel.cu) compiled for sm_100."""
    ref = build_ref.load_ref("ref_iou3d_nms_cuda")
    a, b = cuda(synth.make_clustered_boxes(300, seed=31)), cuda(synth.make_clustered_boxes(400, seed=32))
    for fn_ref, fn in ((ref.boxes_iou_bev_gpu, box_ops.boxes_iou_bev),):
        want = torch.zeros((300, 400), device="cuda")
        fn_ref(a, b, want)
        got = fn(a, b)
        assert (want > 0).sum() > 1000
        assert torch.equal(got, want), "%d IoU values differ" % int((got != want).sum())
    boxes = cuda(synth.make_clustered_boxes(1000, seed=33))
    scores = torch.from_numpy(np.random.default_rng(3).uniform(size=1000).astype(np.float32)).cuda()
    order = scores.sort(0, descending=True)[1]
    sb = boxes[order].contiguous()
    for thresh in (0.1, 0.7):
        keep = torch.zeros(1000, dtype=torch.int64)
        n = ref.nms_gpu(sb, keep, thresh)
        want = order[keep[:n].cuda()]
        got, _ = box_ops.nms_gpu(boxes, scores, thresh)
        assert torch.equal(got, want)
        keep = torch.zeros(1000, dtype=torch.int64)
        n = ref.nms_normal_gpu(sb, keep, thresh)
        got, _ = box_ops.nms_normal_gpu(boxes, scores, thresh)
        assert torch.equal(got, order[keep[:n].cuda()])


@ref_cuda
def test_points_in_boxes_gpu_vs_reference_cuda():
    ref = build_ref.load_ref("ref_roiaware_pool3d_cuda")
    rng = np.random.default_rng(41)
    boxes = np.stack([synth.make_boxes(60, seed=s) for s in (1, 2)])
    pts = (boxes[:, rng.integers(0, 60, 30000), :3] + rng.normal(0, 1.5, size=(2, 30000, 3))).astype(np.float32)
    want = torch.full((2, 30000), -1, dtype=torch.int32, device="cuda")
    ref.points_in_boxes_gpu(cuda(boxes), cuda(pts), want)
    got = box_ops.points_in_boxes_gpu(cuda(pts), cuda(boxes))
    assert (want >= 0).sum() > 3000 and torch.equal(got, want)


def test_nms_wrapper_matches_model_nms_utils_contract():
    """class_agnostic_nms call pattern (model_nms_utils.py:6-25): topk -> nms_gpu -> indices."""
    boxes = cuda(synth.make_clustered_boxes(800, seed=51))
    scores = torch.from_numpy(np.random.default_rng(5).uniform(size=800).astype(np.float32)).cuda()
    top_scores, idx = torch.topk(scores, k=500)
    keep, _ = box_ops.nms_gpu(boxes[idx][:, :7], top_scores, 0.7)
    assert keep.dtype == torch.int64 and keep.is_cuda and keep.max() < 500
    sel = boxes[idx][keep].cpu().numpy()
    iou = oracle.boxes_bev_cpu(sel, sel)
    np.fill_diagonal(iou, 0)
    assert iou.max() <= 0.7 + 1e-3            # survivors do not suppress each other
    empty, _ = box_ops.nms_gpu(boxes[:0], scores[:0], 0.7)
    assert empty.numel() == 0


def test_config4_comaug_collision_checks_full_size():
    """BASELINE configs[3]: COMAug GT sampling — 10 000 candidate boxes against 100 existing boxes (bit-exact vs the
    oracle, which finishes this in < 1 s) and against each other (10k x 10k = 1e8 pairs: 60 random 64x64 blocks
    bit-exact vs the oracle, plus size-independent properties of the whole matrix), then points_in_boxes removal of
    the surviving <= 35 boxes on a 180k-point frame."""
    cand = np.concatenate([synth.make_clustered_boxes(9000, seed=61, centers=400), synth.make_boxes(1000, seed=64)]).astype(np.float32)
    exist = synth.make_clustered_boxes(100, seed=62, centers=400)
    got = box_ops.boxes_bev_iou_cpu(cand, exist)
    want = oracle.boxes_bev_cpu(cand, exist)
    assert (want > 0).sum() > 500
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    full = box_ops.boxes_bev_iou_cpu(cand, cand)                      # (10000, 10000) float32, 400 MB
    assert full.shape == (10000, 10000)
    rng = np.random.default_rng(63)
    for _ in range(60):
        i0, j0 = rng.integers(0, 10000 - 64, size=2)
        if rng.uniform() < 0.5:
            j0 = i0                                                   # diagonal blocks hold the dense overlaps
        blk = oracle.boxes_bev_cpu(cand[i0:i0 + 64], cand[j0:j0 + 64])
        assert np.array_equal(full[i0:i0 + 64, j0:j0 + 64].view(np.uint32), blk.view(np.uint32))
    d = np.diagonal(full)
    assert np.all(np.abs(d - 1.0) < 1e-4)                             # IoU(box, box) = 1 (SURVEY §8c known answer)
    assert np.all(full >= 0) and np.all(full <= 1.0 + 1e-4)
    # boxes whose centres are further apart than the sum of their half diagonals cannot overlap: exactly 0
    c = cand[:2000]
    rad = 0.5 * np.hypot(c[:, 3], c[:, 4])
    dist = np.hypot(c[:, None, 0] - c[None, :, 0], c[:, None, 1] - c[None, :, 1])
    far = dist > (rad[:, None] + rad[None, :]) + 0.05
    assert np.all(full[:2000, :2000][far] == 0)
    # collision-free survivors (what database_sampler_v2.py:600-604 keeps), then remove their points from the frame
    ok = (got.max(axis=1) == 0)
    iou_self = full.copy()
    np.fill_diagonal(iou_self, 0)
    survivors = cand[ok & (iou_self.max(axis=1) == 0)][:35]
    assert len(survivors) > 0
    pts = synth.make_frame(seed=1003)
    survivors[:, 2] = -1.0
    kept = box_ops.remove_points_in_boxes3d(pts, survivors)
    inside = oracle.points_in_boxes_cpu(pts[:, :3].copy(), survivors)
    assert np.array_equal(kept, pts[inside.sum(0) == 0])


def test_nms_api_maximum_size():
    """nms_gpu at the API maximum of class_agnostic_nms (NMS_PRE_MAXSIZE 4096, centerpoint.yaml:61-69): CPU flavour
    vs the oracle's greedy sweep at 4096 boxes, and idempotence (NMS of the kept set keeps everything)."""
    boxes = synth.make_clustered_boxes(4096, seed=71, centers=40)
    trig = cuda(ops.box_trig4_host(boxes))
    for thresh in (0.1, 0.7):
        want = oracle.nms_cpu(boxes, thresh)
        keep, num = ops.nms(cuda(boxes), thresh, rotated=True, flavour="cpu", trig=trig)
        got = keep[: int(num)].cpu().numpy()
        assert np.array_equal(got, want) and 100 < len(want) < 4000
        kb = boxes[got]
        keep2, num2 = ops.nms(cuda(kb), thresh, rotated=True, flavour="cpu", trig=cuda(ops.box_trig4_host(kb)))
        assert int(num2) == len(kb) and np.array_equal(keep2[: int(num2)].cpu().numpy(), np.arange(len(kb)))


@pytest.mark.parametrize("C,dtype", [(128, torch.bfloat16), (128, torch.float32), (5, torch.float32), (33, torch.float32)])
def test_dense_scatter_matches_dense(C, dtype):
    """comb_dense_scatter into a pre-zeroed tensor == comb_dense (bit-exact), incl. a device-side row count."""
    rng = np.random.default_rng(C)
    shape = [2, 47, 45]
    coords = random_coords(rng, 3000, 3, shape)
    coords = coords[np.lexsort((coords[:, 3], coords[:, 2], coords[:, 1], coords[:, 0]))]
    feats = cuda(rng.normal(size=(len(coords), C)).astype(np.float32)).to(dtype)
    want = ops.dense(feats, cuda(coords), 3, shape)
    out = torch.zeros_like(want)
    got = ops.dense_scatter(feats, cuda(coords), 3, shape, out)
    assert torch.equal(got, want)
    n_dev = torch.tensor([1234], dtype=torch.int32, device="cuda")
    out2 = ops.dense_scatter(feats, cuda(coords), 3, shape, torch.zeros_like(want), n_dev=n_dev)
    assert torch.equal(out2, ops.dense(feats[:1234].contiguous(), cuda(coords[:1234]), 3, shape))


# ---- f4: the BEV image in channels-last bf16 ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dense_nhwc_bf16_equals_dense_view_cast(dtype):
    """comb_dense_scatter_nhwc_bf16 == dense().view(N, C*D, H, W).to(bf16) in channels_last memory, bit for bit, and its
    adjoint gathers exactly the cells of the rows (HeightCompression, height_compression.py:21-24)."""
    from com_b200 import ops
    rng = np.random.default_rng(3)
    B, D, H, W, C, n = 3, 2, 47, 52, 128, 2500
    cells = rng.choice(B * D * H * W, size=n, replace=False)
    coords = np.stack(np.unravel_index(cells, (B, D, H, W)), axis=1).astype(np.int32)
    feats = torch.from_numpy(rng.normal(size=(n, C)).astype(np.float32)).cuda().to(dtype)
    cd = torch.from_numpy(coords).cuda()
    want = ops.dense(feats, cd, B, [D, H, W]).view(B, C * D, H, W).to(torch.bfloat16)
    got = ops.dense_nhwc_bf16(feats, cd, B, [D, H, W])
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    assert got.is_contiguous(memory_format=torch.channels_last) and torch.equal(got, want)
    # adjoint, fp32 and bf16 gradients, any memory format
    g = torch.from_numpy(rng.normal(size=(B, C * D, H, W)).astype(np.float32)).cuda()
    want_rows = ops.dense_gather(g.view(B, C, D, H, W).contiguous(), cd, dtype=torch.float32)
    for grad in (g, g.contiguous(memory_format=torch.channels_last)):
        assert torch.equal(ops.dense_gather_nhwc(grad, cd, C, D, dtype=torch.float32), want_rows)
    gb = g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.dense_gather_nhwc(gb, cd, C, D, dtype=torch.float32), ops.dense_gather(
        gb.float().view(B, C, D, H, W).contiguous(), cd, dtype=torch.float32))
    # autograd
    f2 = feats.clone().requires_grad_(True)
    out = ops.DenseNHWCFunction.apply(f2, cd, B, [D, H, W])
    (out.float() * g).sum().backward()
    assert torch.allclose(f2.grad.float(), want_rows.to(dtype).float(), rtol=1e-2, atol=1e-2)
