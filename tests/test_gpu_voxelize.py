"""a1/a2/a4 parity: comb_voxelize (+ fused MeanVFE) through the C-ABI vs the CPU oracle — bit-exact
voxel coordinates, per-voxel counts, point assignment (first-T in point order) and voxel order
(first appearance, first `cap` voxels)."""
import numpy as np
import pytest
import torch

import oracle
from com_b200 import ops, synth, voxel
from util import WAYMO_RANGE, WAYMO_VSIZE

pytestmark = pytest.mark.gpu

SMALL_RANGE, SMALL_VS = [-4, -4, -1, 4, 4, 3], [0.25, 0.25, 0.5]


def run_gpu(frames, vs, rng, T, cap, **kw):
    offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(int).tolist()
    C = frames[0].shape[1]
    pts = torch.from_numpy(np.concatenate(frames, axis=0).reshape(-1, C)).cuda()
    r = ops.voxelize(pts, offs, vs, rng, T, cap, **kw)
    torch.cuda.synchronize()
    return r


def check_frames(frames, vs, rng, T, cap):
    r = run_gpu(frames, vs, rng, T, cap, mean_dtype=torch.float32)
    counts = r["counts"].cpu().numpy()
    base = 0
    for b, f in enumerate(frames):
        v, c, m = oracle.voxelize(f, vs, rng, T, cap)
        n = int(counts[b])
        assert n == len(c), "frame %d: %d voxels vs oracle %d" % (b, n, len(c))
        sl = slice(base, base + n)
        got_c = r["coords"][sl].cpu().numpy()
        assert (got_c[:, 0] == b).all()
        assert np.array_equal(got_c[:, 1:], c)
        assert np.array_equal(r["num_points"][sl].cpu().numpy(), m)
        assert np.array_equal(r["voxels"][sl].cpu().numpy().view(np.uint32), v.view(np.uint32))
        assert np.array_equal(r["mean"][sl].cpu().numpy().view(np.uint32), oracle.mean_vfe(v, m).view(np.uint32))
        base += n
    assert int(counts[-1]) == base
    return base


@pytest.mark.parametrize("n,cap,T", [(3000, 100000, 5), (3000, 200, 5), (5000, 100000, 1), (4000, 100000, 35),
                                     (1, 10, 5), (0, 10, 5), (20000, 700, 3)])
def test_small_cloud(n, cap, T):
    pts = synth.make_small_cloud(n, seed=n + cap + T)
    if n > 100:
        pts[::97, 0] = 100.0        # outside
        pts[5, 2] = np.nan
        pts[7, 0] = 4.0             # exactly on the max edge: falls out
        pts[9, 1] = -4.0            # exactly on the min edge: stays
    check_frames([pts], SMALL_VS, SMALL_RANGE, T, cap)


def test_all_points_in_one_voxel_and_all_outside():
    one = np.tile(np.array([[0.1, 0.1, 0.1, 0.5, 0.5]], np.float32), (5000, 1))
    one[:, 3] = np.arange(5000)
    assert check_frames([one], SMALL_VS, SMALL_RANGE, 5, 100) == 1
    out = np.full((300, 5), 50.0, np.float32)
    assert check_frames([out], SMALL_VS, SMALL_RANGE, 5, 100) == 0


def test_ragged_batch():
    frames = [synth.make_small_cloud(n, seed=s) for s, n in enumerate([4000, 0, 1500, 9000])]
    check_frames(frames, SMALL_VS, SMALL_RANGE, 5, 600)


def test_waymo_frame_full_size():
    pts = synth.make_frame(seed=1000)
    assert 150000 < len(pts) < 220000
    m = check_frames([pts], WAYMO_VSIZE, WAYMO_RANGE, 5, 150000)
    assert m > 50000


def test_waymo_cap_truncation_and_multiframe():
    pts = synth.make_frame(seed=1001, sweeps=2, beams=32, n_az=1500)
    assert pts.shape[1] == 6
    check_frames([pts, pts[::2].copy()], WAYMO_VSIZE, WAYMO_RANGE, 5, 20000)


def test_mean_only_bf16_padded():
    """The fused path of the pipeline: no voxel tensor, mean of channels [0,C) cast to bf16, padded to 16."""
    pts = synth.make_small_cloud(5000, seed=3)
    r = run_gpu([pts], SMALL_VS, SMALL_RANGE, 5, 100000, want_voxels=False, mean_dtype=torch.bfloat16, mean_ld=16)
    v, c, m = oracle.voxelize(pts, SMALL_VS, SMALL_RANGE, 5, 100000)
    n = int(r["counts"][1])
    want = torch.from_numpy(oracle.mean_vfe(v, m)).to(torch.bfloat16)
    got = r["mean"][:n].cpu()
    assert r["voxels"] is None and n == len(c)
    assert torch.equal(got[:, :5], want) and (got[:, 5:] == 0).all()


def test_mean_vfe_standalone():
    rng = np.random.default_rng(0)
    vox = rng.normal(size=(1000, 5, 5)).astype(np.float32)
    num = rng.integers(0, 6, 1000).astype(np.int32)
    for i in range(1000):
        vox[i, num[i]:] = 0
    want = oracle.mean_vfe(vox, num)
    for nt in (torch.from_numpy(num).cuda(), torch.from_numpy(num).float().cuda()):
        got = ops.mean_vfe(torch.from_numpy(vox).cuda(), nt).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_reference_wrapper_interface():
    """VoxelGeneratorWrapper.generate(points) -> numpy triple (data_processor.py:44-60)."""
    pts = synth.make_small_cloud(3000, seed=9)
    g = voxel.VoxelGeneratorWrapper(vsize_xyz=SMALL_VS, coors_range_xyz=SMALL_RANGE, num_point_features=5,
                                    max_num_points_per_voxel=5, max_num_voxels=500)
    v, c, m = g.generate(pts)
    v2, c2, m2 = oracle.voxelize(pts, SMALL_VS, SMALL_RANGE, 5, 500)
    assert isinstance(v, np.ndarray) and v.dtype == np.float32 and c.dtype == np.int32 and m.dtype == np.int32
    assert np.array_equal(v, v2) and np.array_equal(c, c2) and np.array_equal(m, m2)


def test_idempotent_and_deterministic():
    pts = synth.make_small_cloud(30000, seed=21)
    a = run_gpu([pts], SMALL_VS, SMALL_RANGE, 5, 2000)
    b = run_gpu([pts], SMALL_VS, SMALL_RANGE, 5, 2000)
    n = int(a["counts"][1])
    for k in ("voxels", "coords", "num_points"):
        assert torch.equal(a[k][:n], b[k][:n])


def test_device_side_frame_offsets():
    """Frame offsets read from device memory (CUDA-graph friendly launch sequence): same bits as host offsets,
    also when the point buffer is larger than the frames it holds."""
    frames = [synth.make_small_cloud(n, seed=40 + i) for i, n in enumerate([4000, 0, 1500, 9000])]
    offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int32)
    pts = np.concatenate(frames)
    padded = np.concatenate([pts, np.full((3000, 5), 0.123, np.float32)])      # stale rows beyond off[batch]
    a = ops.voxelize(torch.from_numpy(pts).cuda(), offs.tolist(), SMALL_VS, SMALL_RANGE, 5, 600, mean_dtype=torch.float32)
    b = ops.voxelize(torch.from_numpy(padded).cuda(), torch.from_numpy(offs).cuda(), SMALL_VS, SMALL_RANGE, 5, 600,
                     mean_dtype=torch.float32)
    assert torch.equal(a["counts"], b["counts"])
    n = int(a["counts"][-1])
    for k in ("voxels", "coords", "num_points", "mean"):
        assert torch.equal(a[k][:n], b[k][:n])


_sweep3 = {}


@pytest.mark.parametrize("cap", [100000, 180000, 400000])
def test_config5_three_sweep_frame_full_size(cap):
    """BASELINE configs[4]: 3-sweep aggregated frame (~500k points x 6 features incl. the timestamp channel) with the
    multiframe voxel caps of waymo_dataset_multiframe.yaml:83-89 (and a smaller one that truncates) — bit-exact vs the oracle (coordinates, first-5
    point assignment, counts, first-`cap` voxels in order of first appearance)."""
    if "pts" not in _sweep3:
        _sweep3["pts"] = synth.make_frame(seed=1005, sweeps=3)
    pts = _sweep3["pts"]
    assert pts.shape[1] == 6 and 420000 < len(pts) < 600000
    m = check_frames([pts], WAYMO_VSIZE, WAYMO_RANGE, 5, cap)
    assert m == min(cap, 144913)          # 144 913 voxels on this frame: the 100 000 cap truncates, the others do not
