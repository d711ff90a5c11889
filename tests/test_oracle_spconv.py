"""The spconv segment of the oracle is "parity unpinned" (spconv is not vendored / installed and the
reference ships no fixtures).  What CAN be checked is checked here: the restatement against
independent statements of the same published semantics —
  * voxel generator vs a dictionary-based pure-Python walk;
  * sparse conv (SubM and strided) vs torch.nn.functional.conv3d on the zero-filled dense tensor
    (spconv's defining property: a SparseConv3d equals the dense cross-correlation at the active
    output sites; a SubMConv3d equals it restricted to the input sites);
  * dgrad / wgrad vs torch autograd of that dense conv;
  * dense() vs numpy scatter; output shapes vs the arithmetic of SURVEY.md §8(a6)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from com_b200 import synth
from util import clustered_coords, random_coords


def py_voxelize(points, vsize, rng, T, cap):
    vs, rg = np.asarray(vsize, np.float32), np.asarray(rng, np.float32)
    grid = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
    table, coords, vox, num = {}, [], [], []
    for p in points:
        c = np.floor((p[:3] - rg[:3]) / vs)
        if np.any(c < 0) or np.any(c >= grid) or np.any(np.isnan(c)):
            continue
        key = (int(c[2]), int(c[1]), int(c[0]))
        vid = table.get(key)
        if vid is None:
            if len(coords) >= cap:
                continue
            vid = len(coords)
            table[key] = vid
            coords.append(key)
            vox.append(np.zeros((T, points.shape[1]), np.float32))
            num.append(0)
        if num[vid] < T:
            vox[vid][num[vid]] = p
            num[vid] += 1
    return (np.stack(vox) if vox else np.zeros((0, T, points.shape[1]), np.float32),
            np.asarray(coords, np.int32).reshape(-1, 3), np.asarray(num, np.int32))


@pytest.mark.parametrize("n,cap,T", [(3000, 100000, 5), (3000, 200, 5), (2000, 100000, 1), (0, 10, 5)])
def test_voxelize_vs_python_walk(n, cap, T):
    pts = synth.make_small_cloud(n, seed=n + cap)
    if n:
        pts[::97, 0] = 100.0       # out of range
        pts[5, 2] = np.nan
        pts[7, 0] = 4.0            # exactly on the max edge -> rejected
    rng, vs = [-4, -4, -1, 4, 4, 3], [0.25, 0.25, 0.5]
    v, c, m = oracle.voxelize(pts, vs, rng, T, cap)
    v2, c2, m2 = py_voxelize(pts, vs, rng, T, cap)
    assert np.array_equal(c, c2) and np.array_equal(m, m2)
    assert np.array_equal(v.view(np.uint32), v2.view(np.uint32))
    assert len(c) <= cap


def test_voxelize_waymo_frame_properties():
    pts = synth.make_frame(seed=1000, beams=16, n_az=600, side_rays=200)
    v, c, m = oracle.voxelize(pts, synth.VOXEL_SIZE, synth.POINT_CLOUD_RANGE, 5, 150000)
    assert len(np.unique(c, axis=0)) == len(c) and (m >= 1).all() and (m <= 5).all()
    assert (c >= 0).all() and (c[:, 0] < 40).all() and (c[:, 1:] < 1504).all()
    # first point of every voxel is its first point in scan order
    first = {}
    g = np.floor((pts[:, :3] - np.float32([-75.2, -75.2, -2])) / np.float32(synth.VOXEL_SIZE)).astype(np.int64)
    ok = (g >= 0).all(1) & (g[:, 0] < 1504) & (g[:, 1] < 1504) & (g[:, 2] < 40)
    for i in np.nonzero(ok)[0]:
        first.setdefault((g[i, 2], g[i, 1], g[i, 0]), i)
    assert len(first) == len(c)
    for row in range(0, len(c), 97):
        assert np.array_equal(v[row, 0], pts[first[tuple(c[row])]])


def test_backbone_shapes():
    s = [41, 1504, 1504]
    s2 = oracle.conv_out_shape(s, (3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 1, 1))
    s3 = oracle.conv_out_shape(s2, (3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 1, 1))
    s4 = oracle.conv_out_shape(s3, (3, 3, 3), (2, 2, 2), (0, 1, 1), (1, 1, 1))
    s5 = oracle.conv_out_shape(s4, (3, 1, 1), (2, 1, 1), (0, 0, 0), (1, 1, 1))
    assert (s2, s3, s4, s5) == ([21, 752, 752], [11, 376, 376], [5, 188, 188], [2, 188, 188])


def _dense_from(feats, coords, batch, shape):
    return torch.from_numpy(oracle.dense(feats, coords, batch, shape)).double()


CONVS = [
    dict(ks=(3, 3, 3), st=(1, 1, 1), pd=(1, 1, 1), subm=True),
    dict(ks=(3, 3, 3), st=(2, 2, 2), pd=(1, 1, 1), subm=False),
    dict(ks=(3, 3, 3), st=(2, 2, 2), pd=(0, 1, 1), subm=False),
    dict(ks=(3, 1, 1), st=(2, 1, 1), pd=(0, 0, 0), subm=False),
    dict(ks=(1, 3, 3), st=(1, 1, 1), pd=(0, 1, 1), subm=True),
]


@pytest.mark.parametrize("cv", CONVS)
def test_conv_vs_dense_torch(cv):
    rng = np.random.default_rng(7)
    batch, shape, Cin, Cout = 2, [9, 12, 14], 5, 6
    coords = clustered_coords(rng, 300, batch, shape, clusters=6, spread=2.0)
    feats = rng.normal(size=(len(coords), Cin)).astype(np.float32)
    ks, st, pd, dl = cv["ks"], cv["st"], cv["pd"], (1, 1, 1)
    K = int(np.prod(ks))
    W = rng.normal(size=(Cout, K, Cin)).astype(np.float32)
    if cv["subm"]:
        out_shape, out_coords = shape, coords
        nbr = oracle.subm_nbrmap(coords, shape, ks)
    else:
        out_shape = oracle.conv_out_shape(shape, ks, st, pd, dl)
        out_coords = oracle.conv_out_coords(coords, out_shape, ks, st, pd, dl)
        nbr = oracle.nbrmap(out_coords, coords, shape, ks, st, pd, dl)
    got = oracle.conv_fwd(feats, W, nbr)

    x = _dense_from(feats, coords, batch, shape).requires_grad_(True)
    wt = torch.from_numpy(W).double().reshape(Cout, *ks, Cin).permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    y = F.conv3d(x, wt, stride=st, padding=pd)
    assert list(y.shape[2:]) == list(out_shape)
    oc = torch.from_numpy(out_coords).long()
    want = y[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]]
    assert np.allclose(got, want.detach().numpy(), rtol=1e-5, atol=1e-5)
    if not cv["subm"]:
        # the output set is exactly the support of the dense conv of the occupancy
        occ = (_dense_from(np.ones((len(coords), 1), np.float32), coords, batch, shape))
        sup = F.conv3d(occ, torch.ones((1, 1) + tuple(ks), dtype=torch.double), stride=st, padding=pd)[:, 0] > 0
        assert int(sup.sum()) == len(out_coords)
        assert sup[oc[:, 0], oc[:, 1], oc[:, 2], oc[:, 3]].all()
        key = ((out_coords[:, 0].astype(np.int64) * out_shape[0] + out_coords[:, 1]) * out_shape[1]
               + out_coords[:, 2]) * out_shape[2] + out_coords[:, 3]
        assert (np.diff(key) > 0).all()          # canonical ascending order

    # backward: dgrad / wgrad against autograd of the dense conv restricted to the active outputs
    dout = rng.normal(size=got.shape).astype(np.float32)
    (want * torch.from_numpy(dout).double()).sum().backward()
    ic = torch.from_numpy(coords).long()
    din_want = x.grad[ic[:, 0], :, ic[:, 1], ic[:, 2], ic[:, 3]].numpy()
    dw_want = wt.grad.permute(0, 2, 3, 4, 1).reshape(Cout, K, Cin).numpy()
    assert np.allclose(oracle.conv_dgrad(dout, W, nbr, len(coords)), din_want, rtol=1e-5, atol=1e-5)
    assert np.allclose(oracle.conv_wgrad(feats, dout, nbr), dw_want, rtol=1e-5, atol=1e-5)


def test_dense_and_mean_vfe():
    rng = np.random.default_rng(1)
    coords = random_coords(rng, 50, 2, [2, 6, 7])
    feats = rng.normal(size=(50, 4)).astype(np.float32)
    d = oracle.dense(feats, coords, 2, [2, 6, 7])
    assert d.shape == (2, 4, 2, 6, 7) and np.count_nonzero(d) == 200
    assert np.array_equal(d[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]], feats)
    vox = rng.normal(size=(30, 5, 4)).astype(np.float32)
    num = rng.integers(0, 6, 30).astype(np.int32)
    for i in range(30):
        vox[i, num[i]:] = 0
    want = torch.from_numpy(vox).sum(1) / torch.clamp_min(torch.from_numpy(num).float().view(-1, 1), 1.0)
    assert np.array_equal(oracle.mean_vfe(vox, num), want.numpy())    # mean_vfe.py:26-29


def test_fast_cpu_baseline_matches_oracle():
    """bench.py's CPU baseline (oracle/cpu_fast.c) computes the same conv as the fp64 oracle."""
    rng = np.random.default_rng(2)
    coords = clustered_coords(rng, 1500, 2, [10, 30, 30], 8, 2.0)
    nbr = oracle.subm_nbrmap(coords, [10, 30, 30])
    feats = rng.normal(size=(len(coords), 16)).astype(np.float32)
    W = rng.normal(size=(32, 27, 16)).astype(np.float32) / 20
    bias, scale, shift = [rng.normal(size=(32,)).astype(np.float32) for _ in range(3)]
    res = rng.normal(size=(len(coords), 32)).astype(np.float32)
    want = np.maximum(oracle.conv_fwd(feats, W, nbr, bias).astype(np.float64) * scale + shift + res, 0)
    got = oracle.fast_conv_fwd(feats, W, nbr, bias, scale, shift, res, relu=True)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5
    assert oracle.fast().orc_fast_threads() >= 1
