"""a6 parity (fused-pipeline flavour): bitmap-rank grid index through the C-ABI vs the CPU oracle.
Rows live in canonical order (ascending key); coordinate sets and rulebooks are bit-exact."""
import numpy as np
import pytest
import torch

import oracle
from com_b200 import ops, synth
from util import WAYMO_RANGE, WAYMO_VSIZE, clustered_coords, random_coords

pytestmark = pytest.mark.gpu


def key_of(c, shape):
    c = c.astype(np.int64)
    return ((c[:, 0] * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3]


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


CASES = [
    dict(ks=(3, 3, 3), st=(2, 2, 2), pd=(1, 1, 1)),
    dict(ks=(3, 3, 3), st=(2, 2, 2), pd=(0, 1, 1)),
    dict(ks=(3, 1, 1), st=(2, 1, 1), pd=(0, 0, 0)),
    dict(ks=(2, 2, 2), st=(2, 2, 2), pd=(0, 0, 0)),
    dict(ks=(3, 3, 3), st=(1, 1, 1), pd=(1, 1, 1)),
]


@pytest.mark.parametrize("gen", ["clustered", "random", "dense", "single", "empty"])
def test_self_index_sorts_and_ranks(gen):
    rng = np.random.default_rng(3)
    batch, shape = 3, [11, 40, 37]
    coords = {"clustered": lambda: clustered_coords(rng, 4000, batch, shape, 12, 2.5),
              "random": lambda: random_coords(rng, 3000, batch, shape),
              "dense": lambda: random_coords(rng, 3 * 11 * 40 * 37, batch, shape),
              "single": lambda: np.array([[2, 10, 39, 36]], np.int32),
              "empty": lambda: np.zeros((0, 4), np.int32)}[gen]()
    idx = ops.index_build(cuda(coords), batch, shape)
    n = int(idx.count)
    order = np.argsort(key_of(coords, shape), kind="stable")
    assert n == len(coords)
    assert np.array_equal(idx.coords[:n].cpu().numpy(), coords[order])
    rows = ops.index_rank(cuda(coords), idx).cpu().numpy()
    want = np.empty(len(coords), np.int64)
    want[order] = np.arange(len(coords))
    assert np.array_equal(rows, want)
    if len(coords):
        # absent and out-of-grid coordinates rank to -1
        probe = np.array([[0, 0, 0, 0], [batch, 0, 0, 0], [0, shape[0], 0, 0], [-1, 0, 0, 0]], np.int32)
        got = ops.index_rank(cuda(probe), idx).cpu().numpy()
        present = set(map(tuple, coords.tolist()))
        assert got[1] == got[2] == got[3] == -1 and ((got[0] >= 0) == ((0, 0, 0, 0) in present))


@pytest.mark.parametrize("cv", CASES)
@pytest.mark.parametrize("gen", ["clustered", "random"])
def test_conv_output_set_and_rulebook(cv, gen):
    rng = np.random.default_rng(4)
    batch, shape = 2, [11, 40, 37]
    coords = clustered_coords(rng, 4000, batch, shape, 12, 2.5) if gen == "clustered" else random_coords(rng, 3000, batch, shape)
    coords = coords[np.argsort(key_of(coords, shape))]        # the fused pipeline keeps every level in key order
    ks, st, pd, dl = cv["ks"], cv["st"], cv["pd"], (1, 1, 1)
    in_idx = ops.index_build(cuda(coords), batch, shape)
    out_shape = oracle.conv_out_shape(shape, ks, st, pd, dl)
    cap = len(coords) * 27
    out_idx = ops.index_build(cuda(coords), batch, out_shape, conv=(ks, st, pd, dl), out_cap=cap)
    want_c = oracle.conv_out_coords(coords, out_shape, ks, st, pd, dl)
    n = int(out_idx.count)
    assert n == len(want_c) and np.array_equal(out_idx.coords[:n].cpu().numpy(), want_c)
    nbr = ops.nbrmap_build_indexed(out_idx.coords[:n].contiguous(), in_idx, ks, st, pd, dl).cpu().numpy()
    assert np.array_equal(nbr, oracle.nbrmap(want_c, coords, shape, ks, st, pd, dl))
    # submanifold rulebook of the output level against its own index
    sub = ops.nbrmap_build_indexed(out_idx.coords[:n].contiguous(), out_idx, (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1))
    assert np.array_equal(sub.cpu().numpy(), oracle.subm_nbrmap(want_c, out_shape))


def test_out_cap_clamps_count():
    rng = np.random.default_rng(5)
    coords = random_coords(rng, 2000, 1, [9, 30, 30])
    idx = ops.index_build(cuda(coords), 1, [5, 15, 15], conv=((3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 1, 1)), out_cap=100)
    assert int(idx.count) == 100
    want = oracle.conv_out_coords(coords, [5, 15, 15], (3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 1, 1))
    assert np.array_equal(idx.coords[:100].cpu().numpy(), want[:100])


def test_permute_rows_roundtrip():
    rng = np.random.default_rng(6)
    x = torch.from_numpy(rng.normal(size=(5000, 16)).astype(np.float32)).cuda().to(torch.bfloat16)
    perm = torch.from_numpy(rng.permutation(5000).astype(np.int32)).cuda()
    y = ops.permute_rows(x, perm, scatter=True)
    assert torch.equal(y[perm.long()], x)
    assert torch.equal(ops.permute_rows(y, perm, scatter=False), x)
    perm2 = perm.clone()
    perm2[::7] = -1
    z = ops.permute_rows(y, perm2, scatter=False)
    assert (z[::7] == 0).all() and torch.equal(z[1::7], x[1::7])


def test_waymo_frame_full_size():
    pts = synth.make_frame(seed=1000)
    _, c, _ = oracle.voxelize(pts, WAYMO_VSIZE, WAYMO_RANGE, 5, 150000)
    shape = [41, 1504, 1504]
    coords = np.concatenate([np.zeros((len(c), 1), np.int32), c], axis=1)
    idx = ops.index_build(cuda(coords), 1, shape)
    order = np.argsort(key_of(coords, shape))
    sc = coords[order]
    assert int(idx.count) == len(c) and np.array_equal(idx.coords[: len(c)].cpu().numpy(), sc)
    nbr = ops.nbrmap_build_indexed(idx.coords[: len(c)].contiguous(), idx, (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1))
    assert np.array_equal(nbr.cpu().numpy(), oracle.subm_nbrmap(sc, shape))
    ks, st, pd, dl = (3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 1, 1)
    o2 = ops.index_build(idx.coords[: len(c)].contiguous(), 1, [21, 752, 752], conv=(ks, st, pd, dl), out_cap=len(c) * 8)
    want = oracle.conv_out_coords(sc, [21, 752, 752], ks, st, pd, dl)
    n2 = int(o2.count)
    assert n2 == len(want) and np.array_equal(o2.coords[:n2].cpu().numpy(), want)
    nbr2 = ops.nbrmap_build_indexed(o2.coords[:n2].contiguous(), idx, ks, st, pd, dl)
    assert np.array_equal(nbr2.cpu().numpy(), oracle.nbrmap(want, sc, shape, ks, st, pd, dl))


def test_rank_scatter_gives_key_ordered_rows_without_emit():
    """Level-1 path of the fused backbone: index built WITHOUT coordinate emission (prefix finalisation only), rows in
    key order produced by comb_index_rank_scatter from the (unique) voxel list — same rows, same ranks as the emit path,
    incl. a device-side row count smaller than the buffer."""
    rng = np.random.default_rng(8)
    batch, shape = 2, [21, 90, 77]
    coords = random_coords(rng, 20000, batch, shape)
    full = ops.index_build(cuda(coords), batch, shape)
    lean = ops.index_build(cuda(coords), batch, shape, want_coords=False)
    rows, sorted_c = ops.index_rank(cuda(coords), lean, scatter_coords=True)
    n = int(lean.count)
    assert n == int(full.count) == len(coords)
    assert torch.equal(sorted_c[:n], full.coords[:n])
    assert torch.equal(rows, ops.index_rank(cuda(coords), full))
    nblk = (batch * 21 * 90 * 77 + 255) // 256          # rank blocks that hold cells: their global prefixes agree
    assert torch.equal(lean.prefix.view(torch.int32)[:nblk], full.prefix.view(torch.int32)[:nblk])
    nd = torch.tensor([12345], dtype=torch.int32, device="cuda")
    lean2 = ops.index_build(cuda(coords), batch, shape, want_coords=False, n_dev=nd)
    rows2, sorted2 = ops.index_rank(cuda(coords), lean2, n_dev=nd, scatter_coords=True)
    ref2 = ops.index_build(cuda(coords[:12345]), batch, shape)
    assert int(lean2.count) == 12345 and torch.equal(sorted2[:12345], ref2.coords[:12345])
    assert torch.equal(rows2[:12345], ops.index_rank(cuda(coords[:12345]), ref2))
