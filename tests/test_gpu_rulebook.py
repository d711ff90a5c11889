"""a6 parity: coordinate hash table, strided-conv output set and rulebook through the C-ABI vs the
CPU oracle — bit-exact index pairs.  SubM keeps the input row order (fixed by semantics); strided
outputs come in canonical ascending-key order, so nbr tables compare with array_equal; the
ordering-independent (k, in_coord, out_coord) triple set is compared as well (SURVEY hard part 1)."""
import numpy as np
import pytest
import torch

import oracle
from com_b200 import ops, synth
from util import WAYMO_RANGE, WAYMO_VSIZE, canon_pairs, clustered_coords, random_coords

pytestmark = pytest.mark.gpu


def gpu_rulebook(coords, batch, shape, ks, st, pd, dl, subm):
    c = torch.from_numpy(coords).cuda()
    table, slots = ops.hash_build(c, batch, shape)
    if subm:
        out_c, out_shape = c, shape
        pd = [(k // 2) * d for k, d in zip(ks, dl)]
        st = [1, 1, 1]
    else:
        out_shape = ops.conv_out_shape(shape, ks, st, pd, dl)
        cap = max(len(coords) * int(np.prod(ks)), 1)
        out_c, cnt = ops.conv_out_coords(c, batch, out_shape, ks, st, pd, dl, cap)
        out_c = out_c[: int(cnt.item())]
    nbr = ops.nbrmap_build(out_c, table, slots, batch, shape, ks, st, pd, dl)
    torch.cuda.synchronize()
    return out_c, out_shape, nbr


CASES = [
    dict(ks=(3, 3, 3), st=(1, 1, 1), pd=(1, 1, 1), subm=True),
    dict(ks=(3, 3, 3), st=(2, 2, 2), pd=(1, 1, 1), subm=False),
    dict(ks=(3, 3, 3), st=(2, 2, 2), pd=(0, 1, 1), subm=False),
    dict(ks=(3, 1, 1), st=(2, 1, 1), pd=(0, 0, 0), subm=False),
    dict(ks=(1, 3, 3), st=(1, 1, 1), pd=(0, 1, 1), subm=True),
    dict(ks=(2, 2, 2), st=(2, 2, 2), pd=(0, 0, 0), subm=False),
]


@pytest.mark.parametrize("cv", CASES)
@pytest.mark.parametrize("gen", ["clustered", "random", "single", "empty"])
def test_rulebook_vs_oracle(cv, gen):
    rng = np.random.default_rng(5)
    batch, shape = 3, [11, 40, 37]
    coords = {"clustered": lambda: clustered_coords(rng, 4000, batch, shape, 12, 2.5),
              "random": lambda: random_coords(rng, 3000, batch, shape),
              "single": lambda: np.array([[2, 10, 39, 36]], np.int32),
              "empty": lambda: np.zeros((0, 4), np.int32)}[gen]()
    ks, st, pd, dl = cv["ks"], cv["st"], cv["pd"], (1, 1, 1)
    out_c, out_shape, nbr = gpu_rulebook(coords, batch, shape, ks, st, pd, dl, cv["subm"])
    if cv["subm"]:
        want_c, want_nbr = coords, oracle.subm_nbrmap(coords, shape, ks)
    else:
        want_shape = oracle.conv_out_shape(shape, ks, st, pd, dl)
        assert list(out_shape) == want_shape
        want_c = oracle.conv_out_coords(coords, want_shape, ks, st, pd, dl)
        want_nbr = oracle.nbrmap(want_c, coords, shape, ks, st, pd, dl)
    got_c, got_nbr = out_c.cpu().numpy(), nbr.cpu().numpy()
    assert np.array_equal(got_c, want_c)
    assert np.array_equal(got_nbr, want_nbr)
    if len(coords):
        assert np.array_equal(canon_pairs(got_nbr, coords, got_c), canon_pairs(want_nbr, coords, want_c))


def test_pairs_and_transpose_layout():
    rng = np.random.default_rng(6)
    batch, shape = 2, [9, 30, 30]
    coords = clustered_coords(rng, 3000, batch, shape, 8, 2.0)
    ks, st, pd, dl = (3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 1, 1)
    out_c, out_shape, nbr = gpu_rulebook(coords, batch, shape, ks, st, pd, dl, False)
    nbr_np = nbr.cpu().numpy()
    pairs, num = ops.nbrmap_to_pairs(nbr)
    pairs, num = pairs.cpu().numpy(), num.cpu().numpy()
    for k in range(27):
        o = np.nonzero(nbr_np[k] >= 0)[0]
        assert num[k] == len(o)
        assert np.array_equal(pairs[1, k, : len(o)], o) and np.array_equal(pairs[0, k, : len(o)], nbr_np[k, o])
        assert (pairs[:, k, len(o):] == -1).all()
    nbr_t = ops.nbrmap_transpose(nbr, len(coords)).cpu().numpy()
    want_t = np.full((27, len(coords)), -1, np.int32)
    for k in range(27):
        o = np.nonzero(nbr_np[k] >= 0)[0]
        want_t[k, nbr_np[k, o]] = o          # for a fixed k every input row feeds at most one output
    assert np.array_equal(nbr_t, want_t)


def test_device_side_row_count():
    """n_dev smaller than the allocated rows: rows beyond it are ignored (sync-free chaining)."""
    rng = np.random.default_rng(8)
    batch, shape = 1, [8, 20, 20]
    coords = random_coords(rng, 1000, batch, shape)
    padded = np.concatenate([coords, np.full((200, 4), 3, np.int32)])
    c = torch.from_numpy(padded).cuda()
    n_dev = torch.tensor([1000], dtype=torch.int32, device="cuda")
    table, slots = ops.hash_build(c, batch, shape, n_dev=n_dev)
    nbr = ops.nbrmap_build(c, table, slots, batch, shape, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], no_dev=n_dev)
    assert np.array_equal(nbr[:, :1000].cpu().numpy(), oracle.subm_nbrmap(coords, shape))


def test_waymo_frame_full_size():
    pts = synth.make_frame(seed=1000)
    _, c, _ = oracle.voxelize(pts, WAYMO_VSIZE, WAYMO_RANGE, 5, 150000)
    coords = np.concatenate([np.zeros((len(c), 1), np.int32), c], axis=1)
    shape = [41, 1504, 1504]
    out_c, _, nbr = gpu_rulebook(coords, 1, shape, (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1), True)
    assert np.array_equal(nbr.cpu().numpy(), oracle.subm_nbrmap(coords, shape))
    ks, st, pd, dl = (3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 1, 1)
    out_c, out_shape, nbr = gpu_rulebook(coords, 1, shape, ks, st, pd, dl, False)
    want_c = oracle.conv_out_coords(coords, out_shape, ks, st, pd, dl)
    assert list(out_shape) == [21, 752, 752] and np.array_equal(out_c.cpu().numpy(), want_c)
    assert np.array_equal(nbr.cpu().numpy(), oracle.nbrmap(want_c, coords, shape, ks, st, pd, dl))
