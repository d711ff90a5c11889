"""Shared helpers of the test-suite (oracle = checker, never the product)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

WAYMO_RANGE = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]
WAYMO_VSIZE = [0.1, 0.1, 0.15]


def canon_pairs(nbr, in_coords, out_coords):
    """Ordering-independent form of a rulebook: sorted array of (k, in b,z,y,x, out b,z,y,x)."""
    K, no = nbr.shape
    ks, os_ = np.nonzero(nbr >= 0)
    rows = nbr[ks, os_]
    t = np.concatenate([ks[:, None], in_coords[rows], out_coords[os_]], axis=1).astype(np.int64)
    order = np.lexsort(t.T[::-1])
    return t[order]


def random_coords(rng, n, batch, shape):
    """n unique (b,z,y,x) int32 rows in random order."""
    D, H, W = shape
    vol = batch * D * H * W
    n = min(n, vol)
    keys = rng.choice(vol, size=n, replace=False)
    x = keys % W
    r = keys // W
    y = r % H
    r //= H
    z = r % D
    b = r // D
    return np.stack([b, z, y, x], axis=1).astype(np.int32)


def clustered_coords(rng, n, batch, shape, clusters=20, spread=3.0):
    """Unique coords bunched in clusters (dense neighbourhoods -> rich rulebooks)."""
    D, H, W = shape
    ctr = np.stack([rng.integers(0, batch, clusters), rng.integers(0, D, clusters), rng.integers(0, H, clusters),
                    rng.integers(0, W, clusters)], axis=1)
    pick = ctr[rng.integers(0, clusters, n * 2)]
    jit = np.rint(rng.normal(0, spread, size=(n * 2, 3))).astype(np.int64)
    c = pick.copy()
    c[:, 1:] += jit
    ok = (c[:, 1] >= 0) & (c[:, 1] < D) & (c[:, 2] >= 0) & (c[:, 2] < H) & (c[:, 3] >= 0) & (c[:, 3] < W)
    c = c[ok]
    _, first = np.unique((((c[:, 0] * D + c[:, 1]) * H + c[:, 2]) * W + c[:, 3]), return_index=True)
    c = c[np.sort(first)][:n]
    return c.astype(np.int32)
