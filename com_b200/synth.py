"""Seeded synthetic Waymo-shaped LiDAR frames and boxes.

A spinning 64-beam sensor (elevation -17.6..+2.4 deg x 2650 azimuth steps = 169 600 rays) plus four
short-range side sensors, ray-cast against a ground plane, 100 object cuboids and 20 wall segments;
range noise, tanh(intensity), elongation.  The frame is then range-masked in x,y (inclusive max
edge, like common_utils.mask_points_by_range, pcdet/utils/common_utils.py:60-63) and shuffled
(data_processor.py:103-113) with the same seed.  Data only: no reference code is involved.
"""
import numpy as np

POINT_CLOUD_RANGE = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]
VOXEL_SIZE = [0.1, 0.1, 0.15]
MAX_POINTS_PER_VOXEL = 5
MAX_NUMBER_OF_VOXELS = 150000
GRID_SIZE = [1504, 1504, 40]          # x, y, z


def _ray_box(orig, dirs, boxes):
    """Nearest positive hit distance of rays (R,3) from `orig` (3,) with rotated cuboids (B,7)."""
    R = dirs.shape[0]
    best = np.full((R,), np.inf, dtype=np.float64)
    for b in boxes:
        c, s = np.cos(-b[6]), np.sin(-b[6])
        ox, oy = orig[0] - b[0], orig[1] - b[1]
        o = np.array([ox * c - oy * s, ox * s + oy * c, orig[2] - b[2]])
        d = np.stack([dirs[:, 0] * c - dirs[:, 1] * s, dirs[:, 0] * s + dirs[:, 1] * c, dirs[:, 2]], axis=1)
        half = b[3:6] / 2.0
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-half - o) / d
            t2 = (half - o) / d
        tmin = np.nanmax(np.minimum(t1, t2), axis=1)
        tmax = np.nanmin(np.maximum(t1, t2), axis=1)
        hit = (tmax >= tmin) & (tmax > 0.5)
        t = np.where(tmin > 0.5, tmin, tmax)
        best = np.where(hit & (t < best), t, best)
    return best


def make_scene(rng, rng_xy=70.0):
    def place(n, dims, jitter):
        xy = rng.uniform(-rng_xy, rng_xy, size=(n, 2))
        d = np.asarray(dims)[None, :] * rng.uniform(1 - jitter, 1 + jitter, size=(n, 3))
        z = -1.7 + d[:, 2] / 2
        h = rng.uniform(-np.pi, np.pi, size=(n,))
        return np.concatenate([xy, z[:, None], d, h[:, None]], axis=1)

    cars = place(60, (4.5, 1.9, 1.6), 0.15)
    peds = place(30, (0.8, 0.8, 1.7), 0.1)
    cycs = place(10, (1.8, 0.8, 1.7), 0.1)
    walls = place(20, (1.0, 0.3, 1.0), 0.0)
    walls[:, 3] = rng.uniform(10, 40, size=20)
    walls[:, 5] = rng.uniform(3, 8, size=20)
    walls[:, 2] = -1.7 + walls[:, 5] / 2
    return np.concatenate([cars, peds, cycs, walls], axis=0)


def _scan(rng, scene, origin, elev_deg, n_az, max_range):
    el = np.deg2rad(np.asarray(elev_deg))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False) + rng.uniform(0, 2 * np.pi / n_az)
    e, a = np.meshgrid(el, az, indexing="ij")
    dirs = np.stack([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)], axis=-1).reshape(-1, 3)
    t_obj = _ray_box(origin, dirs, scene)
    with np.errstate(divide="ignore"):
        t_gnd = np.where(dirs[:, 2] < -1e-6, (-1.7 - origin[2]) / dirs[:, 2], np.inf)
    t = np.minimum(t_obj, t_gnd)
    ok = np.isfinite(t) & (t < max_range)
    t = t[ok] + rng.normal(0, 0.02, size=int(ok.sum()))
    return origin[None, :] + dirs[ok] * t[:, None]


def make_frame(seed=1000, sweeps=1, beams=64, n_az=2650, side_rays=2500):
    """-> points (N, 5) float32 [x,y,z,intensity,elongation] (or (N,6) with a timestamp when sweeps>1)."""
    rng = np.random.default_rng(seed)
    scene = make_scene(rng)
    pts_all = []
    for k in range(sweeps):
        ego = np.array([0.5 * k, 0.0, 0.0])
        xyz = [_scan(rng, scene, ego, np.linspace(-17.6, 2.4, beams), n_az, 75.0)]
        side_el = np.linspace(-60.0, 20.0, 25)
        for off in ((1.5, 0.0), (-1.5, 0.0), (0.0, 0.8), (0.0, -0.8)):
            o = ego + np.array([off[0], off[1], -0.8])
            xyz.append(_scan(rng, scene, o, side_el, max(side_rays // 25, 1), 20.0))
        xyz = np.concatenate(xyz, axis=0)
        n = xyz.shape[0]
        cols = [xyz, np.tanh(rng.uniform(0, 1, size=(n, 1))), rng.uniform(0, 1.5, size=(n, 1))]
        if sweeps > 1:
            cols.append(np.full((n, 1), 0.1 * k))
        pts_all.append(np.concatenate(cols, axis=1))
    pts = np.concatenate(pts_all, axis=0).astype(np.float32)
    r = POINT_CLOUD_RANGE
    m = (pts[:, 0] >= r[0]) & (pts[:, 0] <= r[3]) & (pts[:, 1] >= r[1]) & (pts[:, 1] <= r[4])
    pts = pts[m]
    return pts[rng.permutation(pts.shape[0])]


def make_boxes(n, seed=0, rng_xy=75.2):
    """Car-sized boxes (n,7) uniformly placed (same recipe as the survey's probe, seed 0)."""
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-rng_xy, rng_xy, size=(n, 2))
    z = rng.uniform(-1.0, 1.0, size=(n, 1))
    d = np.stack([rng.uniform(3.5, 5.5, n), rng.uniform(1.6, 2.2, n), rng.uniform(1.4, 2.0, n)], axis=1)
    h = rng.uniform(-np.pi, np.pi, size=(n, 1))
    return np.concatenate([xy, z, d, h], axis=1).astype(np.float32)


def make_clustered_boxes(n, seed=0, spread=6.0, centers=40, center_seed=12345):
    """Boxes bunched around a few centres so that many pairs overlap (NMS / IoU stress).
    The centres depend on `center_seed` only, so sets drawn with different seeds overlap each other."""
    rng = np.random.default_rng(seed)
    ctr = np.random.default_rng(center_seed).uniform(-60, 60, size=(centers, 2))
    xy = ctr[rng.integers(0, centers, size=n)] + rng.normal(0, spread / 3, size=(n, 2))
    z = rng.uniform(-1.0, 1.0, size=(n, 1))
    d = np.stack([rng.uniform(3.5, 5.5, n), rng.uniform(1.6, 2.2, n), rng.uniform(1.4, 2.0, n)], axis=1)
    h = rng.uniform(-np.pi, np.pi, size=(n, 1))
    return np.concatenate([xy, z, d, h], axis=1).astype(np.float32)


def make_small_cloud(n, seed=0, extent=(8.0, 8.0, 3.0), channels=5):
    """Small dense random cloud for fast tests (ragged voxels, many points per voxel)."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1, 1, size=(n, 3)) * np.asarray(extent)[None, :] / 2
    xyz[:, 2] += 1.0
    rest = rng.uniform(0, 1, size=(n, channels - 3))
    return np.concatenate([xyz, rest], axis=1).astype(np.float32)
