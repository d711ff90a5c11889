"""CPU oracle of the COM hot path — TEST INFRASTRUCTURE, never imported by com_b200.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
See oracle/oracle.c for what is restated, what pins it, and which part is "parity unpinned".
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    fsrc = os.path.join(_HERE, "cpu_fast.c")
    lfast = os.path.join(_HERE, "liboracle_fast.so")
    if force or not os.path.exists(lfast) or os.path.getmtime(lfast) < os.path.getmtime(fsrc):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_fast.so"], stdout=subprocess.DEVNULL)
    return _LIB


_LIB_FAST = os.path.join(_HERE, "liboracle_fast.so")
_fast = None


def fast():
    """CPU-baseline library (OpenMP fp32 sparse conv) — bench.py's cpu_baseline / reference arm."""
    global _fast
    if _fast is None:
        src = os.path.join(_HERE, "cpu_fast.c")
        if not os.path.exists(_LIB_FAST) or os.path.getmtime(_LIB_FAST) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_fast.so"], stdout=subprocess.DEVNULL)
        _fast = ctypes.CDLL(_LIB_FAST)
        _fast.orc_fast_threads.restype = ctypes.c_int
    return _fast


def fast_conv_fwd(feats, weight, nbr, bias=None, scale=None, shift=None, residual=None, relu=False):
    x, w, nb = _f(feats), _f(weight), _i(nbr)
    Cout, K, Cin = w.shape
    no = nb.shape[1]
    out = np.empty((no, Cout), dtype=np.float32)
    opt = [(_f(t) if t is not None else None) for t in (bias, scale, shift, residual)]
    fast().orc_fast_conv_fwd(_p(x), Cin, _p(w), K, Cout, _p(nb), no, _p(opt[0]), _p(opt[1]), _p(opt[2]), _p(opt[3]),
                             int(bool(relu)), _p(out))
    return out


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_voxelize.restype = ctypes.c_int
        _lib.orc_conv_out_coords.restype = ctypes.c_int
        _lib.orc_nms_cpu.restype = ctypes.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _i3(v):
    if isinstance(v, int):
        v = (v, v, v)
    return (ctypes.c_int * 3)(*[int(x) for x in v])


_lut_cache = {}


def voxelize(points, vsize, rng, max_points, max_voxels):
    """-> voxels (M,T,C) f32, coords (M,3) zyx i32, num (M,) i32 — spconv generator semantics."""
    pts = _f(points)
    n, C = pts.shape
    vs = _f(vsize)
    rg = _f(rng)
    grid = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
    vol = int(grid.prod())
    lut = _lut_cache.get(vol)
    if lut is None:
        lut = np.full((vol,), -1, dtype=np.int32)
        _lut_cache.clear()
        _lut_cache[vol] = lut
    voxels = np.empty((max_voxels, max_points, C), dtype=np.float32)
    coords = np.empty((max_voxels, 3), dtype=np.int32)
    num = np.empty((max_voxels,), dtype=np.int32)
    m = lib().orc_voxelize(_p(pts), n, C, _p(vs), _p(rg), int(max_points), int(max_voxels), _p(voxels), _p(coords),
                           _p(num), _p(lut))
    return voxels[:m].copy(), coords[:m].copy(), num[:m].copy()


def mean_vfe(voxels, num):
    v = _f(voxels)
    M, T, C = v.shape
    out = np.empty((M, C), dtype=np.float32)
    n = _i(num)
    lib().orc_mean_vfe(_p(v), _p(n), M, T, C, _p(out))
    return out


def conv_out_shape(in_shape, ks, st, pd, dl):
    return [(int(i) + 2 * p - d * (k - 1) - 1) // s + 1 for i, k, s, p, d in zip(in_shape, ks, st, pd, dl)]


def conv_out_coords(in_coords, out_shape, ks, st, pd, dl, cap=None):
    c = _i(in_coords)
    n = c.shape[0]
    K = int(np.prod(ks))
    cap = int(cap if cap is not None else max(n * K, 1))
    out = np.empty((cap, 4), dtype=np.int32)
    m = lib().orc_conv_out_coords(_p(c), n, _i3(out_shape), _i3(ks), _i3(st), _i3(pd), _i3(dl), _p(out), cap)
    return out[:m].copy()


def nbrmap(out_coords, in_coords, in_shape, ks, st, pd, dl):
    oc, ic = _i(out_coords), _i(in_coords)
    K = int(np.prod(ks))
    nbr = np.empty((K, oc.shape[0]), dtype=np.int32)
    lib().orc_nbrmap(_p(oc), oc.shape[0], _p(ic), ic.shape[0], _i3(in_shape), _i3(ks), _i3(st), _i3(pd), _i3(dl),
                     _p(nbr))
    return nbr


def subm_nbrmap(coords, shape, ks=(3, 3, 3), dl=(1, 1, 1)):
    pd = [(k // 2) * d for k, d in zip(ks, dl)]
    return nbrmap(coords, coords, shape, ks, (1, 1, 1), pd, dl)


def conv_fwd(feats, weight, nbr, bias=None):
    """weight (Cout,K,Cin); fp64 accumulation."""
    x, w, nb = _f(feats), _f(weight), _i(nbr)
    Cout, K, Cin = w.shape
    no = nb.shape[1]
    out = np.empty((no, Cout), dtype=np.float32)
    b = _f(bias) if bias is not None else None
    lib().orc_conv_fwd(_p(x), Cin, _p(w), K, Cout, _p(nb), no, _p(b), _p(out))
    return out


def conv_dgrad(dout, weight, nbr, ni):
    g, w, nb = _f(dout), _f(weight), _i(nbr)
    Cout, K, Cin = w.shape
    din = np.empty((ni, Cin), dtype=np.float32)
    lib().orc_conv_dgrad(_p(g), Cout, _p(w), K, Cin, _p(nb), nb.shape[1], int(ni), _p(din))
    return din


def conv_wgrad(feats, dout, nbr):
    x, g, nb = _f(feats), _f(dout), _i(nbr)
    K, no = nb.shape
    Cin, Cout = x.shape[1], g.shape[1]
    dw = np.empty((Cout, K, Cin), dtype=np.float32)
    lib().orc_conv_wgrad(_p(x), Cin, _p(g), Cout, K, _p(nb), no, _p(dw))
    return dw


def dense(feats, coords, batch, shape):
    x, c = _f(feats), _i(coords)
    D, H, W = [int(v) for v in shape]
    C = x.shape[1]
    out = np.empty((batch, C, D, H, W), dtype=np.float32)
    lib().orc_dense(_p(x), _p(c), x.shape[0], int(batch), C, D, H, W, _p(out))
    return out


def points_in_boxes_cpu(points, boxes):
    """(P,3),(Nb,7) -> (Nb,P) int32 (reference argument order of the Python wrapper)."""
    p, b = _f(points), _f(boxes)
    out = np.empty((b.shape[0], p.shape[0]), dtype=np.int32)
    lib().orc_points_in_boxes_cpu(_p(b), b.shape[0], _p(p), p.shape[0], _p(out))
    return out


def boxes_bev_cpu(a, b, what="iou"):
    a, b = _f(a), _f(b)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().orc_boxes_bev_cpu(_p(a), a.shape[0], _p(b), b.shape[0], 0 if what == "iou" else 1, _p(out))
    return out


def nms_cpu(boxes_sorted, thresh):
    b = _f(boxes_sorted)
    keep = np.empty((max(b.shape[0], 1),), dtype=np.int64)
    n = lib().orc_nms_cpu(_p(b), b.shape[0], ctypes.c_float(thresh), _p(keep))
    return keep[:n].copy()
