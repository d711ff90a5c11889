"""Torch-facing functional wrappers over the C-ABI (device pointers + current stream).

PyTorch is plumbing here: it owns device memory and the stream; all arithmetic happens in
libcomb200.so.  Every function raises if a tensor is not a contiguous CUDA tensor — there is no
CPU fallback.
"""
import ctypes

import numpy as np
import torch
import torch.utils.data

from . import _lib
from ._lib import DT_BF16, DT_F32, EPI_AFFINE, EPI_BIAS, EPI_RELU, EPI_RESIDUAL, check, int3

__all__ = [
    "voxelize", "mean_vfe", "hash_build", "conv_out_coords", "conv_out_shape", "nbrmap_build",
    "nbrmap_transpose", "nbrmap_to_pairs", "spconv_fwd_f32", "spconv_dgrad_f32", "spconv_wgrad_f32", "spconv_wgrad_bf16",
    "pack_weight_bf16", "spconv_fwd_bf16", "affine_relu", "cast_pad", "bn_train_fwd", "bn_train_bwd", "col_sum", "dense", "dense_gather", "DenseFunction", "points_in_boxes_mask",
    "points_in_any_box", "points_in_boxes_index", "boxes_bev", "nms", "centerhead_decode_nms", "dense_nhwc_bf16", "dense_gather_nhwc", "comaug_valid_mask", "centerhead_assign_targets", "centerhead_cluster_groups", "comloss_group_confidence", "comloss_reweight", "box_trig_host", "box_trig4_host",
]


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


# Optional profiler hook (bench.py): an object with begin(tag, **info) / end() called around each
# kernel family on the launching stream.  None in production.
_prof = None


def set_profiler(p):
    global _prof
    _prof = p


class _Scope:
    __slots__ = ("on",)

    def __init__(self, tag, **info):
        self.on = _prof is not None
        if self.on:
            _prof.begin(tag, **info)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.on:
            _prof.end()
        return False


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need(t, dtype, name):
    if t is None:
        return
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (libcomb200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def host_op_device():
    """Device for the `*_cpu`-named entry points (CPU tensors in and out, arithmetic on the GPU).

    Their reference callers run inside DataLoader worker processes (pcdet/datasets/augmentor/
    database_sampler_v2.py:600-604, pcdet/utils/box_utils.py:117-131).  Policy:
      * `spawn` / `forkserver` workers (what the reference selects under --launcher pytorch/slurm,
        pcdet/utils/common_utils.py:172-173): a CUDA context is created lazily in the worker on its rank's GPU
        (LOCAL_RANK, else the current device) the first time an op runs;
      * a worker FORKED from a process that already initialised CUDA cannot use the GPU at all: raise with the remedy
        instead of dying inside the driver.  There is no CPU fallback."""
    import os
    if torch.cuda._is_in_bad_fork():
        raise RuntimeError(
            "com_b200: this process was forked from a parent that had already initialised CUDA, so its `*_cpu` box ops "
            "(points_in_boxes_cpu, boxes_iou_bev_cpu, remove_points_in_boxes3d) cannot reach the GPU.  Start DataLoader "
            "workers with the 'spawn' start method (DataLoader(..., multiprocessing_context='spawn') or "
            "torch.multiprocessing.set_start_method('spawn'), which tools/train.py --launcher pytorch already does), or "
            "run COMAug in the main process (num_workers=0).  com_b200 has no CPU fallback.")
    if not torch.cuda.is_available():
        raise RuntimeError("com_b200 box ops need a CUDA device (no CPU fallback)")
    try:
        in_worker = torch.utils.data.get_worker_info() is not None
    except Exception:
        in_worker = False
    if in_worker and "LOCAL_RANK" in os.environ:
        return torch.device("cuda", int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())
    return torch.device("cuda", torch.cuda.current_device())


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _dt(t):
    if t.dtype == torch.float32:
        return DT_F32
    if t.dtype == torch.bfloat16:
        return DT_BF16
    raise RuntimeError("unsupported dtype %s" % t.dtype)


# --------------------------------------------------------------------------------------------- voxelization
def voxelize(points, frame_offsets, vsize_xyz, range_xyz, max_points, max_voxels, want_voxels=True,
             mean_dtype=None, mean_c0=0, mean_ld=None):
    """Batched hard voxelization (+ optional fused MeanVFE).  See comb_voxelize in include/comb200.h.

    points: (N_total, C) fp32 CUDA, frames concatenated; frame_offsets: python ints, len batch+1.
    Returns dict(voxels, coords (cap,4) b,z,y,x, num_points, counts (batch+1) device int32, mean).
    Rows >= counts[-1] of every output are undefined.
    """
    lib = _lib.load()
    _need(points, torch.float32, "points")
    n_total, C = int(points.shape[0]), int(points.shape[1])
    offs_dev = None
    if isinstance(frame_offsets, torch.Tensor):
        # device-side offsets (int32, batch+1): points.shape[0] is the capacity, off[batch] <= capacity
        _need(frame_offsets, torch.int32, "frame_offsets")
        offs_dev = frame_offsets
        batch = int(frame_offsets.shape[0]) - 1
    else:
        batch = len(frame_offsets) - 1
        assert frame_offsets[-1] == n_total, "frame_offsets[-1] must equal the number of points"
    dev = points.device
    cap = batch * max_voxels
    voxels = torch.empty((cap, max_points, C), dtype=torch.float32, device=dev) if want_voxels else None
    coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    num = torch.empty((cap,), dtype=torch.int32, device=dev)
    counts = torch.empty((batch + 1,), dtype=torch.int32, device=dev)
    mean = None
    mdt = DT_F32
    if mean_dtype is not None:
        mean_ld = mean_ld or (C - mean_c0)
        mean = torch.empty((cap, mean_ld), dtype=mean_dtype, device=dev)
        mdt = _dt(mean)
    ws_bytes = lib.comb_voxelize_workspace_bytes(n_total, batch, max_voxels, max_points)
    ws = _ws(ws_bytes, dev)
    offs = None if offs_dev is not None else (ctypes.c_int * (batch + 1))(*[int(x) for x in frame_offsets])
    vs = (ctypes.c_float * 3)(*[float(np.float32(x)) for x in vsize_xyz])
    rg = (ctypes.c_float * 6)(*[float(np.float32(x)) for x in range_xyz])
    with _Scope("voxelize", n=n_total, C=C, T=int(max_points), counts=counts, batch=batch,
                want_voxels=want_voxels, mean_ld=int(mean_ld or 0), mean_bytes=(mean.element_size() if mean is not None else 0)):
        check(lib.comb_voxelize(_p(points), offs, _p(offs_dev), n_total, batch, C, vs, rg, int(max_points), int(max_voxels), _p(voxels),
                                _p(coords), _p(num), _p(mean), mdt, int(mean_c0), int(mean_ld or 1), _p(counts), _p(ws),
                                ws.numel(), _stream()), "comb_voxelize")
    return dict(voxels=voxels, coords=coords, num_points=num, counts=counts, mean=mean)


def mean_vfe(voxels, num_points):
    lib = _lib.load()
    _need(voxels, torch.float32, "voxels")
    M, T, C = voxels.shape
    if num_points.dtype == torch.float32:
        is_f = 1
    elif num_points.dtype == torch.int32:
        is_f = 0
    else:
        num_points, is_f = num_points.to(torch.int32), 0
    _need(num_points, num_points.dtype, "num_points")
    out = torch.empty((M, C), dtype=torch.float32, device=voxels.device)
    check(lib.comb_mean_vfe(_p(voxels), _p(num_points), is_f, M, T, C, _p(out), _stream()), "comb_mean_vfe")
    return out


# --------------------------------------------------------------------------------------------- rulebook
def hash_build(coords, batch, shape, n_dev=None):
    """Hash table over (b,z,y,x) rows. Returns (table uint8 tensor, slots)."""
    lib = _lib.load()
    _need(coords, torch.int32, "coords")
    _need(n_dev, torch.int32, "n_dev")
    n = int(coords.shape[0])
    slots = lib.comb_hash_slots(n)
    table = torch.empty((slots * 8,), dtype=torch.uint8, device=coords.device)
    D, H, W = [int(x) for x in shape]
    with _Scope("hash_build", n=n, n_dev=n_dev, slots=slots):
        check(lib.comb_hash_build(_p(coords), n, _p(n_dev), int(batch), D, H, W, _p(table), slots, _stream()),
              "comb_hash_build")
    return table, slots


def conv_out_shape(in_shape, ksize, stride, pad, dil):
    return [(int(i) + 2 * p - d * (k - 1) - 1) // s + 1 for i, k, s, p, d in zip(in_shape, ksize, stride, pad, dil)]


def conv_out_coords(in_coords, batch, out_shape, ksize, stride, pad, dil, out_cap, n_dev=None):
    """Canonical (ascending key) output coordinate set of a strided conv.
    Returns (out_coords (out_cap,4) int32, out_count device int32[1])."""
    lib = _lib.load()
    _need(in_coords, torch.int32, "in_coords")
    _need(n_dev, torch.int32, "n_dev")
    dev = in_coords.device
    oD, oH, oW = [int(x) for x in out_shape]
    ws = _ws(lib.comb_outcoords_workspace_bytes(int(batch), oD, oH, oW), dev)
    out = torch.empty((int(out_cap), 4), dtype=torch.int32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    with _Scope("conv_out_coords", n=int(in_coords.shape[0]), n_dev=n_dev, out_count=cnt,
                bitmap_bytes=int(batch) * oD * oH * oW // 8):
        check(lib.comb_conv_out_coords(_p(in_coords), int(in_coords.shape[0]), _p(n_dev), int(batch), oD, oH, oW,
                                       int3(ksize), int3(stride), int3(pad), int3(dil), _p(out), int(out_cap), _p(cnt),
                                       _p(ws), ws.numel(), _stream()), "comb_conv_out_coords")
    return out, cnt


def nbrmap_build(out_coords, table, slots, batch, in_shape, ksize, stride, pad, dil, no_dev=None, ld=None):
    """Gather-form rulebook nbr (K, ld) int32; nbr[k, o] = input row or -1."""
    lib = _lib.load()
    _need(out_coords, torch.int32, "out_coords")
    _need(no_dev, torch.int32, "no_dev")
    no = int(out_coords.shape[0])
    ld = int(ld or no)
    K = int(ksize[0]) * int(ksize[1]) * int(ksize[2])
    nbr = torch.empty((K, ld), dtype=torch.int32, device=out_coords.device)
    iD, iH, iW = [int(x) for x in in_shape]
    with _Scope("nbrmap_build", no=no, no_dev=no_dev, K=K, nbr=nbr):
        check(lib.comb_nbrmap_build(_p(out_coords), no, _p(no_dev), _p(table), int(slots), int(batch), iD, iH, iW,
                                    int3(ksize), int3(stride), int3(pad), int3(dil), _p(nbr), ld, _stream()),
              "comb_nbrmap_build")
    return nbr


def nbrmap_transpose(nbr, ni, no_dev=None):
    lib = _lib.load()
    _need(nbr, torch.int32, "nbr")
    K, ld = int(nbr.shape[0]), int(nbr.shape[1])
    nbr_t = torch.empty((K, int(ni)), dtype=torch.int32, device=nbr.device)
    check(lib.comb_nbrmap_transpose(_p(nbr), K, ld, _p(no_dev), ld, _p(nbr_t), int(ni), int(ni), _stream()),
          "comb_nbrmap_transpose")
    return nbr_t


def nbrmap_to_pairs(nbr, no_dev=None):
    """spconv-style (2, K, ld) pair lists and (K,) pair counts."""
    lib = _lib.load()
    _need(nbr, torch.int32, "nbr")
    K, ld = int(nbr.shape[0]), int(nbr.shape[1])
    pairs = torch.empty((2, K, ld), dtype=torch.int32, device=nbr.device)
    num = torch.empty((K,), dtype=torch.int32, device=nbr.device)
    check(lib.comb_nbrmap_to_pairs(_p(nbr), K, ld, _p(no_dev), ld, _p(pairs), _p(num), _stream()),
          "comb_nbrmap_to_pairs")
    return pairs, num


class GridIndex:
    """Bitmap-rank index of one level (rows in ascending key order). See comb_index_build."""

    def __init__(self, bitmap, prefix, batch, shape, coords, count):
        self.bitmap, self.prefix, self.batch, self.shape = bitmap, prefix, int(batch), [int(x) for x in shape]
        self.coords, self.count = coords, count          # (cap,4) int32 rows in key order, device int32[1]


def index_build(coords, batch, shape, conv=None, out_cap=None, n_dev=None, want_coords=True):
    """Index of `coords` (conv=None; shape = their grid) or of the output set of the strided conv
    conv=(ksize, stride, pad, dil) applied to them (shape = the OUTPUT grid)."""
    lib = _lib.load()
    _need(coords, torch.int32, "coords")
    _need(n_dev, torch.int32, "n_dev")
    dev = coords.device
    D, H, W = [int(x) for x in shape]
    n = int(coords.shape[0])
    bitmap = torch.empty((max(lib.comb_index_bitmap_bytes(int(batch), D, H, W), 256),), dtype=torch.uint8, device=dev)
    prefix = torch.empty((max(lib.comb_index_prefix_bytes(int(batch), D, H, W), 256),), dtype=torch.uint8, device=dev)
    out_cap = int(n if out_cap is None else out_cap)
    out = torch.empty((max(out_cap, 1), 4), dtype=torch.int32, device=dev) if want_coords else None
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)
    cv = [None] * 4 if conv is None else [int3(v) for v in conv]
    with _Scope("index_build", n=n, n_dev=n_dev, out_count=cnt, bitmap_bytes=int(bitmap.numel())):
        check(lib.comb_index_build(_p(coords), n, _p(n_dev), int(batch), D, H, W, cv[0], cv[1], cv[2], cv[3],
                                   _p(bitmap), _p(prefix), _p(out), out_cap, _p(cnt), _stream()), "comb_index_build")
    return GridIndex(bitmap, prefix, batch, shape, out, cnt)


def index_rank(coords, index, n_dev=None, scatter_coords=False):
    """Row of every coordinate in the index (-1 when absent).  scatter_coords=True (unique coords, e.g. a voxel list
    indexed with want_coords=False): also returns the coordinate list in key order, and stores it in index.coords."""
    lib = _lib.load()
    _need(coords, torch.int32, "coords")
    n = int(coords.shape[0])
    rows = torch.empty((n,), dtype=torch.int32, device=coords.device)
    D, H, W = index.shape
    if scatter_coords:
        sorted_coords = torch.empty((max(n, 1), 4), dtype=torch.int32, device=coords.device)
        with _Scope("index_rank", n=n):
            check(lib.comb_index_rank_scatter(_p(coords), n, _p(n_dev), index.batch, D, H, W, _p(index.bitmap),
                                              _p(index.prefix), _p(rows), _p(sorted_coords), n, _stream()),
                  "comb_index_rank_scatter")
        index.coords = sorted_coords
        return rows, sorted_coords
    with _Scope("index_rank", n=n):
        check(lib.comb_index_rank(_p(coords), n, _p(n_dev), index.batch, D, H, W, _p(index.bitmap), _p(index.prefix),
                                  _p(rows), _stream()), "comb_index_rank")
    return rows


def nbrmap_build_indexed(out_coords, index, ksize, stride, pad, dil, no_dev=None, ld=None):
    """Gather-form rulebook against the bitmap-rank index of the INPUT level."""
    lib = _lib.load()
    _need(out_coords, torch.int32, "out_coords")
    _need(no_dev, torch.int32, "no_dev")
    no = int(out_coords.shape[0])
    ld = int(ld or no)
    K = int(ksize[0]) * int(ksize[1]) * int(ksize[2])
    nbr = torch.empty((K, ld), dtype=torch.int32, device=out_coords.device)
    iD, iH, iW = index.shape
    with _Scope("nbrmap_build", no=no, no_dev=no_dev, K=K, nbr=nbr):
        check(lib.comb_nbrmap_build_indexed(_p(out_coords), no, _p(no_dev), _p(index.bitmap), _p(index.prefix),
                                            index.batch, iD, iH, iW, int3(ksize), int3(stride), int3(pad), int3(dil),
                                            _p(nbr), ld, _stream()), "comb_nbrmap_build_indexed")
    return nbr


def permute_rows(x, row_map, scatter, n_out=None, n_dev=None):
    """scatter: out[row_map[r]] = x[r];  gather: out[r] = x[row_map[r]] (negative map entries: skipped / zero)."""
    lib = _lib.load()
    _need(x, x.dtype, "x")
    _need(row_map, torch.int32, "row_map")
    n = int(row_map.shape[0])
    row_bytes = int(x.shape[1]) * x.element_size()
    out = torch.empty((int(n_out if n_out is not None else n), int(x.shape[1])), dtype=x.dtype, device=x.device)
    with _Scope("permute_rows", n=n, row_bytes=row_bytes):
        check(lib.comb_permute_rows(_p(x), _p(row_map), n, _p(n_dev), row_bytes, int(bool(scatter)), _p(out), _stream()),
              "comb_permute_rows")
    return out


# --------------------------------------------------------------------------------------------- sparse conv
def _epi_flags(bias, scale, shift, residual, relu):
    f = 0
    if bias is not None:
        f |= EPI_BIAS
    if scale is not None:
        assert shift is not None
        f |= EPI_AFFINE
    if residual is not None:
        f |= EPI_RESIDUAL
    if relu:
        f |= EPI_RELU
    return f


def spconv_fwd_f32(feats, weight, nbr, bias=None, scale=None, shift=None, residual=None, relu=False, no_dev=None,
                   no=None):
    """weight: (Cout, K, Cin) fp32. feats (Ni, Cin) fp32 -> (No, Cout) fp32."""
    lib = _lib.load()
    _need(feats, torch.float32, "feats")
    _need(weight, torch.float32, "weight")
    _need(nbr, torch.int32, "nbr")
    for n_, t_ in (("bias", bias), ("scale", scale), ("shift", shift), ("residual", residual)):
        _need(t_, torch.float32, n_)
    Cout, K, Cin = [int(x) for x in weight.shape]
    assert int(feats.shape[1]) == Cin and int(nbr.shape[0]) == K
    ld = int(nbr.shape[1])
    no = ld if no is None else int(no)
    out = torch.empty((no, Cout), dtype=torch.float32, device=feats.device)
    check(lib.comb_spconv_fwd_f32(_p(feats), Cin, _p(weight), K, Cout, _p(nbr), ld, no, _p(no_dev),
                                  _epi_flags(bias, scale, shift, residual, relu), _p(bias), _p(scale), _p(shift),
                                  _p(residual), _p(out), _stream()), "comb_spconv_fwd_f32")
    return out


def spconv_dgrad_f32(dout, weight, nbr_t, ni_dev=None):
    lib = _lib.load()
    _need(dout, torch.float32, "dout")
    _need(weight, torch.float32, "weight")
    _need(nbr_t, torch.int32, "nbr_t")
    Cout, K, Cin = [int(x) for x in weight.shape]
    ni = int(nbr_t.shape[1])
    din = torch.empty((ni, Cin), dtype=torch.float32, device=dout.device)
    check(lib.comb_spconv_dgrad_f32(_p(dout), Cout, _p(weight), K, Cin, _p(nbr_t), ni, ni, _p(ni_dev), _p(din),
                                    _stream()), "comb_spconv_dgrad_f32")
    return din


def spconv_wgrad_f32(feats, dout, nbr, no_dev=None):
    lib = _lib.load()
    _need(feats, torch.float32, "feats")
    _need(dout, torch.float32, "dout")
    _need(nbr, torch.int32, "nbr")
    K, ld = int(nbr.shape[0]), int(nbr.shape[1])
    Cin, Cout = int(feats.shape[1]), int(dout.shape[1])
    dw = torch.empty((Cout, K, Cin), dtype=torch.float32, device=feats.device)
    check(lib.comb_spconv_wgrad_f32(_p(feats), Cin, _p(dout), Cout, K, _p(nbr), ld, int(dout.shape[0]), _p(no_dev),
                                    _p(dw), _stream()), "comb_spconv_wgrad_f32")
    return dw


def spconv_wgrad_bf16(feats, dout, nbr, Cin, no_dev=None):
    """Tensor-core wgrad: feats (Ni, Cin_p) bf16 (zero padded beyond Cin), dout (No, Cout) bf16 -> dW (Cout, K, Cin)
    fp32 (bf16 products, fp32 accumulation in tensor memory, deterministic reduction over the row chunks)."""
    lib = _lib.load()
    _need(feats, torch.bfloat16, "feats")
    _need(dout, torch.bfloat16, "dout")
    _need(nbr, torch.int32, "nbr")
    K, ld = int(nbr.shape[0]), int(nbr.shape[1])
    cin_p, Cout, Cin, no = int(feats.shape[1]), int(dout.shape[1]), int(Cin), int(dout.shape[0])
    nbytes = lib.comb_spconv_wgrad_bf16_workspace_bytes(cin_p, Cin, Cout, K, no)
    if nbytes == 0:
        raise RuntimeError("unsupported conv shape for the bf16 wgrad: Cin_p=%d Cin=%d Cout=%d K=%d" % (cin_p, Cin, Cout, K))
    ws = _ws(nbytes, feats.device)
    dw = torch.empty((Cout, K, Cin), dtype=torch.float32, device=feats.device)
    with _Scope("spconv_wgrad_bf16", cin=cin_p, cout=Cout, K=K, nbr=nbr, no=no, no_dev=no_dev, ni=int(feats.shape[0]),
                feats=feats, dout=dout, cin_real=Cin):
        check(lib.comb_spconv_wgrad_bf16(_p(feats), cin_p, Cin, _p(dout), Cout, K, _p(nbr), ld, no, _p(no_dev), _p(dw),
                                         _p(ws), nbytes, _stream()), "comb_spconv_wgrad_bf16")
    return dw


def pad16(c):
    for p in (16, 32, 64, 128):
        if c <= p:
            return p
    raise RuntimeError("channel count %d > 128 not supported by the tensor-core path" % c)


def pack_weight_bf16(weight):
    """(Cout, K, Cin) fp32 -> pre-swizzled bf16 shared-memory image (uint8 tensor)."""
    lib = _lib.load()
    _need(weight, torch.float32, "weight")
    Cout, K, Cin = [int(x) for x in weight.shape]
    cin_p = pad16(Cin)
    nbytes = lib.comb_spconv_packed_bytes(cin_p, K, Cout)
    if nbytes == 0:
        raise RuntimeError("unsupported conv shape for the bf16 path: Cin=%d Cout=%d K=%d" % (Cin, Cout, K))
    out = torch.empty((nbytes,), dtype=torch.uint8, device=weight.device)
    check(lib.comb_spconv_pack_weight_bf16(_p(weight), Cout, K, Cin, cin_p, _p(out), _stream()),
          "comb_spconv_pack_weight_bf16")
    return out


def spconv_fwd_bf16(feats, wpacked, K, Cout, nbr, bias=None, scale=None, shift=None, residual=None, relu=False,
                    no_dev=None, no=None, out_dtype=torch.bfloat16, out=None):
    """feats (Ni, Cin_p) bf16 with Cin_p in {16,32,64,128}; returns (No, Cout) bf16/fp32."""
    lib = _lib.load()
    _need(feats, torch.bfloat16, "feats")
    _need(nbr, torch.int32, "nbr")
    _need(residual, torch.bfloat16, "residual")
    for n_, t_ in (("bias", bias), ("scale", scale), ("shift", shift)):
        _need(t_, torch.float32, n_)
    cin_p = int(feats.shape[1])
    ld = int(nbr.shape[1])
    no = ld if no is None else int(no)
    if out is None:
        out = torch.empty((no, Cout), dtype=out_dtype, device=feats.device)
    with _Scope("spconv_fwd_bf16", cin=cin_p, cout=int(Cout), K=int(K), nbr=nbr, no=no, no_dev=no_dev,
                ni=int(feats.shape[0]), residual=residual is not None, out_bytes=out.element_size()):
        check(lib.comb_spconv_fwd_bf16(_p(feats), cin_p, _p(wpacked), int(K), int(Cout), _p(nbr), ld, no, _p(no_dev),
                                       _epi_flags(bias, scale, shift, residual, relu), _p(bias), _p(scale), _p(shift),
                                       _p(residual), _p(out), _dt(out), _stream()), "comb_spconv_fwd_bf16")
    return out


def affine_relu(x, scale=None, shift=None, residual=None, relu=True, n_dev=None, out=None):
    lib = _lib.load()
    _need(x, x.dtype, "x")
    n, C = int(x.shape[0]), int(x.shape[1])
    out = torch.empty_like(x) if out is None else out
    check(lib.comb_affine_relu(_p(x), _dt(x), n, _p(n_dev), C, _p(scale), _p(shift), _p(residual), int(bool(relu)),
                               _p(out), _stream()), "comb_affine_relu")
    return out


def cast_pad(x, ld, n_dev=None):
    lib = _lib.load()
    _need(x, torch.float32, "x")
    n, C = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((n, int(ld)), dtype=torch.bfloat16, device=x.device)
    check(lib.comb_cast_pad(_p(x), n, _p(n_dev), C, _p(out), int(ld), _stream()), "comb_cast_pad")
    return out


def bn_train_fwd(x, gamma, beta, eps, momentum, running_mean=None, running_var=None, residual=None, relu=True,
                 n_dev=None, out=None):
    """BatchNorm1d(train) + residual + ReLU; x fp32 or bf16 rows.  -> (out bf16, save_mean, save_invstd)."""
    lib = _lib.load()
    _need(x, x.dtype, "x")
    _need(residual, torch.bfloat16, "residual")
    for n_, t_ in (("gamma", gamma), ("beta", beta), ("running_mean", running_mean), ("running_var", running_var)):
        _need(t_, torch.float32, n_)
    n, C = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((n, C), dtype=torch.bfloat16, device=x.device) if out is None else out
    _need(out, torch.bfloat16, "out")
    mean = torch.empty((C,), dtype=torch.float32, device=x.device)
    invstd = torch.empty((C,), dtype=torch.float32, device=x.device)
    nbytes = lib.comb_bn_workspace_bytes(C)
    if nbytes == 0:
        raise RuntimeError("bn_train_fwd: C=%d not in {16,32,64,128}" % C)
    ws = _ws(nbytes, x.device)
    with _Scope("bn_train", n=n, n_dev=n_dev, C=C, passes=3 + (residual is not None)):
        check(lib.comb_bn_train_fwd(_p(x), _dt(x), n, _p(n_dev), C, _p(gamma), _p(beta), float(eps), float(momentum),
                                    _p(running_mean), _p(running_var), _p(residual), int(bool(relu)), _p(out), _p(mean),
                                    _p(invstd), _p(ws), nbytes, _stream()), "comb_bn_train_fwd")
    return out, mean, invstd


def bn_train_bwd(dy, act, x, gamma, mean, invstd, relu=True, want_g=False, n_dev=None):
    """-> (dx bf16, g bf16 or None, dgamma, dbeta)."""
    lib = _lib.load()
    _need(dy, torch.bfloat16, "dy")
    _need(act, torch.bfloat16, "act")
    _need(x, x.dtype, "x")
    n, C = int(x.shape[0]), int(x.shape[1])
    dx = torch.empty_like(dy)
    g = torch.empty_like(dy) if want_g else None
    dgamma = torch.empty((C,), dtype=torch.float32, device=x.device)
    dbeta = torch.empty((C,), dtype=torch.float32, device=x.device)
    nbytes = lib.comb_bn_workspace_bytes(C)
    ws = _ws(nbytes, x.device)
    with _Scope("bn_train", n=n, n_dev=n_dev, C=C, passes=6 + int(relu) * 2 + int(want_g)):
        check(lib.comb_bn_train_bwd(_p(dy), _p(act), _p(x), _dt(x), n, _p(n_dev), C, _p(gamma), _p(mean), _p(invstd),
                                    int(bool(relu)), _p(dx), _p(g), _p(dgamma), _p(dbeta), _p(ws), nbytes, _stream()),
              "comb_bn_train_bwd")
    return dx, g, dgamma, dbeta


def col_sum(x, n_dev=None):
    lib = _lib.load()
    _need(x, torch.bfloat16, "x")
    n, C = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((C,), dtype=torch.float32, device=x.device)
    nbytes = lib.comb_bn_workspace_bytes(C)
    ws = _ws(nbytes, x.device)
    check(lib.comb_col_sum(_p(x), n, _p(n_dev), C, _p(out), _p(ws), nbytes, _stream()), "comb_col_sum")
    return out


def dense(feats, coords, batch, shape, n_dev=None):
    """SparseConvTensor.dense(): (batch, C, D, H, W) fp32, fully written."""
    lib = _lib.load()
    _need(feats, feats.dtype, "feats")
    _need(coords, torch.int32, "coords")
    D, H, W = [int(x) for x in shape]
    n, C = int(feats.shape[0]), int(feats.shape[1])
    out = torch.empty((int(batch), C, D, H, W), dtype=torch.float32, device=feats.device)
    ws = _ws(lib.comb_dense_workspace_bytes(int(batch), D, H, W), feats.device)
    with _Scope("dense", n=n, C=C, cells=int(batch) * D * H * W, in_bytes=feats.element_size()):
        check(lib.comb_dense(_p(feats), _dt(feats), _p(coords), n, _p(n_dev), int(batch), C, D, H, W, _p(out), _p(ws),
                             ws.numel(), _stream()), "comb_dense")
    return out


def dense_scatter(feats, coords, batch, shape, out, n_dev=None):
    """Scatter the rows into `out` (batch, C, D, H, W) fp32, which the caller has ALREADY zeroed (see
    FramePipeline._enqueue: the zero-fill runs early on the side stream)."""
    lib = _lib.load()
    _need(feats, feats.dtype, "feats")
    _need(coords, torch.int32, "coords")
    _need(out, torch.float32, "out")
    D, H, W = [int(x) for x in shape]
    n, C = int(feats.shape[0]), int(feats.shape[1])
    assert tuple(out.shape) == (int(batch), C, D, H, W)
    with _Scope("dense", n=n, n_dev=n_dev, C=C, cells=int(batch) * D * H * W, in_bytes=feats.element_size(), scatter=True):
        check(lib.comb_dense_scatter(_p(feats), _dt(feats), _p(coords), n, _p(n_dev), int(batch), C, D, H, W, _p(out),
                                     _stream()), "comb_dense_scatter")
    return out


def dense_gather(grad_dense, coords, dtype=torch.float32, n_dev=None):
    """Adjoint of dense(): (batch, C, D, H, W) fp32 -> (n, C) rows at `coords`."""
    lib = _lib.load()
    _need(grad_dense, torch.float32, "grad_dense")
    _need(coords, torch.int32, "coords")
    B, C, D, H, W = [int(v) for v in grad_dense.shape]
    n = int(coords.shape[0])
    out = torch.empty((n, C), dtype=dtype, device=grad_dense.device)
    with _Scope("dense", n=n, C=C, cells=B * D * H * W, in_bytes=4, gather=True):
        check(lib.comb_dense_gather(_p(grad_dense), _p(coords), n, _p(n_dev), B, C, D, H, W, _p(out), _dt(out), _stream()),
              "comb_dense_gather")
    return out


class DenseFunction(torch.autograd.Function):
    """SparseConvTensor.dense() with autograd: forward = comb_dense, backward = comb_dense_gather."""

    @staticmethod
    def forward(ctx, feats, coords, batch, shape):
        ctx.save_for_backward(coords)
        ctx.dtype = feats.dtype
        return dense(feats.contiguous(), coords, batch, shape)

    @staticmethod
    def backward(ctx, grad):
        (coords,) = ctx.saved_tensors
        return dense_gather(grad.contiguous().float(), coords, dtype=ctx.dtype), None, None, None


def dense_nhwc_bf16(feats, coords, batch, shape, n_dev=None):
    """f4: the BEV image of HeightCompression for a bf16 NHWC 2D backbone: logical shape (batch, C*D, H, W), bf16,
    channels_last strides (memory [batch][H][W][C*D], channel c*D + z) — comb_dense_scatter_nhwc_bf16."""
    lib = _lib.load()
    _need(feats, feats.dtype, "feats")
    _need(coords, torch.int32, "coords")
    D, H, W = [int(x) for x in shape]
    n, C = int(feats.shape[0]), int(feats.shape[1])
    mem = torch.zeros((int(batch), H, W, C * D), dtype=torch.bfloat16, device=feats.device)
    with _Scope("dense", n=n, n_dev=n_dev, C=C, cells=int(batch) * D * H * W, in_bytes=feats.element_size(), nhwc=True):
        check(lib.comb_dense_scatter_nhwc_bf16(_p(feats), _dt(feats), _p(coords), n, _p(n_dev), int(batch), C, D, H, W,
                                               _p(mem), _stream()), "comb_dense_scatter_nhwc_bf16")
    return mem.permute(0, 3, 1, 2)


def dense_gather_nhwc(grad, coords, C, D, dtype=torch.float32, n_dev=None):
    """Adjoint of dense_nhwc_bf16: grad logical (batch, C*D, H, W) in channels_last memory (fp32 or bf16) -> (n, C)."""
    lib = _lib.load()
    _need(coords, torch.int32, "coords")
    B, CD, H, W = [int(v) for v in grad.shape]
    assert CD == C * D
    mem = grad.permute(0, 2, 3, 1)
    if not mem.is_contiguous():
        mem = mem.contiguous()
    if mem.dtype not in (torch.float32, torch.bfloat16):
        mem = mem.float()
    n = int(coords.shape[0])
    out = torch.empty((n, C), dtype=dtype, device=grad.device)
    with _Scope("dense", n=n, C=C, cells=B * D * H * W, in_bytes=mem.element_size(), gather=True, nhwc=True):
        check(lib.comb_dense_gather_nhwc(_p(mem), _dt(mem), _p(coords), n, _p(n_dev), B, C, D, H, W, _p(out), _dt(out),
                                         _stream()), "comb_dense_gather_nhwc")
    return out


class DenseNHWCFunction(torch.autograd.Function):
    """dense_nhwc_bf16 with autograd (backward = comb_dense_gather_nhwc)."""

    @staticmethod
    def forward(ctx, feats, coords, batch, shape):
        ctx.save_for_backward(coords)
        ctx.dtype, ctx.C, ctx.D = feats.dtype, int(feats.shape[1]), int(shape[0])
        return dense_nhwc_bf16(feats.contiguous(), coords, batch, shape)

    @staticmethod
    def backward(ctx, grad):
        (coords,) = ctx.saved_tensors
        return dense_gather_nhwc(grad, coords, ctx.C, ctx.D, dtype=ctx.dtype), None, None, None


# --------------------------------------------------------------------------------------------- box ops
def box_trig_host(boxes_np):
    """(nb,2) float32 = (cosf(-rz), sinf(-rz)) from the host libm (bit-identical to the reference's)."""
    lib = _lib.load()
    b = np.ascontiguousarray(boxes_np, dtype=np.float32)
    out = np.empty((b.shape[0], 2), dtype=np.float32)
    lib.comb_box_trig_host(b.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), b.shape[0],
                           out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def box_trig4_host(boxes_np):
    lib = _lib.load()
    b = np.ascontiguousarray(boxes_np, dtype=np.float32)
    out = np.empty((b.shape[0], 4), dtype=np.float32)
    lib.comb_box_trig4_host(b.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), b.shape[0],
                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def points_in_boxes_mask(points, boxes, trig, out=None):
    """points (P, >=3) fp32 CUDA (row stride = points.shape[1]), boxes (Nb,7), trig (Nb,2) -> (Nb,P) int32."""
    lib = _lib.load()
    _need(points, torch.float32, "points")
    _need(boxes, torch.float32, "boxes")
    _need(trig, torch.float32, "trig")
    P, nb = int(points.shape[0]), int(boxes.shape[0])
    if out is None:
        out = torch.empty((nb, P), dtype=torch.int32, device=points.device)
    check(lib.comb_points_in_boxes_mask(_p(points), P, int(points.shape[1]), _p(boxes), _p(trig), nb, _p(out),
                                        _stream()), "comb_points_in_boxes_mask")
    return out


def points_in_any_box(points, boxes, trig, out=None):
    """points (P, >=3) fp32 CUDA, boxes (Nb,7), trig (Nb,2) -> (P,) uint8: 1 iff the point lies in any box
    (== (points_in_boxes_mask(...) != 0).any(0), one byte per point instead of Nb ints)."""
    lib = _lib.load()
    _need(points, torch.float32, "points")
    _need(boxes, torch.float32, "boxes")
    _need(trig, torch.float32, "trig")
    P, nb = int(points.shape[0]), int(boxes.shape[0])
    if out is None:
        out = torch.empty((P,), dtype=torch.uint8, device=points.device)
    _need(out, torch.uint8, "out")
    check(lib.comb_points_in_any_box(_p(points), P, int(points.shape[1]), _p(boxes), _p(trig), nb, _p(out),
                                     _stream()), "comb_points_in_any_box")
    return out


def points_in_boxes_index(points, boxes, out=None):
    """points (B,P,3), boxes (B,T,7) -> (B,P) int32 first-hit index or -1."""
    lib = _lib.load()
    _need(points, torch.float32, "points")
    _need(boxes, torch.float32, "boxes")
    B, P, T = int(points.shape[0]), int(points.shape[1]), int(boxes.shape[1])
    if out is None:
        out = torch.empty((B, P), dtype=torch.int32, device=points.device)
    _need(out, torch.int32, "out")
    check(lib.comb_points_in_boxes_index(_p(points), _p(boxes), B, P, T, _p(out), _stream()),
          "comb_points_in_boxes_index")
    return out


def boxes_bev(boxes_a, boxes_b, flavour="gpu", what="iou", trig_a=None, trig_b=None, out=None):
    lib = _lib.load()
    _need(boxes_a, torch.float32, "boxes_a")
    _need(boxes_b, torch.float32, "boxes_b")
    _need(trig_a, torch.float32, "trig_a")
    _need(trig_b, torch.float32, "trig_b")
    na, nb = int(boxes_a.shape[0]), int(boxes_b.shape[0])
    if out is None:
        out = torch.empty((na, nb), dtype=torch.float32, device=boxes_a.device)
    _need(out, torch.float32, "out")
    check(lib.comb_boxes_bev(_p(boxes_a), _p(trig_a), na, _p(boxes_b), _p(trig_b), nb,
                             0 if flavour == "cpu" else 1, 0 if what == "iou" else 1, _p(out), _stream()),
          "comb_boxes_bev")
    return out


def nms(boxes, thresh, rotated=True, flavour="gpu", trig=None):
    """boxes (N,7) sorted by descending score. Returns (keep int64 (N,), num_keep int32 (1,)) on device."""
    lib = _lib.load()
    _need(boxes, torch.float32, "boxes")
    _need(trig, torch.float32, "trig")
    n = int(boxes.shape[0])
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=boxes.device)
    num = torch.empty((1,), dtype=torch.int32, device=boxes.device)
    ws = _ws(lib.comb_nms_workspace_bytes(n), boxes.device)
    check(lib.comb_nms(_p(boxes), _p(trig), n, float(thresh), int(bool(rotated)), 0 if flavour == "cpu" else 1,
                       _p(keep), _p(num), _p(ws), ws.numel(), _stream()), "comb_nms")
    return keep, num


def centerhead_decode_nms(hm, center, center_z, dim, rot, K, feature_map_stride, voxel_size, point_cloud_range,
                          post_center_limit_range, score_thresh, nms_thresh, nms_pre_max, nms_post_max, label_map=None,
                          vel=None):
    """One separate head of CenterHead.generate_predicted_boxes on the device (comb_centerhead_decode_nms[_vel]).
    hm / dim are the RAW head outputs.  -> (boxes (B,K,7), scores (B,K), labels (B,K) int32 1-based, counts (B,) int32),
    capacity-sized with the per-frame counts on the device; with a velocity head (vel (B,2,H,W)) boxes are (B,K,9)."""
    lib = _lib.load()
    for n_, t_ in (("hm", hm), ("center", center), ("center_z", center_z), ("dim", dim), ("rot", rot)):
        _need(t_, torch.float32, n_)
    _need(label_map, torch.int32, "label_map")
    _need(vel, torch.float32, "vel")
    B, C, H, W = [int(v) for v in hm.shape]
    K = int(K)
    dev = hm.device
    boxes = torch.empty((B, K, 7 if vel is None else 9), dtype=torch.float32, device=dev)
    scores = torch.empty((B, K), dtype=torch.float32, device=dev)
    labels = torch.empty((B, K), dtype=torch.int32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    nbytes = lib.comb_centerhead_workspace_bytes(B, K)
    if nbytes == 0:
        raise RuntimeError("centerhead_decode_nms: K=%d outside [1,1024]" % K)
    ws = _ws(nbytes, dev)
    lim = (ctypes.c_float * 6)(*[float(v) for v in post_center_limit_range])
    with _Scope("centerhead_decode_nms", B=B, C=C, H=H, W=W, K=K):
        check(lib.comb_centerhead_decode_nms_vel(
            _p(hm), _p(center), _p(center_z), _p(dim), _p(rot), _p(vel), B, C, H, W, K, float(feature_map_stride),
            float(voxel_size[0]), float(voxel_size[1]), float(point_cloud_range[0]), float(point_cloud_range[1]), lim,
            float(score_thresh), _p(label_map), float(nms_thresh), int(nms_pre_max), int(nms_post_max), _p(boxes),
            _p(scores), _p(labels), _p(counts), _p(ws), nbytes, _stream()), "comb_centerhead_decode_nms_vel")
    return boxes, scores, labels, counts


# ---------------------------------------------------------------------------------------------------------------------
# f1 — CenterHead target assignment / COM loss re-weighting (targets.cu)
_GTAB_RMAX = 48
_gtab_cache = {}


def _gaussian_tables_host():
    """Gaussian windows of radius 0.._GTAB_RMAX, built with the formula of centernet_utils.gaussian2D
    (pcdet/models/model_utils/centernet_utils.py:78-84: float64 exp, eps cut, cast to fp32) so that the device heat maps
    carry the reference's values bit for bit.  -> (list of (2r+1, 2r+1) fp32 arrays, offsets int32[rmax+2])."""
    tabs, offs, o = [], [], 0
    for r in range(_GTAB_RMAX + 1):
        d = 2 * r + 1
        m = n = (d - 1.) / 2.
        y, x = np.ogrid[-m:m + 1, -n:n + 1]
        sigma = d / 6
        h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
        h[h < np.finfo(h.dtype).eps * h.max()] = 0
        tabs.append(h.astype(np.float32))
        offs.append(o)
        o += d * d
    offs.append(o)
    return tabs, np.asarray(offs, dtype=np.int32)


def _gaussian_tables(device):
    key = str(device)
    if key not in _gtab_cache:
        tabs, offs = _gaussian_tables_host()
        flat = np.concatenate([t.reshape(-1) for t in tabs])
        _gtab_cache[key] = (torch.from_numpy(flat).to(device), torch.from_numpy(offs).to(device))
    return _gtab_cache[key]


def centerhead_assign_targets(gt_boxes, npgt, group, cls_map, num_classes_head, feature_map_size, feature_map_stride,
                              point_cloud_range, voxel_size, num_max_objs=500, gaussian_overlap=0.1, min_radius=2,
                              filter_points=False, min_points=1, relabel_in_place=False):
    """assign_target_of_single_head for every frame of ONE separate head (comb_centerhead_assign_targets).
    gt_boxes (B,M,C) fp32 cuda, npgt (B,M), group (B,M) int64 or None, cls_map (n_cls+1,) int32: global class id ->
    index inside the head or -1; feature_map_size = (W, H).  relabel_in_place: overwrite the class column of this head's
    boxes in `gt_boxes` with the head-local id, as the reference does while it walks a head (the next head sees it).
    -> heatmap (B,Ch,H,W), ret_boxes (B,N,C), inds (B,N) int64, mask (B,N) fp32, radius_map (B,N,4|5) int64."""
    lib = _lib.load()
    _need(gt_boxes, torch.float32, "gt_boxes")
    _need(cls_map, torch.int32, "cls_map")
    B, M, C = [int(v) for v in gt_boxes.shape]
    dev = gt_boxes.device
    npgt = npgt.to(device=dev, dtype=torch.float32).contiguous()
    if group is not None:
        group = group.to(device=dev, dtype=torch.int64).contiguous()
    W, H = int(feature_map_size[0]), int(feature_map_size[1])
    R = 5 if group is not None else 4
    N = int(num_max_objs)
    heatmap = torch.zeros((B, num_classes_head, H, W), dtype=torch.float32, device=dev)
    ret_boxes = torch.zeros((B, N, C), dtype=torch.float32, device=dev)
    inds = torch.zeros((B, N), dtype=torch.int64, device=dev)
    mask = torch.zeros((B, N), dtype=torch.float32, device=dev)
    radius_map = torch.zeros((B, N, R), dtype=torch.int64, device=dev)
    gtab, goff = _gaussian_tables(dev)
    import numpy as np
    f32 = lambda v: float(np.float32(v))
    with _Scope("centerhead_assign_targets", B=B, M=M):
        check(lib.comb_centerhead_assign_targets(
            _p(gt_boxes), _p(npgt), _p(group), _p(cls_map), int(cls_map.numel()) - 1, B, M, C, f32(point_cloud_range[0]),
            f32(point_cloud_range[1]), f32(voxel_size[0]), f32(voxel_size[1]), f32(feature_map_stride), W, H, N,
            float(gaussian_overlap), int(min_radius), 1 if filter_points else 0, float(min_points), _p(gtab), _p(goff),
            _GTAB_RMAX, int(num_classes_head), _p(heatmap), _p(ret_boxes), _p(inds), _p(mask), _p(radius_map), R,
            1 if relabel_in_place else 0, _stream()),
            "comb_centerhead_assign_targets")
    return heatmap, ret_boxes, inds, mask, radius_map


def centerhead_cluster_groups(gt_boxes, true_object, occupancy_ratio, facade_type):
    """CurriculumCenterHead.cluster on the device -> group (B,M) int64."""
    lib = _lib.load()
    _need(gt_boxes, torch.float32, "gt_boxes")
    B, M, C = [int(v) for v in gt_boxes.shape]
    dev = gt_boxes.device
    conv = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
    to, oc, fa = conv(true_object), conv(occupancy_ratio), conv(facade_type)
    group = torch.empty((B, M), dtype=torch.int64, device=dev)
    check(lib.comb_centerhead_cluster_groups(_p(gt_boxes), B * M, C, _p(to), _p(oc), _p(fa), _p(group), _stream()),
          "comb_centerhead_cluster_groups")
    return group


def comloss_group_confidence(pred, radius_map, n_class, n_group):
    """FocalLossCenterCurriculum.confidence_of_all_groups -> (confidence_all, num_all), both (n_class, n_group) fp32."""
    lib = _lib.load()
    _need(pred, torch.float32, "pred")
    _need(radius_map, torch.int64, "radius_map")
    B, Ch, H, W = [int(v) for v in pred.shape]
    nobj, R = int(radius_map.shape[1]), int(radius_map.shape[2])
    conf = torch.empty((n_class, n_group), dtype=torch.float32, device=pred.device)
    num = torch.empty((n_class, n_group), dtype=torch.float32, device=pred.device)
    check(lib.comb_comloss_group_confidence(_p(pred), B, Ch, H, W, _p(radius_map), nobj, R, int(n_class), int(n_group),
                                            _p(conf), _p(num), _stream()), "comb_comloss_group_confidence")
    return conf, num


def comloss_reweight(pred, radius_map, box_mask, mask, threshold, elongation, height, K=1.0, mode=0, fixed_radius=0,
                     add_radius=0, only_center=False, active=True):
    """The object loop of FocalLossCenterCurriculum.neg_loss: box_mask (B,N) and mask (B,Ch,H,W) are updated IN PLACE."""
    lib = _lib.load()
    _need(pred, torch.float32, "pred")
    _need(radius_map, torch.int64, "radius_map")
    _need(box_mask, torch.float32, "box_mask")
    _need(mask, torch.float32, "mask")
    B, Ch, H, W = [int(v) for v in pred.shape]
    nobj, R = int(radius_map.shape[1]), int(radius_map.shape[2])
    with _Scope("comloss_reweight", B=B, N=nobj):
        check(lib.comb_comloss_reweight(_p(pred), B, Ch, H, W, _p(radius_map), nobj, R, float(threshold), float(elongation),
                                        float(height), float(K), int(mode), int(fixed_radius), int(add_radius),
                                        1 if only_center else 0, 1 if active else 0, _p(box_mask), _p(mask), _stream()),
              "comb_comloss_reweight")
    return box_mask, mask


# ---------------------------------------------------------------------------------------------------------------------
# f3 — COMAug placement test (boxes.cu)
def comaug_valid_mask(sampled_boxes, existed_boxes):
    """database_sampler_v2.py:600-604 in one device round trip: numpy (S,7+) sampled and (E,7+) existing boxes in,
    numpy bool (S,) out — True where a sampled box overlaps neither an existing box nor another sampled one.  The two
    IoU matrices (CPU-flavour arithmetic, bit-exact with boxes_bev_iou_cpu) never leave the device."""
    import numpy as np
    lib = _lib.load()
    dev = host_op_device()
    sb = np.ascontiguousarray(np.asarray(sampled_boxes)[:, 0:7], dtype=np.float32)
    eb = np.ascontiguousarray(np.asarray(existed_boxes)[:, 0:7], dtype=np.float32)
    S, E = int(sb.shape[0]), int(eb.shape[0])
    if S == 0:
        return np.zeros((0,), dtype=bool)
    with torch.cuda.device(dev):
        a, ta = torch.from_numpy(sb).to(dev), torch.from_numpy(box_trig4_host(sb)).to(dev)
        iou2 = boxes_bev(a, a, flavour="cpu", what="iou", trig_a=ta, trig_b=ta)
        iou1 = None
        if E > 0:
            b, tb = torch.from_numpy(eb).to(dev), torch.from_numpy(box_trig4_host(eb)).to(dev)
            iou1 = boxes_bev(a, b, flavour="cpu", what="iou", trig_a=ta, trig_b=tb)
        valid = torch.empty((S,), dtype=torch.uint8, device=dev)
        check(lib.comb_comaug_valid_mask(_p(iou1), _p(iou2), S, E, _p(valid), _stream()), "comb_comaug_valid_mask")
        return valid.cpu().numpy().astype(bool)
