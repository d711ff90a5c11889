"""ctypes binding of libcomb200.so (the C-ABI declared in include/comb200.h).

The product path has NO CPU fallback: if the shared library is missing it is built with nvcc (works
without a GPU); if it cannot be loaded, or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_double, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcomb200.so")

DT_F32, DT_BF16 = 0, 1
EPI_BIAS, EPI_AFFINE, EPI_RESIDUAL, EPI_RELU = 1, 2, 4, 8

_lib = None

# name -> (restype, argtypes); MUST list every symbol include/comb200.h declares (tests check this)
_P = c_void_p
_PI = POINTER(c_int)
_PF = POINTER(c_float)
SIGNATURES = {
    "comb_version": (c_int, []),
    "comb_last_error": (c_char_p, []),
    "comb_sm_count": (c_int, []),
    "comb_launch_count": (c_longlong, []),
    "comb_voxelize_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "comb_voxelize": (c_int, [_P, _PI, _P, c_int, c_int, c_int, _PF, _PF, c_int, c_int, _P, _P, _P, _P, c_int, c_int, c_int,
                              _P, _P, c_size_t, _P]),
    "comb_mean_vfe": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "comb_hash_slots": (c_int, [c_int]),
    "comb_hash_build": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "comb_outcoords_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "comb_conv_out_coords": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _PI, _PI, _PI, _PI, _P, c_int, _P,
                                     _P, c_size_t, _P]),
    "comb_nbrmap_build": (c_int, [_P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _PI, _PI, _PI, _PI, _P,
                                  c_int, _P]),
    "comb_nbrmap_transpose": (c_int, [_P, c_int, c_int, _P, c_int, _P, c_int, c_int, _P]),
    "comb_nbrmap_to_pairs": (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, _P]),
    "comb_index_bitmap_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "comb_index_prefix_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "comb_index_build": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _PI, _PI, _PI, _PI, _P, _P, _P, c_int, _P,
                                 _P]),
    "comb_index_rank": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "comb_index_rank_scatter": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P]),
    "comb_nbrmap_build_indexed": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _PI, _PI, _PI, _PI, _P,
                                          c_int, _P]),
    "comb_permute_rows": (c_int, [_P, _P, c_int, _P, c_int, c_int, _P, _P]),
    "comb_spconv_fwd_f32": (c_int, [_P, c_int, _P, c_int, c_int, _P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "comb_spconv_dgrad_f32": (c_int, [_P, c_int, _P, c_int, c_int, _P, c_int, c_int, _P, _P, _P]),
    "comb_spconv_wgrad_f32": (c_int, [_P, c_int, _P, c_int, c_int, _P, c_int, c_int, _P, _P, _P]),
    "comb_spconv_packed_bytes": (c_size_t, [c_int, c_int, c_int]),
    "comb_spconv_pack_weight_bf16": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "comb_spconv_fwd_bf16": (c_int, [_P, c_int, _P, c_int, c_int, _P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P,
                                     c_int, _P]),
    "comb_spconv_wgrad_bf16_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "comb_spconv_wgrad_bf16": (c_int, [_P, c_int, c_int, _P, c_int, c_int, _P, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "comb_debug_conv_trace": (c_int, [_P]),
    "comb_affine_relu": (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, _P, c_int, _P, _P]),
    "comb_cast_pad": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P]),
    "comb_bn_workspace_bytes": (c_size_t, [c_int]),
    "comb_bn_train_fwd": (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, c_float, c_float, _P, _P, _P, c_int, _P, _P, _P, _P,
                                  c_size_t, _P]),
    "comb_bn_train_bwd": (c_int, [_P, _P, _P, c_int, c_int, _P, c_int, _P, _P, _P, c_int, _P, _P, _P, _P, _P, c_size_t,
                                  _P]),
    "comb_col_sum": (c_int, [_P, c_int, _P, c_int, _P, _P, c_size_t, _P]),
    "comb_dense": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "comb_dense_scatter": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "comb_dense_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "comb_dense_gather": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "comb_box_trig_host": (None, [_PF, c_int, _PF]),
    "comb_points_in_boxes_mask": (c_int, [_P, c_int, c_int, _P, _P, c_int, _P, _P]),
    "comb_points_in_boxes_index": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "comb_points_in_any_box": (c_int, [_P, c_int, c_int, _P, _P, c_int, _P, _P]),
    "comb_box_trig4_host": (None, [_PF, c_int, _PF]),
    "comb_boxes_bev": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, _P, _P]),
    "comb_nms_workspace_bytes": (c_size_t, [c_int]),
    "comb_nms": (c_int, [_P, _P, c_int, c_float, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "comb_nms_dev": (c_int, [_P, _P, c_int, _P, c_float, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "comb_centerhead_workspace_bytes": (c_size_t, [c_int, c_int]),
    "comb_centerhead_decode_nms": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float,
                                           c_float, c_float, _PF, c_float, _P, c_float, c_int, c_int, _P, _P, _P, _P, _P,
                                           c_size_t, _P]),
    "comb_centerhead_decode_nms_vel": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                               c_float, c_float, c_float, _PF, c_float, _P, c_float, c_int, c_int, _P, _P,
                                               _P, _P, _P, c_size_t, _P]),
    "comb_dense_scatter_nhwc_bf16": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "comb_dense_gather_nhwc": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "comb_comaug_valid_mask": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "comb_centerhead_assign_targets": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float,
                                               c_float, c_int, c_int, c_int, c_double, c_int, c_int, c_float, _P, _P, c_int,
                                               c_int, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    "comb_centerhead_cluster_groups": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P]),
    "comb_comloss_group_confidence": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "comb_comloss_reweight": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_double, c_double, c_double,
                                      c_double, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
}


def load(build_if_missing=True):
    """Load (building first if needed) libcomb200.so and declare all signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError("libcomb200.so not built: run `python -m com_b200.build`")
        from . import build as _build
        _build.build()
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # fail loudly: there is no fallback path
        raise RuntimeError("cannot load %s: %s" % (LIB_PATH, e)) from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().comb_last_error()
        raise RuntimeError("libcomb200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def int3(v):
    if isinstance(v, int):
        v = (v, v, v)
    assert len(v) == 3
    return (c_int * 3)(*[int(x) for x in v])
