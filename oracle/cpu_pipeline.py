"""CPU restatement of the whole frame pipeline (voxelize -> MeanVFE -> VoxelResBackBone8x (eval) ->
HeightCompression) on top of the oracle's C functions.  TEST INFRASTRUCTURE: the checker for
tests/test_gpu_backbone.py and the CPU baseline that bench.py times.  Never imported by com_b200.

Follows pcdet/models/backbones_3d/spconv_backbone.py:183-293 (layer order, indice keys, padding),
:34-66 (SparseBasicBlock) and height_compression.py:10-26; takes the weights from a state_dict with
the reference's key names.  `conv` selects the convolution arithmetic: oracle.conv_fwd (fp64
accumulate = "truth") or oracle.fast_conv_fwd (OpenMP fp32 = the timed CPU baseline)."""
import numpy as np

import oracle

BN_EPS = 1e-3          # norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01), spconv_backbone.py:186


def _np(t):
    return t.detach().cpu().float().numpy() if hasattr(t, "detach") else np.asarray(t, dtype=np.float32)


def fold(sd, conv_key, bn_key):
    """(W (Cout,K,Cin), scale, shift) of conv (+bias) followed by eval-mode BatchNorm1d."""
    w = _np(sd[conv_key + ".weight"])
    cout, cin = w.shape[0], w.shape[-1]
    w3 = w.reshape(cout, -1, cin)
    g, b = _np(sd[bn_key + ".weight"]).astype(np.float64), _np(sd[bn_key + ".bias"]).astype(np.float64)
    mu, var = _np(sd[bn_key + ".running_mean"]).astype(np.float64), _np(sd[bn_key + ".running_var"]).astype(np.float64)
    scale = g / np.sqrt(var + BN_EPS)
    shift = b - mu * scale
    if conv_key + ".bias" in sd:
        shift = shift + _np(sd[conv_key + ".bias"]).astype(np.float64) * scale
    return w3, scale.astype(np.float32), shift.astype(np.float32)


def _apply(conv, x, w, nbr, scale, shift, residual, rnd):
    if conv is oracle.fast_conv_fwd:
        y = conv(x, w, nbr, None, scale, shift, residual, True)
    else:
        y = conv(x, w, nbr).astype(np.float64) * scale + shift
        if residual is not None:
            y = y + residual
        y = np.maximum(y, 0).astype(np.float32)
    return rnd(y) if rnd else y


def backbone_forward(feats, coords, batch, sparse_shape, sd, conv=None, rnd=None, last_pad=0):
    """feats (N,Cin) f32, coords (N,4) b,z,y,x int32.  Returns [(feats, coords, shape)] for x_conv1..4
    and the encoded output (5 levels).  `rnd` (optional) rounds every stored activation and weight —
    pass a bf16 round-trip to emulate the storage precision of the tensor-core path."""
    conv = conv or oracle.conv_fwd
    r = (lambda a: a) if rnd is None else rnd
    x = r(np.ascontiguousarray(feats, dtype=np.float32))
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    shape = [int(s) for s in sparse_shape]
    ones, k3 = (1, 1, 1), (3, 3, 3)
    levels = []
    nbr = oracle.subm_nbrmap(coords, shape, k3)
    w, sc, sh = fold(sd, "conv_input.0", "conv_input.1")
    x = _apply(conv, x, r(w), nbr, sc, sh, None, rnd)
    downs = {2: ((2, 2, 2), (1, 1, 1)), 3: ((2, 2, 2), (1, 1, 1)), 4: ((2, 2, 2), (0, 1, 1))}
    for li in (1, 2, 3, 4):
        blocks = (0, 1)
        if li > 1:
            st, pd = downs[li]
            oshape = oracle.conv_out_shape(shape, k3, st, pd, ones)
            ocoords = oracle.conv_out_coords(coords, oshape, k3, st, pd, ones)
            nbr_d = oracle.nbrmap(ocoords, coords, shape, k3, st, pd, ones)
            w, sc, sh = fold(sd, "conv%d.0.0" % li, "conv%d.0.1" % li)
            x = _apply(conv, x, r(w), nbr_d, sc, sh, None, rnd)
            coords, shape = ocoords, oshape
            nbr = oracle.subm_nbrmap(coords, shape, k3)
            blocks = (1, 2)
        for bi in blocks:
            p = "conv%d.%d." % (li, bi)
            w1, sc1, sh1 = fold(sd, p + "conv1", p + "bn1")
            w2, sc2, sh2 = fold(sd, p + "conv2", p + "bn2")
            y = _apply(conv, x, r(w1), nbr, sc1, sh1, None, rnd)
            x = _apply(conv, y, r(w2), nbr, sc2, sh2, x, rnd)
        levels.append((x, coords, list(shape)))
    ks, st, pd = (3, 1, 1), (2, 1, 1), (int(last_pad),) * 3
    oshape = oracle.conv_out_shape(shape, ks, st, pd, ones)
    ocoords = oracle.conv_out_coords(coords, oshape, ks, st, pd, ones)
    nbr_d = oracle.nbrmap(ocoords, coords, shape, ks, st, pd, ones)
    w, sc, sh = fold(sd, "conv_out.0", "conv_out.1")
    x = _apply(conv, x, r(w), nbr_d, sc, sh, None, rnd)
    levels.append((x, ocoords, list(oshape)))
    return levels


def frame_forward(frames, sd, vsize, rng, max_points, max_voxels, conv=None, rnd=None, want_dense=True):
    """Whole hot path on the CPU for a list of frames (points (N,C) each).
    -> (levels, spatial_features (B, C*D, H, W) or None, voxel coords (M,4))."""
    vs, rg = np.asarray(vsize, np.float32), np.asarray(rng, np.float32)
    grid = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
    sparse_shape = [int(grid[2]) + 1, int(grid[1]), int(grid[0])]          # spconv_backbone.py:189
    feats, coords = [], []
    for b, pts in enumerate(frames):
        v, c, m = oracle.voxelize(pts, vsize, rng, max_points, max_voxels)
        feats.append(oracle.mean_vfe(v, m))
        coords.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], axis=1))
    feats, coords = np.concatenate(feats), np.concatenate(coords)
    levels = backbone_forward(feats, coords, len(frames), sparse_shape, sd, conv=conv, rnd=rnd)
    sf = None
    if want_dense:
        x, c, shape = levels[-1]
        d = oracle.dense(x, c, len(frames), shape)
        sf = d.reshape(d.shape[0], d.shape[1] * d.shape[2], d.shape[3], d.shape[4])
    return levels, sf, coords
