"""a7/a8/a9 parity: sparse convolution through the C-ABI vs the CPU oracle (fp64 accumulation).
Tolerances (BASELINE.json north_star): fp32 check mode max relative error 1e-4; bf16 tensor-core
path 2e-2.  "max relative error" = max|got - want| / max|want| over the output tensor; the bf16 path is
additionally held to an element-wise bound."""
import numpy as np
import pytest
import torch

import oracle
from com_b200 import ops, sparse
from util import clustered_coords, random_coords

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-4
TOL_BF16 = 2e-2


def rel_err(got, want):
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max() / max(np.abs(want).max(), 1e-30))


def make_case(seed, n, Cin, Cout, ks=(3, 3, 3), st=(1, 1, 1), pd=(1, 1, 1), subm=True, batch=2, shape=(12, 40, 40)):
    rng = np.random.default_rng(seed)
    shape = list(shape)
    coords = clustered_coords(rng, n, batch, shape, clusters=10, spread=2.5)
    dl = (1, 1, 1)
    if subm:
        out_coords, nbr = coords, oracle.subm_nbrmap(coords, shape, ks)
    else:
        oshape = oracle.conv_out_shape(shape, ks, st, pd, dl)
        out_coords = oracle.conv_out_coords(coords, oshape, ks, st, pd, dl)
        nbr = oracle.nbrmap(out_coords, coords, shape, ks, st, pd, dl)
    K = int(np.prod(ks))
    feats = rng.normal(size=(len(coords), Cin)).astype(np.float32)
    W = (rng.normal(size=(Cout, K, Cin)) / np.sqrt(K * Cin)).astype(np.float32)
    return coords, out_coords, nbr, feats, W, rng


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("Cin,Cout", [(5, 16), (16, 16), (16, 32), (32, 32), (64, 64), (64, 128), (128, 128), (3, 7)])
@pytest.mark.parametrize("subm", [True, False])
def test_fwd_f32(Cin, Cout, subm):
    coords, out_coords, nbr, feats, W, rng = make_case(Cin * 131 + Cout, 2500, Cin, Cout, st=(1, 1, 1) if subm else (2, 2, 2), subm=subm)
    bias = rng.normal(size=(Cout,)).astype(np.float32)
    want = oracle.conv_fwd(feats, W, nbr, bias)
    got = ops.spconv_fwd_f32(cuda(feats), cuda(W), cuda(nbr), bias=cuda(bias)).cpu().numpy()
    assert got.shape == want.shape and rel_err(got, want) < TOL_F32


def test_fwd_f32_epilogue():
    coords, out_coords, nbr, feats, W, rng = make_case(3, 2000, 16, 32)
    no = nbr.shape[1]
    bias, scale, shift = [rng.normal(size=(32,)).astype(np.float32) for _ in range(3)]
    res = rng.normal(size=(no, 32)).astype(np.float32)
    want = np.maximum((oracle.conv_fwd(feats, W, nbr, bias).astype(np.float64)) * scale + shift + res, 0)
    got = ops.spconv_fwd_f32(cuda(feats), cuda(W), cuda(nbr), bias=cuda(bias), scale=cuda(scale), shift=cuda(shift),
                             residual=cuda(res), relu=True).cpu().numpy()
    assert rel_err(got, want.astype(np.float32)) < TOL_F32


@pytest.mark.parametrize("Cin,Cout", [(5, 16), (16, 32), (64, 64), (128, 128)])
@pytest.mark.parametrize("subm", [True, False])
def test_bwd_f32(Cin, Cout, subm):
    coords, out_coords, nbr, feats, W, rng = make_case(Cin + 7 * Cout, 2000, Cin, Cout, st=(1, 1, 1) if subm else (2, 2, 2), subm=subm)
    dout = rng.normal(size=(nbr.shape[1], Cout)).astype(np.float32)
    nbr_d = cuda(nbr)
    nbr_t = ops.nbrmap_transpose(nbr_d, len(coords))
    din = ops.spconv_dgrad_f32(cuda(dout), cuda(W), nbr_t).cpu().numpy()
    dw = ops.spconv_wgrad_f32(cuda(feats), cuda(dout), nbr_d).cpu().numpy()
    assert rel_err(din, oracle.conv_dgrad(dout, W, nbr, len(coords))) < TOL_F32
    assert rel_err(dw, oracle.conv_wgrad(feats, dout, nbr)) < TOL_F32


def bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("Cin,Cout", [(5, 16), (16, 16), (16, 32), (32, 16), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128)])
@pytest.mark.parametrize("subm", [True, False])
def test_fwd_bf16_tensor_core(Cin, Cout, subm):
    coords, out_coords, nbr, feats, W, rng = make_case(Cin * 17 + Cout, 3000, Cin, Cout, st=(1, 1, 1) if subm else (2, 2, 2), subm=subm)
    fb, Wb = bf16_round(feats), bf16_round(W)
    want = oracle.conv_fwd(fb, Wb, nbr)                 # same bf16-rounded operands, fp64 accumulate
    x = ops.cast_pad(cuda(feats), ops.pad16(Cin))
    wp = ops.pack_weight_bf16(cuda(W))
    K = W.shape[1]
    got32 = ops.spconv_fwd_bf16(x, wp, K, Cout, cuda(nbr), out_dtype=torch.float32).cpu().numpy()
    assert rel_err(got32, want) < 1e-4                  # fp32 accumulation of exact bf16 products
    got16 = ops.spconv_fwd_bf16(x, wp, K, Cout, cuda(nbr), out_dtype=torch.bfloat16).float().cpu().numpy()
    assert rel_err(got16, want) < TOL_BF16
    assert np.all(np.abs(got16 - want) <= 2.0 ** -8 * np.abs(want) + 1e-5 * np.abs(want).max())   # one bf16 rounding
    # against the un-rounded fp32 problem the bf16 path stays inside the stated tolerance as well
    assert rel_err(got16, oracle.conv_fwd(feats, W, nbr)) < TOL_BF16


def test_fwd_bf16_epilogue_and_device_count():
    coords, out_coords, nbr, feats, W, rng = make_case(11, 3000, 32, 32)
    no = nbr.shape[1]
    scale, shift = [rng.normal(size=(32,)).astype(np.float32) for _ in range(2)]
    res = bf16_round(rng.normal(size=(no, 32)).astype(np.float32))
    fb, Wb = bf16_round(feats), bf16_round(W)
    want = np.maximum(oracle.conv_fwd(fb, Wb, nbr).astype(np.float64) * scale + shift + res, 0).astype(np.float32)
    x = ops.cast_pad(cuda(feats), 32)
    wp = ops.pack_weight_bf16(cuda(W))
    n_dev = torch.tensor([no - 77], dtype=torch.int32, device="cuda")
    out = torch.full((no, 32), -7.0, dtype=torch.bfloat16, device="cuda")
    ops.spconv_fwd_bf16(x, wp, 27, 32, cuda(nbr), scale=cuda(scale), shift=cuda(shift),
                        residual=cuda(res).to(torch.bfloat16), relu=True, no_dev=n_dev, out=out)
    got = out.float().cpu().numpy()
    assert rel_err(got[: no - 77], want[: no - 77]) < TOL_BF16
    assert (got[no - 77:] == -7.0).all()                # rows beyond the device-side count are untouched


@pytest.mark.parametrize("Cin,Cout,ks,n", [(64, 64, (3, 3, 3), 13000), (128, 128, (3, 3, 3), 13000), (64, 128, (3, 3, 3), 9000),
                                           (128, 128, (3, 1, 1), 13000), (128, 64, (3, 3, 3), 2000), (64, 64, (1, 1, 3), 500)])
def test_ts_single_tile_passes_split_k(Cin, Cout, ks, n):
    """Fewer row tiles than SMs (level 4 of the bench frames: 101 tiles): every pass of conv_ts is a single-tile pass and
    its K range is split over the two halves of the pipeline (two accumulators summed in the epilogue, the streamed-weight
    ring interleaved).  Streamed (27 taps) and resident (3 taps) weight images, odd and even stage counts, the full
    epilogue (affine + residual + ReLU) and a device-side row count that cuts the last tile."""
    rng = np.random.default_rng(Cin + 3 * Cout + n)
    coords = random_coords(rng, n, 4, [16, 48, 48])
    nbr = oracle.subm_nbrmap(coords, [16, 48, 48], ks)
    K = int(np.prod(ks))
    feats = rng.normal(size=(len(coords), Cin)).astype(np.float32)
    W = (rng.normal(size=(Cout, K, Cin)) / np.sqrt(K * Cin)).astype(np.float32)
    no = nbr.shape[1]
    assert no == n and no < 148 * 128
    cut = 57
    scale, shift = [rng.normal(size=(Cout,)).astype(np.float32) for _ in range(2)]
    res = bf16_round(rng.normal(size=(no, Cout)).astype(np.float32))
    want = np.maximum(oracle.fast_conv_fwd(bf16_round(feats), bf16_round(W), nbr).astype(np.float64) * scale + shift + res, 0)
    x = ops.cast_pad(cuda(feats), Cin)
    wp = ops.pack_weight_bf16(cuda(W))
    out = torch.full((no, Cout), -7.0, dtype=torch.float32, device="cuda")
    ops.spconv_fwd_bf16(x, wp, K, Cout, cuda(nbr), scale=cuda(scale), shift=cuda(shift), residual=cuda(res).to(torch.bfloat16),
                        relu=True, no_dev=torch.tensor([no - cut], dtype=torch.int32, device="cuda"), out=out)
    got = out.cpu().numpy()
    assert rel_err(got[: no - cut], want[: no - cut].astype(np.float32)) < 1e-4
    assert (got[no - cut:] == -7.0).all()
    plain = ops.spconv_fwd_bf16(x, wp, K, Cout, cuda(nbr), out_dtype=torch.float32)
    again = ops.spconv_fwd_bf16(x, wp, K, Cout, cuda(nbr), out_dtype=torch.float32)
    assert torch.equal(plain, again)                    # fixed summation order: bit-reproducible


@pytest.mark.parametrize("C,n", [(16, 50000), (32, 50000), (64, 90000), (128, 90000)])
def test_bf16_many_tiles_persistent_loop(C, n):
    """More (super-)tiles than SMs: every CTA walks several of them (slot / weight-stage / accumulator / index
    rings wrap; streamed weights at C >= 64 use 256-row super-tiles, odd tile counts leave a half-empty one)."""
    rng = np.random.default_rng(12 + C)
    coords = random_coords(rng, n, 4, [16, 96, 96])
    nbr = oracle.subm_nbrmap(coords, [16, 96, 96])
    feats = rng.normal(size=(len(coords), C)).astype(np.float32)
    W = (rng.normal(size=(C, 27, C)) / np.sqrt(27 * C)).astype(np.float32)
    assert nbr.shape[1] > 148 * 128 * (2 if C < 64 else 4) and (nbr >= 0).mean() > 0.08
    want = oracle.fast_conv_fwd(bf16_round(feats), bf16_round(W), nbr)
    got = ops.spconv_fwd_bf16(ops.cast_pad(cuda(feats), C), ops.pack_weight_bf16(cuda(W)), 27, C, cuda(nbr),
                              out_dtype=torch.float32).cpu().numpy()
    assert rel_err(got, want) < 1e-4


@pytest.mark.parametrize("Cin,Cout", [(5, 16), (16, 16), (16, 32), (32, 16), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128)])
@pytest.mark.parametrize("subm", [True, False])
def test_wgrad_bf16_tensor_core(Cin, Cout, subm):
    """a8 wgrad on tcgen05 (MN-major operands, conv_wgrad.cu) vs the oracle: 1e-4 on the same bf16-rounded operands
    (exact bf16 products, fp32 accumulation), 2e-2 against the un-rounded fp32 problem."""
    coords, out_coords, nbr, feats, W, rng = make_case(Cin * 19 + Cout, 3000, Cin, Cout, st=(1, 1, 1) if subm else (2, 2, 2), subm=subm)
    dout = rng.normal(size=(nbr.shape[1], Cout)).astype(np.float32)
    want = oracle.conv_wgrad(bf16_round(feats), bf16_round(dout), nbr)
    x = ops.cast_pad(cuda(feats), ops.pad16(Cin))
    g = ops.cast_pad(cuda(dout), Cout)
    got = ops.spconv_wgrad_bf16(x, g, cuda(nbr), Cin)
    assert tuple(got.shape) == (Cout, 27, Cin)
    got = got.cpu().numpy()
    assert rel_err(got, want) < 1e-4
    assert rel_err(got, oracle.conv_wgrad(feats, dout, nbr)) < TOL_BF16
    # deterministic: fixed-order reduction over the row chunks, no atomics
    assert np.array_equal(got, ops.spconv_wgrad_bf16(x, g, cuda(nbr), Cin).cpu().numpy())


def test_wgrad_bf16_kernel_311_and_device_count():
    """spconv_down2 shape (k=(3,1,1), stride (2,1,1), 128 -> 128) and a device-side row count below the bound."""
    rng = np.random.default_rng(5)
    shape = [12, 40, 40]
    coords = clustered_coords(rng, 3000, 2, shape, clusters=10, spread=2.5)
    ks, st, pd, dl = (3, 1, 1), (2, 1, 1), (0, 0, 0), (1, 1, 1)
    oshape = oracle.conv_out_shape(shape, ks, st, pd, dl)
    out_coords = oracle.conv_out_coords(coords, oshape, ks, st, pd, dl)
    nbr = oracle.nbrmap(out_coords, coords, shape, ks, st, pd, dl)
    no = nbr.shape[1]
    feats = rng.normal(size=(len(coords), 128)).astype(np.float32)
    dout = rng.normal(size=(no, 128)).astype(np.float32)
    x, g = ops.cast_pad(cuda(feats), 128), ops.cast_pad(cuda(dout), 128)
    got = ops.spconv_wgrad_bf16(x, g, cuda(nbr), 128).cpu().numpy()
    assert rel_err(got, oracle.conv_wgrad(bf16_round(feats), bf16_round(dout), nbr)) < 1e-4
    cut = no - 101
    n_dev = torch.tensor([cut], dtype=torch.int32, device="cuda")
    got = ops.spconv_wgrad_bf16(x, g, cuda(nbr), 128, no_dev=n_dev).cpu().numpy()
    want = oracle.conv_wgrad(bf16_round(feats), bf16_round(dout)[:cut], np.ascontiguousarray(nbr[:, :cut]))
    assert rel_err(got, want) < 1e-4


@pytest.mark.parametrize("C,n", [(16, 50000), (32, 50000), (64, 60000), (128, 60000)])
def test_wgrad_bf16_many_tiles(C, n):
    """Every CTA walks several 64-row tiles: the stage ring wraps, accumulators stay in tensor memory across tiles."""
    rng = np.random.default_rng(40 + C)
    coords = random_coords(rng, n, 4, [16, 96, 96])
    nbr = oracle.subm_nbrmap(coords, [16, 96, 96])
    feats = rng.normal(size=(len(coords), C)).astype(np.float32)
    dout = rng.normal(size=(len(coords), C)).astype(np.float32)
    want = oracle.conv_wgrad(bf16_round(feats), bf16_round(dout), nbr)
    got = ops.spconv_wgrad_bf16(ops.cast_pad(cuda(feats), C), ops.cast_pad(cuda(dout), C), cuda(nbr), C).cpu().numpy()
    assert rel_err(got, want) < 1e-4


def test_module_autograd_matches_oracle():
    """SubMConv3d / SparseConv3d modules (spconv API) incl. backward through torch autograd."""
    coords, out_coords, nbr, feats, W, rng = make_case(13, 1500, 16, 32, st=(2, 2, 2), subm=False)
    conv = sparse.SparseConv3d(16, 32, 3, stride=2, padding=1, bias=True, indice_key="sp").cuda()
    with torch.no_grad():
        conv.weight.copy_(cuda(W).reshape(32, 3, 3, 3, 16))
    bias = conv.bias.detach().cpu().numpy()
    x = cuda(feats).requires_grad_(True)
    t = sparse.SparseConvTensor(x, cuda(coords), [12, 40, 40], 2)
    y = conv(t)
    assert np.array_equal(y.indices.cpu().numpy(), out_coords) and y.spatial_shape == [6, 20, 20]
    assert rel_err(y.features.detach().cpu().numpy(), oracle.conv_fwd(feats, W, nbr, bias)) < TOL_F32
    dout = rng.normal(size=tuple(y.features.shape)).astype(np.float32)
    y.features.backward(cuda(dout))
    assert rel_err(x.grad.cpu().numpy(), oracle.conv_dgrad(dout, W, nbr, len(coords))) < TOL_F32
    assert rel_err(conv.weight.grad.reshape(32, 27, 16).cpu().numpy(), oracle.conv_wgrad(feats, dout, nbr)) < TOL_F32
    assert rel_err(conv.bias.grad.cpu().numpy(), dout.sum(0)) < TOL_F32
    # SparseInverseConv3d re-uses the rulebook and returns to the input index set
    inv = sparse.SparseInverseConv3d(32, 16, 3, indice_key="sp", bias=False).cuda()
    z = inv(y)
    assert np.array_equal(z.indices.cpu().numpy(), coords) and tuple(z.features.shape) == (len(coords), 16)
    Wi = inv.weight.detach().reshape(16, 27, 32).cpu().numpy()
    want = oracle.conv_fwd(y.features.detach().cpu().numpy(), Wi, ops.nbrmap_transpose(cuda(nbr), len(coords)).cpu().numpy())
    assert rel_err(z.features.detach().cpu().numpy(), want) < TOL_F32


@pytest.mark.parametrize("env", [{"COMB_CONV_IMPL": "ss"}, {"COMB_CONV_IMPL": "ts"}, {"COMB_CONV_IMPL": "tr"}, {"COMB_PDL": "1"},
                                 {"COMB_CONV_IMPL": "ts", "COMB_TS_BLOCKED": "1", "COMB_TS_NI": "8", "COMB_TS_NB": "8"},
                                 {"COMB_CONV_IMPL": "ts", "COMB_TS_SPLIT": "0"}, {"COMB_CONV_IMPL": "ts", "COMB_TS_NB": "3"},
                                 {"COMB_CONV_IMPL": "ts", "COMB_TS_WIDE": "1"}, {"COMB_CONV_IMPL": "ts", "COMB_TS_WIDE": "0"}])
def test_alternate_conv_kernels_stay_correct(env):
    """The documented A/B switches (shared-memory-A tcgen05 kernel; conv_ts or conv_tr forced for every layer shape;
    programmatic dependent launch; blocked tile assignment with deep rings) are read once per process, so each is
    exercised in a child process: bf16 forward vs the oracle, 1e-4."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "tests")
import oracle
from com_b200 import ops
from util import clustered_coords
rng = np.random.default_rng(5)
shape = [12, 40, 40]
coords = clustered_coords(rng, 3000, 2, shape, clusters=10, spread=2.5)
nbr = oracle.subm_nbrmap(coords, shape)
bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
for cin, cout in ((16, 16), (32, 32), (64, 64)):
    feats = rng.normal(size=(len(coords), cin)).astype(np.float32)
    W = (rng.normal(size=(cout, 27, cin)) / np.sqrt(27 * cin)).astype(np.float32)
    want = oracle.conv_fwd(bf(feats), bf(W), nbr)
    got = ops.spconv_fwd_bf16(ops.cast_pad(torch.from_numpy(feats).cuda(), cin), ops.pack_weight_bf16(torch.from_numpy(W).cuda()),
                              27, cout, torch.from_numpy(nbr).cuda(), out_dtype=torch.float32).cpu().numpy()
    err = float(np.abs(got - want).max() / np.abs(want).max())
    assert err < 1e-4, (cin, cout, err)
print("alt-ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, PYTHONPATH=root, **env),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "alt-ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("Cin,Cout,subm", [(5, 16, True), (16, 32, False), (64, 64, True), (128, 128, False)])
def test_module_autograd_bf16_training_form(Cin, Cout, subm):
    """config.compute = "bf16" with autograd: forward and dgrad on the tcgen05 kernel (dgrad = the same gather-GEMM over
    the transposed rulebook with W^T), wgrad on the tcgen05 MN-major kernel.  vs the oracle: 2e-2 for the bf16-operand
    results (forward, dgrad, wgrad), 1e-4 for the bias gradient (fp32 arithmetic on fp32 operands); with
    config.wgrad = "f32" the weight gradient comes from the fp32 check kernel and meets 1e-4."""
    st = (1, 1, 1) if subm else (2, 2, 2)
    coords, out_coords, nbr, feats, W, rng = make_case(Cin * 3 + Cout, 2500, Cin, Cout, st=st, subm=subm)
    mod = (sparse.SubMConv3d(Cin, Cout, 3, bias=True, indice_key="k") if subm else
           sparse.SparseConv3d(Cin, Cout, 3, stride=2, padding=1, bias=True, indice_key="k")).cuda()
    with torch.no_grad():
        mod.weight.copy_(cuda(W).reshape(Cout, 3, 3, 3, Cin))
    bias = mod.bias.detach().cpu().numpy()
    old = sparse.config.compute
    sparse.config.compute = "bf16"
    try:
        x = cuda(feats).requires_grad_(True)
        y = mod(sparse.SparseConvTensor(x, cuda(coords), [12, 40, 40], 2))
        dout = rng.normal(size=tuple(y.features.shape)).astype(np.float32)
        y.features.backward(cuda(dout))
    finally:
        sparse.config.compute = old
    assert y.features.dtype == torch.float32 and x.grad.dtype == torch.float32
    assert rel_err(y.features.detach().cpu().numpy(), oracle.conv_fwd(feats, W, nbr, bias)) < TOL_BF16
    assert rel_err(x.grad.cpu().numpy(), oracle.conv_dgrad(dout, W, nbr, len(coords))) < TOL_BF16
    assert rel_err(mod.weight.grad.reshape(Cout, 27, Cin).cpu().numpy(), oracle.conv_wgrad(feats, dout, nbr)) < TOL_BF16
    assert rel_err(mod.bias.grad.cpu().numpy(), dout.sum(0)) < TOL_F32
    old_w = sparse.config.wgrad
    sparse.config.compute, sparse.config.wgrad = "bf16", "f32"
    try:
        mod.weight.grad = None
        x2 = cuda(feats).requires_grad_(True)
        mod(sparse.SparseConvTensor(x2, cuda(coords), [12, 40, 40], 2)).features.backward(cuda(dout))
    finally:
        sparse.config.compute, sparse.config.wgrad = old, old_w
    assert rel_err(mod.weight.grad.reshape(Cout, 27, Cin).cpu().numpy(), oracle.conv_wgrad(feats, dout, nbr)) < TOL_F32
