"""a9, training form: comb_bn_train_fwd / comb_bn_train_bwd / comb_col_sum against torch's own nn.BatchNorm1d
(+ ReLU + residual, the chain of pcdet/models/backbones_3d/spconv_backbone.py:21-25,50-66) and its autograd, on the
same bf16-rounded inputs.  Bars: batch / running statistics 1e-4 relative, activations and gradients within bf16
rounding of the fp32 result (storage is bf16), parameter gradients 1e-3 (bf16 inputs, fp32/fp64 accumulation)."""
import numpy as np
import pytest
import torch

from com_b200 import ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("C", [16, 32, 64, 128])
@pytest.mark.parametrize("n,cap", [(50000, 50000), (1237, 4096), (1, 64)])
@pytest.mark.parametrize("residual,relu", [(False, True), (True, True), (False, False)])
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
def test_bn_train_fwd_bwd_vs_torch(C, n, cap, residual, relu, xdt):
    g = torch.Generator(device="cuda").manual_seed(C * 7 + n)
    x = (torch.randn((cap, C), device="cuda", generator=g) * 1.7 + 0.3).to(xdt)
    res = torch.randn((cap, C), device="cuda", generator=g).to(torch.bfloat16) if residual else None
    dy = torch.randn((cap, C), device="cuda", generator=g).to(torch.bfloat16)
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, device="cuda", generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, device="cuda", generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(C, device="cuda", generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(C, device="cuda", generator=g) + 0.5)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    out, mean, invstd = ops.bn_train_fwd(x, bn.weight.detach(), bn.bias.detach(), 1e-3, 0.01, rm, rv, residual=res,
                                         relu=relu, n_dev=n_dev)
    # torch on the same (bf16-rounded) numbers in fp32
    xf = x[:n].float().requires_grad_(True)
    rf = res[:n].float().requires_grad_(True) if residual else None
    bn.train()
    if n > 1:
        y = bn(xf)
    else:                                   # torch refuses a single row in train mode; the formula still holds
        y = (xf - xf.mean(0)) / torch.sqrt(xf.var(0, unbiased=False) + 1e-3) * bn.weight + bn.bias
    if residual:
        y = y + rf
    if relu:
        y = torch.relu(y)
    assert rel(mean, xf.detach().mean(0)) < 1e-4
    assert rel(invstd, 1.0 / torch.sqrt(xf.detach().var(0, unbiased=False) + 1e-3)) < 1e-4
    if n > 1:
        assert rel(rm, bn.running_mean) < 1e-4 and rel(rv, bn.running_var) < 1e-4
    assert torch.equal(out[:n], y.detach().to(torch.bfloat16)) or rel(out[:n].float(), y.detach()) < 8e-3
    # backward
    y.backward(dy[:n].float())
    dx, gres, dgamma, dbeta = ops.bn_train_bwd(dy, out, x, bn.weight.detach(), mean, invstd, relu=relu,
                                               want_g=residual, n_dev=n_dev)
    if n > 1:
        assert rel(dgamma, bn.weight.grad) < 2e-3 and rel(dbeta, bn.bias.grad) < 2e-3
        assert rel(dx[:n].float(), xf.grad) < 1e-2
    if residual:
        assert rel(gres[:n].float(), rf.grad) < 8e-3
    cs = ops.col_sum(dy, n_dev=n_dev)
    assert rel(cs, dy[:n].float().sum(0)) < 1e-5
    # rows beyond the live count are never touched
    if cap > n:
        probe = torch.full((cap, C), 7.0, dtype=torch.bfloat16, device="cuda")
        ops.bn_train_fwd(x, bn.weight.detach(), bn.bias.detach(), 1e-3, 0.01, None, None, n_dev=n_dev, out=probe)
        assert bool((probe[n:] == 7.0).all())


def test_bn_deterministic():
    x = torch.randn((200000, 32), device="cuda").to(torch.bfloat16)
    w, b = torch.ones(32, device="cuda"), torch.zeros(32, device="cuda")
    a = ops.bn_train_fwd(x, w, b, 1e-3, 0.01)
    c = ops.bn_train_fwd(x, w, b, 1e-3, 0.01)
    assert all(torch.equal(p, q) for p, q in zip(a, c))
