"""Layout probe for the tcgen05 wgrad kernel (conv_wgrad.cu): one-hot operands show where a single product lands.
usage: python scripts/wgrad_diag.py   (prints, for a few (row, ci, co) probes, the non-zero entries of dW)"""
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from com_b200 import ops  # noqa: E402


def probe(C, Co, K, n, r0, k0, ci0, co0):
    x = torch.zeros((n, C), dtype=torch.bfloat16, device="cuda")
    g = torch.zeros((n, Co), dtype=torch.bfloat16, device="cuda")
    nbr = torch.full((K, n), -1, dtype=torch.int32, device="cuda")
    nbr[k0] = torch.arange(n, dtype=torch.int32, device="cuda")
    x[r0, ci0] = 1.0
    g[r0, co0] = 1.0
    dw = ops.spconv_wgrad_bf16(x, g, nbr, C).cpu().numpy()
    nz = np.argwhere(dw != 0)
    print("C=%d Co=%d K=%d n=%d probe row=%d k=%d ci=%d co=%d -> expect dW[%d,%d,%d]=1; got %d nonzero: %s" % (
        C, Co, K, n, r0, k0, ci0, co0, co0, k0, ci0, len(nz),
        [(int(a), int(b), int(c), float(dw[a, b, c])) for a, b, c in nz[:8]]))


for C, Co in ((64, 64), (16, 16), (128, 128)):
    for (r0, k0, ci0, co0) in ((0, 0, 0, 0), (0, 0, 1, 0), (0, 0, 8, 0), (0, 0, 0, 1), (0, 0, 0, 8), (1, 0, 0, 0),
                               (8, 0, 0, 0), (9, 0, 3, 5), (17, 1, 11, 13), (40, 2, C - 1, Co - 1), (100, 2, 5, 9)):
        probe(C, Co, 3, 128, r0, k0, ci0, co0)

# dense random integer case: exact in fp32
rng = np.random.default_rng(0)
for C, Co in ((64, 64), (16, 16), (32, 32), (128, 128)):
    n, K = 200, 3
    x = torch.from_numpy(rng.integers(-3, 4, size=(n, C)).astype(np.float32)).cuda()
    g = torch.from_numpy(rng.integers(-3, 4, size=(n, Co)).astype(np.float32)).cuda()
    nbr = torch.from_numpy(rng.integers(-1, n, size=(K, n)).astype(np.int32)).cuda()
    want = ops.spconv_wgrad_f32(x, g, nbr).cpu().numpy()
    got = ops.spconv_wgrad_bf16(x.bfloat16(), g.bfloat16(), nbr, C).cpu().numpy()
    print("C=%d Co=%d integer case: max abs diff %.3f (max |want| %.1f), equal fraction %.3f" % (
        C, Co, np.abs(got - want).max(), np.abs(want).max(), (got == want).mean()))
