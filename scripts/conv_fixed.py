"""Fixed cost of one sparse-conv launch: the 16x16 / 64x64 kernel on 1, 2, 4, 8, 17 tiles per SM of a dense-local
rulebook: full, with every pipeline piece switched off (COMB_TS_ABLATE=29), and set-up + tear-down only (64, conv_tr),
CUDA-graph timed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from com_b200 import ops
from conv_floor import timed

if __name__ == "__main__":
    for cin, cout in ((16, 16), (64, 64)):
        w = ops.pack_weight_bf16(torch.randn((cout, 27, cin), device="cuda") / 20)
        for tps in (1, 2, 4, 8, 17):
            n = 148 * 128 * tps
            o = torch.arange(n, device="cuda", dtype=torch.int32)
            nbr = torch.stack([(o + k - 13).clamp(0, n - 1) for k in range(27)]).contiguous()
            x = torch.randn((n, cin), device="cuda").to(torch.bfloat16)
            res = {}
            for m in (0, 1, 4 | 16, 29, 64):
                os.environ["COMB_TS_ABLATE"] = str(m)
                res[m] = round(timed(lambda: ops.spconv_fwd_bf16(x, w, 27, cout, nbr)), 1)
            os.environ["COMB_TS_ABLATE"] = "0"
            print("%dx%d %2d tiles/SM: us by mask %s" % (cin, cout, tps, res))
    # a trivial kernel through the same graph-timing path
    a = torch.zeros(1024, device="cuda")
    print("torch a.add_(1) on 1024 floats: %.1f us" % timed(lambda: a.add_(1)))
