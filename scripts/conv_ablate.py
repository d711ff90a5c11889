"""Which piece of the tcgen05 gather-GEMM pipeline bounds it?  The production kernel on the real level rulebooks of the
bench batch with pieces switched off (COMB_TS_ABLATE bit mask; results are garbage, only the time is read):
 1 no tcgen05.mma, 2 no tcgen05.st, 4 no index LDS / LDG / tcgen05.st, 8 no epilogue (tcgen05.ld + stores),
 16 no index-tile cp.async, 32 no look-ahead barrier probes in the MMA threads."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from com_b200 import ops
import conv_trace_ts as ct
from conv_floor import timed  # noqa: E402  (runs nothing at import: guarded below)

MASKS = [0, 1, 2, 4, 8, 16, 32, 1 | 8, 4 | 16, 1 | 4 | 16, 1 | 4 | 8 | 16]
if os.environ.get("COMB_CONV_IMPL") == "tr":
    MASKS = [0, 1, 2, 4, 8, 16, 4 | 16, 1 | 4 | 16, 1 | 4 | 8 | 16]
if __name__ == "__main__":
    for cin, cout, lv in ((16, 16, 1), (32, 32, 2), (64, 64, 3), (128, 128, 4)):
        cd, idx = ct.level_coords(lv)
        n = int(cd.shape[0])
        real = ops.nbrmap_build_indexed(cd, idx, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1])
        x = torch.randn((n, cin), device="cuda").to(torch.bfloat16)
        w = ops.pack_weight_bf16(torch.randn((cout, 27, cin), device="cuda") / 20)
        tiles = (n + 127) // 128
        out = {}
        for m in MASKS:
            os.environ["COMB_TS_ABLATE"] = str(m)
            us = timed(lambda: ops.spconv_fwd_bf16(x, w, 27, cout, real))
            out[m] = (round(us, 1), int(us * 1e-6 * 1.9e9 / (tiles / 148.0)))
        os.environ["COMB_TS_ABLATE"] = "0"
        print("level %d %dx%d rows %d tiles/SM %.1f: mask -> (us, cycles per tile per SM) %s" % (lv, cin, cout, n, tiles / 148.0, out))
