"""Where does the time of the tcgen05 gather-GEMM go?  The same kernel on three synthetic rulebooks of the level sizes of
the bench batch: (a) every neighbour absent (no global load at all: barriers + tcgen05.st of zeros + MMA), (b) every
neighbour = row (o mod 64) (all loads hit L1, one or two lines per instruction), (c) every neighbour present and
pointing at the output row itself + tap (a dense, perfectly local rulebook), next to (d) the real level rulebook."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from com_b200 import ops
import conv_trace_ts as ct


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("COMB_TS_PIPE=%s" % os.environ.get("COMB_TS_PIPE", "0"))
for cin, cout, lv in ((16, 16, 1), (32, 32, 2), (64, 64, 3), (128, 128, 4)):
    cd, idx = ct.level_coords(lv)
    n = int(cd.shape[0])
    real = ops.nbrmap_build_indexed(cd, idx, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1])
    x = torch.randn((n, cin), device="cuda").to(torch.bfloat16)
    w = ops.pack_weight_bf16(torch.randn((cout, 27, cin), device="cuda") / 20)
    o = torch.arange(n, device="cuda", dtype=torch.int32)
    maps = {"absent": torch.full_like(real, -1),
            "l1_hits": (o % 64).repeat(27, 1).contiguous(),
            "dense_local": torch.stack([(o + k - 13).clamp(0, n - 1) for k in range(27)]).contiguous(),
            "real": real}
    fill = float((real >= 0).float().mean())
    res = {k: timed(lambda m=m: ops.spconv_fwd_bf16(x, w, 27, cout, m)) for k, m in maps.items()}
    tiles = (n + 127) // 128
    per_tile = {k: v * 1e-6 * 1.9e9 / (tiles / 148.0) for k, v in res.items()}
    print("level %d %dx%d rows %d fill %.2f: us %s | cycles per 128-row tile per SM %s" % (
        lv, cin, cout, n, fill, {k: round(v, 1) for k, v in res.items()}, {k: int(v) for k, v in per_tile.items()}))
