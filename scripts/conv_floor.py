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
    """us per call, GPU time: `reps` calls captured in ONE CUDA graph and replayed (a python loop of ctypes launches is
    CPU-bound at ~21 us per call on the bench box, which hid every kernel time below that in the first r2 runs)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


if __name__ == "__main__":
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

