"""Pipeline trace of the tensor-memory-A sparse-conv kernel (CTA 0): where does a chunk's time go?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ctypes
import numpy as np, torch
from com_b200 import _lib, ops
import conv_trace as ct_mod  # noqa: E402  (re-uses real_coords)


def level_coords(level):
    """key-ordered voxels of 4 synthetic frames at backbone level 1..4 (index sets via the strided-conv output sets)"""
    cd, idx = ct_mod.real_coords()
    shape = [41, 1504, 1504]
    batch = 4
    pads = {2: [1, 1, 1], 3: [1, 1, 1], 4: [0, 1, 1]}
    for lv in range(2, level + 1):
        n_in = int(cd.shape[0])
        oshape = ops.conv_out_shape(shape, [3, 3, 3], [2, 2, 2], pads[lv], [1, 1, 1])
        idx = ops.index_build(cd, batch, oshape, conv=([3, 3, 3], [2, 2, 2], pads[lv], [1, 1, 1]), out_cap=n_in)
        m = int(idx.count.item())
        cd = idx.coords[:m].contiguous()
        shape = oshape
    return cd, idx


def run(cin, cout, level):
    cd, idx = level_coords(level)
    n = int(cd.shape[0])
    nbr = ops.nbrmap_build_indexed(cd, idx, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1])
    x = torch.randn((n, cin), device="cuda").to(torch.bfloat16)
    w = ops.pack_weight_bf16(torch.randn((cout, 27, cin), device="cuda") / 20)
    lib = _lib.load()
    buf = torch.zeros((512 * 16 + 256 * 4,), dtype=torch.int64, device="cuda")
    for _ in range(3):
        ops.spconv_fwd_bf16(x, w, 27, cout, nbr)
    torch.cuda.synchronize()
    lib.comb_debug_conv_trace(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.spconv_fwd_bf16(x, w, 27, cout, nbr); e1.record()
    torch.cuda.synchronize()
    lib.comb_debug_conv_trace(None)
    full = buf.cpu().numpy()
    ct = full[512 * 16:].reshape(256, 4)
    ct = ct[ct[:, 0] > 0]
    base = ct[:, 0].min()
    print("== level %d cin %d cout %d rows %d: kernel %.1f us, fill %.2f, CTAs %d" % (
        level, cin, cout, n, e0.elapsed_time(e1) * 1e3, float((nbr >= 0).float().mean()), len(ct)))
    print('  per-CTA wall clock (us): start max %.1f, setup done max %.1f, mma done min/median/max %.1f/%.1f/%.1f, exit max %.1f' % (
        (ct[:, 0].max() - base) / 1e3, (ct[:, 1].max() - base) / 1e3, (ct[:, 2].min() - base) / 1e3,
        float(np.median(ct[:, 2]) - base) / 1e3, (ct[:, 2].max() - base) / 1e3, (ct[:, 3].max() - base) / 1e3))
    t = full[:512 * 16].reshape(512, 16)
    ok = (t[:, 1] > 0) & (t[:, 2] > 0)
    G = np.nonzero(ok)[0]
    t0 = t[ok][:, 2].min()
    print("    G  gath:start   gath:idx gath:empty gath:stored |  mma:top  mma:full mma:issued")
    for g in G[:12]:
        r = t[g]
        print("%5d %10d %10d %10d %10d | %8d %9d %9d" % (g, r[2] - t0, r[7] - t0, r[3] - t0, r[4] - t0, r[6] - t0, r[0] - t0, r[1] - t0))
    iss = t[ok][:, 1]
    print("  mma issue-to-issue: %.0f cycles/STAGE (4 chunks) over %d stages -> %.0f cycles/chunk" % (np.diff(iss).mean(), len(iss), np.diff(iss).mean() / 4))
    print("  mma: top->full seen %.0f (waiting for data), full seen->issued+committed %.0f" % (
        (t[ok][:, 0] - t[ok][:, 6]).mean(), (t[ok][:, 1] - t[ok][:, 0]).mean()))
    print("  gather: start->idx %.0f, idx->empty seen (index LDS + LDG issue + slot wait) %.0f, empty->stored+arrived (LDG data + tcgen05.st) %.0f, stored->mma saw full %.0f" % (
        (t[ok][:, 7] - t[ok][:, 2]).mean(), (t[ok][:, 3] - t[ok][:, 7]).mean(), (t[ok][:, 4] - t[ok][:, 3]).mean(), (t[ok][:, 0] - t[ok][:, 4]).mean()))
    epi = t[:, 5][t[:, 5] > 0]
    if len(epi) > 2: print("  epilogue tile-to-tile: %.0f cycles" % np.diff(epi).mean())


if __name__ == "__main__":
    for cin, cout, lv in ((16, 16, 1), (32, 32, 2), (64, 64, 3), (128, 128, 4)):
        run(cin, cout, lv)
