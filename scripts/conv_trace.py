"""Pipeline trace of the tcgen05 sparse-conv kernel (CTA 0): where does a chunk's time go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np, torch
import oracle
from com_b200 import _lib, ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import random_coords

_real = {}


def real_coords():
    """key-ordered level-1 voxels of 4 synthetic Waymo frames"""
    if "c" not in _real:
        from com_b200 import synth
        fr = [synth.make_frame(seed=1000 + b) for b in range(4)]
        offs = np.concatenate([[0], np.cumsum([len(f) for f in fr])]).astype(int).tolist()
        r = ops.voxelize(torch.from_numpy(np.concatenate(fr)).cuda(), offs, synth.VOXEL_SIZE, synth.POINT_CLOUD_RANGE, 5,
                         150000, want_voxels=False)
        m = int(r["counts"][4])
        idx = ops.index_build(r["coords"][:m].contiguous(), 4, [41, 1504, 1504])
        _real["c"] = (idx.coords[:m].contiguous(), idx)
    return _real["c"]


def run(cin, cout, n=300000, real=False):
    rng = np.random.default_rng(0)
    if real:
        cd, idx = real_coords()
        n = int(cd.shape[0])
    else:
        shape = [16, 400, 400]
        coords = random_coords(rng, n, 1, shape)
        c = coords.astype(np.int64)
        coords = coords[np.argsort(((c[:, 0] * 16 + c[:, 1]) * 400 + c[:, 2]) * 400 + c[:, 3])]
        cd = torch.from_numpy(coords).cuda()
        idx = ops.index_build(cd, 1, shape)
    nbr = ops.nbrmap_build_indexed(cd, idx, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1])
    x = torch.randn((n, cin), device="cuda").to(torch.bfloat16)
    w = ops.pack_weight_bf16(torch.randn((cout, 27, cin), device="cuda") / 20)
    lib = _lib.load()
    buf = torch.zeros((512 * 8 + 256 * 4,), dtype=torch.int64, device="cuda")
    for _ in range(3):
        ops.spconv_fwd_bf16(x, w, 27, cout, nbr)
    torch.cuda.synchronize()
    lib.comb_debug_conv_trace(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.spconv_fwd_bf16(x, w, 27, cout, nbr); e1.record()
    torch.cuda.synchronize()
    lib.comb_debug_conv_trace(None)
    full = buf.cpu().numpy()
    ct = full[512 * 8:].reshape(256, 4)
    ct = ct[ct[:, 0] > 0]
    base = ct[:, 0].min()
    print('  per-CTA wall clock (us): start min/max %.1f/%.1f, setup done max %.1f, mma done min/median/max %.1f/%.1f/%.1f, exit max %.1f' % (
        0, (ct[:, 0].max() - base) / 1e3, (ct[:, 1].max() - base) / 1e3, (ct[:, 2].min() - base) / 1e3, float(np.median(ct[:, 2]) - base) / 1e3, (ct[:, 2].max() - base) / 1e3, (ct[:, 3].max() - base) / 1e3))
    slow = np.argsort(ct[:, 2])[-5:]
    print('  slowest CTAs (index: mma done us):', [(int(i), round(float(ct[i, 2] - base) / 1e3, 1)) for i in slow], ' CTA0:', round(float(ct[0, 2] - base) / 1e3, 1))
    t = full[:512 * 8].reshape(512, 8)
    mm = np.nonzero(t[:, 0] > 0)[0]            # chunk slots at which the MMA warp stamped (one per stage)
    S = int(mm[1] - mm[0]) if len(mm) > 1 else 1
    pr = t[:, 2] > 0
    t0 = t[pr][:, 2].min()
    print("== cin %d cout %d: kernel %.1f us, fill %.2f, stage = %d chunks, traced stages %d" % (
        cin, cout, e0.elapsed_time(e1) * 1e3, float((nbr >= 0).float().mean()), S, len(mm)))
    print("  G   prod:loads  prod:empty  prod:stored |  mma:full  mma:issued")
    for G in range(0, min(40, 512)):
        r = t[G]
        if r[2] == 0: continue
        extra = "  %10d %10d" % (r[0] - t0, r[1] - t0) if r[0] > 0 else ""
        print("%4d %11d %11d %11d |%s" % (G, r[2] - t0, r[3] - t0, r[4] - t0, extra))
    iss = t[mm][:, 1]
    print("  mean cycles/stage (mma issue to issue): %.0f  -> %.0f per chunk" % (np.diff(iss).mean(), np.diff(iss).mean() / S))
    print("  mma: full seen -> issued %.0f cycles (of which proxy fence %.0f); issued -> next full seen %.0f" % (
        (t[mm][:, 1] - t[mm][:, 0]).mean(), (t[mm][:, 6] - t[mm][:, 0]).mean(), (t[mm][1:, 0] - t[mm][:-1, 1]).mean()))
    print("  prod: loads issued->empty seen %.0f, empty seen->stored+arrived %.0f" % (
        (t[pr][:, 3] - t[pr][:, 2]).mean(), (t[pr][:, 4] - t[pr][:, 3]).mean()))
    epi = t[:, 5][t[:, 5] > 0]
    if len(epi) > 2: print("  epilogue tile-to-tile: %.0f cycles" % np.diff(epi).mean())

if __name__ == "__main__":
    for cin, cout in ((16, 16), (64, 64)):
        run(cin, cout, real=True)
