"""Small run of the atomics-based kernels (voxelizer: hash claim + atomicMin chains; bitmap index: atomicOr marks,
block scans; bitmap rulebook; hash rulebook) for compute-sanitizer --tool racecheck / synccheck / memcheck."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from com_b200 import ops, synth

rng_, vs = [-8.0, -8.0, -2.0, 8.0, 8.0, 4.0], [0.1, 0.1, 0.15]
frames = [synth.make_small_cloud(6000, seed=s, extent=(16.0, 16.0, 3.0)) for s in (1, 2)]
offs = [0, 6000, 12000]
pts = torch.from_numpy(np.concatenate(frames)).cuda()
r = ops.voxelize(pts, offs, vs, rng_, 5, 8000, mean_dtype=torch.bfloat16, mean_ld=16)
m = int(r["counts"][2])
coords = r["coords"][:m].contiguous()
shape = [41, 160, 160]
idx = ops.index_build(coords, 2, shape, want_coords=False)
perm, _ = ops.index_rank(coords, idx, scatter_coords=True)
nbr = ops.nbrmap_build_indexed(idx.coords, idx, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], no_dev=idx.count)
cv = ([3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1])
oshape = ops.conv_out_shape(shape, *cv)
oidx = ops.index_build(idx.coords, 2, oshape, conv=cv, out_cap=m, n_dev=idx.count)
nbr_d = ops.nbrmap_build_indexed(oidx.coords, idx, *cv, no_dev=oidx.count)
table, slots = ops.hash_build(coords, 2, shape)
nbr_h = ops.nbrmap_build(coords, table, slots, 2, shape, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1])
oc, cnt = ops.conv_out_coords(coords, 2, oshape, *cv, m * 8)
x = ops.permute_rows(r["mean"], perm, scatter=True)
torch.cuda.synchronize()
print("racecheck target ok: %d voxels, %d level-2 rows, %d pairs" % (m, int(oidx.count.item()), int((nbr >= 0).sum())))

# r2 additions: the row-per-thread tcgen05 conv (mbarrier / tensor-memory hand-shakes), target assignment (atomicMax
# heat maps, block scan), COM loss re-weighting (object-ordered window writes), COMAug placement reduce
feat16 = ops.cast_pad(torch.randn((int(idx.count.item()), 16), device="cuda"), 16)
w16 = ops.pack_weight_bf16(torch.randn((16, 27, 16), device="cuda") / 20)
y = ops.spconv_fwd_bf16(feat16, w16, 27, 16, nbr[:, : feat16.shape[0]].contiguous())
gt = torch.zeros((2, 40, 8), device="cuda")
gt[:, :30, 0:2] = (torch.rand((2, 30, 2), device="cuda") - 0.5) * 100
gt[:, :30, 3:6] = torch.tensor([4.5, 2.0, 1.6], device="cuda")
gt[:, :30, 7] = torch.randint(1, 4, (2, 30), device="cuda").float()
grp = ops.centerhead_cluster_groups(gt, torch.ones((2, 40), device="cuda"), torch.rand((2, 40), device="cuda"),
                                    torch.randint(0, 4, (2, 40), device="cuda"))
cls_map = torch.tensor([-1, 0, 1, 2], dtype=torch.int32, device="cuda")
hm, rb, inds, mask, rmap = ops.centerhead_assign_targets(gt, torch.full((2, 40), 9.0, device="cuda"), grp, cls_map, 3, (188, 188), 8,
                                                         [-75.2, -75.2, -2, 75.2, 75.2, 4], [0.1, 0.1, 0.15])
pred = torch.rand((2, 3, 188, 188), device="cuda").clamp(1e-4, 1 - 1e-4)
ops.comloss_group_confidence(pred, rmap, 3, 96)
ops.comloss_reweight(pred, rmap, mask.clone(), torch.ones_like(hm), 0.05, -10.0, 1.0)
ops.comaug_valid_mask(synth.make_clustered_boxes(40, seed=1), synth.make_boxes(60, seed=2))
torch.cuda.synchronize()
print("r2 kernels ok: conv_tr rows %d, targets %d" % (int(y.shape[0]), int(mask.sum())))

# late r2: conv_ts with streamed weights — full two-tile passes followed by split single-tile passes (two accumulators,
# interleaved weight ring) at 64x64, and a resident-weight split pass at 128x128 K=3; the velocity-head decode
n3 = 150 * 128 * 2 + 700                                     # 148 full passes, then a wave of single-tile passes
nb3 = torch.randint(-1, n3, (27, n3), device="cuda", dtype=torch.int32)
nb3[nb3 % 3 != 0] = -1
f64 = ops.cast_pad(torch.randn((n3, 64), device="cuda"), 64)
y3 = ops.spconv_fwd_bf16(f64, ops.pack_weight_bf16(torch.randn((64, 27, 64), device="cuda") / 40), 27, 64, nb3)
n4 = 3000
nb4 = torch.randint(-1, n4, (3, n4), device="cuda", dtype=torch.int32)
f128 = ops.cast_pad(torch.randn((n4, 128), device="cuda"), 128)
y4 = ops.spconv_fwd_bf16(f128, ops.pack_weight_bf16(torch.randn((128, 3, 128), device="cuda") / 20), 3, 128, nb4)
hmv = torch.randn((2, 3, 64, 64), device="cuda") - 2.0
bv = ops.centerhead_decode_nms(hmv, torch.rand((2, 2, 64, 64), device="cuda"), torch.zeros((2, 1, 64, 64), device="cuda"),
                               torch.zeros((2, 3, 64, 64), device="cuda"), torch.randn((2, 2, 64, 64), device="cuda"), 100, 8,
                               [0.1, 0.1, 0.15], [-25.6, -25.6, -2, 25.6, 25.6, 4], [-30, -30, -5, 30, 30, 5], 0.1, 0.7, 4096, 83,
                               label_map=torch.arange(3, dtype=torch.int32, device="cuda"),
                               vel=torch.randn((2, 2, 64, 64), device="cuda"))
torch.cuda.synchronize()
print("late r2 kernels ok: conv_ts rows %d + %d, decode counts %s" % (int(y3.shape[0]), int(y4.shape[0]), bv[3].tolist()))
