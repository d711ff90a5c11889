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
