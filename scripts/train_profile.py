"""Where the train-mode step of the sparse part goes (configs[2], batch 2): phase times with CUDA events, host wall
time per phase (launch-bound or not) and the kernel table of one step from torch.profiler.
usage: python scripts/train_profile.py [f32|bf16] > gpurun_out/train_profile.txt"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
import bench  # noqa: E402
from com_b200 import models, ops, sparse, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
sparse.config.compute = mode
dev = torch.device("cuda", 0)
fr = bench.make_frames([1000, 1001])
offs = np.concatenate([[0], np.cumsum([len(f) for f in fr])]).astype(int).tolist()
pts = torch.from_numpy(np.concatenate(fr, axis=0)).to(dev)
torch.manual_seed(0)
net = models.VoxelResBackBone8x(None, 5, synth.GRID_SIZE).to(dev).train()
net.fused = False
opt = torch.optim.SGD(net.parameters(), lr=1e-3)
vfe = models.MeanVFE(None, 5)


def step(marks=None):
    def mark(name):
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e, time.perf_counter()))
    mark("start")
    r = ops.voxelize(pts, offs, synth.VOXEL_SIZE, synth.POINT_CLOUD_RANGE, synth.MAX_POINTS_PER_VOXEL,
                     synth.MAX_NUMBER_OF_VOXELS)
    m = int(r["counts"][2])
    bd = vfe({"voxels": r["voxels"][:m], "voxel_num_points": r["num_points"][:m]})
    mark("voxelize+vfe")
    bd = net({"batch_size": 2, "voxel_features": bd["voxel_features"], "voxel_coords": r["coords"][:m].float()})
    loss = bd["encoded_spconv_tensor"].features.float().square().mean()
    mark("forward")
    opt.zero_grad(set_to_none=True)
    loss.backward()
    mark("backward")
    opt.step()
    mark("sgd")
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
for rep in range(2):
    marks = []
    step(marks)
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    print("mode %s rep %d" % (mode, rep))
    for (n0, e0, t0), (n1, e1, t1) in zip(marks[:-1], marks[1:]):
        print("  %-14s device %8.3f ms   host %8.3f ms" % (n1, e0.elapsed_time(e1), (t1 - t0) * 1e3))
    print("  total device %.3f ms, host until sync %.3f ms" % (marks[0][1].elapsed_time(marks[-1][1]),
                                                                (t_end - marks[0][2]) * 1e3))

if len(sys.argv) > 2 and sys.argv[2] == "noprof":      # under ncu: the steps above are all that is needed
    sys.exit(0)
try:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
except Exception as e:  # CUPTI may be unavailable on the box
    print("torch.profiler unavailable: %s: %s" % (type(e).__name__, e))
