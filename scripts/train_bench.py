"""Full-size fused train-mode step (2 Waymo-shaped frames): timing, or CUDA_LAUNCH_BLOCKING=1 fault finding."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from com_b200 import _lib, models, ops, sparse, synth

fused = "--module" not in sys.argv
sparse.config.compute, sparse.config.wgrad = "bf16", "bf16"
fr = [synth.make_frame(seed=1000 + b) for b in range(2)]
offs = np.concatenate([[0], np.cumsum([len(f) for f in fr])]).astype(int).tolist()
pts = torch.from_numpy(np.concatenate(fr, axis=0)).cuda()
torch.manual_seed(0)
net = models.VoxelResBackBone8x(None, 5, synth.GRID_SIZE).cuda().train()
net.fused = fused
opt = torch.optim.SGD(net.parameters(), lr=1e-3)


def step():
    r = ops.voxelize(pts, offs, synth.VOXEL_SIZE, synth.POINT_CLOUD_RANGE, 5, 150000, want_voxels=False,
                     mean_dtype=torch.float32)
    m = int(r["counts"][2])
    bd = net({"batch_size": 2, "voxel_features": r["mean"][:m, :5].contiguous(), "voxel_coords": r["coords"][:m].float()})
    loss = bd["encoded_spconv_tensor"].features.float().square().mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss.detach()


lib = _lib.load()
for i in range(3):
    l = step()
    torch.cuda.synchronize()
    print("warm-up step", i, "loss", float(l), flush=True)
n0 = lib.comb_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(10):
    l = step()
e1.record()
torch.cuda.synchronize()
print("%s train step: %.3f ms per step (wall %.3f), %d libcomb200 launches per step, loss %.5f" % (
    "fused" if fused else "module", e0.elapsed_time(e1) / 10, (time.perf_counter() - t0) * 100, (lib.comb_launch_count() - n0) // 10, float(l)))
if "--prof" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
