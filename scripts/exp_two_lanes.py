"""Experiment: two whole-step graphs in flight on two streams (steps of DIFFERENT batches are independent) vs one."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from com_b200 import pipeline

frames = bench.make_frames([1000 + b for b in range(4)])
offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(int).tolist()
pts = torch.from_numpy(np.concatenate(frames, axis=0)).cuda()
NL = 4
pipes = [pipeline.FramePipeline(seed=0, use_graph=True)]
for _ in range(NL - 1):
    pipes.append(pipes[0].clone_lane())
streams = [torch.cuda.Stream() for _ in range(NL)]
for p, s in zip(pipes, streams):
    with torch.cuda.stream(s):
        for _ in range(3):
            p.forward_device(pts, offs)
torch.cuda.synchronize()
K = 40
for lanes in (1, 2, 3, 4, 2, 3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams[:lanes]:
        s.wait_event(e0)
    for k in range(K):
        i = k % lanes
        with torch.cuda.stream(streams[i]):
            pipes[i].enqueue_device(pts, offs)
    for s in streams[:lanes]:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("lanes %d: %.4f ms/step -> %.1f frames/s (no L2 flush)" % (lanes, ms, 4e3 / ms))
