import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from com_b200 import pipeline
frames = bench.make_frames([1000, 1001])
for B in (1, 8):
    fr = [frames[i % 2] for i in range(B)]
    eager = pipeline.FramePipeline(seed=0)
    graph = pipeline.FramePipeline(seed=0, use_graph=True)
    graph.backbone = eager.backbone
    e = eager.forward_host(fr); g = graph.forward_host(fr); g = graph.forward_host(fr)
    ok = torch.equal(e["encoded_spconv_tensor"].features, g["encoded_spconv_tensor"].features) and torch.equal(e["spatial_features"], g["spatial_features"])
    print("batch", B, "voxels", int(e["voxel_coords"].shape[0]), "encoded rows", int(e["encoded_spconv_tensor"].features.shape[0]), "graph==eager", ok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    offs = np.concatenate([[0], np.cumsum([len(f) for f in fr])]).astype(int).tolist()
    pts = torch.from_numpy(np.concatenate(fr)).cuda()
    e0.record()
    for _ in range(10): h = graph.enqueue_device(pts, offs)
    e1.record(); torch.cuda.synchronize()
    print("   %.3f ms/step -> %.0f frames/s (one lane)" % (e0.elapsed_time(e1)/10, B*1e4/e0.elapsed_time(e1)))
    del eager, graph; torch.cuda.empty_cache()
