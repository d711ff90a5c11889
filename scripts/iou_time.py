"""Device time of the rotated BEV IoU matrix kernel on the COMAug 10k x 10k set and the 500-box NMS set."""
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from com_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
for name, b_np in (("10k x 10k", np.concatenate([synth.make_clustered_boxes(9000, seed=61, centers=400),
                                                 synth.make_boxes(1000, seed=64)]).astype(np.float32)),
                   ("500 x 500", synth.make_clustered_boxes(500, seed=21))):
    b = torch.from_numpy(b_np).to(dev)
    t = torch.from_numpy(ops.box_trig4_host(b_np)).to(dev)
    out = torch.empty((len(b_np), len(b_np)), dtype=torch.float32, device=dev)
    for flavour in ("cpu", "gpu"):
        fn = lambda: ops.boxes_bev(b, b, flavour=flavour, what="iou", trig_a=t, trig_b=t, out=out)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        torch.cuda._sleep(10_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("%s flavour %s: %.4f ms, %d pairs > 0" % (name, flavour, e0.elapsed_time(e1) / 10, int((out > 0).sum())))
