"""Dataset used by the DataLoader-worker tests: every item runs the reference's COMAug box checks
(pcdet/datasets/augmentor/database_sampler_v2.py:535-539,600-604 call these two functions) inside the WORKER process,
through the reference wrappers over the drop-in pybind modules.  Lives in its own module so that spawned workers can
unpickle it."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class BoxOpDataset(torch.utils.data.Dataset):
    def __init__(self, n=4):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        from com_b200 import synth
        from oracle import ref_py
        iou = ref_py.load("pcdet.ops.iou3d_nms.iou3d_nms_utils")
        bu = ref_py.load("pcdet.utils.box_utils")
        pts = synth.make_small_cloud(20000, seed=10 + i, extent=(60.0, 60.0, 3.0))
        boxes = synth.make_boxes(35, seed=i, rng_xy=30.0)
        cand = synth.make_clustered_boxes(64, seed=100 + i)
        kept = bu.remove_points_in_boxes3d(pts, boxes)
        m = iou.boxes_bev_iou_cpu(cand[:, 0:7], boxes[:, 0:7])
        return {"i": i, "kept": int(kept.shape[0]), "kept_sum": float(kept[:, :3].astype(np.float64).sum()),
                "iou": torch.from_numpy(np.ascontiguousarray(m)), "pid": os.getpid(),
                "cuda_inited": bool(torch.cuda.is_initialized())}


def first(batch):
    """collate_fn (module level: picklable for spawned workers)"""
    return batch[0]
