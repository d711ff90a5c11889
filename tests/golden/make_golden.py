"""Generate tests/golden/box_ops_ref.npz from the REFERENCE's own compiled CPU entry points
(oracle/_ref, built by oracle/build_ref.py from /root/reference) on seeded inputs.

Run here (needs /root/reference or a prebuilt oracle/_ref):  python tests/golden/make_golden.py
The fixture pins oracle/oracle.c's box-op restatement wherever /root/reference is absent.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from com_b200 import synth  # noqa: E402
from oracle import build_ref  # noqa: E402


def main():
    if not build_ref.available():
        build_ref.build()
    iou = build_ref.load_ref("ref_iou3d_nms_cuda")
    roi = build_ref.load_ref("ref_roiaware_pool3d_cuda")
    out = {}

    # rotated BEV IoU: uniform boxes (mostly disjoint) and clustered boxes (many overlaps)
    a = synth.make_boxes(96, seed=0)
    b = synth.make_clustered_boxes(128, seed=1)
    c = synth.make_clustered_boxes(96, seed=2)
    for name, (x, y) in {"uc": (a, b), "cc": (c, b), "self": (c, c)}.items():
        res = torch.zeros((x.shape[0], y.shape[0]), dtype=torch.float32)
        iou.boxes_iou_bev_cpu(torch.from_numpy(x), torch.from_numpy(y), res)
        out["iou_%s_a" % name], out["iou_%s_b" % name], out["iou_%s" % name] = x, y, res.numpy()

    # known answers recorded in SURVEY.md §8c
    ka = np.array([[0, 0, 0, 4, 2, 1.5, 0.3]], dtype=np.float32)
    kb = np.array([[0, 0, 0, 4, 2, 1.5, 0.3], [1, 0.5, 0, 4, 2, 1.5, 0.3], [50, 50, 0, 4, 2, 1.5, 0.3]], dtype=np.float32)
    res = torch.zeros((1, 3), dtype=torch.float32)
    iou.boxes_iou_bev_cpu(torch.from_numpy(ka), torch.from_numpy(kb), res)
    out["iou_known_a"], out["iou_known_b"], out["iou_known"] = ka, kb, res.numpy()

    # points in boxes: points sampled around the boxes so that the mask is not trivially empty
    rng = np.random.default_rng(3)
    boxes = synth.make_boxes(48, seed=4)
    near = boxes[rng.integers(0, 48, 6000), :3] + rng.normal(0, 1.5, size=(6000, 3)).astype(np.float32)
    far = rng.uniform(-75, 75, size=(2000, 3)).astype(np.float32)
    pts = np.concatenate([near, far], axis=0).astype(np.float32)
    mask = torch.zeros((48, pts.shape[0]), dtype=torch.int32)
    roi.points_in_boxes_cpu(torch.from_numpy(boxes), torch.from_numpy(pts), mask)
    out["pib_boxes"], out["pib_points"] = boxes, pts
    out["pib_mask_packed"] = np.packbits(mask.numpy().astype(np.uint8), axis=1)
    out["pib_hits"] = np.array([int(mask.sum())])

    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "box_ops_ref.npz"), **out)
    print("wrote box_ops_ref.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
