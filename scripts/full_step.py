"""BASELINE configs[2], the FULL training step as SURVEY §8(d) defines it: the reference's own CenterPoint detector
(tools/cfgs/waymo_models/centerpoint.yaml) with DENSE_HEAD.NAME CurriculumCenterHead_x5 and the LOSS_CURRICULUM block of
tools/cfgs/waymo_models/com/centercurriculum_pillar_3cls_b2_com.yaml:168-173, built by the reference's
Detector3DTemplate.build_networks out of the registry dictionaries, running UNMODIFIED on the drop-ins:
MeanVFE -> VoxelResBackBone8x (fused train step) -> HeightCompression (autograd dense) -> BaseBEVBackbone (torch/cuDNN,
out of scope) -> CurriculumCenterHead_x5 + COMLoss (reference Python, out of scope) -> backward -> Adam step.
Batch 2 Waymo-shaped frames per GPU, synthetic ground truth.  Prints one JSON object with a per-stage breakdown.

The reference modules come from oracle/ref_py.py (the by-path loader of the reference tree / its byte-code): this is a
measurement of the reference's detector on top of this library, not part of bench.py's product path."""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from com_b200 import ops, synth
from oracle import ref_py


def model_cfg(E):
    # tools/cfgs/waymo_models/centerpoint.yaml:8-72 with the COM head of centercurriculum_pillar_3cls_b2_com.yaml:124-173
    return E(
        NAME='CenterPoint', VFE=E(NAME='MeanVFE'), BACKBONE_3D=E(NAME='VoxelResBackBone8x'),
        MAP_TO_BEV=E(NAME='HeightCompression', NUM_BEV_FEATURES=256),
        BACKBONE_2D=E(NAME='BaseBEVBackbone', LAYER_NUMS=[5, 5], LAYER_STRIDES=[1, 2], NUM_FILTERS=[128, 256],
                      UPSAMPLE_STRIDES=[1, 2], NUM_UPSAMPLE_FILTERS=[256, 256]),
        DENSE_HEAD=E(
            NAME='CurriculumCenterHead_x5', CLASS_AGNOSTIC=False,
            CLASS_NAMES_EACH_HEAD=[['Vehicle', 'Pedestrian', 'Cyclist']], SHARED_CONV_CHANNEL=64,
            USE_BIAS_BEFORE_NORM=True, NUM_HM_CONV=2,
            SEPARATE_HEAD_CFG=E(HEAD_ORDER=['center', 'center_z', 'dim', 'rot'],
                                HEAD_DICT={'center': {'out_channels': 2, 'num_conv': 2},
                                           'center_z': {'out_channels': 1, 'num_conv': 2},
                                           'dim': {'out_channels': 3, 'num_conv': 2},
                                           'rot': {'out_channels': 2, 'num_conv': 2}}),
            TARGET_ASSIGNER_CONFIG=E(FEATURE_MAP_STRIDE=8, NUM_MAX_OBJS=500, GAUSSIAN_OVERLAP=0.1, MIN_RADIUS=2, MIN_POINTS=0),
            LOSS_CONFIG=E(LOSS_WEIGHTS={'cls_weight': 1.0, 'loc_weight': 2.0, 'code_weights': [1.0] * 8}),
            POST_PROCESSING=E(SCORE_THRESH=0.1, POST_CENTER_LIMIT_RANGE=[-75.2, -75.2, -2, 75.2, 75.2, 4],
                              MAX_OBJ_PER_SAMPLE=500,
                              NMS_CONFIG=E(NMS_TYPE='nms_gpu', NMS_THRESH=0.7, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500)),
            LOSS_CURRICULUM=E(UCL=False, THRESHOLD=0.2, ELONGATION=-10, HEIGHT=1, FIX=True)),
        POST_PROCESSING=E(RECALL_THRESH_LIST=[0.3, 0.5, 0.7], EVAL_METRIC='waymo'))


def build_detector():
    E = ref_py.EasyDict
    ref_py.registry()
    cp = ref_py.load('pcdet.models.detectors.centerpoint')
    dataset = types.SimpleNamespace(
        class_names=['Vehicle', 'Pedestrian', 'Cyclist'], point_feature_encoder=types.SimpleNamespace(num_point_features=5),
        grid_size=np.array(synth.GRID_SIZE), point_cloud_range=np.array(synth.POINT_CLOUD_RANGE, dtype=np.float32),
        voxel_size=list(synth.VOXEL_SIZE), depth_downsample_factor=None)
    torch.manual_seed(0)
    return cp.CenterPoint(model_cfg=model_cfg(E), num_class=3, dataset=dataset).cuda()


def make_batch(frames, G=60, seed=0):
    """Voxelized batch in the reference's collate layout + synthetic ground truth with the COM extras
    (curriculum_center_head.py:486-492: true_object, occupancy_ratio, facade_type, num_points_in_gt)."""
    offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(int).tolist()
    pts = torch.from_numpy(np.concatenate(frames, axis=0)).cuda()
    r = ops.voxelize(pts, offs, synth.VOXEL_SIZE, synth.POINT_CLOUD_RANGE, 5, 150000)
    m = int(r["counts"][len(frames)])
    rng = np.random.default_rng(seed)
    B = len(frames)
    gt = np.zeros((B, G, 8), dtype=np.float32)
    for b in range(B):
        bx = synth.make_boxes(G, seed=seed * 10 + b, rng_xy=70.0)
        gt[b, :, :7] = bx
        gt[b, :, 7] = rng.integers(1, 4, G)
    cuda = lambda a: torch.from_numpy(a).cuda()
    return {"batch_size": B, "voxels": r["voxels"][:m], "voxel_num_points": r["num_points"][:m].float(),
            "voxel_coords": r["coords"][:m].float(), "gt_boxes": cuda(gt),
            "num_points_in_gt": cuda(rng.integers(5, 500, (B, G)).astype(np.float32)),
            "true_object": cuda(rng.integers(1, 3, (B, G)).astype(np.float32)),
            "occupancy_ratio": cuda(rng.uniform(0, 1, (B, G)).astype(np.float32)),
            "facade_type": cuda(rng.integers(0, 4, (B, G)).astype(np.float32))}


def main():
    assert torch.cuda.is_available()
    frames = [synth.make_frame(seed=1000 + b) for b in range(2)]
    model = build_detector().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    batch = make_batch(frames)
    stages = {}

    def timed_modules(bd):
        """forward of the reference's CenterPoint.forward loop (centerpoint.py:9-11) with an event pair per module"""
        evs = []
        for mod in model.module_list:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bd = mod(bd)
            e1.record()
            evs.append((type(mod).__name__, e0, e1))
        return bd, evs

    def step(timed=False):
        bd = dict(batch)
        opt.zero_grad(set_to_none=True)
        if timed:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            bd, evs = timed_modules(bd)
            e[0].record()
            loss, tb, _ = model.get_training_loss()
            e[1].record()
            loss.backward()
            e[2].record()
            opt.step()
            e[3].record()
            torch.cuda.synchronize()
            for name, a, b in evs:
                stages[name] = stages.get(name, 0.0) + a.elapsed_time(b)
            stages["COMLoss (get_training_loss)"] = stages.get("COMLoss (get_training_loss)", 0.0) + e[0].elapsed_time(e[1])
            stages["backward (all modules)"] = stages.get("backward (all modules)", 0.0) + e[1].elapsed_time(e[2])
            stages["optimizer"] = stages.get("optimizer", 0.0) + e[2].elapsed_time(e[3])
            return loss
        ret, tb, _ = model(bd)
        ret["loss"].backward()
        opt.step()
        return ret["loss"]

    out = {"workload": "configs[2] full step: reference CenterPoint + CurriculumCenterHead_x5 + COMLoss on the drop-ins, "
                       "batch 2 Waymo-shaped frames, fwd + bwd + AdamW", "modules": [type(m).__name__ for m in model.module_list]}
    # fused_train_step: everything on; reference_head_loops: the reference's own Python target assignment / COM loss loops
    # (COMB_FUSED_TARGETS=0) on the fused backbone; module_path: additionally the per-module backbone
    # fused_train_step_bev_bf16: additionally the 2D backbone on a channels-last bf16 image (f4, sparse.config.bev)
    from com_b200 import sparse as _sparse
    for label, env, tgt, bev in (("fused_train_step", "1", "1", "f32"), ("fused_train_step_bev_bf16", "1", "1", "bf16"),
                                 ("reference_head_loops", "1", "0", "f32"), ("module_path", "0", "0", "f32")):
        os.environ["COMB_FUSED_TRAIN"] = env
        os.environ["COMB_FUSED_TARGETS"] = tgt
        _sparse.config.bev = bev
        for _ in range(3):
            loss = step()
        torch.cuda.synchronize()
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            loss = step()
        torch.cuda.synchronize()
        out[label] = {"ms_per_step": 1e3 * (time.perf_counter() - t0) / reps, "loss": float(loss)}
        stages.clear()
        for _ in range(5):
            step(timed=True)
        out[label]["breakdown_ms"] = {k: round(v / 5, 3) for k, v in stages.items()}
    os.environ.pop("COMB_FUSED_TRAIN", None)
    os.environ.pop("COMB_FUSED_TARGETS", None)
    _sparse.config.bev = "f32"
    # eval: forward + fused post-processing
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            preds, _ = model(dict(batch))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            preds, _ = model(dict(batch))
        torch.cuda.synchronize()
    out["eval_forward_ms"] = 1e3 * (time.perf_counter() - t0) / 10
    _sparse.config.bev = "bf16"
    with torch.no_grad():
        for _ in range(3):
            preds, _ = model(dict(batch))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            preds, _ = model(dict(batch))
        torch.cuda.synchronize()
    out["eval_forward_bev_bf16_ms"] = 1e3 * (time.perf_counter() - t0) / 10
    _sparse.config.bev = "f32"
    out["eval_detections"] = [int(p["pred_boxes"].shape[0]) for p in preds]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
