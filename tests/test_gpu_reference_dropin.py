"""The REFERENCE's own Python, unmodified, executed on top of the drop-ins (com_b200.install_dropins()).

What runs here is the reference code itself (loaded by oracle/ref_py.py from /root/reference when it is mounted,
else from the byte-code built into oracle/_ref/pcdet_bc): pcdet/models/backbones_3d/spconv_backbone.py
(VoxelResBackBone8x through the registry dictionary), vfe/mean_vfe.py, map_to_bev/height_compression.py,
pcdet/ops/iou3d_nms/iou3d_nms_utils.py, pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py,
pcdet/utils/box_utils.py, pcdet/datasets/processor/data_processor.py and
pcdet/models/model_utils/model_nms_utils.py.  Only `spconv`, `cumm` and the two pybind modules underneath are ours.
Results are held against the CPU oracle at the usual bars: indices / masks / keep lists bit-exact, features 1e-4
(fp32 check mode) and 2e-2 (bf16 tensor-core path)."""
import os

import numpy as np
import pytest
import torch

import oracle
from com_b200 import models, pipeline, sparse, synth
from oracle import build_ref, cpu_pipeline, ref_py

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_py.available(), reason="reference Python not available (oracle/_ref/pcdet_bc)")]

RANGE, VSIZE = [-12.8, -12.8, -2.0, 12.8, 12.8, 4.0], [0.1, 0.1, 0.15]      # grid 256 x 256 x 40
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "box_ops_ref.npz")


def rel_err(got, want):
    return float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


def key_order(coords, shape):
    c = coords.astype(np.int64)
    return np.argsort(((c[:, 0] * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3], kind="stable")


def randomize_bn(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.5)


@pytest.fixture(scope="module")
def ref():
    """The reference modules + the registries its package __init__s would define."""
    reg = ref_py.registry()
    return {
        "reg": reg, "E": ref_py.EasyDict,
        "data_processor": ref_py.load("pcdet.datasets.processor.data_processor"),
        "iou": ref_py.load("pcdet.ops.iou3d_nms.iou3d_nms_utils"),
        "roi": ref_py.load("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils"),
        "box_utils": ref_py.load("pcdet.utils.box_utils"),
        "nms_utils": ref_py.load("pcdet.models.model_utils.model_nms_utils"),
    }


@pytest.fixture(scope="module")
def scene(ref):
    """Two small frames -> the reference's DataProcessor (VoxelGeneratorWrapper over the spconv.utils drop-in) ->
    the reference's collate layout -> the reference's MeanVFE; weights shared with the mirror; CPU oracle beside."""
    E = ref["E"]
    frames = [synth.make_small_cloud(n, seed=s, extent=(25.0, 25.0, 4.0)) for s, n in ((1, 30000), (2, 22000))]
    for f in frames:
        f[:, 2] = f[:, 2] * 0.4
    dp_mod = ref["data_processor"]
    proc = dp_mod.DataProcessor(
        dataset_cfg=E(), point_cloud_range=np.array(RANGE, dtype=np.float32), training=False, num_point_features=5,
        processor_configs=[E(NAME="transform_points_to_voxels", VOXEL_SIZE=VSIZE, MAX_POINTS_PER_VOXEL=5,
                             MAX_NUMBER_OF_VOXELS={"train": 40000, "test": 40000})])
    per_frame = []
    for f in frames:
        d = {"points": f, "use_lead_xyz": True}
        for fn in proc.data_processor_queue:
            d = fn(data_dict=d)
        per_frame.append(d)
    # DatasetTemplate.collate_batch, voxel branch (pcdet/datasets/dataset.py:252-259)
    voxels = np.concatenate([d["voxels"] for d in per_frame], axis=0)
    num = np.concatenate([d["voxel_num_points"] for d in per_frame], axis=0)
    coords = np.concatenate([np.pad(d["voxel_coords"], ((0, 0), (1, 0)), mode="constant", constant_values=i)
                             for i, d in enumerate(per_frame)], axis=0)
    # the mirror provides weights (same state_dict keys as the reference class)
    pipe = pipeline.FramePipeline(point_cloud_range=RANGE, voxel_size=VSIZE, max_voxels=40000, seed=3)
    randomize_bn(pipe.backbone, 4)
    sd = {k: v.detach().cpu() for k, v in pipe.backbone.state_dict().items()}
    ref_levels, ref_sf, ref_coords = cpu_pipeline.frame_forward(frames, sd, VSIZE, RANGE, 5, 40000)
    return dict(frames=frames, per_frame=per_frame, voxels=voxels, num=num, coords=coords, proc=proc, pipe=pipe, sd=sd,
                ref_levels=ref_levels, ref_sf=ref_sf, ref_coords=ref_coords)


def batch_dict(scene):
    """load_data_to_gpu (pcdet/models/__init__.py:23-37): everything .float().cuda(), coords included."""
    return {"batch_size": 2, "voxels": torch.from_numpy(scene["voxels"]).float().cuda(),
            "voxel_num_points": torch.from_numpy(scene["num"]).float().cuda(),
            "voxel_coords": torch.from_numpy(scene["coords"]).float().cuda()}


def build_reference_modules(ref, scene):
    E, reg = ref["E"], ref["reg"]
    vfe = reg["vfe"].__all__["MeanVFE"](model_cfg=E(NAME="MeanVFE"), num_point_features=5)
    # exactly what Detector3DTemplate.build_backbone_3d does (detector3d_template.py:75-90): the class comes out of
    # the registry dictionary, grid_size is the DataProcessor's numpy array
    bb = reg["backbones_3d"].__all__["VoxelResBackBone8x"](
        model_cfg=E(NAME="VoxelResBackBone8x"), input_channels=5, grid_size=scene["proc"].grid_size,
        voxel_size=VSIZE, point_cloud_range=np.array(RANGE, dtype=np.float32)).cuda()
    missing = bb.load_state_dict(scene["sd"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    hc = reg["map_to_bev"].__all__["HeightCompression"](model_cfg=E(NAME="HeightCompression", NUM_BEV_FEATURES=256))
    return vfe, bb, hc


def test_data_processor_voxels_bit_exact(ref, scene):
    """pcdet/datasets/processor/data_processor.py:15-60,125-153 (unmodified) over spconv.utils.Point2VoxelCPU3d /
    cumm.tensorview of the drop-in."""
    assert type(scene["proc"].voxel_generator).__module__ == "pcdet.datasets.processor.data_processor"
    assert list(scene["proc"].grid_size) == [256, 256, 40]
    for f, d in zip(scene["frames"], scene["per_frame"]):
        v, c, m = oracle.voxelize(f, VSIZE, RANGE, 5, 40000)
        assert d["voxels"].dtype == np.float32 and d["voxel_coords"].dtype == np.int32
        assert np.array_equal(d["voxel_coords"], c) and np.array_equal(d["voxel_num_points"], m)
        assert np.array_equal(d["voxels"], v)
    assert np.array_equal(scene["coords"], scene["ref_coords"])


@pytest.mark.parametrize("compute", ["f32", "bf16"])
def test_reference_backbone_module_path(ref, scene, compute):
    """Reference VFE + VoxelResBackBone8x.forward + HeightCompression line by line (COMB_FUSED=0: every SubMConv3d /
    SparseConv3d / BatchNorm1d / ReLU is called by the reference's own forward) in both arithmetic forms."""
    vfe, bb, hc = build_reference_modules(ref, scene)
    bb.eval()
    old, old_env = sparse.config.compute, os.environ.get("COMB_FUSED")
    sparse.config.compute, os.environ["COMB_FUSED"] = compute, "0"
    try:
        with torch.no_grad():
            bd = hc(bb(vfe(batch_dict(scene))))
    finally:
        sparse.config.compute = old
        os.environ.pop("COMB_FUSED") if old_env is None else os.environ.__setitem__("COMB_FUSED", old_env)
    tol = 1e-4 if compute == "f32" else 2e-2
    names = ["x_conv1", "x_conv2", "x_conv3", "x_conv4"]
    got = [bd["multi_scale_3d_features"][n] for n in names] + [bd["encoded_spconv_tensor"]]
    for t, (wf, wc, wshape) in zip(got, scene["ref_levels"]):
        assert [int(s) for s in t.spatial_shape] == wshape
        assert np.array_equal(t.indices.cpu().numpy(), wc)                       # voxel order kept by SubM
        assert rel_err(t.features.float().cpu().numpy(), wf) < tol
    sf = bd["spatial_features"].cpu().numpy()
    assert sf.shape == scene["ref_sf"].shape == (2, 256, 32, 32) and rel_err(sf, scene["ref_sf"]) < tol
    assert bd["spatial_features_stride"] == 8


def test_reference_backbone_reaches_fused_path(ref, scene):
    """A backbone instantiated from the reference's registry takes the fused bf16 tensor-core path in eval mode
    (com_b200.models.patch_reference_backbone, installed by install_dropins' post-import hook) and equals the mirror
    bit for bit; rows are in key order; 2e-2 against the fp32 oracle."""
    vfe, bb, hc = build_reference_modules(ref, scene)
    assert getattr(type(bb), "_comb_fused_patch", False) and models._fusable(bb)
    bb.eval()
    with torch.no_grad():
        bd = hc(bb(vfe(batch_dict(scene))))
    mirror = scene["pipe"].forward_host(scene["frames"])
    names = ["x_conv1", "x_conv2", "x_conv3", "x_conv4"]
    got = [bd["multi_scale_3d_features"][n] for n in names] + [bd["encoded_spconv_tensor"]]
    mir = [mirror["multi_scale_3d_features"][n] for n in names] + [mirror["encoded_spconv_tensor"]]
    for t, m, (wf, wc, wshape) in zip(got, mir, scene["ref_levels"]):
        assert t.features.dtype == torch.bfloat16
        o = key_order(wc, wshape)
        assert np.array_equal(t.indices.cpu().numpy(), wc[o])
        assert rel_err(t.features.float().cpu().numpy(), wf[o]) < 2e-2
        assert torch.equal(t.indices, m.indices)
        # the reference MeanVFE (eager torch, fp32) feeds the reference-instantiated backbone, the mirror's input
        # comes from the fused voxelizer mean in bf16: identical after the bf16 rounding of the first layer's input
        assert rel_err(t.features.float().cpu().numpy(), m.features.float().cpu().numpy()) < 1e-2
    sf = bd["spatial_features"].cpu().numpy()
    assert rel_err(sf, scene["ref_sf"]) < 2e-2
    # training mode / autograd keeps the reference's own forward
    bb.train()
    bd2 = bb(vfe(batch_dict(scene)))
    assert bd2["encoded_spconv_tensor"].features.dtype == torch.float32
    assert bd2["encoded_spconv_tensor"].features.requires_grad


@pytest.mark.parametrize("compute,fused", [("f32", "0"), ("bf16", "0"), ("bf16", "1")])
def test_reference_train_step_through_height_compression(ref, scene, compute, fused, monkeypatch):
    """train(): loss taken on spatial_features, i.e. THROUGH HeightCompression -> SparseConvTensor.dense(): every
    backbone parameter receives a gradient (dense() is an autograd op).  On the module path (COMB_FUSED_TRAIN=0) the
    gradients equal the mirror's (same modules underneath, 1e-5); on the fused train step (bf16 tensor cores, batch
    statistics in the library) they meet the bf16 tolerance of north_star, 2e-2, against the fp32-accumulating mirror."""
    monkeypatch.setenv("COMB_FUSED_TRAIN", fused)
    tol = 1e-5
    vfe, bb, hc = build_reference_modules(ref, scene)
    bb.train()
    old = sparse.config.compute
    sparse.config.compute = compute
    try:
        bd = hc(bb(vfe(batch_dict(scene))))
        sf = bd["spatial_features"]
        assert sf.requires_grad and sf.grad_fn is not None
        g = torch.Generator(device="cuda").manual_seed(0)
        wgt = torch.randn(sf.shape, device="cuda", generator=g)
        (sf * wgt).sum().backward()
        grads = {k: p.grad.detach().clone() for k, p in bb.named_parameters()}
        assert all(p.grad is not None for p in bb.parameters())
        # a convolution bias that feeds a BatchNorm in train mode has a mathematically ZERO gradient (the batch mean
        # absorbs it): rounding noise in the module path, exact zeros in the fused train step — both are right
        before_bn = lambda k: k.endswith(("conv1.bias", "conv2.bias"))
        bad = [k for k, v in grads.items() if not torch.isfinite(v).all() or (not before_bn(k) and float(v.abs().max()) == 0)]
        assert not bad, bad
        scale = max(float(v.abs().max()) for v in grads.values())
        # the mirror class in module mode: same modules underneath, so the same numbers
        mb = models.VoxelResBackBone8x(None, 5, [256, 256, 40], fused=False).cuda()
        mb.load_state_dict(scene["sd"])
        mb.train()
        bd_m = models.HeightCompression(None)(mb(vfe(batch_dict(scene))))
        (bd_m["spatial_features"] * wgt).sum().backward()
        if fused == "0":
            for k, p in mb.named_parameters():
                denom = float(grads[k].abs().max()) if not before_bn(k) else scale
                assert float((p.grad - grads[k]).abs().max()) <= tol * denom, (k, float((p.grad - grads[k]).abs().max()), denom)
        else:
            # bf16 activations / activation gradients across 21 layers with ReLU masks that flip under rounding: the
            # bar of tests/test_gpu_train_fused.py::test_fused_train_step_matches_module_path (direction of the gradient)
            ga = [p.grad.flatten() for k, p in mb.named_parameters() if not before_bn(k)]
            gb = [grads[k].flatten() for k, p in mb.named_parameters() if not before_bn(k)]
            cos_all = float(torch.nn.functional.cosine_similarity(torch.cat(ga), torch.cat(gb), dim=0))
            cos_each = [float(torch.nn.functional.cosine_similarity(a, b, dim=0)) for a, b in zip(ga, gb) if float(a.norm()) > 1e-6]
            # (this scene with a random-sign loss weight measures 0.973 overall, 0.885 worst tensor)
            assert cos_all > 0.95 and float(np.median(cos_each)) > 0.95 and min(cos_each) > 0.8, (cos_all, min(cos_each))
        # dense() backward against autograd of an index_put formulation of the same scatter
        enc = bd["encoded_spconv_tensor"]
        x = enc.features.detach().clone().requires_grad_(True)
        idx = enc.indices.long()
        dense = torch.zeros((2, 128, *[int(s) for s in enc.spatial_shape]), device="cuda")
        dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = x
        (dense.view(sf.shape) * wgt).sum().backward()
        from com_b200 import ops
        got = ops.dense_gather(wgt.view(dense.shape).contiguous(), enc.indices)
        assert torch.equal(got, x.grad)
    finally:
        sparse.config.compute = old


def test_reference_bev_backbone_on_bf16_nhwc(ref, scene):
    """f4: with sparse.config.bev = "bf16" the reference's own HeightCompression hands the reference's own
    BaseBEVBackbone (base_bev_backbone.py:81-112) a channels-last bf16 image and the backbone runs under bf16 autocast;
    spatial_features_2d comes back as fp32 within the bf16 tolerance (2e-2) of the fp32 NCHW run, in eval and — with
    gradients reaching the sparse backbone — in train mode."""
    E, reg = ref["E"], ref["reg"]
    vfe, bb, hc = build_reference_modules(ref, scene)
    torch.manual_seed(0)
    bev = reg["backbones_2d"].__all__["BaseBEVBackbone"](
        model_cfg=E(NAME="BaseBEVBackbone", LAYER_NUMS=[2, 2], LAYER_STRIDES=[1, 2], NUM_FILTERS=[64, 128],
                    UPSAMPLE_STRIDES=[1, 2], NUM_UPSAMPLE_FILTERS=[128, 128]), input_channels=256).cuda()
    assert type(hc).forward._comb and type(bev).forward._comb                # the post-import hooks are in place
    old = sparse.config.bev
    try:
        outs = {}
        for mode in ("f32", "bf16"):
            sparse.config.bev = mode
            bb.eval(); bev.eval()
            with torch.no_grad():
                bd = bev(hc(bb(vfe(batch_dict(scene)))))
            sf, out = bd["spatial_features"], bd["spatial_features_2d"]
            assert out.dtype == torch.float32 and out.shape[1] == 256
            if mode == "bf16":
                assert sf.dtype == torch.bfloat16 and sf.is_contiguous(memory_format=torch.channels_last)
                assert torch.equal(sf, outs["sf"].to(torch.bfloat16))        # same image, other layout / precision
            else:
                outs["sf"] = sf
            outs[mode] = out
        err = float((outs["bf16"] - outs["f32"]).abs().max() / outs["f32"].abs().max())
        assert err < 2e-2, err
        # train mode: gradients flow through the bf16 image back into the sparse backbone
        sparse.config.bev = "bf16"
        bb.train(); bev.train()
        bd = bev(hc(bb(vfe(batch_dict(scene)))))
        bd["spatial_features_2d"].square().mean().backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in bb.parameters())
        assert all(p.grad is not None and p.grad.dtype == torch.float32 for p in bev.parameters())
        assert float(next(bb.parameters()).grad.abs().max()) > 0
    finally:
        sparse.config.bev = old


def test_reference_box_op_wrappers(ref):
    """iou3d_nms_utils.py:12-116, roiaware_pool3d_utils.py:9-41, box_utils.py:117-131, model_nms_utils.py:6-25 — the
    reference wrappers themselves, over the drop-in pybind modules."""
    iou, roi, bu, nms_utils = ref["iou"], ref["roi"], ref["box_utils"], ref["nms_utils"]
    assert iou.iou3d_nms_cuda.__name__.startswith("com_b200") and roi.roiaware_pool3d_cuda.__name__.startswith("com_b200")
    g = np.load(GOLDEN)
    # golden vectors generated from the compiled reference (tests/golden/make_golden.py)
    for name in ("uc", "cc", "self", "known"):
        a, b, want = g["iou_%s_a" % name], g["iou_%s_b" % name], g["iou_%s" % name]
        got = iou.boxes_bev_iou_cpu(torch.from_numpy(a), torch.from_numpy(b))
        assert isinstance(got, torch.Tensor) and np.array_equal(got.numpy(), want), name
    got_np = iou.boxes_bev_iou_cpu(g["iou_cc_a"], g["iou_cc_b"])                  # numpy in -> numpy out
    assert isinstance(got_np, np.ndarray) and np.array_equal(got_np, g["iou_cc"])
    m = roi.points_in_boxes_cpu(torch.from_numpy(g["pib_points"]), torch.from_numpy(g["pib_boxes"]))
    assert m.dtype == torch.int32 and int(m.sum()) == int(g["pib_hits"][0])
    assert np.array_equal(np.packbits(m.numpy().astype(np.uint8), axis=1), g["pib_mask_packed"])
    # remove_points_in_boxes3d: the accelerated function (any-box kernel) == the reference's own body == the oracle
    pts = synth.make_frame(seed=5)[:60000]
    boxes = synth.make_boxes(35, seed=2)
    fast = bu.remove_points_in_boxes3d(pts, boxes)
    slow = bu.remove_points_in_boxes3d.reference(pts, boxes)
    want = pts[oracle.points_in_boxes_cpu(pts[:, :3].copy(), boxes).sum(0) == 0]
    assert isinstance(fast, np.ndarray) and np.array_equal(fast, slow) and np.array_equal(fast, want)
    assert 0 < len(fast) < len(pts)
    # device flavour: NMS through class_agnostic_nms with the reference's config object
    E = ref["E"]
    bx = torch.from_numpy(synth.make_clustered_boxes(600, seed=7)).cuda()
    sc = torch.rand(600, generator=torch.Generator().manual_seed(1)).cuda()
    cfg = E(NMS_TYPE="nms_gpu", NMS_THRESH=0.7, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500)
    sel, sel_scores = nms_utils.class_agnostic_nms(sc, bx, cfg, score_thresh=0.1)
    if build_ref.available():
        ref_iou = build_ref.load_ref("ref_iou3d_nms_cuda")
        mask = sc >= 0.1
        s2, b2 = sc[mask], bx[mask]
        top, ind = torch.topk(s2, k=min(4096, s2.shape[0]))
        order = top.sort(0, descending=True)[1]
        sb = b2[ind][order].contiguous()
        keep = torch.empty(sb.shape[0], dtype=torch.int64)
        n = ref_iou.nms_gpu(sb, keep, 0.7)
        want_sel = mask.nonzero().view(-1)[ind[order[keep[:n].cuda()][:500]]]
        assert torch.equal(sel, want_sel)
    assert torch.equal(sel_scores, sc[sel]) and 0 < sel.numel() <= 500
    # boxes_iou3d_gpu / boxes_iou_bev: diagonal of a self comparison is 1
    iou3d = iou.boxes_iou3d_gpu(bx[:50], bx[:50])
    assert torch.allclose(torch.diagonal(iou3d), torch.ones(50, device="cuda"), atol=1e-4)
    pig = roi.points_in_boxes_gpu(torch.from_numpy(pts[None, :5000, :3].copy()).cuda(),
                                  torch.from_numpy(boxes[None]).cuda())
    assert pig.shape == (1, 5000) and int(pig.max()) < 35 and int(pig.min()) >= -1


@pytest.mark.parametrize("S,E,seed", [(20, 60, 0), (15, 0, 1), (64, 400, 2), (1, 5, 3)])
def test_comaug_placement_and_scene_update_vs_reference_lines(ref, S, E, seed):
    """f3: the device side of a COMAug sampler step.  The reference lines (database_sampler_v2.py:600-611 and
    :535-539) are executed here verbatim through the reference's own wrappers (boxes_bev_iou_cpu, enlarge_box3d,
    remove_points_in_boxes3d); the fused helpers must select the same boxes and produce the same point cloud."""
    from com_b200.pcdet_ops import box_ops
    iou, bu = ref["iou"], ref["box_utils"]
    sampled_boxes = synth.make_clustered_boxes(S, seed=50 + seed).astype(np.float32)
    existed_boxes = synth.make_boxes(E, seed=seed, rng_xy=40.0).astype(np.float32) if E else np.zeros((0, 7), np.float32)
    if S > 2 and E:
        sampled_boxes[1, :7] = existed_boxes[0, :7]                      # a certain collision with the scene
    # --- reference lines 600-611
    iou1 = iou.boxes_bev_iou_cpu(sampled_boxes[:, 0:7], existed_boxes[:, 0:7])
    iou2 = iou.boxes_bev_iou_cpu(sampled_boxes[:, 0:7], sampled_boxes[:, 0:7])
    iou2[range(sampled_boxes.shape[0]), range(sampled_boxes.shape[0])] = 0
    iou1 = iou1 if iou1.shape[1] > 0 else iou2
    valid_mask = ((iou1.max(axis=1) + iou2.max(axis=1)) == 0)
    want_idx = valid_mask.nonzero()[0]
    want_existed = np.concatenate((existed_boxes, sampled_boxes[want_idx][:, :existed_boxes.shape[-1]]), axis=0)
    got_idx, got_existed = box_ops.comaug_place_sampled_boxes(sampled_boxes, existed_boxes)
    assert np.array_equal(got_idx, want_idx) and np.array_equal(got_existed, want_existed)
    if S > 2 and E:
        assert len(want_idx) < S and 1 not in want_idx
    # --- reference lines 535-539
    pts = synth.make_small_cloud(40000, seed=seed, extent=(60.0, 60.0, 3.0))
    obj_points = synth.make_small_cloud(500, seed=90 + seed, extent=(5.0, 5.0, 2.0))
    sampled_gt_boxes = want_existed[existed_boxes.shape[0]:, :]
    if len(sampled_gt_boxes):
        large = bu.enlarge_box3d(sampled_gt_boxes[:, 0:7], extra_width=[0.2, 0.2, 0.2])
        want_pts = bu.remove_points_in_boxes3d.reference(pts, large)
        want_pts = np.concatenate([obj_points[:, :want_pts.shape[-1]], want_pts], axis=0)
        got_pts = box_ops.comaug_add_to_scene(pts, sampled_gt_boxes, obj_points, extra_width=[0.2, 0.2, 0.2])
        assert got_pts.dtype == want_pts.dtype and np.array_equal(got_pts, want_pts)


def _worker_expect(i):
    pts = synth.make_small_cloud(20000, seed=10 + i, extent=(60.0, 60.0, 3.0))
    boxes = synth.make_boxes(35, seed=i, rng_xy=30.0)
    cand = synth.make_clustered_boxes(64, seed=100 + i)
    kept = pts[oracle.points_in_boxes_cpu(pts[:, :3].copy(), boxes).sum(0) == 0]
    return kept, oracle.boxes_bev_cpu(cand, boxes)


def test_cpu_named_ops_in_spawned_dataloader_workers():
    """Worker-process policy of the `*_cpu` entry points (ops.host_op_device): under the `spawn` start method — what
    the reference selects for --launcher pytorch/slurm (pcdet/utils/common_utils.py:172-173) — every worker creates its
    CUDA context lazily and the reference's COMAug calls return the oracle's results."""
    from worker_ds import BoxOpDataset, first
    torch.zeros(1).cuda()                                   # the parent owns a context, as in training
    dl = torch.utils.data.DataLoader(BoxOpDataset(4), batch_size=1, num_workers=2, multiprocessing_context="spawn",
                                     collate_fn=first)
    pids = set()
    for item in dl:
        kept, iou = _worker_expect(item["i"])
        assert item["kept"] == kept.shape[0] and item["kept_sum"] == float(kept[:, :3].astype(np.float64).sum())
        assert np.array_equal(item["iou"].numpy() == 0, iou == 0)      # the consumer only tests == 0 (:604)
        assert np.array_equal(item["iou"].numpy(), iou)
        assert item["cuda_inited"] and item["pid"] != os.getpid()
        pids.add(item["pid"])
    assert len(pids) == 2


def test_cpu_named_ops_in_forked_worker_fail_with_remedy():
    """A worker FORKED from a process that already initialised CUDA cannot reach the GPU: the drop-ins raise a
    RuntimeError that names the remedy (spawn start method) instead of crashing inside the driver."""
    from worker_ds import BoxOpDataset, first
    torch.zeros(1).cuda()
    dl = torch.utils.data.DataLoader(BoxOpDataset(2), batch_size=1, num_workers=1, multiprocessing_context="fork",
                                     collate_fn=first)
    with pytest.raises(RuntimeError) as e:
        for _ in dl:
            pass
    assert "spawn" in str(e.value) and "forked" in str(e.value)
