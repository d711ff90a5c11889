"""CPU checks of the machinery that runs the reference's own Python on the drop-ins (oracle/ref_py.py +
com_b200.install_dropins): imports resolve to the mirrors, post-import hooks patch the reference classes, the
byte-code build works without the source tree.  The numerical tests of the same modules are the -m gpu tests in
test_gpu_reference_dropin.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_py.available(), reason="reference Python not available")


def test_reference_modules_resolve_to_dropins():
    reg = ref_py.registry()
    sb = ref_py.load("pcdet.models.backbones_3d.spconv_backbone")
    from com_b200 import models, sparse
    assert sb.spconv.SubMConv3d is sparse.SubMConv3d and sb.spconv.SparseConvTensor is sparse.SparseConvTensor
    cls = reg["backbones_3d"].__all__["VoxelResBackBone8x"]
    assert cls.__module__ == "pcdet.models.backbones_3d.spconv_backbone" and cls._comb_fused_patch
    bb = cls(ref_py.EasyDict(NAME="VoxelResBackBone8x"), 5, np.array([1504, 1504, 40]))
    assert isinstance(bb.conv_input[0], sparse.SubMConv3d) and models._fusable(bb)
    assert [int(v) for v in bb.sparse_shape] == [41, 1504, 1504] and bb.out_spatial_shape() == [2, 188, 188]
    mirror = models.VoxelResBackBone8x(None, 5, [1504, 1504, 40])
    assert list(bb.state_dict().keys()) == list(mirror.state_dict().keys())
    assert all(a.shape == b.shape for a, b in zip(bb.state_dict().values(), mirror.state_dict().values()))
    # spconv_utils.find_all_spconv_keys (spconv_utils.py:10-25) sees the conv weights by name
    su = ref_py.load("pcdet.utils.spconv_utils")
    keys = su.find_all_spconv_keys(bb)
    assert "conv_input.0.weight" in keys and "conv_out.0.weight" in keys and len(keys) == 21
    # non-residual backbone of the same file: not fusable, keeps the reference forward
    vb = reg["backbones_3d"].__all__["VoxelBackBone8x"](ref_py.EasyDict(), 5, np.array([1504, 1504, 40]))
    assert not models._fusable(vb)
    iou = ref_py.load("pcdet.ops.iou3d_nms.iou3d_nms_utils")
    roi = ref_py.load("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils")
    assert iou.iou3d_nms_cuda.__name__ == "com_b200.pcdet_ops.iou3d_nms_cuda"
    assert roi.roiaware_pool3d_cuda.__name__ == "com_b200.pcdet_ops.roiaware_pool3d_cuda"
    bu = ref_py.load("pcdet.utils.box_utils")
    assert bu.remove_points_in_boxes3d._comb and callable(bu.remove_points_in_boxes3d.reference)
    dp = ref_py.load("pcdet.datasets.processor.data_processor")
    from com_b200 import voxel
    assert dp.tv.from_numpy is voxel.from_numpy      # the generator class itself is imported lazily (data_processor.py:17-26)
    assert "CurriculumCenterHead_x5" in reg["dense_heads"].__all__
    assert ref_py.load("pcdet.models.detectors.centerpoint").CenterPoint.__name__ == "CenterPoint"


def test_cpu_tensors_raise_no_fallback():
    """Without a GPU the `*_cpu` entry points raise — there is no host implementation behind them."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    iou = ref_py.load("pcdet.ops.iou3d_nms.iou3d_nms_utils")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        iou.boxes_bev_iou_cpu(np.zeros((2, 7), np.float32), np.zeros((2, 7), np.float32))


def test_bytecode_build_loads_without_source_tree():
    if not ref_py.source_available():
        pytest.skip("byte-code is built where the reference tree is mounted")
    assert ref_py.build_pyc() >= 19
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import ref_py\n"
            "assert not ref_py.source_available() and ref_py.pyc_available()\n"
            "reg = ref_py.registry()\n"
            "m = ref_py.load('pcdet.models.backbones_3d.spconv_backbone')\n"
            "assert m.__file__.endswith('.bc') and m.VoxelResBackBone8x._comb_fused_patch\n"
            "assert 'CurriculumCenterHead_x5' in reg['dense_heads'].__all__\n"
            "print('ok')\n" % ROOT)
    env = dict(os.environ, COM_REFERENCE="/nonexistent")
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_round2_hooks_attach_and_leave_cpu_tensors_to_the_reference():
    """The §8(f) hooks: every patched method keeps the reference's own function as `.reference`, and inputs the device
    kernels do not cover (CPU tensors here) run the reference method unchanged — the drop-in never silently computes
    on another path."""
    import torch
    E = ref_py.EasyDict
    ch = ref_py.load("pcdet.models.dense_heads.curriculum_center_head")
    chp = ref_py.load("pcdet.models.dense_heads.center_head")
    lu = ref_py.load("pcdet.utils.loss_utils")
    hc = ref_py.load("pcdet.models.backbones_2d.map_to_bev.height_compression")
    bev = ref_py.load("pcdet.models.backbones_2d.base_bev_backbone")
    for fn in (ch.CurriculumCenterHead.assign_targets, ch.CurriculumCenterHead.cluster,
               ch.CurriculumCenterHead.generate_predicted_boxes, chp.CenterHead.assign_targets,
               chp.CenterHead.generate_predicted_boxes, lu.FocalLossCenterCurriculum.neg_loss,
               hc.HeightCompression.forward, bev.BaseBEVBackbone.forward):
        assert fn._comb and callable(fn.reference) and not getattr(fn.reference, "_comb", False)
    from com_b200.pcdet_ops import center_targets
    head = object.__new__(ch.CurriculumCenterHead)
    cfg = E(TARGET_ASSIGNER_CONFIG=E(FEATURE_MAP_STRIDE=8, NUM_MAX_OBJS=500, GAUSSIAN_OVERLAP=0.1, MIN_RADIUS=2))
    head.__dict__.update(model_cfg=cfg, class_names=["Vehicle", "Pedestrian", "Cyclist"],
                         class_names_each_head=[["Vehicle", "Pedestrian", "Cyclist"]],
                         point_cloud_range=np.array([-75.2, -75.2, -2, 75.2, 75.2, 4], dtype=np.float32),
                         voxel_size=[0.1, 0.1, 0.15], epoch=0, epoch_thredhold=100, min_points=1)
    gt = torch.zeros((1, 4, 8))
    gt[0, :3] = torch.tensor([[10.0, 5.0, 0.0, 4.5, 2.0, 1.6, 0.3, 1.0], [-20.0, 7.0, 0.0, 0.9, 0.8, 1.7, 1.0, 2.0],
                              [30.0, -40.0, 0.5, 1.8, 0.8, 1.7, -2.0, 3.0]])
    assert not center_targets.supported_head(head, gt)                    # CPU tensor: the reference method runs
    npgt = torch.full((1, 4), 9.0)
    grp = ch.CurriculumCenterHead.cluster(head, gt.clone(), torch.ones((1, 4)), torch.rand((1, 4)), torch.zeros((1, 4)))
    out = ch.CurriculumCenterHead.assign_targets(head, gt.clone(), feature_map_size=(188, 188), npgt=npgt, true_object=grp)
    assert float(out["masks"][0].sum()) == 3.0 and float(out["heatmaps"][0].max()) == 1.0
    assert out["radius_map"][0].shape == (1, 500, 5)
    # the Gaussian table the device kernels read is the reference's own formula, bit for bit
    cu = ref_py.load("pcdet.models.model_utils.centernet_utils")
    from com_b200 import ops
    tabs, offs = ops._gaussian_tables_host()
    assert len(tabs) == ops._GTAB_RMAX + 1 and int(offs[-1]) == sum((2 * r + 1) ** 2 for r in range(len(tabs)))
    for r, tab in enumerate(tabs):
        d = 2 * r + 1
        want = torch.from_numpy(cu.gaussian2D((d, d), sigma=d / 6)).float().numpy()
        assert tab.dtype == np.float32 and np.array_equal(tab, want), r


def test_fused_postprocessing_coverage_rule():
    """Which CenterHead configurations the fused decode + NMS covers (everything else keeps the reference method):
    rotated `nms_gpu`, at most 1024 candidates per frame; heads with a velocity branch included (nine-column boxes)."""
    import types
    from com_b200.pcdet_ops import center_decode
    E = ref_py.EasyDict

    def head(nms_type="nms_gpu", K=500, order=("center", "center_z", "dim", "rot")):
        cfg = E(POST_PROCESSING=E(MAX_OBJ_PER_SAMPLE=K, NMS_CONFIG=E(NMS_TYPE=nms_type)))
        return types.SimpleNamespace(model_cfg=cfg, separate_head_cfg=E(HEAD_ORDER=list(order)))

    assert center_decode.supported(head())
    assert center_decode.supported(head(order=("center", "center_z", "dim", "rot", "vel")))      # nuScenes / 4-frame Waymo
    assert not center_decode.supported(head(nms_type="circle_nms"))
    assert not center_decode.supported(head(K=4096))
    assert not center_decode.supported(types.SimpleNamespace(model_cfg=E()))                       # no POST_PROCESSING
    # every reference config with a CenterHead falls on one side of the rule without raising
    import glob
    import os
    import yaml
    root = os.path.join(ref_py.REF, "tools", "cfgs")
    if os.path.isdir(root):                                   # only where the reference tree is present (this container)
        seen = 0
        for f in sorted(glob.glob(os.path.join(root, "*_models", "*.yaml"))):
            y = yaml.safe_load(open(f))
            dh = (y.get("MODEL") or {}).get("DENSE_HEAD") or {}
            if "CenterHead" not in str(dh.get("NAME", "")) or "POST_PROCESSING" not in dh:
                continue
            pp = dh["POST_PROCESSING"]
            h = head(pp["NMS_CONFIG"]["NMS_TYPE"], pp["MAX_OBJ_PER_SAMPLE"], dh["SEPARATE_HEAD_CFG"]["HEAD_ORDER"])
            assert center_decode.supported(h) == (pp["NMS_CONFIG"]["NMS_TYPE"] == "nms_gpu" and pp["MAX_OBJ_PER_SAMPLE"] <= 1024), f
            seen += 1
        assert seen >= 5
