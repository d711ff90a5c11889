"""Host-side logic that needs no GPU: drop-in import surface, module/state-dict naming, shape
arithmetic, error behaviour of the product path without CUDA, synthetic data determinism."""
import numpy as np
import pytest
import torch

import com_b200
from com_b200 import models, ops, sparse, synth
from com_b200.pcdet_ops import box_ops, iou3d_nms_cuda, roiaware_pool3d_cuda


def test_dropin_import_surface():
    """Names the reference imports: pcdet/utils/spconv_utils.py:3-21, data_processor.py:8-26."""
    com_b200.install_dropins()
    import spconv.pytorch as spconv
    from spconv.utils import Point2VoxelCPU3d  # noqa: F401
    from cumm import tensorview as tv
    for n in ("SparseConvTensor", "SubMConv3d", "SparseConv3d", "SparseInverseConv3d", "SparseSequential",
              "SparseModule"):
        assert hasattr(spconv, n)
    assert issubclass(spconv.SubMConv3d, spconv.conv.SparseConvolution)
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    assert np.array_equal(tv.from_numpy(a).numpy(), a)
    import sys
    assert sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_cuda"] is iou3d_nms_cuda
    assert sys.modules["pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda"] is roiaware_pool3d_cuda


def test_backbone_state_dict_matches_reference_naming():
    """Keys/shapes a reference checkpoint has (spconv_backbone.py:183-232; weight layout
    (Cout,kz,ky,kx,Cin), detector3d_template.py:337-348)."""
    m = models.VoxelResBackBone8x(None, input_channels=5, grid_size=[1504, 1504, 40])
    sd = m.state_dict()
    assert m.sparse_shape == [41, 1504, 1504]
    assert tuple(sd["conv_input.0.weight"].shape) == (16, 3, 3, 3, 5)
    assert tuple(sd["conv1.0.conv1.weight"].shape) == (16, 3, 3, 3, 16) and "conv1.0.conv1.bias" in sd
    assert tuple(sd["conv2.0.0.weight"].shape) == (32, 3, 3, 3, 16) and "conv2.0.0.bias" not in sd
    assert tuple(sd["conv4.2.conv2.weight"].shape) == (128, 3, 3, 3, 128)
    assert tuple(sd["conv_out.0.weight"].shape) == (128, 3, 1, 1, 128)
    assert "conv3.1.bn1.running_var" in sd
    nconv = sum(1 for k in sd if k.endswith(".weight") and sd[k].dim() == 5)
    assert nconv == 21
    nparam = sum(v.numel() for k, v in sd.items() if sd[k].dim() == 5)
    assert nparam == 2691696                                  # 2 691 696 conv weights
    # find_all_spconv_keys (spconv_utils.py:10-25) relies on isinstance(..., SparseConvolution) + '.weight'
    keys = {n + ".weight" for n, mod in m.named_modules() if isinstance(mod, sparse.SparseConvolution)}
    assert len(keys) == 21 and "conv_input.0.weight" in keys


def test_conv_out_shape():
    assert ops.conv_out_shape([41, 1504, 1504], [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1]) == [21, 752, 752]
    assert ops.conv_out_shape([5, 188, 188], [3, 1, 1], [2, 1, 1], [0, 0, 0], [1, 1, 1]) == [2, 188, 188]


def test_sparse_sequential_applies_plain_modules_to_features():
    t = sparse.SparseConvTensor(torch.tensor([[-1.0, 2.0]]), torch.zeros((1, 4), dtype=torch.int32), [1, 1, 1], 1)
    out = sparse.SparseSequential(torch.nn.ReLU())(t)
    assert out.features.tolist() == [[0.0, 2.0]] and out.indices is t.indices
    assert t.replace_feature(t.features * 2).indice_dict is t.indice_dict


def test_no_cpu_fallback():
    """CPU tensors must raise, never silently compute on the host."""
    x = torch.zeros((4, 5))
    with pytest.raises(RuntimeError):
        ops.spconv_fwd_f32(x, torch.zeros((16, 27, 5)), torch.zeros((27, 4), dtype=torch.int32))
    conv = sparse.SubMConv3d(5, 16, 3, indice_key="k")
    t = sparse.SparseConvTensor(x, torch.zeros((4, 4), dtype=torch.int32), [4, 4, 4], 1)
    with pytest.raises(RuntimeError):
        conv(t)
    with pytest.raises(RuntimeError):
        iou3d_nms_cuda.nms_gpu(torch.zeros((3, 7)), torch.zeros(3, dtype=torch.int64), 0.5)
    with pytest.raises(AssertionError):
        box_ops.boxes_bev_iou_cpu(np.zeros((2, 6), np.float32), np.zeros((2, 7), np.float32))
    with pytest.raises(NotImplementedError):
        roiaware_pool3d_cuda.forward()


def test_synthetic_frame_is_seeded_and_waymo_shaped():
    a = synth.make_frame(seed=1000, beams=8, n_az=300, side_rays=100)
    b = synth.make_frame(seed=1000, beams=8, n_az=300, side_rays=100)
    c = synth.make_frame(seed=1001, beams=8, n_az=300, side_rays=100)
    assert a.dtype == np.float32 and a.shape[1] == 5 and np.array_equal(a, b) and not np.array_equal(a[:100], c[:100])
    assert np.abs(a[:, :2]).max() <= 75.2 + 1e-4
    m = synth.make_frame(seed=5, sweeps=3, beams=4, n_az=200, side_rays=50)
    assert m.shape[1] == 6 and sorted(float(v) for v in np.unique(m[:, 5].astype(np.float64)).round(2)) == [0.0, 0.1, 0.2]
    assert synth.make_boxes(10, seed=0).shape == (10, 7)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): stdout carries exactly ONE JSON line with the
    contract's keys (metric/unit of BASELINE.json, impl, cpu_baseline, e2e with zero copy bytes); anything a native
    library prints goes to stderr."""
    import json
    import os
    import subprocess
    import sys
    from util import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and base["metric"].startswith(d["metric"])
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
