"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"`: oracle vs reference / golden vectors, host logic, C-ABI symbol check, gloo sharding.
`-m gpu`      : parity tests proper — CUDA path through the C-ABI vs the CPU oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _fp32_check_mode_by_default():
    """Tests run the module path in the fp32 CHECK arithmetic (1e-4 bar) unless they select the bf16 tensor-core form
    themselves (the product default is bf16, com_b200/sparse.py)."""
    from com_b200 import sparse
    old = (sparse.config.compute, sparse.config.wgrad)
    sparse.config.compute = "f32"
    yield
    sparse.config.compute, sparse.config.wgrad = old
