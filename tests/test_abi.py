"""The C-ABI library loads and exports every symbol include/comb200.h declares; the ctypes table in
com_b200._lib lists exactly those symbols.  No compute calls (no GPU needed)."""
import ctypes
import os
import re

from com_b200 import _lib, build
from util import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "comb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(comb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libcomb200.so does not export %s" % s


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_library_calls_without_gpu():
    lib = _lib.load()
    assert lib.comb_version() >= 100
    assert lib.comb_hash_slots(1000) == 2048 and lib.comb_hash_slots(0) == 1024
    assert lib.comb_nms_workspace_bytes(500) >= 500 * 8 * 8
    assert lib.comb_spconv_packed_bytes(16, 27, 16) == 7 * 16 * 128
    assert lib.comb_spconv_packed_bytes(128, 27, 128) == 54 * 128 * 128
    assert lib.comb_spconv_packed_bytes(24, 27, 16) == 0
    assert lib.comb_voxelize_workspace_bytes(180000, 1, 150000, 5) > 0
    # wgrad workspace = (row chunks) x (Cout*K*Cin) fp32 partials; one chunk per SM (148) when a CTA holds all 27 offsets
    assert lib.comb_spconv_wgrad_bf16_workspace_bytes(16, 16, 16, 27, 160000) == 148 * 16 * 27 * 16 * 4
    assert lib.comb_spconv_wgrad_bf16_workspace_bytes(16, 5, 16, 27, 160000) == 148 * 16 * 27 * 5 * 4
    tiles = -(-15000 // 64)                      # 7 offset groups at 128 channels: 148 // 7 chunks of whole 64-row tiles
    chunks = -(-tiles // -(-tiles // (148 // 7)))
    assert lib.comb_spconv_wgrad_bf16_workspace_bytes(128, 128, 128, 27, 15000) == chunks * 128 * 27 * 128 * 4
    assert lib.comb_spconv_wgrad_bf16_workspace_bytes(24, 24, 16, 27, 100) == 0      # unsupported channel count


def test_argument_errors_are_status_codes():
    lib = _lib.load()
    rc = lib.comb_hash_build(None, 10, None, 1, 4, 4, 4, None, 64, None)
    assert rc == -1 and b"comb_hash_build" in lib.comb_last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under com_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "com_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
