"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import vectorx_b200 as vx
from vectorx_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vectorx_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(vx.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vectorx_b200.h but not exported"


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_constants_entry_point_needs_no_gpu():
    rc = vx.poseidon_round_constants()
    assert hex(int(rc[0])) == "0xb585f766f2144405" and rc.shape == (360,)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vx.VxError, match="no CUDA device|no CPU fallback"):
        vx.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vectorx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
