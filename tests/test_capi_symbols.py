"""The C-ABI library loads on a machine without a GPU and exports every symbol include/amg1d.h
declares (no compute calls here)."""
import ctypes as C
import os
import re

from agglomerationmultigrid1d_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "amg1d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amg1d_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_bindings_and_library(lib):
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted(capi.PROTOTYPES) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_loads_without_gpu_and_reports_errors(lib):
    assert lib.amg1d_version() >= 100
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        rc = lib.amg1d_create(C.byref(h), 2, 0, None)
        assert rc in (capi.ERR_CUDA, capi.ERR_ARG)
        assert len(lib.amg1d_last_error(None)) > 0
        assert not h.value


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "agglomerationmultigrid1d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src, f
