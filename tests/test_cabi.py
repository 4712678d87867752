"""CPU checks of the drop-in boundary: the library builds, loads, exports every symbol that
include/dpmm_b200.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import numpy as np
import pytest

import dpmm_pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    return dpmm_pkg.load()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "dpmm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dpmm_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound(pkg):
    lib = pkg._lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dpmm_b200.h but not exported"
    assert sorted(pkg._lib.SIGNATURES) == declared, "ctypes binding and header disagree"


def test_limits_and_timing_names(pkg):
    import ctypes as C
    lib = pkg._lib.load()
    out = (C.c_int32 * 3)()
    assert lib.dpmm_limits(out) == 0
    assert out[0] >= 64 and out[2] >= 256
    names = [lib.dpmm_timing_name(i).decode() for i in range(lib.dpmm_timing_kinds())]
    assert names[:4] == ["label", "sort", "sublabel", "stats"]


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg._lib.DpmmError) as ei:
        pkg.GpuSweep(np.zeros((2, 8), np.float32), pkg.NIW)
    assert ei.value.code == pkg._lib.ECUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "dpmmsubclusters.jl_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"
