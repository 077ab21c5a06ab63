"""CPU: the C-ABI library builds, loads and exports every symbol include/genpf.h declares;
without a CUDA device compute calls fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "genpf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(genpf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    import genpf_b200
    lib = genpf_b200.load()
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"libgenpf_cuda.so does not export {n}"
    assert set(names) == set(genpf_b200._lib.SIGNATURES), "python binding and header disagree"
    assert lib.genpf_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import genpf_b200
    lw = np.zeros(16)
    with pytest.raises(genpf_b200.GenPFError) as e:
        genpf_b200.logsumexp_host(lw)
    assert e.value.status == genpf_b200._lib.ERR_CUDA


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "genparticlefilters.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"
