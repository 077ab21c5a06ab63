import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A machine WITHOUT a CUDA device skips the gpu-marked tests (they used to error).  A machine WITH one never
    skips: a missing or broken libgenpf_cuda.so must fail loudly there (fixture `g`), not hide behind a skip."""
    if "not gpu" in (config.getoption("markexpr", "") or ""):
        return  # the CPU suite deselects them anyway: no need to import torch
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have = False
    if not have:
        skip = pytest.mark.skip(reason="no CUDA device on this machine (the library has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.load()
    return oracle


@pytest.fixture(scope="session")
def g():
    """The product package; GPU tests only.  Fails loudly if the CUDA library is missing."""
    import genpf_b200
    genpf_b200.load()
    return genpf_b200


GOLDEN_CASES = ["n100_s1", "n1000_s5", "n2048_s2", "n3000_s1"]
