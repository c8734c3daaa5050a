import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) when selected on a box without a GPU only if FRB_REQUIRE_GPU=1;
    otherwise they are skipped so a plain `pytest tests/` works on the CPU box."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or os.environ.get("FRB_REQUIRE_GPU") == "1":
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cosine_golden():
    return np.load(os.path.join(GOLDEN, "cosine_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def lbph_golden():
    return np.load(os.path.join(GOLDEN, "lbph_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle_lbph():
    from oracle import lbph
    lbph.build()
    return lbph
