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


def pytest_sessionstart(session):
    """A fresh clone has no libfrb200.so (built files are git-ignored): build it once with the same recipe as
    __graft_entry__.build() so the C-ABI symbol test and the host-logic tests can load it (nvcc cross-compiles sm_100a
    without a GPU).  The package itself never builds or falls back: importing it without the library raises."""
    lib = os.path.join(ROOT, "facerecognition_b200", "libfrb200.so")
    if os.path.exists(lib):
        return
    import shutil
    import subprocess
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        return                                         # the tests that need the library will say so
    print("\n[conftest] libfrb200.so is missing: building it (make -C facerecognition_b200/csrc, about a minute)", flush=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "facerecognition_b200", "csrc"), "-j8"], check=True,
                   stdout=subprocess.DEVNULL)


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
