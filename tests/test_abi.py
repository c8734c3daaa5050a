"""The C-ABI library loads on a CPU-only box and exports every symbol include/frb200.h declares.
No compute entry point is driven here (that is what -m gpu is for); only argument validation that
returns before any CUDA call."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "frb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from facerecognition_b200 import _native as N
    names = _declared()
    assert len(names) >= 13
    raw = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in frb200.h but not exported by libfrb200.so"
        assert n in N.SIGNATURES, f"{n} has no ctypes signature in _native.py"
    assert sorted(N.SIGNATURES) == names


def test_version_and_argument_errors():
    from facerecognition_b200 import _native as N
    assert N.lib.frb_version() >= 100
    # invalid arguments are rejected before anything touches the device
    assert N.lib.frb_topk_merge(None, None, 1, 4, 0, 1, None, None, None) == N.FRB_ERR_INVALID
    assert "k=0" in N.last_error()
    assert N.lib.frb_lbp_hist_u8(None, 1, 100, 100, 2, 9, 8, 8, None, None, None) == N.FRB_ERR_UNSUPPORTED   # > 8 neighbours
    assert "neighbors 1..8" in N.last_error()
    assert N.lib.frb_lbp_hist_u8_counts8(None, 1, 600, 600, 1, 8, 8, 8, None, None, None) == N.FRB_ERR_UNSUPPORTED  # cells > 255 px
    assert N.lib.frb_chisq_top1_filtered_g8(None, 1, None, 1, 100, 144, 0, None, None, None, None, None, 0, None) == N.FRB_ERR_INVALID
    assert N.lib.frb_chisq_top1_filtered_g8(None, 1, None, 1, 16384, 300, 0, None, None, None, None, None, 0, None) == N.FRB_ERR_INVALID
    assert N.lib.frb_cosine_rescore_topk(None, 1, None, 1, 512, None, None, 0, None, None, 4, 5, 0.0, 0.0, 0, None, None, None, None, None) == N.FRB_ERR_INVALID
    assert N.lib.frb_lbp_hist_u8(None, 1, 2, 100, 1, 8, 8, 8, None, None, None) == N.FRB_ERR_INVALID
    assert N.lib.frb_chisq_topk(None, 1, 144, None, 1, 100, 144, 1, 0, None, None, None, 0, None) == N.FRB_ERR_INVALID
    assert N.lib.frb_cosine_topk(None, 1, None, 0, 1, 100, None, None, 0, 0, 1, 0, None, None, None, 0, None) == N.FRB_ERR_INVALID
    assert N.lib.frb_cosine_topk(None, 1, None, 0, 1, 512, None, None, 0, 0, 65, 0, None, None, None, 0, None) == N.FRB_ERR_INVALID
    # empty work is a no-op success
    assert N.lib.frb_row_norms_f32(None, 0, 512, None, None) == N.FRB_OK
    assert N.lib.frb_lbp_codes_u8(None, 0, 100, 100, 1, 8, None, None) == N.FRB_OK


def test_python_layer_refuses_cpu_tensors():
    import pytest
    import torch
    from facerecognition_b200 import ops
    with pytest.raises(ValueError, match="no CPU path"):
        ops.lbp_hist(torch.zeros((1, 10, 10), dtype=torch.uint8))
    with pytest.raises(ValueError, match="no CPU path"):
        ops.cosine_topk(torch.zeros((1, 512)), torch.zeros((4, 512)), 1)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under facerecognition_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "facerecognition_b200")
    pat = re.compile(r"(^|\n)\s*(from|import)\s+oracle\b|liblbph_oracle|oracle/_ref|frb_oracle_")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{os.path.join(dirpath, f)} reaches into oracle/"
