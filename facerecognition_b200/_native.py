"""ctypes binding of libfrb200.so (C ABI: include/frb200.h).

The library is the product: if it is missing or fails to load, importing this module raises —
there is no CPU or PyTorch fallback for any of the kernels.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libfrb200.so")

FRB_OK, FRB_ERR_INVALID, FRB_ERR_UNSUPPORTED, FRB_ERR_CUDA, FRB_ERR_WORKSPACE = 0, -1, -2, -3, -4
FRB_F32, FRB_BF16, FRB_F16 = 0, 1, 2
FRB_QNORM_NONE, FRB_QNORM_CLAMP, FRB_QNORM_EPS = 0, 1, 2
FRB_SCORE_IP, FRB_SCORE_REF_COSINE = 0, 1
FRB_MAX_K = 64
FRB_EXCHANGE_MAX_WORLD, FRB_IPC_HANDLE_BYTES = 8, 64

_STATUS_NAMES = {0: "FRB_OK", -1: "FRB_ERR_INVALID", -2: "FRB_ERR_UNSUPPORTED", -3: "FRB_ERR_CUDA",
                 -4: "FRB_ERR_WORKSPACE"}


class FrbError(RuntimeError):
    """A libfrb200 entry point returned a negative frb_status."""

    def __init__(self, fn: str, status: int, message: str):
        self.fn, self.status, self.message = fn, status, message
        super().__init__(f"{fn} -> {_STATUS_NAMES.get(status, status)}: {message}")


# name -> (restype, argtypes); every symbol include/frb200.h declares
SIGNATURES = {
    "frb_version": (c_int, []),
    "frb_last_error": (c_char_p, []),
    "frb_device_info": (c_int, [c_void_p, c_void_p, c_void_p]),
    "frb_profile_enable": (c_int, [c_int]),
    "frb_profile_read": (c_int, [c_int, c_void_p, c_void_p]),
    "frb_row_norms_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "frb_normalize_rows": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]),
    "frb_cosine_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "frb_cosine_topk": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_int, c_int,
                                c_int, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "frb_cosine_topk_bf16q": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    "frb_cosine_rescore_topk": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                        c_int, c_int, c_float, c_float, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "frb_topk_merge": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "frb_bgr2gray_u8": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "frb_resize_linear_u8": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "frb_group_mean_renorm": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "frb_topk_merge_strided": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p]),
    "frb_exchange_create": (c_int, [c_int, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "frb_exchange_open": (c_int, [c_void_p, c_void_p]),
    "frb_exchange_destroy": (c_int, [c_void_p]),
    "frb_exchange_topk_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "frb_exchange_status": (c_int, [c_void_p, c_void_p, c_void_p]),
    "frb_exchange_reset": (c_int, [c_void_p]),
    "frb_exchange_emulate": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "frb_lbp_codes_u8": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "frb_lbp_hist_u8": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p]),
    "frb_lbp_hist_u8_counts8": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p]),
    "frb_counts_u16_to_u8": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "frb_index_remap": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "frb_chisq_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "frb_chisq_topk": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_void_p]),
    "frb_chisq_dist": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "frb_chisq_topk_g8": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "frb_chisq_dist_g8": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "frb_chisq_filter_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "frb_chisq_top1_filtered_g8": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "frb_chisq_filter_tables": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
}


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C facerecognition_b200/csrc`).  facerecognition_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


def last_error() -> str:
    msg = lib.frb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(fn: str, status: int) -> None:
    if status != FRB_OK:
        raise FrbError(fn, status, last_error())


def call(fn: str, *args) -> None:
    """Call an frb_* function that returns frb_status; raise FrbError on failure."""
    check(fn, getattr(lib, fn)(*args))


K_COSINE_TC, K_COSINE_SIMT, K_LBP_HIST, K_CHISQ, K_BGR2GRAY, K_COSINE_GEMV, K_RESIZE, K_CHISQ_FILTER = 0, 1, 2, 3, 4, 5, 6, 7


def profile_enable(on: bool) -> None:
    call("frb_profile_enable", 1 if on else 0)


def profile_read(kernel: int):
    """(summed device ms, launches) of `kernel` since the last read; waits for those launches."""
    ms, n = c_float(0.0), c_int(0)
    call("frb_profile_read", kernel, ctypes.byref(ms), ctypes.byref(n))
    return ms.value, n.value


def device_info():
    sm, major, minor = c_int(0), c_int(0), c_int(0)
    call("frb_device_info", ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor))
    return sm.value, major.value, minor.value
