"""oracle/resize.py — CPU restatement of cv2.resize(img, dsize) at its default INTER_LINEAR for 8-bit images.
TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's CPU leg may import it, nothing else.

The reference resizes before the LBPH path: `cv2.resize(image, target_size)` followed by
`cv2.cvtColor(cropped, cv2.COLOR_BGR2GRAY)` (models/lbphmodel/train_lbph_script.py:67-72, web_app.py:472-475,
484-486).  OpenCV itself is a dependency, not vendored; this restates imgproc/src/resize.cpp (4.x) for CV_8U:

* `cv::resize`: inv_scale = dsize / ssize (double); `hal::resize`: scale = 1. / inv_scale;
* per destination coordinate d: f = (float)((d + 0.5) * scale - 0.5); s = cvFloor(f); f -= s;
  weights = saturate_cast<short>((1.f - f) * 2048), saturate_cast<short>(f * 2048)   (cvRound: half to even);
* columns: s < 0 -> (s, f) = (0, 0); s >= src - 1 -> (src - 1, 0); rows: weights kept, row indices clamped
  (`clip(sy + k, 0, ssize.height)` in resizeGeneric_Invoker), so a border row is blended with itself;
* HResizeLinear: r = S[s] * a0 + S[s + 1] * a1 (int); VResizeLinear<uchar, int, short, ...>:
  dst = ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
* INTER_LINEAR with both axes shrinking by exactly 2 is rerouted to the INTER_AREA fast path
  (`if (interpolation == INTER_LINEAR && is_area_fast && iscale_x == 2 && iscale_y == 2) interpolation = INTER_AREA`):
  (p00 + p01 + p10 + p11 + 2) >> 2.

PARITY PINNED: cv2.resize is in the installed opencv-python-headless (4.13.0, here and on the GPU box);
tests/test_oracle_lbph.py compares this file with it bit for bit over seeded size pairs (up- and down-scaling,
1 and 3 channels, the exact-halving case, 1-pixel sides).
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = np.float32(1 << COEF_BITS)


def resize_taps(src: int, dst: int, clamp_weights: bool):
    """(first source index int64 [dst], w0 int32 [dst], w1 int32 [dst]) for one axis."""
    scale = 1.0 / (float(dst) / float(src))
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        lo = s < 0
        s[lo], f[lo] = 0, 0
        hi = s >= src - 1
        s[hi], f[hi] = src - 1, 0
    w0 = np.rint(((np.float32(1.0) - f).astype(np.float32) * COEF_SCALE).astype(np.float32)).astype(np.int32)
    w1 = np.rint((f * COEF_SCALE).astype(np.float32)).astype(np.int32)
    return s, w0, w1


def resize_linear_u8(src: np.ndarray, dst_cols: int, dst_rows: int) -> np.ndarray:
    """src u8 [H, W] or [H, W, C] -> u8 [dst_rows, dst_cols(, C)]; argument order as cv2.resize's dsize = (cols, rows)."""
    assert src.dtype == np.uint8 and src.ndim in (2, 3)
    sh, sw = src.shape[:2]
    s = src.astype(np.int32)
    if sw == 2 * dst_cols and sh == 2 * dst_rows:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xo, a0, a1 = resize_taps(sw, dst_cols, True)
    yo, b0, b1 = resize_taps(sh, dst_rows, False)
    tail = (1,) * (src.ndim - 2)
    x1 = np.minimum(xo + 1, sw - 1)
    rows = s[:, xo] * a0.reshape((1, -1) + tail) + s[:, x1] * a1.reshape((1, -1) + tail)
    y0, y1 = np.clip(yo, 0, sh - 1), np.clip(yo + 1, 0, sh - 1)
    b0, b1 = b0.reshape((-1, 1) + tail), b1.reshape((-1, 1) + tail)
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def bgr2gray_u8(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(bgr, COLOR_BGR2GRAY) for 8-bit images (OpenCV 4.x fixed point, imgproc/src/color_rgb.simd.hpp)."""
    b, g, r = (bgr[..., c].astype(np.int64) for c in range(3))
    return ((3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15).astype(np.uint8)


def preprocess_for_lbph(image_bgr: np.ndarray, target_size=(100, 100), grayscale: bool = True) -> np.ndarray:
    """_preprocess_image_for_lbph without a detector (train_lbph_script.py:49-76): resize, then gray."""
    out = resize_linear_u8(image_bgr, target_size[0], target_size[1])
    return bgr2gray_u8(out) if grayscale else out
