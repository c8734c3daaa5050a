"""oracle/lbph.py — Python face of the LBPH CPU oracle.  TEST INFRASTRUCTURE ONLY.

Two independent restatements of OpenCV-contrib's LBPHFaceRecognizer arithmetic
(the reference calls it at models/lbphmodel/train_lbph.py:24-35 and
models/lbphmodel/inference_lbph.py:5; opencv-contrib itself is not vendored):

* ``c_*``  — ctypes bindings of oracle/lbph_oracle.c (fast; used for parity at
  size and as the CPU baseline);
* ``np_*`` — a vectorised NumPy float32 restatement written separately, used by
  the tests to cross-check the C one (two restatements, one answer).

PARITY UNPINNED for the LBP-code/histogram stage (no cv2.face to run);
chi-square is pinned on the installed ``cv2.compareHist`` in tests.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblbph_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/lbph_oracle.c with the committed Makefile."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "lbph_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liblbph_oracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c = ctypes
        L.frb_oracle_elbp.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
        L.frb_oracle_elbp.restype = None
        L.frb_oracle_lbp_hist.argtypes = [c.c_void_p] + [c.c_int] * 6 + [c.c_void_p]
        L.frb_oracle_lbp_hist.restype = c.c_int
        L.frb_oracle_lbp_hist_batch.argtypes = [c.c_void_p] + [c.c_int] * 7 + [c.c_void_p]
        L.frb_oracle_lbp_hist_batch.restype = None
        L.frb_oracle_hist_to_f32.argtypes = [c.c_void_p, c.c_size_t, c.c_int, c.c_void_p]
        L.frb_oracle_hist_to_f32.restype = None
        L.frb_oracle_chisq_alt.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t]
        L.frb_oracle_chisq_alt.restype = c.c_double
        L.frb_oracle_chisq_scan.argtypes = [c.c_void_p, c.c_size_t, c.c_size_t, c.c_void_p, c.c_void_p]
        L.frb_oracle_chisq_scan.restype = None
        L.frb_oracle_chisq_scan_u16.argtypes = [c.c_void_p, c.c_size_t, c.c_size_t, c.c_int, c.c_void_p,
                                                c.c_int, c.c_void_p]
        L.frb_oracle_chisq_scan_u16.restype = None
        L.frb_oracle_predict.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t, c.c_size_t, c.c_void_p,
                                         c.c_double, c.c_void_p, c.c_void_p]
        L.frb_oracle_predict.restype = c.c_long
        L.frb_oracle_lbp_taps.argtypes = [c.c_int, c.c_int, c.c_void_p]
        L.frb_oracle_lbp_taps.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------- C restatement
def c_taps(radius: int = 1, neighbors: int = 8):
    """(fx, fy, cx, cy, w1..w4) per sample point, as elbp_ computes them."""
    dt = np.dtype([("fx", "i4"), ("fy", "i4"), ("cx", "i4"), ("cy", "i4"),
                   ("w1", "f4"), ("w2", "f4"), ("w3", "f4"), ("w4", "f4")])
    out = np.zeros(neighbors, dt)
    lib().frb_oracle_lbp_taps(radius, neighbors, _p(out))
    return out


def c_elbp(img: np.ndarray, radius: int = 1, neighbors: int = 8) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h - 2 * radius, w - 2 * radius), np.int32)
    lib().frb_oracle_elbp(_p(img), h, w, radius, neighbors, _p(out))
    return out


def c_lbp_hist(imgs: np.ndarray, radius=1, neighbors=8, grid_x=8, grid_y=8):
    """u8 [B,H,W] -> (u16 [B, gx*gy*2^P], cell_px)."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    if imgs.ndim == 2:
        imgs = imgs[None]
    b, h, w = imgs.shape
    L = grid_x * grid_y * (1 << neighbors)
    out = np.zeros((b, L), np.uint16)
    lib().frb_oracle_lbp_hist_batch(_p(imgs), b, h, w, radius, neighbors, grid_x, grid_y, _p(out))
    cell_px = ((w - 2 * radius) // grid_x) * ((h - 2 * radius) // grid_y)
    return out, cell_px


def hist_to_f32(hist_u16: np.ndarray, cell_px: int) -> np.ndarray:
    """OpenCV's normalised view: float32(count) * float32(1.0/cell_px)."""
    return hist_u16.astype(np.float32) * np.float32(1.0 / cell_px)


def c_chisq_alt(h1: np.ndarray, h2: np.ndarray) -> float:
    h1 = np.ascontiguousarray(h1, np.float32).ravel()
    h2 = np.ascontiguousarray(h2, np.float32).ravel()
    return float(lib().frb_oracle_chisq_alt(_p(h1), _p(h2), h1.size))


def c_chisq_scan(gallery_f32: np.ndarray, query_f32: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(gallery_f32, np.float32)
    q = np.ascontiguousarray(query_f32, np.float32).ravel()
    out = np.zeros(g.shape[0], np.float64)
    lib().frb_oracle_chisq_scan(_p(g), g.shape[0], g.shape[1], _p(q), _p(out))
    return out


def c_chisq_scan_u16(gallery_u16: np.ndarray, gallery_cell_px: int, query_u16: np.ndarray,
                     query_cell_px: int) -> np.ndarray:
    g = np.ascontiguousarray(gallery_u16, np.uint16)
    q = np.ascontiguousarray(query_u16, np.uint16).ravel()
    out = np.zeros(g.shape[0], np.float64)
    lib().frb_oracle_chisq_scan_u16(_p(g), g.shape[0], g.shape[1], gallery_cell_px, _p(q), query_cell_px, _p(out))
    return out


def c_predict(gallery_f32: np.ndarray, labels: np.ndarray, query_f32: np.ndarray,
              threshold: float = np.finfo(np.float64).max):
    """LBPH::predict -> (label, distance, winner_index)."""
    g = np.ascontiguousarray(gallery_f32, np.float32)
    lab = np.ascontiguousarray(labels, np.int32)
    q = np.ascontiguousarray(query_f32, np.float32).ravel()
    ol = ctypes.c_int32(-1)
    od = ctypes.c_double(0.0)
    idx = lib().frb_oracle_predict(_p(g), _p(lab), g.shape[0], g.shape[1], _p(q), threshold,
                                   ctypes.byref(ol), ctypes.byref(od))
    return int(ol.value), float(od.value), int(idx)


class OracleLBPH:
    """cv2.face.LBPHFaceRecognizer protocol on the CPU oracle (train/update/predict)."""

    def __init__(self, radius=1, neighbors=8, grid_x=8, grid_y=8, threshold=np.finfo(np.float64).max):
        self.radius, self.neighbors, self.grid_x, self.grid_y = radius, neighbors, grid_x, grid_y
        self.threshold = threshold
        self.hists = np.zeros((0, grid_x * grid_y * (1 << neighbors)), np.float32)
        self.labels = np.zeros((0,), np.int32)

    def _hist(self, img):
        h, px = c_lbp_hist(np.asarray(img), self.radius, self.neighbors, self.grid_x, self.grid_y)
        return hist_to_f32(h, px)

    def train(self, faces, labels):
        self.hists = np.zeros((0, self.hists.shape[1]), np.float32)
        self.labels = np.zeros((0,), np.int32)
        self.update(faces, labels)

    def update(self, faces, labels):
        hs = [self._hist(f) for f in faces]
        if hs:
            self.hists = np.concatenate([self.hists] + hs, 0)
            self.labels = np.concatenate([self.labels, np.asarray(labels, np.int32).ravel()])

    def predict(self, img):
        label, dist, _ = c_predict(self.hists, self.labels, self._hist(img), self.threshold)
        return label, dist


# ------------------------------------------------------------------------- NumPy restatement
def np_taps(radius: int = 1, neighbors: int = 8):
    taps = []
    for n in range(neighbors):
        ang = 2.0 * np.pi * n / float(np.float32(neighbors))
        x = np.float32(radius * np.cos(ang))
        y = np.float32(-radius * np.sin(ang))
        fx, fy = int(np.floor(x)), int(np.floor(y))
        cx, cy = int(np.ceil(x)), int(np.ceil(y))
        ty = np.float32(y - np.float32(fy))
        tx = np.float32(x - np.float32(fx))
        one = np.float32(1)
        w1 = np.float32((one - tx) * (one - ty))
        w2 = np.float32(tx * (one - ty))
        w3 = np.float32((one - tx) * ty)
        w4 = np.float32(tx * ty)
        taps.append((fx, fy, cx, cy, w1, w2, w3, w4))
    return taps


def np_elbp(img: np.ndarray, radius: int = 1, neighbors: int = 8) -> np.ndarray:
    """Vectorised float32 elbp_: each product and each sum rounded to float32, left to right."""
    src = np.asarray(img, np.uint8).astype(np.float32)
    h, w = src.shape
    r = radius
    out = np.zeros((h - 2 * r, w - 2 * r), np.int32)
    c = src[r:h - r, r:w - r]

    def sh(dy, dx):
        return src[r + dy:h - r + dy, r + dx:w - r + dx]

    eps = np.float32(np.finfo(np.float32).eps)
    for n, (fx, fy, cx, cy, w1, w2, w3, w4) in enumerate(np_taps(radius, neighbors)):
        t = (w1 * sh(fy, fx)).astype(np.float32)
        t = (t + (w2 * sh(fy, cx)).astype(np.float32)).astype(np.float32)
        t = (t + (w3 * sh(cy, fx)).astype(np.float32)).astype(np.float32)
        t = (t + (w4 * sh(cy, cx)).astype(np.float32)).astype(np.float32)
        bit = (t > c) | (np.abs((t - c).astype(np.float32)) < eps)
        out += bit.astype(np.int32) << n
    return out


def np_spatial_hist(codes: np.ndarray, grid_x=8, grid_y=8, num_patterns=256):
    orows, ocols = codes.shape
    width, height = ocols // grid_x, orows // grid_y
    hist = np.zeros((grid_y * grid_x, num_patterns), np.uint16)
    for i in range(grid_y):
        for j in range(grid_x):
            cell = codes[i * height:(i + 1) * height, j * width:(j + 1) * width]
            hist[i * grid_x + j] = np.bincount(cell.ravel(), minlength=num_patterns)[:num_patterns]
    return hist.reshape(-1), width * height


def np_chisq_alt(h1: np.ndarray, h2: np.ndarray) -> float:
    a = h1.astype(np.float64) - h2.astype(np.float64)
    b = h1.astype(np.float64) + h2.astype(np.float64)
    m = np.abs(b) > np.finfo(np.float64).eps
    return float(2.0 * np.sum(a[m] * a[m] / b[m]))
