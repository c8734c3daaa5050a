"""oracle/chisq_filter.py — CPU statement of the candidate filter in front of the chi-square scan
(facerecognition_b200/csrc/chisq_filter.cu).  TEST INFRASTRUCTURE ONLY; nothing in the product imports it.

chi-square over integer counts of equal cell size:
    d(g, q) = sum_j (g_j - q_j)^2 / (g_j + q_j) = sum g + sum q - 4 S(g, q),   S = sum_j f(g_j, q_j),   f(a, b) = a b / (a + b)
(f = 0 at a + b = 0), so the nearest row maximises  score(g) = S(g, q) - (sum g) / 4.
f on [0, cell_px]^2 is a symmetric table F.  The product factors W F W, w(a) = (1 + a)^-1.5, keeps the 8 eigen-pairs of
largest |eigenvalue| and unscales:  u_m(a) = sqrt|lam_m| e_m(a) / w(a),  v_m(b) = sign(lam_m) u_m(b), both rounded to
fp16, so  F~(a, b) = <u(a), v(b)>  and  S~(g, q) = sum_j F~(g_j, q_j)  is an inner product of length 8 * hist_len — what
the tensor cores compute.  With E = F~ - F (float64, exact for the rounded tables):

    |S~(g, q) - S(g, q)| <= e(q) = sum_j max_a |E(a, q_j)|          for EVERY gallery row g,

so a row can only be the exact scan's answer if  score~(g) >= max_rows score~ - 2 e(q)  (the product adds an allowance for
the fp32 accumulation inside the tensor core and for the exact scan's own rounding to e).  The exact scan over those
survivors returns exactly what the exact scan over all rows returns (ties: lowest row).

This module restates the tables independently (numpy eigh) and states the survivor rule; tests/test_oracle_lbph.py
checks the LIBRARY's tables (frb_chisq_filter_tables, a host function) against it: the library's error bounds must
dominate the true error of the library's own fp16 tables, and the rule must keep the true nearest neighbour.
"""
from __future__ import annotations

import numpy as np

RANK = 8
WEIGHT_POWER = 1.5


def f_table(cell_px: int) -> np.ndarray:
    a = np.arange(cell_px + 1, dtype=np.float64)
    s = a[:, None] + a[None, :]
    return np.where(s > 0, a[:, None] * a[None, :] / np.maximum(s, 1.0), 0.0)


def feature_tables(cell_px: int, rank: int = RANK, dtype=np.float16):
    """(u [cell_px+1, rank], v [cell_px+1, rank], E [cell_px+1, cell_px+1]) with u, v rounded to `dtype` (returned as
    float64) and E = u v^T - F for exactly those rounded tables."""
    F = f_table(cell_px)
    w = (1.0 + np.arange(cell_px + 1, dtype=np.float64)) ** -WEIGHT_POWER
    lam, vec = np.linalg.eigh(w[:, None] * F * w[None, :])
    order = np.argsort(-np.abs(lam))[:rank]
    u = np.sqrt(np.abs(lam[order]))[None, :] * vec[:, order] / w[:, None]
    v = u * np.sign(lam[order])[None, :]
    u[0] = 0.0
    v[0] = 0.0
    u = u.astype(dtype).astype(np.float64)
    v = v.astype(dtype).astype(np.float64)
    return u, v, u @ v.T - F


def eps_bound(q_hist: np.ndarray, E: np.ndarray) -> float:
    """Rigorous bound on |S~ - S| for this query against ANY gallery row (S units; the distance moves by 4x as much)."""
    return float(np.abs(E).max(axis=0)[q_hist.astype(np.int64)].sum())


def approx_scores(gallery: np.ndarray, q_hist: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """score~ = <U(g), V(q)> - (sum g) / 4 for every gallery row (float64 accumulation)."""
    G = gallery.astype(np.int64)
    M = u @ v.T                                           # [a, b]
    S = M[G, q_hist.astype(np.int64)[None, :]].sum(1)
    return S - G.sum(1) / 4.0


def exact_distances(gallery: np.ndarray, q_hist: np.ndarray) -> np.ndarray:
    """sum_j (g_j - q_j)^2 / (g_j + q_j) in count units (OpenCV's CHISQR_ALT distance is 2 / cell_px times this)."""
    G = gallery.astype(np.int64)
    q = q_hist.astype(np.int64)
    s = G + q
    return np.where(s > 0, (G - q) ** 2 / np.maximum(s, 1), 0.0).sum(1)


def filtered_nearest(gallery: np.ndarray, q_hist: np.ndarray, u, v, E, extra: float = 0.0):
    """(row, exact distance, survivors): nearest neighbour through the filter; equals the unfiltered answer."""
    score = approx_scores(gallery, q_hist, u, v)
    e = eps_bound(q_hist, E) + extra
    keep = np.flatnonzero(score >= score.max() - 2.0 * e)
    d = exact_distances(gallery[keep], q_hist)
    j = int(np.argmin(d))                                 # first minimum: lowest row wins ties
    return int(keep[j]), float(d[j]), int(keep.size)
