"""oracle/chisq_filter.py — CPU statement of the candidate-filter algorithm planned in front of K3.
TEST INFRASTRUCTURE ONLY; nothing in the product imports it.  It exists so that a future tensor-core filter kernel has
an oracle: the feature tables, the rigorous error bound and the survivor rule are all defined (and tested) here.

chi-square over integer counts:  d(g, q) = sum_j (g_j - q_j)^2 / (g_j + q_j) = sum g + sum q - 4 sum_j f(g_j, q_j),
f(a, b) = a b / (a + b) (0 at a + b = 0).  f on [0, cell_px]^2 is a table F; its rank-M truncated SVD gives per-count
feature vectors u(a), v(b) in R^M with F ~ u v^T, so sum_j f(g_j, q_j) ~ <U(g), V(q)> with U(g) = concat_j u(g_j):
an inner product of length hist_len * M — what a GEMM computes.  With features rounded to `dtype`:

    |approx(g, q) - d(g, q)| <= eps(q) = 4 * sum_j max_a |F~ - F|(a, q_j)          for EVERY gallery row g,

so every row whose approximate distance exceeds (approximate minimum + 2 eps) cannot be the nearest neighbour, and the
exact scan over the remaining rows returns exactly what the exact scan over all rows returns (ties: lowest row).
(A kernel accumulating in fp32 must add its accumulation error bound to eps; this module accumulates in float64.)
"""
from __future__ import annotations

import numpy as np


def f_table(cell_px: int) -> np.ndarray:
    a = np.arange(cell_px + 1, dtype=np.float64)
    s = a[:, None] + a[None, :]
    return np.where(s > 0, a[:, None] * a[None, :] / np.maximum(s, 1.0), 0.0)


def feature_tables(cell_px: int, rank: int = 8, dtype=np.float16):
    """(u [cell_px+1, rank], v [cell_px+1, rank], err [cell_px+1, cell_px+1]) with u, v rounded to `dtype` (returned as
    float64) and err = |u v^T - F| for exactly those rounded tables."""
    F = f_table(cell_px)
    U, S, Vt = np.linalg.svd(F)
    u = (U[:, :rank] * np.sqrt(S[:rank])).astype(dtype).astype(np.float64)
    v = (Vt[:rank].T * np.sqrt(S[:rank])).astype(dtype).astype(np.float64)
    return u, v, np.abs(u @ v.T - F)


def eps_bound(q_hist: np.ndarray, err: np.ndarray) -> float:
    """Rigorous bound on |approx - exact| for this query against ANY gallery row (count units)."""
    return float(4.0 * err.max(axis=0)[q_hist.astype(np.int64)].sum())


def approx_distances(gallery: np.ndarray, q_hist: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    """sum g + sum q - 4 <U(g), V(q)> for every gallery row (float64 accumulation)."""
    G = gallery.astype(np.int64)
    q = q_hist.astype(np.int64)
    vq = v[q]                                             # [L, M]
    dots = np.empty(G.shape[0])
    for lo in range(0, G.shape[0], 256):                  # bounded temporaries
        dots[lo:lo + 256] = np.einsum("nlm,lm->n", u[G[lo:lo + 256]], vq)
    return G.sum(1) + q.sum() - 4.0 * dots


def exact_distances(gallery: np.ndarray, q_hist: np.ndarray) -> np.ndarray:
    G = gallery.astype(np.int64)
    q = q_hist.astype(np.int64)
    s = G + q
    return np.where(s > 0, (G - q) ** 2 / np.maximum(s, 1), 0.0).sum(1)


def filtered_nearest(gallery: np.ndarray, q_hist: np.ndarray, u, v, err):
    """(row, exact distance, survivors): nearest neighbour through the filter; equals the unfiltered answer."""
    approx = approx_distances(gallery, q_hist, u, v)
    eps = eps_bound(q_hist, err)
    keep = np.flatnonzero(approx <= approx.min() + 2.0 * eps)
    d = exact_distances(gallery[keep], q_hist)
    j = int(np.argmin(d))                                 # first minimum: lowest row wins ties
    return int(keep[j]), float(d[j]), int(keep.size)
