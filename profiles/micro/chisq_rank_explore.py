"""Which (rank, weighting) gives the smallest provable bound eps(q) for fp16 feature tables?  CPU only.
eps(q) = 4 sum_j max_a |F~ - F|(a, q_j); weighted SVD: factor F diag(w), w(b) = (1+b)^-p, then v /= w."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lbph as OL

px = 169; side = 112
a = np.arange(px + 1, dtype=np.float64)
S_ = a[:, None] + a[None, :]
F = np.where(S_ > 0, a[:, None] * a[None, :] / np.maximum(S_, 1), 0.0)

def faces(n, seed, blocky=True):
    r = np.random.default_rng(seed)
    if not blocky:
        return r.integers(0, 256, (n, side, side)).astype(np.uint8)
    base = r.integers(0, 256, (n, side // 4 + 2, side // 4 + 2)).astype(np.float32)
    up = np.kron(base, np.ones((4, 4), np.float32))[:, :side, :side]
    return np.clip(up + r.normal(0, 12, (n, side, side)), 0, 255).astype(np.uint8)

for blocky in (True, False):
    G, _ = OL.c_lbp_hist(faces(600, 1, blocky)); Q, _ = OL.c_lbp_hist(faces(32, 2, blocky))
    G = G.astype(np.int64); Q = Q.astype(np.int64)
    gmax = G.max(0)
    exact = np.stack([np.where(G + q > 0, (G - q) ** 2 / np.maximum(G + q, 1), 0.0).sum(1) for q in Q])
    print(f"blocky={blocky}: count hist of query bins: {np.bincount(Q.ravel())[:12]} max {Q.max()}; dist min {exact.min(1).mean():.0f} "
          f"median {np.median(exact):.0f} std {exact.std(1).mean():.0f}")
    for M in (4, 5, 6, 8):
        for p in (0.0, 0.5, 1.0, 1.5):
            wa = (1 + a) ** -p
            U, S, Vt = np.linalg.svd(wa[:, None] * F * wa[None, :])
            u = (U[:, :M] * np.sqrt(S[:M])) / wa[:, None]
            v = (Vt[:M].T * np.sqrt(S[:M])) / wa[:, None]
            for rounded in ("exact", "fp16"):
                uu, vv = (u, v) if rounded == "exact" else (u.astype(np.float16).astype(np.float64), v.astype(np.float16).astype(np.float64))
                E = uu @ vv.T - F
                err = np.abs(E)
                eps = 4.0 * err.max(0)[Q].sum(1)
                # gallery-aware: max over a <= gmax_j only
                cm = np.maximum.accumulate(err, axis=0)      # cm[a, b] = max_{a' <= a} err(a', b)
                eps_g = 4.0 * cm[gmax[None, :], Q].sum(1)
                # signed window
                hi = 4.0 * np.maximum.accumulate(E, axis=0)[gmax[None, :], Q].sum(1)
                lo = 4.0 * np.minimum.accumulate(E, axis=0)[gmax[None, :], Q].sum(1)
                fg = uu[G]
                approx = np.stack([G.sum(1) + q.sum() - 4.0 * np.einsum("nlm,lm->n", fg, vv[q]) for q in Q])
                derr = np.abs(approx - exact).max()
                print(f"  M={M} p={p} {rounded:5s}: eps {eps.mean():7.1f}  gallery-aware {eps_g.mean():7.1f}  signed window/2 {((hi - lo) / 2).mean():7.1f}  actual max err {derr:.2f}")
