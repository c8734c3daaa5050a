"""Groundwork for a tensor-core candidate filter in front of K3: how well is the chi-square
term separable?

    d(g, q) = sum_j (g_j - q_j)^2 / (g_j + q_j) = sum_j g_j + sum_j q_j - 4 sum_j f(g_j, q_j),   f(a, b) = a b / (a + b)

Counts are integers in [0, cell_px], so f is a (cell_px + 1)^2 table F.  A rank-M factorisation F ~ sum_m u_m(a) v_m(b)
turns sum_j f(g_j, q_j) into M inner products of per-count feature vectors, i.e. GEMM work.  This script reports, on
the CPU (numpy only):
  1. the singular values of F and the worst / rms entry error of its rank-M truncations (exact features);
  2. the same with features rounded to bf16 (what tcgen05 kind::f16 would multiply);
  3. on LBP histograms of synthetic faces (oracle/lbph.py): the error of the approximate distance relative to the
     spread of true distances, and how many candidates a filter must keep so that the true nearest neighbour is
     always among them.
Run: python profiles/micro/chisq_lowrank.py [cell_px] [n_gallery] [n_query]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def main():
    px = int(sys.argv[1]) if len(sys.argv) > 1 else 169
    n_gal = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    n_q = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    a = np.arange(px + 1, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        F = np.where((a[:, None] + a[None, :]) > 0, a[:, None] * a[None, :] / (a[:, None] + a[None, :]), 0.0)
    U, S, Vt = np.linalg.svd(F)
    print(f"f(a,b) = ab/(a+b) on [0,{px}]^2: singular values {S[:12].round(3)}")

    from oracle import lbph as OL
    rng = np.random.default_rng(0)
    side = 112 if px == 169 else 100
    def faces(n, seed):
        r = np.random.default_rng(seed)
        base = r.integers(0, 256, (n, side // 4 + 2, side // 4 + 2)).astype(np.float32)
        up = np.kron(base, np.ones((4, 4), np.float32))[:, :side, :side]            # blocky structure
        return np.clip(up + r.normal(0, 12, (n, side, side)), 0, 255).astype(np.uint8)
    gal_faces = faces(n_gal, 1)
    q_faces = gal_faces[rng.integers(0, n_gal, n_q)].astype(np.int16) + rng.integers(-6, 7, (n_q, side, side))   # noisy re-shots
    q_faces = np.clip(q_faces, 0, 255).astype(np.uint8)
    if os.environ.get("FRESH_QUERIES"):                      # queries with no planted match: small nearest / second gaps
        q_faces = faces(n_q, 2)
    G, gpx = OL.c_lbp_hist(gal_faces)
    Q, qpx = OL.c_lbp_hist(q_faces)
    assert gpx == qpx == px, (gpx, qpx, px)
    G = G.astype(np.int64); Q = Q.astype(np.int64)
    # exact distances in count units: sum (g-q)^2/(g+q)
    exact = np.empty((n_q, n_gal))
    for i in range(n_q):
        s = G + Q[i]
        d = (G - Q[i]) ** 2
        exact[i] = np.where(s > 0, d / np.maximum(s, 1), 0.0).sum(1)
    nn = exact.argmin(1)
    const = G.sum(1)[None, :] + Q.sum(1)[:, None]
    print(f"{n_q} queries x {n_gal} gallery faces ({side}x{side}, cell_px {px}); exact distance: nearest {exact.min(1).mean():.0f}, "
          f"median row {np.median(exact):.0f} (count units)")
    for M in (4, 6, 8, 12):
        for rounded in ("exact", "bf16", "fp16", "fp16 hi+lo"):
            u = U[:, :M] * np.sqrt(S[:M])
            v = Vt[:M].T * np.sqrt(S[:M])
            if rounded == "bf16":
                u, v = bf16(u).astype(np.float64), bf16(v).astype(np.float64)
            elif rounded == "fp16":
                u, v = u.astype(np.float16).astype(np.float64), v.astype(np.float16).astype(np.float64)
            elif rounded == "fp16 hi+lo":          # two fp16 planes per feature: 3 products (hi hi, hi lo, lo hi) per pair
                uh, vh = u.astype(np.float16).astype(np.float64), v.astype(np.float16).astype(np.float64)
                ul, vl = (u - uh).astype(np.float16).astype(np.float64), (v - vh).astype(np.float16).astype(np.float64)
                u, v = np.concatenate([uh, uh, ul], 1), np.concatenate([vh, vl, vh], 1)
            Fm = u @ v.T
            err = np.abs(Fm - F)
            # approximate distances through the features: sum_j f(g_j,q_j) ~ sum_m <u_m(g), v_m(q)>
            fg = u[G]                     # [n_gal, L, M]
            approx = np.empty((n_q, n_gal))
            for i in range(n_q):
                fq = v[Q[i]]              # [L, M]
                approx[i] = const[i] - 4.0 * np.einsum("nlm,lm->n", fg, fq)
            derr = np.abs(approx - exact)
            # smallest candidate list that always contains the true nearest neighbour
            rank_of_nn = np.array([(approx[i] < approx[i, nn[i]]).sum() for i in range(n_q)])
            # a RIGOROUS per-query bound on |approx - exact| over any gallery row: 4 * sum_j max_a |Fm - F|(a, q_j);
            # every row whose approximate distance is within 2 eps of the approximate minimum must be re-scored exactly
            eps = 4.0 * err.max(0)[Q].sum(1)
            keep = np.array([(approx[i] <= approx[i].min() + 2 * eps[i]).sum() for i in range(n_q)])
            assert all(approx[i, nn[i]] <= approx[i].min() + 2 * eps[i] for i in range(n_q))
            print(f"  rank {M:2d} {rounded + ' features':20s}: table error max {err.max():.3g} rms {np.sqrt((err**2).mean()):.3g}; "
                  f"distance error max {derr.max():.3g} ({100 * derr.max() / np.median(exact):.3f} % of the median distance); "
                  f"true NN found within the best {rank_of_nn.max() + 1} approximate candidates; provable bound eps {eps.mean():.0f}: "
                  f"rows to re-score exactly mean {keep.mean():.1f} max {keep.max()} of {n_gal}")


if __name__ == "__main__":
    main()
