"""Seeded sweeps over ragged shapes: the kernels pick different code paths by shape (TMA-pipelined vs plain LBP
kernel, full vs partial chi-square coverage, row-streaming vs tiled vs tensor-core cosine), so every sweep checks
the same contract against the oracle on shapes nobody chose by hand."""
import numpy as np
import pytest
import torch

from oracle import cosine as OC

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_lbp_hist_random_shapes_and_grids(oracle_lbph):
    """LBP histograms bit-exact for 40 random (count, rows, cols, grid) — multiples of 16 bytes (pipelined kernel)
    and not (plain kernel), odd widths (byte loads), odd grids (a band without a partner), tiny images."""
    from facerecognition_b200 import ops
    rng = np.random.default_rng(20261018)
    for case in range(40):
        rows, cols = int(rng.integers(3, 140)), int(rng.integers(3, 140))
        if case % 3 == 0:                       # force a TMA-eligible size
            cols = max(4, cols // 4 * 4)
            rows = max(4, rows // 4 * 4)
        gx, gy = int(rng.integers(1, 10)), int(rng.integers(1, 10))
        n = int(rng.integers(1, 9)) if case % 5 else int(rng.integers(300, 700))   # some batches larger than the grid
        kind = case % 3
        if kind == 0:
            faces = rng.integers(0, 256, (n, rows, cols), dtype=np.uint8)
        elif kind == 1:
            faces = rng.integers(0, 4, (n, rows, cols), dtype=np.uint8)             # ties and the centre == 1 threshold
        else:
            faces = np.repeat(np.repeat(rng.integers(0, 256, (n, (rows + 7) // 8, (cols + 7) // 8), dtype=np.uint8), 8, 1), 8, 2)[:, :rows, :cols]
        faces = np.ascontiguousarray(faces)
        want, wpx = oracle_lbph.c_lbp_hist(faces, 1, 8, gx, gy)
        got, px = ops.lbp_hist(dev(faces), 1, 8, gx, gy)
        assert px == wpx, (rows, cols, gx, gy)
        np.testing.assert_array_equal(got.cpu().numpy(), want, err_msg=f"case {case}: {n}x{rows}x{cols} grid {gx}x{gy}")
        m = min(n, 3)
        np.testing.assert_array_equal(ops.lbp_codes(dev(faces[:m])).cpu().numpy(), np.stack([oracle_lbph.c_elbp(f) for f in faces[:m]]))


def test_chisq_random_shapes(oracle_lbph):
    from facerecognition_b200 import ops
    rng = np.random.default_rng(7)
    for case in range(12):
        L = int(rng.integers(1, 2049)) * 8
        N, Q, k = int(rng.integers(1, 300)), int(rng.integers(1, 9)), int(rng.integers(1, 6))
        px = int(rng.integers(1, 400))
        gal = rng.integers(0, px + 1, (N, L)).astype(np.uint16)
        gal[rng.random((N, L)) < 0.5] = 0
        q = gal[rng.integers(0, N, Q)].copy()
        q[:, : L // 2] = rng.integers(0, px + 1, (Q, L // 2)).astype(np.uint16)
        q_px = px if case % 2 == 0 else px + 1 + int(rng.integers(0, 50))
        ref = np.stack([oracle_lbph.c_chisq_scan_u16(gal, px, qq, q_px) for qq in q])
        d = ops.chisq_dist(dev(q), q_px, dev(gal), px).cpu().numpy().astype(np.float64)
        np.testing.assert_allclose(d, ref, rtol=1e-5, atol=0, err_msg=f"case {case}: L={L} N={N} Q={Q}")
        dist, idx = ops.chisq_topk(dev(q), q_px, dev(gal), px, k=k)
        kk = min(k, N)
        order = np.argsort(ref, axis=1, kind="stable")[:, :kk]
        np.testing.assert_allclose(dist.cpu().numpy()[:, :kk], np.take_along_axis(ref, order, 1), rtol=1e-5)
        assert bool((idx[:, kk:] == -1).all())


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_cosine_random_shapes(dtype):
    """Q across the kernel-selection thresholds (1, 2, 3, 16, 17, 64, 65, 129...), N across tile boundaries, any k."""
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(11 if dtype == "f32" else 12)
    tol = 1e-5 if dtype == "f32" else 1e-3
    for Q in [1, 2, 3, 4, 5, 16, 17, 63, 64, 65, 127, 129, 257]:
        N, k = int(rng.integers(1, 3000)), int(rng.integers(1, 9))
        gal = rng.standard_normal((N, 512)).astype(np.float32)
        gal /= np.linalg.norm(gal, axis=1, keepdims=True)
        q = gal[rng.integers(0, N, Q)] + 0.05 * rng.standard_normal((Q, 512)).astype(np.float32)
        g = dev(gal) if dtype == "f32" else ops.normalize_rows(dev(gal), NV.FRB_QNORM_NONE, torch.bfloat16)
        s, i = ops.cosine_topk(dev(q), g, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
        s, i = s.cpu().numpy(), i.cpu().numpy()
        ref = OC.l2_normalize(q).astype(np.float64) @ gal.T.astype(np.float64)
        kk = min(k, N)
        want = -np.sort(-ref, axis=1)[:, :kk]
        np.testing.assert_allclose(s[:, :kk], want, atol=tol, rtol=0, err_msg=f"Q={Q} N={N} k={k}")
        picked = np.take_along_axis(ref, np.clip(i[:, :kk], 0, N - 1), 1)
        assert np.all(np.abs(picked - want) <= 2 * tol) and np.all(i[:, kk:] == -1) and np.all(i[:, :kk] >= 0)


def test_cosine_other_embedding_widths():
    """The reference's embeddings are 512-d, but nothing in the ABI fixes that: fp32 takes any multiple of 8, the
    tensor-core path multiples of 64 up to 512 (anything else is refused, not silently mis-computed)."""
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(5)
    for dtype, dims in (("f32", [8, 24, 200, 256, 384, 1024]), ("bf16", [64, 128, 256, 384])):
        for D in dims:
            for Q in (1, 3, 40):
                N, k = 777, 4
                gal = rng.standard_normal((N, D)).astype(np.float32)
                gal /= np.linalg.norm(gal, axis=1, keepdims=True)
                q = gal[rng.integers(0, N, Q)] + 0.05 * rng.standard_normal((Q, D)).astype(np.float32)
                g = dev(gal) if dtype == "f32" else ops.normalize_rows(dev(gal), NV.FRB_QNORM_NONE, torch.bfloat16)
                s, i = ops.cosine_topk(dev(q), g, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
                ref = OC.l2_normalize(q).astype(np.float64) @ gal.T.astype(np.float64)
                want = -np.sort(-ref, axis=1)[:, :k]
                # bf16: 1e-3 is the bar at the reference's 512-d; with fewer, larger components the rounding of a
                # unit vector's entries averages out less (measured 1.2e-3 at 64-d)
                tol = 1e-5 if dtype == "f32" else (1e-3 if D >= 256 else 3e-3)
                np.testing.assert_allclose(s.cpu().numpy(), want, atol=tol, rtol=0, err_msg=f"{dtype} D={D} Q={Q}")
    with pytest.raises(NV.FrbError):                       # 96 is not a multiple of 64: refused by the bf16 path
        g = torch.zeros((10, 96), dtype=torch.bfloat16, device="cuda")
        ops.cosine_topk(torch.zeros((40, 96), device="cuda"), g, 1)
