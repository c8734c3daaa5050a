"""oracle/lbph_oracle.c + oracle/lbph.py: two independent restatements of OpenCV-contrib's LBPH agree with each
other and with the committed fixtures; the chi-square stage is pinned on the REAL cv2.compareHist.
(The LBP-code/histogram stage is PARITY UNPINNED: no cv2.face exists here to run.)"""
import cv2
import numpy as np
import pytest

SIZES = ["s100", "s112", "s57x83"]


def test_tap_constants(oracle_lbph):
    O = oracle_lbph
    taps = O.c_taps(1, 8)
    hexw = lambda v: int(np.float32(v).view(np.uint32))
    A, B, C = 0x3E5413CD, 0x3EFFFFFF, 0x3DAFB0CE       # the constants hard-coded in csrc/lbp_hist.cu
    assert [hexw(taps[1][w]) for w in ("w1", "w2", "w3", "w4")] == [A, B, C, A]
    assert [hexw(taps[3][w]) for w in ("w1", "w2", "w3", "w4")] == [B, A, A, C]
    assert [hexw(taps[5][w]) for w in ("w1", "w2", "w3", "w4")] == [A, C, B, A]
    assert [hexw(taps[7][w]) for w in ("w1", "w2", "w3", "w4")] == [C, A, A, B]
    assert [(int(t["fy"]), int(t["fx"])) for t in taps] == [(0, 1), (-1, 0), (-1, 0), (-1, -1), (-1, -1), (0, -1), (1, -1), (0, 0)]
    for n, (fx, fy, cx, cy, w1, w2, w3, w4) in enumerate(O.np_taps(1, 8)):
        t = taps[n]
        assert (fx, fy, cx, cy) == (t["fx"], t["fy"], t["cx"], t["cy"])
        assert [hexw(x) for x in (w1, w2, w3, w4)] == [hexw(t[w]) for w in ("w1", "w2", "w3", "w4")]


def test_compare_identities_used_by_the_kernel():
    """(t > c) || |t - c| < FLT_EPSILON  <=>  t >= thr(c), thr(1) = 1 - 2^-24, thr(c) = c otherwise;
    and the axis bits reduce to neighbour >= centre.  Both are relied on by csrc/lbp_hist.cu."""
    eps = np.float32(np.finfo(np.float32).eps)
    for e in range(256):
        c = np.float32(e)
        base = int(c.view(np.uint32))
        near = np.array([base + d for d in range(-64, 65) if base + d >= 0], np.uint32).view(np.float32)
        t = np.concatenate([near, np.random.default_rng(e).uniform(0, 256, 4000).astype(np.float32)])
        ref = (t > c) | (np.abs((t - c).astype(np.float32)) < eps)
        thr = np.float32(1 - 2.0 ** -24) if e == 1 else c
        assert np.array_equal(ref, t >= thr)
    # the packed kernel forms the threshold as fl(c - 2^-24) and takes bits from the SIGN of a subtraction:
    # fl(c - 2^-24) is thr(c) for c >= 1 and negative for c = 0 (t >= 0 always), sign(t - thr) set <=> t < thr,
    # and a product by +0-added FMA equals the rounded product for the non-negative operands involved
    d24 = np.float32(2.0 ** -24)
    for e in range(256):
        c = np.float32(e)
        thr = (c - d24).astype(np.float32)
        assert thr == (np.float32(1 - 2.0 ** -24) if e == 1 else c) if e >= 1 else thr < 0
        nthr = (d24 - c).astype(np.float32)                       # what the kernel holds: -(thr)
        assert nthr == -thr
        base = int(c.view(np.uint32))
        near = np.array([base + d for d in range(-64, 65) if base + d >= 0], np.uint32).view(np.float32)
        r = (near + nthr).astype(np.float32)
        assert np.array_equal(np.signbit(r), near < thr)
    w2 = np.uint32(0x248D3132).view(np.float32)
    b, c, e = np.meshgrid(*(np.arange(256, dtype=np.float32),) * 3, indexing="ij")
    t = (b + (w2 * c).astype(np.float32)).astype(np.float32)
    assert np.array_equal((t > e) | (np.abs((t - e).astype(np.float32)) < eps), b >= e)


@pytest.mark.parametrize("tag", SIZES)
def test_c_and_numpy_restatements_agree_with_fixture(oracle_lbph, lbph_golden, tag):
    O, g = oracle_lbph, lbph_golden
    faces = g[f"{tag}_faces"]
    hist, px = O.c_lbp_hist(faces)
    assert px == int(g[f"{tag}_cell_px"])
    np.testing.assert_array_equal(hist, g[f"{tag}_hist"])
    np.testing.assert_array_equal(O.c_elbp(faces[0]), g[f"{tag}_codes0"])
    for f in faces:
        codes = O.np_elbp(f)
        np.testing.assert_array_equal(codes, O.c_elbp(f))
        h, p = O.np_spatial_hist(codes)
        assert p == px
        np.testing.assert_array_equal(h, O.c_lbp_hist(f)[0][0])
    # every cell holds exactly cell_px pixels
    assert np.all(hist.reshape(len(faces), 64, 256).sum(-1) == px)


def test_flat_images_give_gray_level_dependent_codes(oracle_lbph, lbph_golden):
    O, g = oracle_lbph, lbph_golden
    codes = np.array([O.c_elbp(np.full((5, 5), v, np.uint8))[1, 1] for v in range(256)])
    np.testing.assert_array_equal(codes, g["flat_codes"])
    assert {0: 255, 2: 223, 3: 247, 7: 221, 127: 221, 128: 223, 254: 221, 255: 255}.items() <= dict(enumerate(codes.tolist())).items()


@pytest.mark.parametrize("tag", SIZES)
def test_chisq_is_pinned_on_cv2_comparehist(oracle_lbph, lbph_golden, tag):
    O, g = oracle_lbph, lbph_golden
    hf = O.hist_to_f32(g[f"{tag}_hist"], int(g[f"{tag}_cell_px"]))
    n = hf.shape[0]
    live = np.array([[cv2.compareHist(hf[i], hf[j], cv2.HISTCMP_CHISQR_ALT) for j in range(n)] for i in range(n)])
    np.testing.assert_array_equal(live, g[f"{tag}_cv2_chisq"])          # fixture == this box's OpenCV
    mine = np.stack([O.c_chisq_scan(hf, hf[i]) for i in range(n)])
    np.testing.assert_allclose(mine, live, rtol=1e-12, atol=0)
    assert np.all(np.diag(mine) == 0.0)                                  # self-match is exactly 0
    u16 = np.stack([O.c_chisq_scan_u16(g[f"{tag}_hist"], int(g[f"{tag}_cell_px"]), g[f"{tag}_hist"][i],
                                       int(g[f"{tag}_cell_px"])) for i in range(n)])
    np.testing.assert_allclose(u16, live, rtol=1e-12)
    assert abs(O.np_chisq_alt(hf[0], hf[1]) - live[0, 1]) <= 1e-12 * live[0, 1]


def test_predict_semantics(oracle_lbph, lbph_golden):
    O, g = oracle_lbph, lbph_golden
    faces = g["s100_faces"]
    labels = np.arange(len(faces), dtype=np.int32) + 100
    m = O.OracleLBPH()
    m.train(list(faces), labels)
    for i, f in enumerate(faces):
        assert m.predict(f) == (100 + i, 0.0)
    # duplicate gallery rows: strict '<' keeps the FIRST one
    m2 = O.OracleLBPH()
    m2.train([faces[3], faces[5], faces[3]], np.array([7, 8, 9], np.int32))
    assert m2.predict(faces[3]) == (7, 0.0)
    # model threshold rejects everything -> (-1, DBL_MAX)
    m2.threshold = 0.0
    lab, d = m2.predict(faces[3])
    assert lab == -1 and d == np.finfo(np.float64).max
    # update() appends
    m2.threshold = np.finfo(np.float64).max
    m2.update([faces[6]], np.array([11], np.int32))
    assert m2.predict(faces[6]) == (11, 0.0) and m2.hists.shape[0] == 4


def test_resize_restatement_is_bit_exact_with_the_installed_cv2():
    """oracle/resize.py vs the REAL cv2.resize (default INTER_LINEAR) and cv2.cvtColor of the installed OpenCV core:
    up- and down-scaling, 1 and 3 channels, the exact-halving case (cv2 reroutes it to the INTER_AREA average),
    1-pixel sides, and the reference's two target sizes from camera-like frames.  This pins the oracle."""
    import cv2
    from oracle import resize as OR
    rng = np.random.default_rng(99)
    cases = [(480, 640, 100, 100), (250, 250, 112, 112), (50, 40, 100, 100), (200, 200, 100, 100), (224, 224, 112, 112),
             (100, 100, 100, 100), (37, 53, 112, 112), (720, 1280, 112, 112), (101, 99, 100, 100), (1, 1, 7, 5), (9, 1, 4, 6),
             (1, 13, 3, 40), (2, 2, 1, 1), (300, 200, 100, 100), (64, 48, 32, 24)]
    cases += [tuple(int(v) for v in rng.integers(1, 320, 4)) for _ in range(60)]
    for sh, sw, dh, dw in cases:
        for ch in (1, 3):
            img = rng.integers(0, 256, (sh, sw, ch) if ch == 3 else (sh, sw), dtype=np.uint8)
            ref = cv2.resize(img, (dw, dh))
            got = OR.resize_linear_u8(img, dw, dh)
            assert np.array_equal(got.reshape(ref.shape), ref), (sh, sw, dh, dw, ch)
            if ch == 3:
                want = cv2.cvtColor(ref.reshape(dh, dw, 3), cv2.COLOR_BGR2GRAY)
                assert np.array_equal(OR.preprocess_for_lbph(img, (dw, dh)), want), (sh, sw, dh, dw)
    # smooth images (long runs of equal weights) and saturated ones
    ramp = np.add.outer(np.arange(240), np.arange(320)).astype(np.uint8)
    for dsize in [(100, 100), (112, 112), (333, 17)]:
        assert np.array_equal(OR.resize_linear_u8(ramp, *dsize), cv2.resize(ramp, dsize))
        full = np.full((77, 91, 3), 255, np.uint8)
        assert np.array_equal(OR.resize_linear_u8(full, *dsize), cv2.resize(full, dsize))


def _library_filter_tables(px):
    """frb_chisq_filter_tables: a HOST function of libfrb200 (no GPU needed) -> (u, v fp16 tables as float64, emax, absmax)."""
    import ctypes
    from facerecognition_b200 import _native as N
    u = np.zeros((256, 8), np.uint16)
    v = np.zeros((256, 8), np.uint16)
    em = np.zeros(256, np.float32)
    am = np.zeros(256, np.float32)
    N.call("frb_chisq_filter_tables", px, *(x.ctypes.data_as(ctypes.c_void_p) for x in (u, v, em, am)))
    return u.view(np.float16).astype(np.float64), v.view(np.float16).astype(np.float64), em, am


def test_chisq_filter_statement_is_complete(oracle_lbph):
    """oracle/chisq_filter.py restates the tensor-core candidate filter of csrc/chisq_filter.cu on the CPU.  Checked here
    without a GPU: (1) the LIBRARY's feature tables (host code: Jacobi eigen-decomposition) — its error bound emax
    dominates the true error of its own fp16 tables against f(a, b) = ab/(a+b), computed in float64, for 100x100 and
    112x112 faces; empty bins are exact; (2) the oracle's independent tables (numpy eigh) fit f as well; (3) with either
    pair of tables the bound holds for every (query, row) pair and the filtered nearest neighbour equals the exhaustive
    one — planted queries, unplanted ones, exact duplicates (lowest row wins) and rows of unequal mass."""
    from oracle import chisq_filter as CF
    for px in (144, 169):
        F = CF.f_table(px)
        lu, lv, em, am = _library_filter_tables(px)
        E_lib = lu[:px + 1] @ lv[:px + 1].T - F
        assert np.all(em[:px + 1] >= np.abs(E_lib).max(0)), "emax must dominate the table error it is used to bound"
        assert np.all(am[:px + 1] >= (np.abs(lu[:px + 1])[:, None, :] * np.abs(lv[:px + 1])[None, :, :]).sum(2).max(0) - 1e-6)
        assert np.abs(E_lib[0]).max() == 0 and np.abs(E_lib[:, 0]).max() == 0 and not lu[px + 1:].any() and not lv[px + 1:].any()
        u, v, E = CF.feature_tables(px)
        assert np.abs(E).max() < 0.1 and np.abs(E_lib).max() < 0.1
        # same construction, two eigen-solvers: the per-count bounds agree to fp16 rounding noise
        assert abs(np.abs(E).max(0).sum() - np.abs(E_lib).max(0).sum()) <= 0.25 * np.abs(E).max(0).sum()
    rng = np.random.default_rng(12)
    faces = rng.integers(0, 256, (260, 100, 100), dtype=np.uint8)
    faces[:, 20:60] //= 3                                    # some structure: darker band
    hist, px = oracle_lbph.c_lbp_hist(faces)
    gallery, fresh = hist[:200].copy(), hist[200:]
    gallery[150] = gallery[7]                                # duplicate row
    gallery[60, 4096:] = 0                                   # rows of unequal mass (not LBPH histograms any more)
    gallery[61, :8192] //= 2
    queries = [gallery[7], gallery[33], gallery[60]] + list(fresh[:10])
    lu, lv, em, _ = _library_filter_tables(px)
    for name, (u, v, E) in (("oracle", CF.feature_tables(px)), ("library", (lu[:px + 1], lv[:px + 1], lu[:px + 1] @ lv[:px + 1].T - CF.f_table(px)))):
        survivors = []
        for q in queries:
            exact = CF.exact_distances(gallery, q)
            score = CF.approx_scores(gallery, q, u, v)
            exact_score = (q.astype(np.int64).sum() - exact) / 4.0                 # S - (sum g) / 4
            assert np.abs(score - exact_score).max() <= CF.eps_bound(q, E) + 1e-6, name
            row, d, kept = CF.filtered_nearest(gallery, q, u, v, E)
            assert row == int(np.argmin(exact)) and d == exact.min(), name
            survivors.append(kept)
        assert survivors[0] <= 3 and survivors[1] <= 3 and survivors[2] <= 3    # planted: the duplicate pair / the row itself survive
    # the count-unit distance is the kernels' distance up to the 2 / cell_px scale
    ref = oracle_lbph.c_chisq_scan_u16(gallery, px, queries[3], px)
    np.testing.assert_allclose(CF.exact_distances(gallery, queries[3]) * (2.0 / px), ref, rtol=1e-6)   # the C oracle works on OpenCV's float32 view


def test_third_party_pins():
    """tests/golden/verify_with_contrib.py under pytest: pins the LBP histogram stage on the real cv2.face and the FAISS
    file layout / search order on the real faiss wherever those modules are installed; skipped (stages stay "parity
    unpinned") where they are not — as in the authoring image and on the GPU box."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("verify_with_contrib", os.path.join(os.path.dirname(__file__), "golden", "verify_with_contrib.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ran = []
    if mod.have_cv2_face():
        ran += mod.check_lbph_against_contrib()
    if mod.have_faiss():
        ran += mod.check_faiss_against_real()
    if not ran:
        pytest.skip("neither cv2.face (opencv-contrib-python) nor faiss is installed")
    assert all(ok for _, ok, _ in ran), [r for r in ran if not r[1]]
