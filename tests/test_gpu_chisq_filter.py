"""The tensor-core candidate filter in front of the chi-square scan (frb_chisq_top1_filtered_g8) must return, for every
query, exactly what the exact scan returns: same fp32 distance bits, same row, first row wins ties — the reference's
predict() answer (web_app.py:587, models/lbphmodel/evaluate_lbph.py:31-33) is the exact scan's.  Also checked: the GEMM
itself against float64 arithmetic on the same fp16 feature tables (the accumulation allowance lives on that margin)."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def faces_gpu(n, side, seed, blocky=True):
    """Synthetic gray faces on the device: blocky structure + noise (more realistic count spread) or pure noise."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    if not blocky:
        return torch.randint(0, 256, (n, side, side), generator=g, device="cuda", dtype=torch.uint8)
    base = torch.randint(0, 256, (n, side // 4 + 2, side // 4 + 2), generator=g, device="cuda").float()
    up = base.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :side, :side]
    return (up + 12.0 * torch.randn((n, side, side), generator=g, device="cuda")).clamp(0, 255).to(torch.uint8)


def hists(imgs, **kw):
    from facerecognition_b200 import ops
    h, px = ops.lbp_hist(imgs.contiguous(), **kw)
    return h, px


def exact_top1(qh, g8, px):
    from facerecognition_b200 import ops
    saved = ops.FILTER_ENABLED
    ops.FILTER_ENABLED = False
    try:
        return ops.chisq_topk(qh, px, g8, px, 1)
    finally:
        ops.FILTER_ENABLED = saved


def tables(px):
    from facerecognition_b200 import _native as N
    u = np.zeros((256, 8), np.uint16)
    v = np.zeros((256, 8), np.uint16)
    em = np.zeros(256, np.float32)
    am = np.zeros(256, np.float32)
    N.call("frb_chisq_filter_tables", px, *(x.ctypes.data_as(ctypes.c_void_p) for x in (u, v, em, am)))
    return u.view(np.float16).astype(np.float64), v.view(np.float16).astype(np.float64), em, am


def test_filter_gemm_against_float64_on_the_same_features():
    """approx_scores (TMEM accumulators) vs the float64 inner product of the same fp16 features; two query tiles (one
    ragged) share each generated gallery stage, three gallery tiles (one ragged)."""
    from facerecognition_b200 import ops
    gal, px = hists(faces_gpu(600, 112, 1))
    qh, _ = hists(faces_gpu(130, 112, 2))
    g8 = ops.compact_histograms(gal, px)
    d, i, s = ops.chisq_top1_filtered(qh, g8, px, want_scores=True)
    u, v, em, am = tables(px)
    M = u @ v.T                                                  # [a, b] approximate f
    G = gal.cpu().numpy().astype(np.int64)
    Q = qh.cpu().numpy().astype(np.int64)
    rows = np.r_[0:12, 118:130]
    cols = np.arange(G.shape[1])
    worst, worst_tab = 0.0, 0.0
    for r in rows:
        ref = M[G, Q[r][None, :]].sum(1)                         # float64 <U(g), V(q)>
        got = s[r].cpu().numpy().astype(np.float64)
        worst = max(worst, float(np.abs(got - ref).max()))
        exact = np.where(G + Q[r] > 0, G * Q[r] / np.maximum(G + Q[r], 1), 0.0).sum(1)
        e_tab = float(em[Q[r]].astype(np.float64).sum())
        assert np.abs(ref - exact).max() <= e_tab               # the table bound is rigorous
        worst_tab = max(worst_tab, float(np.abs(ref - exact).max()) / e_tab)
        allowance = 1e-4 * float(am[Q[r]].astype(np.float64).sum())   # cf_acc_rel() in chisq_filter.cu
        assert np.abs(got - ref).max() <= 0.25 * allowance, (np.abs(got - ref).max(), allowance)
    print(f"\n[filter] tensor-core accumulation error vs float64: max {worst:.4f} (S units); table error / bound max {worst_tab:.3f}")
    want_d, want_i = exact_top1(qh, g8, px)
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32))


@pytest.mark.parametrize("side,blocky,n_gal", [(112, True, 20_000), (100, False, 9_000)])
def test_filtered_top1_is_bit_identical_to_the_exact_scan(side, blocky, n_gal):
    """configs[4]-shaped data: planted queries, queries with NO match in the gallery, exact duplicates of gallery rows
    and duplicate rows inside the gallery (first row wins)."""
    from facerecognition_b200 import ops
    gimg = faces_gpu(n_gal, side, 10, blocky)
    gimg[n_gal - 5] = gimg[17]                                   # duplicate rows: the lower one must win
    gimg[4000] = gimg[17]
    gal, px = hists(gimg)
    g8 = ops.compact_histograms(gal, px)
    gen = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randint(0, n_gal, (100,), generator=gen, device="cuda")
    noisy = (gimg[src].float() + 6.0 * torch.randn((100, side, side), generator=gen, device="cuda")).clamp(0, 255).to(torch.uint8)
    fresh = faces_gpu(150, side, 11, blocky)
    dup = gimg[torch.tensor([17, 4000, 3, n_gal - 1, n_gal - 5], device="cuda")]
    qh, qpx = hists(torch.cat([noisy, fresh, dup, gimg[100:145]], 0))
    assert qpx == px and qh.shape[0] == 300
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    d, i = ops.chisq_top1_filtered(qh, g8, px, idx_base=7_000_000_000, stats=stats)
    want_d, want_i = exact_top1(qh, g8, px)
    assert torch.equal(i - 7_000_000_000, want_i)
    assert torch.equal(d.view(torch.int32), want_d.view(torch.int32))
    assert [int(x) - 7_000_000_000 for x in i[250:255, 0]] == [17, 17, 3, n_gal - 1, 17]
    assert float(d[250:255].abs().max()) == 0.0
    st = stats.cpu().numpy()
    print(f"\n[filter] {side}x{side} N={n_gal}: fallback queries {st[0]}, survivors/query {st[1] / 300:.1f}, raw/query {st[2] / 300:.1f}")
    assert st[0] == 0 and st[3] == 0 and st[1] < 300 * 400      # no fallback, no audited row outside the bound
    # the dispatching wrapper takes the same path
    d2, i2 = ops.chisq_topk(qh, px, g8, px, 1)
    assert torch.equal(i2, want_i) and torch.equal(d2.view(torch.int32), want_d.view(torch.int32))


def test_overflowing_queries_fall_back_to_the_exact_scan():
    from facerecognition_b200 import ops
    gal, px = hists(faces_gpu(9000, 112, 21))
    qh, _ = hists(faces_gpu(70, 112, 22))
    g8 = ops.compact_histograms(gal, px)
    want_d, want_i = exact_top1(qh, g8, px)
    os.environ["FRB_CHISQ_FILTER_ACC_REL"] = "10.0"              # a window that keeps every row: 9000 > cap
    try:
        stats = torch.zeros(4, dtype=torch.int32, device="cuda")
        d, i = ops.chisq_top1_filtered(qh, g8, px, stats=stats)
    finally:
        del os.environ["FRB_CHISQ_FILTER_ACC_REL"]
    assert int(stats[0]) == 70
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32))


@pytest.mark.parametrize("Q,N,side,grid", [(1, 1, 100, 8), (64, 255, 100, 8), (129, 300, 34, 4), (5, 0, 100, 8), (257, 513, 18, 2)])
def test_small_and_ragged_shapes(Q, N, side, grid):
    """Direct calls below the dispatch thresholds: single rows, ragged tiles, an empty gallery, shorter histograms
    (grid 4x4 -> 4096 bins, 2x2 -> 1024 bins; 8x8-pixel cells)."""
    from facerecognition_b200 import ops
    qh, px = hists(faces_gpu(Q, side, 31), grid_x=grid, grid_y=grid)
    assert px <= 255
    gal, _ = hists(faces_gpu(max(N, 1), side, 32), grid_x=grid, grid_y=grid)
    gal = gal[:N].contiguous()
    g8 = ops.compact_histograms(gal, px) if N else torch.empty((0, gal.shape[1]), dtype=torch.uint8, device="cuda")
    d, i = ops.chisq_top1_filtered(qh, g8, px)
    if N == 0:
        assert bool((i == -1).all()) and bool(torch.isinf(d).all())
        return
    want_d, want_i = exact_top1(qh, g8, px)
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32))


def test_rows_with_unequal_totals_and_arbitrary_counts():
    """The C ABI takes any u8 count matrix, not only LBPH histograms (whose rows all sum to cells x cell_px): the kernel
    ranks by sum_j f - (row total) / 4, so rows of different mass must still come out exactly as in the exact scan."""
    from facerecognition_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(77)
    n, nq, L, px = 6000, 90, 16384, 200
    dens = torch.rand((n, 1), generator=gen, device="cuda") * 0.12 + 0.02          # 2 % .. 14 % of the bins occupied
    g = (torch.rand((n, L), generator=gen, device="cuda") < dens) * torch.randint(1, 40, (n, L), generator=gen, device="cuda")
    g8 = g.to(torch.uint8).contiguous()
    q = g8[torch.randint(0, n, (nq,), generator=gen, device="cuda")].to(torch.int16)
    q = (q + (torch.rand(q.shape, generator=gen, device="cuda") < 0.01).to(torch.int16) * 3).clamp_(0, px)
    q[:10] = g8[:10].to(torch.int16)                                               # exact copies: distance 0
    qh = q.contiguous().view(torch.uint16)
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    d, i = ops.chisq_top1_filtered(qh, g8, px, stats=stats)
    want_d, want_i = exact_top1(qh, g8, px)
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32))
    assert int(stats[3]) == 0 and float(d[:10].abs().max()) == 0.0
    print(f"\n[filter] unequal totals: fallback {int(stats[0])}, survivors/query {int(stats[1]) / nq:.1f}")


def test_configs4_share_at_full_shape_is_identical_to_the_exact_scan():
    """BASELINE configs[4] as one GPU of eight sees it: 1024 frames against 125 000 gallery histograms, the bench's own data
    recipe (bench.c5_gallery / c5_step).  The filter's answers must equal the exact scan's bit for bit, with no fallback
    and no audited row outside the bound."""
    import bench
    from facerecognition_b200 import ops
    g8, px = bench.c5_gallery(torch, ops, torch.device("cuda"), 0, 125_000)
    frames = torch.cat([faces_gpu(256, 112, 41), bench.blocky_faces(torch, 768, 112, torch.device("cuda"), 31337)], 0)
    chunk0 = bench.blocky_faces(torch, bench.C5_CHUNK, 112, torch.device("cuda"), 500)
    frames[:128] = chunk0[1000:1128]                                    # exact re-shots: distance 0 at rows 1000..1127
    qh, qpx = hists(frames)
    assert qpx == px and g8.dtype == torch.uint8 and g8.shape == (125_000, 16384)
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    d, i = ops.chisq_top1_filtered(qh, g8, px, stats=stats)
    want_d, want_i = exact_top1(qh, g8, px)
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32))
    assert i[:128, 0].tolist() == list(range(1000, 1128)) and float(d[:128].abs().max()) == 0.0
    st = stats.cpu().tolist()
    assert st[0] == 0 and st[3] == 0
    print(f"\n[filter] configs[4] share: survivors/query {st[1] / 1024:.1f}, raw candidates/query {st[2] / 1024:.1f}")


def test_more_queries_than_one_pass_and_edge_cell_sizes():
    """1100 queries = two passes of the 1024-query workspace; cell_px = 255 is the largest cell a u8 gallery can hold
    (4-copy table); counts above the declared cell size (a contract violation) must not crash."""
    from facerecognition_b200 import ops
    gal, px = hists(faces_gpu(9000, 112, 51))
    qh, _ = hists(faces_gpu(1100, 112, 52))
    g8 = ops.compact_histograms(gal, px)
    d, i = ops.chisq_top1_filtered(qh, g8, px)
    want_d, want_i = exact_top1(qh, g8, px)
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32))
    gen = torch.Generator(device="cuda").manual_seed(4)
    g255 = (torch.rand((8300, 4096), generator=gen, device="cuda") < 0.05) * torch.randint(1, 256, (8300, 4096), generator=gen, device="cuda")
    g255 = g255.to(torch.uint8).contiguous()
    q255 = g255[torch.randint(0, 8300, (20,), generator=gen, device="cuda")].to(torch.int16).contiguous().view(torch.uint16)
    d, i = ops.chisq_top1_filtered(q255, g255, 255)
    want_d, want_i = exact_top1(q255, g255, 255)
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32)) and float(d.max()) == 0.0
    d, i = ops.chisq_top1_filtered(q255, g255, 100)          # counts up to 255 declared as cell_px = 100: defined behaviour, no fault
    torch.cuda.synchronize()
    assert i.shape == (20, 1)
    # a QUERY count above cell_px is outside the tables and their bound: such queries get the exact scan inside the call
    gal, px = hists(faces_gpu(9000, 112, 21))
    g8 = ops.compact_histograms(gal, px)
    qn = gal[:40].cpu().numpy().copy()
    qn[::2, 7] = px + 30
    qh = torch.from_numpy(qn).cuda()
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    d, i = ops.chisq_top1_filtered(qh, g8, px, stats=stats)
    want_d, want_i = exact_top1(qh, g8, px)
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32))
    assert int(stats[0]) == 20 and int(stats[3]) == 0


def test_faces_with_flat_and_saturated_regions():
    """SURVEY §8d's face mix (noise + a 255 stripe, blurred noise, piecewise-flat / saturated patches): cells whose pixels
    all fall into one or two bins (counts up to the whole cell) stress the large-count end of the feature tables."""
    import bench
    from facerecognition_b200 import ops
    dev = torch.device("cuda")
    gimg = bench.synthetic_faces(torch, 12_000, 112, 112, dev, seed=7)
    gal, px = hists(gimg)
    g8 = ops.compact_histograms(gal, px)
    qimg = torch.cat([gimg[torch.arange(0, 12_000, 100, device=dev)], bench.synthetic_faces(torch, 180, 112, 112, dev, seed=8)], 0)
    qimg[5, 40:60, 40:60] = 0                                             # a modified re-shot
    qh, _ = hists(qimg)
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    d, i = ops.chisq_top1_filtered(qh, g8, px, stats=stats)
    want_d, want_i = exact_top1(qh, g8, px)
    assert torch.equal(i, want_i) and torch.equal(d.view(torch.int32), want_d.view(torch.int32))
    st = stats.cpu().tolist()
    assert st[3] == 0
    print(f"\n[filter] flat / saturated mix: max count {int(gal.cpu().numpy().max())}, fallback {st[0]} of 300, survivors/query {st[1] / 300:.1f}")
