"""BASELINE configs[3] and configs[4] AT SHAPE on one GPU, checked against the oracle on samples the host finishes in
seconds (the whole 100M x 512 / 1M x 16384 problems are hours of CPU work):

* configs[3]: 100M x 512 bf16 rows (102.4 GB, one B200), 4096 queries, top-5.  Size-independent properties over the
  whole answer (planted source first, sorted, unique, in range) + for a 16-query slice (i) every returned (row, score)
  re-computed in float64 from the stored bf16 operands and (ii) the float64 oracle over SAMPLED 250k-row blocks: no
  sampled row may beat the returned 5th score (a missed row would), and every sampled row above it must be in the list.
* configs[4]: 1024 frames of 112x112 against 1M u8 LBPH histograms (16.4 GB) through the tensor-core filter: planted
  frames return their source row; for a 32-query slice the answer is bit-identical to the exact scan over all 1M rows;
  for 8 queries the oracle (oracle/lbph_oracle.c, float64 compareHist form) over sampled rows + the returned row
  confirms distance (1e-5 relative) and that no sampled row is nearer.

Skipped when the GPU has less free memory than the shape needs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DIM = 512


def _free_gib():
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    return free / 2**30


def test_configs3_100m_rows_sampled_oracle():
    from facerecognition_b200 import ops, _native as NV
    N, Q, k, BLOCK = 100_000_000, 4096, 5, 1_000_000
    if _free_gib() < N * DIM * 2 / 2**30 + 12:
        pytest.skip(f"needs {N * DIM * 2 / 2**30 + 12:.0f} GiB of free HBM, {_free_gib():.0f} free")
    gal = torch.empty((N, DIM), dtype=torch.bfloat16, device="cuda")
    for b in range(N // BLOCK):
        gen = torch.Generator(device="cuda").manual_seed(9000 + b)
        gal[b * BLOCK:(b + 1) * BLOCK] = ops.normalize_rows(torch.randn((BLOCK, DIM), generator=gen, device="cuda"),
                                                            NV.FRB_QNORM_CLAMP, torch.bfloat16)
    gen = torch.Generator(device="cuda").manual_seed(99)
    src = torch.randint(0, N, (Q,), generator=gen, device="cuda")
    q = gal[src].float() + 0.03 * torch.randn((Q, DIM), generator=gen, device="cuda")
    n_rand = Q // 10
    q[:n_rand] = torch.randn((n_rand, DIM), generator=gen, device="cuda")
    q = q.contiguous()
    s, i = ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
    assert torch.equal(i[n_rand:, 0], src[n_rand:])                         # planted source wins among 100M rows
    assert bool((s[n_rand:, 0] > 0.75).all()) and bool((s[:n_rand, 0] < 0.5).all())
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    assert bool(((i >= 0) & (i < N)).all())
    srt = i.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())

    # the oracle's arithmetic on a slice: float64 inner products of the same bf16-rounded operands
    sl = slice(400, 416)                                                     # 10 random + 6 planted queries
    q16 = ops.normalize_rows(q[sl].contiguous(), NV.FRB_QNORM_CLAMP, torch.bfloat16).float().cpu().numpy().astype(np.float64)
    got_s, got_i = s[sl].cpu().numpy(), i[sl].cpu().numpy()
    rows = gal[torch.from_numpy(got_i.reshape(-1)).cuda()].float().cpu().numpy().astype(np.float64).reshape(16, k, DIM)
    want = np.einsum("qd,qkd->qk", q16, rows)
    assert np.abs(got_s - want).max() <= 2e-6, np.abs(got_s - want).max()   # only the fp32 accumulation order differs
    rng = np.random.default_rng(5)
    kth = got_s[:, k - 1]
    checked = 0
    for lo in sorted(int(x) * 250_000 for x in rng.choice(N // 250_000, 12, replace=False)):
        chunk = gal[lo:lo + 250_000].float().cpu().numpy().astype(np.float64)
        S = q16 @ chunk.T                                                     # the oracle's scores of the sampled rows
        for r in range(16):
            above = np.nonzero(S[r] > kth[r] + 2e-6)[0] + lo                  # must all have been returned
            assert set(above.tolist()) <= set(got_i[r].tolist()), (r, lo, above[:4])
            inside = (got_i[r] >= lo) & (got_i[r] < lo + 250_000)
            assert np.abs(S[r][got_i[r][inside] - lo] - got_s[r][inside]).max(initial=0.0) <= 2e-6
        checked += chunk.shape[0]
    assert checked == 3_000_000
    # one query (the recognize() path: row-streaming kernel) over the same 100M rows gives the batched answer
    s1, i1 = ops.cosine_topk(q[n_rand:n_rand + 1].contiguous(), gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
    assert torch.equal(i1[0], i[n_rand]) and float((s1[0] - s[n_rand]).abs().max()) <= 2e-6


def test_configs4_1m_histograms_filter_vs_exact_and_oracle(oracle_lbph):
    from facerecognition_b200 import ops
    from test_gpu_chisq_filter import faces_gpu, exact_top1
    N, Q, CH = 1_000_000, 1024, 50_000
    if _free_gib() < 30:
        pytest.skip(f"needs 30 GiB of free HBM, {_free_gib():.0f} free")
    parts, px = [], None
    for c in range(N // CH):
        h8, px = ops.lbp_hist(faces_gpu(CH, 112, 500 + c), counts8=True)
        parts.append(h8)
    g8 = torch.cat(parts, 0)
    del parts
    assert g8.shape == (N, 16384) and g8.dtype == torch.uint8 and px == 169
    gen = torch.Generator(device="cuda").manual_seed(77)
    n_plant = Q // 4
    src = torch.randint(0, N, (n_plant,), generator=gen, device="cuda")
    src_faces = torch.empty((n_plant, 112, 112), dtype=torch.uint8, device="cuda")
    for c in torch.unique(src // CH).tolist():                                # regenerate the chunks the sources live in
        m = (src // CH) == c
        src_faces[m] = faces_gpu(CH, 112, 500 + c)[src[m] - c * CH]
    planted = (src_faces.float() + 6.0 * torch.randn((n_plant, 112, 112), generator=gen, device="cuda")).clamp(0, 255).to(torch.uint8)
    frames = torch.cat([planted, faces_gpu(Q - n_plant, 112, 31337)], 0).contiguous()
    qh, qpx = ops.lbp_hist(frames)
    assert qpx == px
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    d, i = ops.chisq_top1_filtered(qh, g8, px, stats=stats)
    assert torch.equal(i[:n_plant, 0], src)                                   # noisy re-shots find their gallery face
    assert int(stats[3]) == 0                                                 # audit: no survivor outside the proven bound
    # the exact scan over all 1M rows, 32 queries (16 planted + 16 unmatched): identical bits
    pick = torch.cat([torch.arange(0, 16), torch.arange(n_plant, n_plant + 16)]).cuda()
    want_d, want_i = exact_top1(qh.view(torch.int16)[pick].view(torch.uint16).contiguous(), g8, px)   # (no u16 index kernel in torch)
    assert torch.equal(i[pick], want_i) and torch.equal(d[pick].view(torch.int32), want_d.view(torch.int32))
    # the oracle on samples: the returned row's distance, and no sampled row nearer than it
    rng = np.random.default_rng(11)
    for r in [0, 1, 2, 3, n_plant, n_plant + 1, n_plant + 2, Q - 1]:
        row = int(i[r, 0])
        lo = int(rng.integers(0, N - 20_000))
        sample = torch.cat([g8[row:row + 1], g8[lo:lo + 20_000]], 0).cpu().numpy().astype(np.uint16)
        ref = oracle_lbph.c_chisq_scan_u16(sample, px, qh[r].cpu().numpy(), px)
        got = float(d[r, 0])
        assert abs(got - ref[0]) <= 1e-5 * ref[0] + 1e-12, (r, got, ref[0])
        nearer = np.nonzero(ref[1:] < ref[0] * (1 - 1e-5))[0]
        assert nearer.size == 0, (r, nearer[:4] + lo, ref[1:][nearer[:4]], ref[0])
        ties_before = np.nonzero((ref[1:] == ref[0]) & (np.arange(lo, lo + 20_000) < row))[0]
        assert ties_before.size == 0                                          # first row wins exact ties
