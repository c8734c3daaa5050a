"""Cosine-path parity on the GPU through the C ABI: the reference's own outputs (golden fixtures),
the oracle on seeded inputs, and size-independent properties at BASELINE's full sizes.
Tolerances (north_star): 1e-5 for fp32 scores, 1e-3 for bf16; top-1 identical except stated near-ties."""
import numpy as np
import pytest
import torch

from oracle import cosine as OC

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_BF16 = 1e-3


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def unit(x):
    return (x / np.linalg.norm(x, axis=-1, keepdims=True)).astype(np.float32)


def check_topk(scores, idx, ref_scores_full, k, tol):
    """scores/idx [Q,k] vs the full reference score matrix: same values within tol; same ids except near-ties."""
    order = np.lexsort((np.broadcast_to(np.arange(ref_scores_full.shape[1]), ref_scores_full.shape), -ref_scores_full), axis=1)[:, :k]
    want = np.take_along_axis(ref_scores_full, order, 1)
    np.testing.assert_allclose(scores, want, atol=tol, rtol=0)
    mism = idx != order
    for r, j in zip(*np.nonzero(mism)):
        assert abs(ref_scores_full[r, idx[r, j]] - want[r, j]) <= 2 * tol, (r, j, idx[r, j], order[r, j])
    return order


def test_golden_reference_outputs_through_engine(cosine_golden):
    """The real reference's recognize_with_db outputs, replayed through the drop-in RecognitionEngine."""
    import facerecognition_b200 as F
    g = cosine_golden
    eng = F.RecognitionEngine(model_path=None, threshold=float(g["a_threshold"]), use_face_detection=False)
    eng.db = {str(n): v for n, v in zip(g["a_names"], g["a_gallery"])}
    S = g["a_scores"]
    batched = eng.recognize_embeddings(g["a_queries"])
    for i, e in enumerate(g["a_queries"]):
        name, score, top = eng.recognize_with_db(e)
        # one query runs the row-streaming kernel, the batch the tiled one: same answer up to fp32 summation order
        bname, bscore, btop = batched[i]
        assert abs(score - bscore) <= 2e-6 and len(top) == len(btop)
        assert name == bname or abs(score - float(g["a_threshold"])) <= 2e-6
        np.testing.assert_allclose([t[1] for t in top], [t[1] for t in btop], atol=2e-6, rtol=0)
        if [t[0] for t in top] != [t[0] for t in btop]:      # names may swap only between scores that close
            assert all(abs(a[1] - b[1]) <= 2e-6 for a, b in zip(top, btop))
        assert abs(score - g["a_best_scores"][i]) <= TOL_F32
        np.testing.assert_allclose([t[1] for t in top], g["a_top_scores"][i], atol=TOL_F32, rtol=0)
        ref_names = [str(x) for x in g["a_top_names"][i]]
        got_names = [t[0] for t in top]
        if got_names != ref_names:           # only allowed where the reference's own scores are near-tied
            row = dict(zip([str(n) for n in g["a_names"]], S[i]))
            assert all(abs(row[a] - row[b]) <= 2 * TOL_F32 for a, b in zip(got_names, ref_names))
        assert name == str(g["a_best_names"][i]) or abs(g["a_best_scores"][i] - float(g["a_threshold"])) <= TOL_F32
    assert [t[0] for t in eng.recognize_with_db(g["a_gallery"][3])[2][:2]] == ["id_00003", "id_00007"]  # stable tie order
    # zero query: every score 0.0 -> first five identities in insertion order, "Unknown"
    name, score, top = eng.recognize_with_db(g["a_queries"][2])
    assert name == "Unknown" and score == 0.0 and [t[0] for t in top] == [f"id_{i:05d}" for i in range(5)]


def test_engine_sentinels_small_db_and_mutation(cosine_golden, tmp_path):
    import facerecognition_b200 as F
    g = cosine_golden
    eng = F.RecognitionEngine(model_path=None, threshold=0.65, use_face_detection=False)
    assert eng.recognize_with_db(g["b_query"]) == (str(g["b_sentinel_name"]), float(g["b_sentinel_score"]), [])
    assert eng.recognize_with_faiss(g["b_query"]) == ("No FAISS index", 0.0, [])
    assert eng.recognize("x.jpg")["status"] == "error"                       # no embedder, like no checkpoint
    eng.db = {f"p{i}": v for i, v in enumerate(g["b_gallery"])}
    name, score, top = eng.recognize_with_db(g["b_query"])
    assert name == str(g["b_best_name"]) and abs(score - float(g["b_best_score"])) <= TOL_F32 and len(top) == 3
    assert [t[0] for t in top] == [str(x) for x in g["b_top_names"]]
    # callers mutate engine.db in place (recognition_engine.py:419); the device copy must follow
    eng.db["new"] = g["b_query"]
    assert eng.recognize_with_db(g["b_query"])[0] == "new"
    # embedder-driven recognize()/add_to_db()/save_db()
    table = {"img_a": g["b_gallery"][0], "img_b": g["b_gallery"][2], "bad": None}
    eng2 = F.RecognitionEngine(model_path=None, threshold=0.5, use_face_detection=False, embedder=table.get)
    assert eng2.recognize("img_a") == {"identity": "Unknown", "confidence": 0.0, "top_k": [], "embedding": table["img_a"],
                                       "status": "error", "message": "No database loaded"}
    assert eng2.add_to_db("alice", ["img_a", "bad"]) and not eng2.add_to_db("nobody", ["bad"])
    r = eng2.recognize("img_a")
    assert r["identity"] == "alice" and abs(r["confidence"] - 1.0) < 1e-5 and r["status"] == "success"
    assert [d["identity"] for d in eng2.recognize_batch(["img_a", "img_b", "bad"])][:2] == ["alice", "Unknown"]
    p = str(tmp_path / "db.npy")
    eng2.save_db(p)
    eng3 = F.RecognitionEngine(model_path=None, db_path=p, use_face_detection=False)
    assert eng3.get_db_identities() == ["alice"]
    assert abs(F.cosine_similarity(g["b_gallery"][0] * 2, g["b_gallery"][1]) - OC.cosine_similarity(g["b_gallery"][0] * 2, g["b_gallery"][1])) < TOL_F32
    assert F.cosine_similarity(np.zeros(512), g["b_gallery"][1]) == 0.0


@pytest.mark.parametrize("Q,N,k", [(1, 1, 1), (3, 5, 5), (64, 128, 1), (65, 129, 5), (256, 10000, 1), (256, 10000, 5),
                                   (7, 4099, 16), (130, 700, 64)])
def test_fp32_kernel_vs_oracle(Q, N, k):
    """frb_cosine_topk, fp32 gallery; (256, 10000) is BASELINE configs[1]."""
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(Q * 1000 + N)
    gal = unit(rng.standard_normal((N, 512)))
    q = unit(gal[rng.integers(0, N, Q)] + 0.03 * rng.standard_normal((Q, 512)).astype(np.float32))
    q[::7] = rng.standard_normal((len(q[::7]), 512)).astype(np.float32) * 3          # un-normalised
    if N > 10:
        gal[N // 2] = gal[1]                                                         # exact duplicate
        gal[3] *= 2.0
        gal[4] = 0
    for mode in ("ip", "ref", "eps"):
        if mode == "ip":
            s, i = ops.cosine_topk(dev(q), dev(gal), k)
            ref = q.astype(np.float64) @ gal.T.astype(np.float64)
        elif mode == "eps":
            s, i = ops.cosine_topk(dev(q), dev(gal), k, qnorm_mode=NV.FRB_QNORM_EPS)
            qn = q / (np.linalg.norm(q, axis=1, keepdims=True) + 1e-8)
            ref = qn.astype(np.float64) @ gal.T.astype(np.float64)
        else:
            s, i = ops.cosine_topk(dev(q), dev(gal), k, score_mode=NV.FRB_SCORE_REF_COSINE,
                                   q_norms=ops.row_norms(dev(q)), g_norms=ops.row_norms(dev(gal)))
            ref = np.array([[OC.cosine_similarity(a, b) for b in gal] for a in q[:16]]) if N <= 5000 else None
        s, i = s.cpu().numpy(), i.cpu().numpy()
        kk = min(k, N)
        assert np.all(i[:, kk:] == -1) and np.all(np.isinf(s[:, kk:]))
        if ref is None:
            continue
        rows = slice(0, ref.shape[0])
        check_topk(s[rows, :kk], i[rows, :kk], ref, kk, TOL_F32 * (10 if mode == "ip" else 1))
    if N > 10:  # duplicate rows: the lower index comes first
        s, i = ops.cosine_topk(dev(gal[1:2]), dev(gal), min(2, k) if k > 1 else 1)
        assert i[0, 0].item() == 1 and (k == 1 or i[0, 1].item() == N // 2)


def bf16_round(x):
    return torch.from_numpy(x).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("Q,N,k", [(1, 1, 1), (5, 255, 5), (128, 256, 5), (129, 257, 5), (300, 5000, 5), (64, 70000, 1),
                                   (1000, 33333, 16), (200, 20000, 9), (150, 9000, 33), (140, 30000, 64),
                                   # one (query tile, gallery group) unit supplies ALL k results with k == the register list
                                   # length: a queued admission must not evict a top-k row (round-1 advisor finding)
                                   (128, 256, 16), (128, 256, 32), (128, 256, 64), (100, 200, 64), (19000, 16000, 16)])
def test_bf16_tensor_core_kernel_vs_oracle(Q, N, k):
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(Q + N)
    gal = unit(rng.standard_normal((N, 512)))
    q = gal[rng.integers(0, N, Q)] + 0.03 * rng.standard_normal((Q, 512)).astype(np.float32)
    q[::5] = rng.standard_normal((len(q[::5]), 512)).astype(np.float32)
    q *= rng.uniform(0.5, 4.0, (Q, 1)).astype(np.float32)                           # the kernel normalises
    gal_bf16 = ops.normalize_rows(dev(gal), NV.FRB_QNORM_NONE, torch.bfloat16)
    s, i = ops.cosine_topk(dev(q), gal_bf16, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    kk = min(k, N)
    assert np.all(i[:, kk:] == -1)
    # (1) against the fp32 reference scores (what the reference's numpy would give): 1e-3
    ref32 = OC.l2_normalize(q).astype(np.float64) @ gal.T.astype(np.float64)
    check_topk(s[:, :kk], i[:, :kk], ref32, kk, TOL_BF16)
    # (2) against exact arithmetic on the bf16-rounded operands: only fp32 accumulation error remains
    q16 = ops.normalize_rows(dev(q), NV.FRB_QNORM_CLAMP, torch.bfloat16).float().cpu().numpy()   # the prologue's own rounding
    assert np.abs(q16 - OC.l2_normalize(q)).max() <= 2.0 ** -8
    ref16 = q16.astype(np.float64) @ gal_bf16.float().cpu().numpy().T.astype(np.float64)
    check_topk(s[:, :kk], i[:, :kk], ref16, kk, 2e-6)


def test_bf16_ties_and_idx_base():
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(3)
    gal = unit(rng.standard_normal((3000, 512)))
    gal[2900] = gal[17]
    gal[300] = gal[17]
    g16 = ops.normalize_rows(dev(gal), NV.FRB_QNORM_NONE, torch.bfloat16)
    s, i = ops.cosine_topk(dev(gal[17:18]), g16, 3, qnorm_mode=NV.FRB_QNORM_CLAMP, idx_base=10_000_000_000)
    assert [int(x) - 10_000_000_000 for x in i[0]] == [17, 300, 2900] and s[0, 0] == s[0, 1] == s[0, 2]


def test_faiss_mode_and_facenet_matcher(tmp_path):
    import facerecognition_b200 as F
    rng = np.random.default_rng(12)
    emb = rng.standard_normal((500, 512)).astype(np.float32) * rng.uniform(0.5, 2, (500, 1)).astype(np.float32)
    p = str(tmp_path / "arcface_index.faiss")
    index = F.build_faiss_index(emb, p)
    assert index.ntotal == 500
    eng = F.RecognitionEngine(model_path=None, faiss_index_path=p, threshold=0.5, use_face_detection=False)
    rows = OC.build_flat_ip(emb)
    for e in [emb[7] * 0.3, emb[400] + 0.05 * rng.standard_normal(512).astype(np.float32), rng.standard_normal(512).astype(np.float32)]:
        name, score, res = eng.recognize_with_faiss(e, 5)
        rname, rscore, rres = OC.recognize_with_faiss(rows, None, e, 5, 0.5)
        assert name == rname and abs(score - rscore) <= TOL_F32 and [r[0] for r in res] == [r[0] for r in rres]
    s, i = index.search(rows[:3], 2)
    assert list(i[:, 0]) == [0, 1, 2]
    db = {f"n{i}": emb[i] for i in range(60)}
    for e in [emb[9] * 2, rng.standard_normal(512).astype(np.float32)]:
        got = F.match_facenet(db, e, 0.5)
        name, score, dist, top = OC.facenet_match(db, e, 0.5)
        assert got["identity"] == name and abs(got["confidence"] - score) <= TOL_F32
        assert [t[0] for t in got["top_k"]] == [t[0] for t in top]
        np.testing.assert_allclose([t[2] for t in got["top_k"]], [t[2] for t in top], atol=1e-4)


def test_full_size_1m_gallery_properties():
    """BASELINE configs[2]: 1M x 512 bf16 gallery, 4096 queries, top-5 — checked through size-independent
    properties: planted queries return their source row first; scores are sorted; ids are unique and in
    range; a 64-query slice agrees with the float64 oracle computed on the host over all 1M rows."""
    from facerecognition_b200 import ops, _native as NV
    N, Q, k = 1_000_000, 4096, 5
    gen = torch.Generator(device="cuda").manual_seed(1234)
    gal = torch.randn((N, 512), generator=gen, device="cuda")
    gal16 = ops.normalize_rows(gal, NV.FRB_QNORM_CLAMP, torch.bfloat16)
    gen = torch.Generator(device="cuda").manual_seed(4321)
    src = torch.randint(0, N, (Q,), generator=gen, device="cuda")
    q = ops.normalize_rows(gal[src].contiguous(), NV.FRB_QNORM_CLAMP) + 0.03 * torch.randn((Q, 512), generator=gen, device="cuda")
    n_rand = Q // 10
    q[:n_rand] = torch.randn((n_rand, 512), generator=gen, device="cuda")
    del gal
    s, i = ops.cosine_topk(q.contiguous(), gal16, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
    assert torch.equal(i[n_rand:, 0], src[n_rand:])                         # planted source wins
    assert bool((s[n_rand:, 0] > 0.75).all()) and bool((s[:n_rand, 0] < 0.5).all())
    assert bool((s[:, :-1] >= s[:, 1:]).all())                              # sorted
    assert bool(((i >= 0) & (i < N)).all())
    srt = i.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                          # unique ids per query
    # a 64-query slice against the ORACLE at full size: float64 inner products of the same bf16-rounded operands over all
    # 1M rows on the host (oracle.cosine.batched_topk semantics: descending score, ties -> lowest row), 65 GFLOP of dgemm
    sub = q[1000:1064].contiguous()
    q16 = ops.normalize_rows(sub, NV.FRB_QNORM_CLAMP, torch.bfloat16).float().cpu().numpy().astype(np.float64)
    best_s = np.full((64, 0), -np.inf)
    best_i = np.zeros((64, 0), np.int64)
    for lo in range(0, N, 125_000):
        chunk = gal16[lo:lo + 125_000].float().cpu().numpy().astype(np.float64)
        S = q16 @ chunk.T
        part = np.argpartition(-S, 8, axis=1)[:, :8]
        best_s = np.concatenate([best_s, np.take_along_axis(S, part, 1)], 1)
        best_i = np.concatenate([best_i, part + lo], 1)
    order = np.lexsort((best_i, -best_s), axis=1)[:, :k]
    want_s, want_i = np.take_along_axis(best_s, order, 1), np.take_along_axis(best_i, order, 1)
    got_s, got_i = s[1000:1064].cpu().numpy(), i[1000:1064].cpu().numpy()
    assert np.abs(got_s - want_s).max() <= 2e-6                              # only the fp32 accumulation differs
    mism = got_i != want_i
    assert np.all(np.abs(got_s - want_s)[mism] <= 2e-6) and mism.sum() <= 2  # ids may swap only between scores that close
    assert np.array_equal(got_i[:, 0], want_i[:, 0])


def test_gallery_builders_match_the_reference_outputs():
    """K4 (frb_group_mean_renorm) behind compute_prototypes / add_to_db / build_db_from_embeddings vs the REAL
    reference's outputs (tests/golden/gallery_golden.npz).  Tolerance 1e-6 absolute on unit-norm rows."""
    import os
    import facerecognition_b200 as F
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gallery_golden.npz"))
    emb, labels = g["emb"], g["labels"]
    protos = F.compute_prototypes(emb, labels)
    assert protos.dtype == np.float32 and protos.shape == g["prototypes"].shape
    np.testing.assert_allclose(protos, g["prototypes"], rtol=0, atol=1e-6)
    eng = F.RecognitionEngine(model_path=None, use_face_detection=False, embedder=lambda x: x)
    for c in g["add_classes"]:
        assert eng.add_to_db(f"id_{c}", list(emb[labels == c]))
    assert eng.add_to_db("nobody", []) is False
    assert list(eng.db.keys()) == list(g["add_names"])
    np.testing.assert_allclose(np.stack([eng.db[k] for k in eng.db]), g["add_rows"], rtol=0, atol=1e-6)
    names = [f"id_{c}" for c in range(40)]                     # 37 classes present, 3 identities without samples
    db = F.build_db_from_embeddings(names, emb, labels)
    assert list(db.keys()) == names[:37]
    np.testing.assert_allclose(np.stack(list(db.values())), g["prototypes"], rtol=0, atol=1e-6)
    # device-level: bf16 copy for the tensor-core gallery, empty groups all-zero
    from facerecognition_b200 import ops
    order, offsets = F.group_plan(labels, 40)
    o32, o16 = ops.group_mean_renorm(torch.from_numpy(emb).cuda(), torch.from_numpy(order).cuda(),
                                     torch.from_numpy(offsets).cuda(), want_bf16=True)
    assert torch.equal(o16, o32.to(torch.bfloat16)) and float(o32[37:].abs().max()) == 0.0


def test_topk_accuracy_matches_the_notebook_form():
    """evaluation.topk_accuracy (fused K1) vs the notebooks' np.dot + argmax / argsort on the golden gallery."""
    import os
    from facerecognition_b200 import evaluation as EV
    from oracle import cosine as OC
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gallery_golden.npz"))
    emb, labels, protos = g["emb"], g["labels"], g["prototypes"]
    S = emb @ protos.T                                                   # evaluate_arcface_kaggle.ipynb:618
    top1 = float((np.argmax(S, 1) == labels).mean())
    top5 = float((np.argsort(S, 1)[:, -5:] == labels[:, None]).any(1).mean())   # :713
    for bf16 in (False, True):
        r = EV.topk_accuracy(emb, labels, protos, (1, 5), bf16=bf16)
        assert r["top1_accuracy"] == pytest.approx(top1, abs=(0 if not bf16 else 2 / len(labels)))
        assert r["top5_accuracy"] == pytest.approx(top5, abs=(0 if not bf16 else 2 / len(labels)))
        np.testing.assert_allclose(r["similarities"], S.max(1), atol=1e-5 if not bf16 else 1e-3)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("Q", [1, 2, 3, 4])
def test_small_batches_take_the_row_streaming_kernel_and_agree_with_the_batched_one(Q, dtype):
    """Few queries go to cosine_gemv_kernel; the same queries inside a 64-query batch go to the tiled kernels.
    Same operand rounding => same ids (ties included) and scores within fp32 summation-order noise."""
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(100 + Q)
    N, k = 30_011, 7
    gal = unit(rng.standard_normal((N, 512)))
    gal[20_000] = gal[11]                                               # duplicate rows: lowest id first
    q = gal[rng.integers(0, N, 64)] + 0.05 * rng.standard_normal((64, 512)).astype(np.float32)
    q[0] = gal[11]
    q *= rng.uniform(0.5, 3.0, (64, 1)).astype(np.float32)
    if dtype == "bf16":
        g = ops.normalize_rows(dev(gal), NV.FRB_QNORM_NONE, torch.bfloat16)
        kw = dict(qnorm_mode=NV.FRB_QNORM_CLAMP)
    else:
        g = dev(gal)
        kw = dict(qnorm_mode=NV.FRB_QNORM_EPS)
    NV.profile_enable(True)
    NV.profile_read(NV.K_COSINE_GEMV)
    s1, i1 = ops.cosine_topk(dev(q[:Q]), g, k, idx_base=5, **kw)
    _, launches = NV.profile_read(NV.K_COSINE_GEMV)
    NV.profile_enable(False)
    expect = 1 if (dtype == "f32" or Q <= 2) else 0     # bf16: the tcgen05 path wins from 3 queries up (cosine_gemv.cu)
    assert launches == expect, "kernel selection by batch size changed"
    s64, i64 = ops.cosine_topk(dev(q), g, k, idx_base=5, **kw)
    assert torch.equal(i1, i64[:Q]) and float((s1 - s64[:Q]).abs().max()) <= 2e-6
    assert [int(x) - 5 for x in i1[0, :2]] == [11, 20_000]
    if dtype == "f32":                                                   # the reference's cosine rule with explicit norms
        qn, gn = ops.row_norms(dev(q)), ops.row_norms(g)
        r1 = ops.cosine_topk(dev(q[:Q]), g, k, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn[:Q].contiguous(), g_norms=gn)
        r64 = ops.cosine_topk(dev(q), g, k, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=gn)
        assert torch.equal(r1[1], r64[1][:Q]) and float((r1[0] - r64[0][:Q]).abs().max()) <= 2e-6
    # fewer rows than k, and an empty gallery
    s, i = ops.cosine_topk(dev(q[:Q]), g[:3].contiguous(), k, **kw)
    assert bool((i[:, 3:] == -1).all()) and bool((i[:, :3] >= 0).all())
    s, i = ops.cosine_topk(dev(q[:Q]), g[:0].contiguous(), k, **kw)
    assert bool((i == -1).all())


def test_entry_points_are_reentrant_from_python_threads():
    """The reference's callers are Flask request threads and daemon threads (web_app.py:1028, database_builder.py:114):
    concurrent calls on one engine / one recognizer must not interfere (per-call workspaces, thread-local error state)."""
    import threading
    import facerecognition_b200 as F
    rng = np.random.default_rng(9)
    gal = unit(rng.standard_normal((5000, 512)))
    eng = F.RecognitionEngine(model_path=None, threshold=0.3, use_face_detection=False)
    eng.db = {f"id_{i:05d}": g for i, g in enumerate(gal)}
    faces = rng.integers(0, 256, (64, 100, 100), dtype=np.uint8)
    model = F.train_lbph_model(list(faces), np.arange(64, dtype=np.int32))
    errors = []

    def worker(t):
        try:
            for j in range(25):
                r = (t * 25 + j) % 5000
                q = gal[r] + 0.01 * rng.standard_normal(512).astype(np.float32)
                assert eng.recognize_with_db(q)[0] == f"id_{r:05d}"
                got = eng.recognize_embeddings(np.stack([gal[r], gal[(r + 1) % 5000]]))
                assert [g[0] for g in got] == [f"id_{r:05d}", f"id_{(r + 1) % 5000:05d}"]
                lab, dist = model.predict(faces[(t + j) % 64])
                assert lab == (t + j) % 64 and dist == 0.0
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


@pytest.mark.parametrize("first_pass", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,Q,k", [(10_000, 256, 5), (3000, 40, 1), (70_000, 333, 16),
                                   (66_000, 38_100, 5)])   # 298 query tiles: the first pass has ONE main gallery group
def test_tensor_core_first_pass_plus_exact_rescore_equals_the_exact_kernel(N, Q, k, first_pass):
    """ops.cosine_topk_refined / cosine_topk_exact (fp16 or bf16 tcgen05 first pass, exact fp32 re-score, completeness
    proof) must return exactly what the fp32 kernel returns; (10000, 256, 5) is configs[1]."""
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(N + Q)
    gal = unit(rng.standard_normal((N, 512)))
    gal[5] *= 1.0004                                     # cosine_similarity()'s raw-dot window
    gal[6] *= 1.7                                        # division branch
    gal[N - 1] = gal[7]                                  # duplicate rows: lowest id first
    q = gal[rng.integers(0, N, Q)] + 0.03 * rng.standard_normal((Q, 512)).astype(np.float32)
    q[0] = gal[7]
    q[1] *= 2.5
    q[2] = gal[5] * 0.9995                               # both norms inside the raw-dot window
    g = dev(gal)
    g16 = ops.normalize_rows(g, NV.FRB_QNORM_CLAMP, first_pass)
    qn, gn = ops.row_norms(dev(q)), ops.row_norms(g)
    want_s, want_i = ops.cosine_topk(dev(q), g, k, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=gn)
    s, i, fail, flags = ops.cosine_topk_refined(dev(q), g, g16, k, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=gn)
    assert int(fail.item()) == int(flags.sum().item())
    ok = flags == 0                                      # proven lists: identical to the exact kernel
    if first_pass == torch.float16 or k == 1:
        assert int(fail.item()) <= Q // 50, "planted queries over a random gallery leave a wide margin"
    assert float((s - want_s)[ok].abs().max()) <= 2e-6
    mism = (i != want_i)
    assert bool((((s - want_s).abs() <= 2e-6) | ~mism)[ok].all())      # ids may differ only between scores that close
    assert int(i[0, 0]) == 7 and (k == 1 or int(i[0, 1]) == N - 1)
    s2, i2 = ops.cosine_topk_exact(dev(q), g, g16, k, q_norms=qn, g_norms=gn)   # unproven queries re-run exactly
    assert float((s2 - want_s).abs().max()) <= 2e-6 and bool((((s2 - want_s).abs() <= 2e-6) | (i2 == want_i)).all())
    # a query with no margin (all scores zero) must be reported, not silently answered
    z = torch.zeros((17, 512), device="cuda")
    _, _, fail, flags = ops.cosine_topk_refined(z, g, g16, k, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=ops.row_norms(z), g_norms=gn)
    assert int(fail.item()) == 17 and int(flags.sum().item()) == 17
    sz, iz = ops.cosine_topk_exact(z, g, g16, k, q_norms=ops.row_norms(z), g_norms=gn)
    assert float(sz.abs().max()) == 0.0 and iz[:, 0].tolist() == [0] * 17    # zero scores: first rows in order, like the reference


def test_fp16_first_pass_bound_holds_on_adversarial_rows():
    """REFINE_EPS_F16 must dominate |first-pass score - true cosine| also for vectors whose mass sits in a few large
    components or in thousands of tiny ones (fp16 subnormals), and the unproven near-tie must be flagged."""
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(99)
    N = 20_000
    gal = unit(rng.standard_normal((N, 512)))
    gal[1] = unit(np.r_[np.ones(3), np.full(509, 1e-6)][None].astype(np.float32))[0]      # tiny tail: fp16 subnormals
    gal[2] = unit(np.r_[1.0, np.zeros(511)][None].astype(np.float32))[0]
    gal[3] = unit((rng.standard_normal(512) ** 5)[None].astype(np.float32))[0]            # heavy-tailed components
    q = np.concatenate([gal[1:4] * 3.0, gal[rng.integers(0, N, 61)] + 0.2 * rng.standard_normal((61, 512)).astype(np.float32)])
    g = dev(gal)
    g16 = ops.normalize_rows(g, NV.FRB_QNORM_CLAMP, torch.float16)
    approx, cand = ops.cosine_topk(dev(q), g16, 16, qnorm_mode=NV.FRB_QNORM_CLAMP)
    true = OC.l2_normalize(q).astype(np.float64) @ OC.l2_normalize(gal).astype(np.float64).T
    got = approx.cpu().numpy().astype(np.float64)
    ref = np.take_along_axis(true, cand.cpu().numpy(), 1)
    assert np.abs(got - ref).max() <= ops.REFINE_EPS_F16, np.abs(got - ref).max()
    print(f"\n[refine] fp16 first pass |score - cosine| max {np.abs(got - ref).max():.2e} (bound {ops.REFINE_EPS_F16:.1e})")


def test_engine_batches_on_a_large_gallery_use_the_first_pass_and_agree_with_single_queries():
    """RecognitionEngine.recognize_embeddings with a batch (>= ops.REFINE_MIN_QUERIES) over a 70k-identity dict DB takes the tensor-core
    first pass + exact re-score; each answer must equal the single-query (row-streaming, exact) answer."""
    import facerecognition_b200 as F
    from facerecognition_b200 import ops
    rng = np.random.default_rng(1)
    n, nq = 70_000, 1100
    assert nq >= ops.REFINE_MIN_QUERIES and n >= ops.REFINE_MIN_ROWS
    gal = unit(rng.standard_normal((n, 512)))
    eng = F.RecognitionEngine(model_path=None, threshold=0.4, use_face_detection=False)
    eng.db = {f"id_{i:06d}": g for i, g in enumerate(gal)}
    src = rng.integers(0, n, nq)
    q = gal[src] + 0.03 * rng.standard_normal((nq, 512)).astype(np.float32)
    batched = eng.recognize_embeddings(q)
    assert all(b[0] == f"id_{s:06d}" for b, s in zip(batched, src))
    for j in (0, 7, nq - 1):
        name, score, top = eng.recognize_with_db(q[j])
        assert batched[j][0] == name == f"id_{src[j]:06d}" and abs(batched[j][1] - score) <= 2e-6
        assert [t[0] for t in batched[j][2]] == [t[0] for t in top]
    # 5-7 queries (the smallest batches that take the first pass, from 64k rows) against the exact single-query answers
    assert ops.refine_applicable(5, n, 512, 5) and not ops.refine_applicable(3, n, 512, 5)
    for nb in (5, 7):
        small = eng.recognize_embeddings(q[20:20 + nb])
        for j in range(nb):
            name, score, top = eng.recognize_with_db(q[20 + j])
            assert small[j][0] == name and abs(small[j][1] - score) <= 2e-6 and [t[0] for t in small[j][2]] == [t[0] for t in top]


def test_prenormalised_bf16_queries_entry_equals_the_fp32_query_entry():
    """frb_cosine_topk_bf16q (what a sharded search calls after gathering the ranks' normalised bf16 query slices) vs
    frb_cosine_topk with the prologue: identical results, through the tcgen05 kernel and the row-streaming one."""
    from facerecognition_b200 import ops, _native as NV
    rng = np.random.default_rng(8)
    gal = ops.normalize_rows(dev(unit(rng.standard_normal((20_011, 512)))), NV.FRB_QNORM_NONE, torch.bfloat16)
    for Q in (1, 2, 5, 300):
        q = dev(rng.standard_normal((Q, 512)).astype(np.float32) * 3)
        want = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP, idx_base=11)
        q16 = ops.normalize_rows(q, NV.FRB_QNORM_CLAMP, torch.bfloat16)
        got = ops.cosine_topk_bf16q(q16, gal, 5, idx_base=11)
        assert torch.equal(got[1], want[1]) and torch.equal(got[0], want[0]), Q


def test_any_embedding_dimension_and_batched_recognize_batch(tmp_path):
    """The reference's matchers take any embedding length (numpy loops); here rows and queries are zero-padded to the
    kernels' vector width.  recognize_batch matches all images in one call and equals per-image recognize()."""
    import facerecognition_b200 as F
    rng = np.random.default_rng(21)
    for dim in (100, 129, 512):
        gal = unit(rng.standard_normal((300, dim)))
        db = {f"p{i:03d}": g for i, g in enumerate(gal)}
        q = gal[rng.integers(0, 300, 6)] + 0.02 * rng.standard_normal((6, dim)).astype(np.float32)
        table = {f"img{j}": e for j, e in enumerate(q)}
        table["bad"] = None
        eng = F.RecognitionEngine(model_path=None, threshold=0.5, use_face_detection=False, embedder=table.get)
        eng.db = db
        for e in q[:3]:
            name, score, top = eng.recognize_with_db(e)
            rname, rscore, rtop = OC.recognize_with_db(db, e, 0.5)
            assert name == rname and abs(score - rscore) <= TOL_F32 and [t[0] for t in top] == [t[0] for t in rtop]
        imgs = ["img0", "bad", "img3", "img5"]
        batch = eng.recognize_batch(imgs)
        single = [eng.recognize(x) for x in imgs]
        assert [b["status"] for b in batch] == ["success", "error", "success", "success"]
        for b, s1 in zip(batch, single):
            assert b["identity"] == s1["identity"] and b["status"] == s1["status"] and abs(b["confidence"] - s1["confidence"]) <= 2e-6
            assert [t[0] for t in b["top_k"]] == [t[0] for t in s1["top_k"]]
        # device tensors in, device tensors out
        s_dev, i_dev = eng.recognize_embeddings_device(torch.from_numpy(q).cuda(), 5)
        assert s_dev.is_cuda and [eng.gallery().names[int(j)] for j in i_dev[:, 0]] == [r[0] if r[0] != "Unknown" else eng.gallery().names[int(i_dev[n, 0])]
                                                                                       for n, r in enumerate(eng.recognize_embeddings(q))]
        # FAISS-style index with the same dimension
        index = F.build_faiss_index(gal, str(tmp_path / f"i{dim}.faiss"))
        back = F.FlatIPIndex.from_file(str(tmp_path / f"i{dim}.faiss"))
        s1, i1 = index.search(q, 3)
        s2, i2 = back.search(torch.from_numpy(q).cuda(), 3)
        rs, ri = OC.flat_ip_search(OC.build_flat_ip(gal), q, 3)
        assert np.array_equal(i1, ri) and np.array_equal(i2, ri) and np.abs(s1 - rs).max() <= TOL_F32 * 10
        got = F.match_facenet(db, q[0], 0.5)
        name, score, dist, top = OC.facenet_match(db, q[0], 0.5)
        assert got["identity"] == name and abs(got["confidence"] - score) <= TOL_F32
    # in-place dict mutations that used to go unnoticed
    eng.db |= {"zz_new": unit(rng.standard_normal((1, 512)))[0]}
    assert "zz_new" in eng.gallery().names
    eng.db["zz_new"][:] = eng.db["p000"]
    eng.db.invalidate()
    assert eng.recognize_with_db(eng.db["p000"])[2][1][0] == "zz_new"
