"""oracle/cosine.py pinned on outputs of the REAL reference (tests/golden/cosine_golden.npz, produced by
tests/golden/make_golden.py importing /root/reference/inference/recognition_engine.py)."""
import numpy as np

from oracle import cosine as O


def _db(g, prefix):
    return {str(n): v for n, v in zip(g[f"{prefix}_names"], g[f"{prefix}_gallery"])}


def test_cosine_similarity_matches_reference_scores(cosine_golden):
    g = cosine_golden
    gal, q, S = g["a_gallery"], g["a_queries"], g["a_scores"]
    got = np.array([[O.cosine_similarity(e, r) for r in gal] for e in q[:12]])
    # same numpy calls as the reference -> bit-equal
    np.testing.assert_array_equal(got, S[:12])
    assert S[2].max() == 0.0 and S[2].min() == 0.0          # zero query -> 0.0 everywhere
    assert np.all(S[:, 11] == 0.0)                            # zero gallery row -> 0.0


def test_recognize_with_db_matches_reference(cosine_golden):
    g = cosine_golden
    db = _db(g, "a")
    for i, e in enumerate(g["a_queries"]):
        name, score, top = O.recognize_with_db(db, e, float(g["a_threshold"]))
        assert name == str(g["a_best_names"][i])
        assert score == g["a_best_scores"][i]
        assert [t[0] for t in top] == [str(x) for x in g["a_top_names"][i]]
        np.testing.assert_array_equal([t[1] for t in top], g["a_top_scores"][i])
    # planted queries are recognised, random ones fall under the threshold
    assert (g["a_best_names"] == "Unknown").sum() >= 6
    assert (g["a_best_names"] != "Unknown").sum() >= 30


def test_duplicate_rows_keep_insertion_order(cosine_golden):
    g = cosine_golden
    db = _db(g, "a")
    q = g["a_gallery"][3]
    _, _, top = O.recognize_with_db(db, q, 0.5)
    assert [t[0] for t in top[:2]] == ["id_00003", "id_00007"]   # equal scores: first inserted wins


def test_small_gallery_and_sentinels(cosine_golden):
    g = cosine_golden
    db = {f"p{i}": v for i, v in enumerate(g["b_gallery"])}
    name, score, top = O.recognize_with_db(db, g["b_query"], 0.65)
    assert name == str(g["b_best_name"]) and score == float(g["b_best_score"])
    assert [t[0] for t in top] == [str(x) for x in g["b_top_names"]] and len(top) == 3
    assert O.recognize_with_db(None, g["b_query"], 0.65) == (str(g["b_sentinel_name"]), float(g["b_sentinel_score"]), [])


def test_flat_ip_and_faiss_restatement():
    rng = np.random.default_rng(7)
    emb = rng.standard_normal((50, 64)).astype(np.float32)
    idx = O.build_flat_ip(emb)
    np.testing.assert_allclose(np.linalg.norm(idx, axis=1), 1.0, atol=1e-6)
    s, i = O.flat_ip_search(idx, idx[:4], 3)
    assert list(i[:, 0]) == [0, 1, 2, 3] and np.all(np.diff(s, axis=1) <= 0)
    s, i = O.flat_ip_search(idx[:2], idx[:1], 5)             # k > ntotal -> (-inf, -1) padding like faiss
    assert list(i[0]) == [0, 1, -1, -1, -1] and np.isinf(s[0, 2:]).all()
    name, score, res = O.recognize_with_faiss(idx, None, emb[5] * 3.0, 5, 0.5)
    assert name == "ID_5" and abs(score - 1.0) < 1e-6 and len(res) == 5
    assert O.recognize_with_faiss(None, None, emb[0], 5, 0.5) == ("No FAISS index", 0.0, [])
    assert O.recognize_with_faiss(idx, None, -emb[5], 1, 0.5)[0] == "Unknown"


def test_facenet_matcher_and_prototypes():
    rng = np.random.default_rng(8)
    db = {f"n{i}": rng.standard_normal(512).astype(np.float32) * (1 + i) for i in range(9)}
    name, score, dist, top = O.facenet_match(db, db["n4"] * 0.2, 0.5)
    assert name == "n4" and abs(score - 1.0) < 1e-6 and dist < 1e-3 and len(top) == 5
    assert all(abs(d - np.sqrt(max(2 - 2 * s, 0))) < 2e-3 for _, s, d in top)
    e = np.stack([O.l2_normalize(rng.standard_normal(512).astype(np.float32)) for _ in range(6)])
    lab = np.array([0, 0, 1, 1, 1, 2])
    P = O.compute_prototypes(e, lab)
    np.testing.assert_allclose(np.linalg.norm(P, axis=1), 1.0, atol=1e-6)
    np.testing.assert_allclose(P[1], O.mean_prototype(list(e[2:5])), atol=1e-7)
    s, i = O.batched_topk(e, P, 2)
    assert list(i[:, 0]) == list(lab)


def test_oracle_prototypes_match_the_real_reference():
    """oracle.cosine.compute_prototypes / mean_prototype vs outputs of the reference's compute_prototypes
    (inference/extract_embeddings.py:555-592) and add_to_db (recognition_engine.py:391-422) in gallery_golden.npz."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gallery_golden.npz"))
    from oracle import cosine as OC
    np.testing.assert_allclose(OC.compute_prototypes(g["emb"], g["labels"]), g["prototypes"], rtol=0, atol=1e-7)
    for c, row in zip(g["add_classes"], g["add_rows"]):
        np.testing.assert_allclose(OC.mean_prototype(list(g["emb"][g["labels"] == c])), row, rtol=0, atol=1e-7)
