"""Host-side logic that needs no GPU: shard partitioning, the gather/merge plumbing under gloo with
world_size 2, gallery file formats, the mutation-tracking gallery dict."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_rows_exactly_once():
    from facerecognition_b200.sharded import shard_bounds
    for n in [0, 1, 7, 8, 9, 1000, 1_000_000, 100_000_000]:
        for r in [1, 2, 3, 4, 8]:
            spans = [shard_bounds(n, r, i) for i in range(r)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= lo <= hi for lo, hi in spans)
            assert max(hi - lo for lo, hi in spans) == (n + r - 1) // r if n else True


def test_balanced_bounds_cover_rows_exactly_once_and_follow_the_weights():
    from facerecognition_b200.sharded import balanced_bounds, shard_bounds
    for n in [0, 1, 255, 256, 1000, 125_000, 1_000_000, 100_000_000]:
        for w in ([1.0], [1, 1], [1.14, 1.05, 0.96, 0.89], [0.7, 1.3, 1.0, 1.0, 1.1, 0.9, 1.0, 1.0], [0, 2, 1], [5, 0, 0]):
            spans = [balanced_bounds(n, w, r) for r in range(len(w))]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and all(0 <= lo <= hi for lo, hi in spans)
            assert all(lo % 256 == 0 for lo, _ in spans[1:] if lo < n)       # inner boundaries sit on gallery tiles
            if n >= 100_000 and sum(w) > 0:
                for (lo, hi), wi in zip(spans, w):
                    assert abs((hi - lo) / n - wi / sum(w)) <= 256 * 2 / n + 1e-12
    assert [balanced_bounds(100, [0, 0], r) for r in range(2)] == [shard_bounds(100, 2, r) for r in range(2)]
    # the sharded answer does not depend on the split: same merged top-k from equal and from skewed shards
    from oracle import cosine as OC
    rng = np.random.default_rng(8)
    gal = rng.standard_normal((1031, 32)).astype(np.float32)
    gal[900] = gal[3]
    qs = rng.standard_normal((9, 32)).astype(np.float32)
    qs[0] = gal[3]
    ref_s, ref_i = OC.flat_ip_search(gal, qs, 5)
    for w in ([1, 1, 1], [0.5, 2.0, 1.0]):
        parts_s, parts_i = [], []
        for r in range(3):
            lo, hi = balanced_bounds(len(gal), w, r, align=16)
            s_, i_ = OC.flat_ip_search(gal[lo:hi], qs, 5)
            parts_s.append(torch.from_numpy(s_))
            parts_i.append(torch.from_numpy(np.where(i_ >= 0, i_ + lo, i_)))
        ms, mi = _np_merge(torch.stack(parts_s), torch.stack(parts_i), True)
        assert np.array_equal(mi.numpy(), ref_i) and np.allclose(ms.numpy(), ref_s, atol=1e-6)
    assert [int(x) for x in ref_i[0, :2]] == [3, 900]


def _np_merge(all_s, all_i, largest):
    """numpy stand-in for frb_topk_merge (same total order: key, then lowest id; id < 0 is padding)."""
    R, Q, k = all_s.shape
    s = all_s.permute(1, 0, 2).reshape(Q, R * k).numpy()
    i = all_i.permute(1, 0, 2).reshape(Q, R * k).numpy()
    key = np.where(i < 0, np.inf, -s if largest else s)
    order = np.lexsort((i, key), axis=1)[:, :k]
    return torch.from_numpy(np.take_along_axis(s, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))


def _worker(rank, world, port, q_out):
    try:
        _worker_body(rank, world, port, q_out)
    except Exception as e:  # surface the failure instead of a queue timeout
        q_out.put((rank, repr(e)))


def _worker_body(rank, world, port, q_out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from facerecognition_b200.sharded import ShardedSearch, shard_bounds
    from oracle import cosine as OC
    rng = np.random.default_rng(5)
    gal = rng.standard_normal((203, 64)).astype(np.float32)
    gal[150] = gal[20]                                   # tie across shards: lowest global id must win
    qs = np.concatenate([gal[[20, 199, 0]], rng.standard_normal((5, 64)).astype(np.float32)])
    lo, hi = shard_bounds(len(gal), world, rank)

    def local(q, k):
        s, i = OC.flat_ip_search(gal[lo:hi], q.numpy(), k)
        return torch.from_numpy(s), torch.from_numpy(np.where(i >= 0, i + lo, i))

    out_s, out_i = ShardedSearch(local, _np_merge, True).search(torch.from_numpy(qs), 5)
    ref_s, ref_i = OC.flat_ip_search(gal, qs, 5)
    ok = bool(np.array_equal(out_i.numpy(), ref_i) and np.allclose(out_s.numpy(), ref_s, atol=1e-6))
    ok = ok and out_i[0, 0].item() == 20 and out_i[0, 1].item() == 150
    # collective calibration: rank 1's probe step is 3x as slow -> its weight is the smaller one, both within 1 +- clamp
    import time
    from facerecognition_b200.sharded import measure_rank_weights
    w = measure_rank_weights(lambda: time.sleep(0.003 if rank == 1 else 0.001), lambda: None, seconds=0.3)
    ok = ok and len(w) == world and w[0] > w[1] and all(0.7 - 1e-9 <= x <= 1.3 + 1e-9 for x in w)
    # bench.py's warm-up: ranks that start at different times (and whose steps do not take equally long here, unlike a
    # sharded step) must still run the SAME number of steps, or one of them would wait in the exchange for a step its
    # peer never launches
    import bench

    def any_rank_wants_more(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(int(t.item()))

    time.sleep(0.04 * rank)
    n = bench.warm_in_lockstep(lambda: time.sleep(0.0004 * (1 + rank)), lambda: None, 3, 0.12, any_rank_wants_more)
    counts = [None] * world
    dist.all_gather_object(counts, n)
    ok = ok and len(set(counts)) == 1 and n >= 16 and n % 16 == 0
    q_out.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_search_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def _engine_worker(rank, world, port, q_out):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        sys.path.insert(0, ROOT)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import facerecognition_b200 as F
        from facerecognition_b200 import recognition_engine as RE
        from facerecognition_b200.sharded import ShardedSearch, shard_bounds
        from oracle import cosine as OC
        from oracle import lbph as OL
        rng = np.random.default_rng(11)
        gal = rng.standard_normal((37, 64)).astype(np.float32)
        gal /= np.linalg.norm(gal, axis=1, keepdims=True)
        gal[30] = gal[4]                                       # a tie across the two shards: the first identity wins
        db = {f"id_{i:03d}": g for i, g in enumerate(gal)}
        qs = np.concatenate([gal[[4, 36, 0]] * 1.7, rng.standard_normal((3, 64)).astype(np.float32)])

        # ---- RecognitionEngine(group=True): only the three device-touching seams are replaced by CPU stand-ins ----
        RE._queries_to_dev = lambda emb, dim, dev: torch.from_numpy(np.asarray(emb, np.float32).reshape(-1, dim))
        eng = F.RecognitionEngine(model_path=None, threshold=0.5, use_face_detection=False, group=True)
        eng.db = db
        lo, hi = shard_bounds(len(db), world, rank)

        class Gallery:
            names, dim = list(db.keys()), 64
        eng.gallery = lambda: Gallery

        def local(q, k):                                      # this rank's rows only, global ids
            S = np.array([[OC.cosine_similarity(a, b) for b in gal[lo:hi]] for a in q.numpy()], np.float32).reshape(len(q), hi - lo)
            order = np.argsort(-S, axis=1, kind="stable")[:, :k]
            return torch.from_numpy(np.take_along_axis(S, order, 1)), torch.from_numpy(order + lo)
        eng._local_topk = local
        eng._make_sharded = lambda: ShardedSearch(local, _np_merge, True, None)
        got = eng.recognize_embeddings(qs)
        ok = True
        for e, g in zip(qs, got):
            name, score, top = OC.recognize_with_db(db, e, 0.5)
            ok = ok and g[0] == name and abs(g[1] - score) < 1e-6 and [t[0] for t in g[2]] == [t[0] for t in top]
        ok = ok and [t[0] for t in got[0][2][:2]] == ["id_004", "id_030"]

        # ---- LBPHFaceRecognizer(group=True): train() keeps this rank's slice, predict merges across ranks ----
        faces = rng.integers(0, 256, (11, 40, 40), dtype=np.uint8)
        faces[9] = faces[2]                                    # duplicate faces on different ranks: the lower row wins
        labels = np.arange(100, 111, dtype=np.int32)
        ref = OL.OracleLBPH()
        ref.train(list(faces), labels)
        model = F.LBPHFaceRecognizer_create(group=True)
        flo, fhi = shard_bounds(len(faces), world, rank)
        kept = {}

        def fake_add(src, lab, what):                          # host part of _add with the K2 call replaced by the oracle
            from facerecognition_b200.lbph import _Group
            fs = list(src)
            a, b = shard_bounds(len(fs), world, rank)
            h, px = OL.c_lbp_hist(np.stack(fs[a:b])) if b > a else (np.zeros((0, 16384), np.uint16), 0)
            kept["hist"], kept["px"], kept["lo"] = h, px, model.size + a
            model._labels = np.concatenate([model._labels, np.asarray(lab, np.int32)])
        model._add = fake_add
        model.train(list(faces), labels)

        def local_lbph(qh, kk):
            d = np.stack([OL.c_chisq_scan_u16(kept["hist"], kept["px"], q.astype(np.uint16), kept["px"]) for q in qh.numpy()])
            d = d.astype(np.float32).reshape(len(qh), len(kept["hist"]))
            order = np.argsort(d, axis=1, kind="stable")[:, :kk]
            return torch.from_numpy(np.take_along_axis(d, order, 1)), torch.from_numpy(order + kept["lo"])
        model._make_sharded = lambda: ShardedSearch(None, _np_merge, False, None)
        model._search_local = lambda qh, qpx, kk: local_lbph(qh, kk)
        qh, qpx = OL.c_lbp_hist(faces[[2, 9, 5]])
        d, i = model._search(torch.from_numpy(qh.astype(np.int32)), qpx, 1)
        want = [ref.predict(f) for f in faces[[2, 9, 5]]]
        ok = ok and [int(model._labels[j]) for j in i[:, 0]] == [w[0] for w in want] == [102, 102, 105]
        ok = ok and all(abs(float(a) - w[1]) <= 1e-5 * max(w[1], 1e-30) for a, w in zip(d[:, 0], want))
        q_out.put((rank, bool(ok)))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q_out.put((rank, traceback.format_exc()))


def test_sharded_engine_and_lbph_classes_gloo_world2():
    """RecognitionEngine(group=...) and LBPHFaceRecognizer(group=...) shard their galleries by shard_bounds and route the
    batched entry points through ShardedSearch; with the CUDA seams replaced by the oracle, two gloo ranks must give the
    single-process reference answers (names, scores, tie order, labels, distances)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)], results


def test_faiss_flat_ip_file_roundtrip(tmp_path):
    from facerecognition_b200 import formats
    rows = np.random.default_rng(0).standard_normal((37, 512)).astype(np.float32)
    p = str(tmp_path / "arcface_index.faiss")
    formats.write_faiss_flat_ip(p, rows)
    raw = open(p, "rb").read()
    assert raw[:4] == b"IxFI" and len(raw) == 4 + 4 + 8 * 3 + 1 + 4 + 8 + rows.nbytes
    np.testing.assert_array_equal(formats.read_faiss_flat_ip(p), rows)
    open(p, "wb").write(b"IxHN" + raw[4:])
    with pytest.raises(ValueError, match="unsupported FAISS index"):
        formats.read_faiss_flat_ip(p)


def test_embedding_db_format_roundtrip(tmp_path):
    from facerecognition_b200 import formats
    db = {"alice": np.ones(512, np.float32), "bob": np.arange(512, dtype=np.float32)}
    p = str(tmp_path / "sub" / "arcface_embeddings_db.npy")
    formats.save_embedding_db(p, db)
    back = np.load(p, allow_pickle=True).item()          # how the reference loads it (recognition_engine.py:135)
    assert list(back) == ["alice", "bob"] and np.array_equal(back["bob"], db["bob"])
    assert list(formats.load_embedding_db(p)) == ["alice", "bob"]


def test_infer_cell_px_from_opencv_float_histograms(oracle_lbph, lbph_golden):
    from facerecognition_b200.formats import _infer_cell_px
    for tag in ["s100", "s112", "s57x83"]:
        px = int(lbph_golden[f"{tag}_cell_px"])
        hf = oracle_lbph.hist_to_f32(lbph_golden[f"{tag}_hist"], px)
        for row in hf:
            assert _infer_cell_px(row) == px
            np.testing.assert_array_equal(np.round(row.astype(np.float64) * px).astype(np.uint16), np.round(row * px))


def test_gallery_dict_tracks_mutations():
    from facerecognition_b200.recognition_engine import _GalleryDict
    d = _GalleryDict({"a": 1})
    v = d.version
    d["b"] = 2; assert d.version > v; v = d.version
    del d["a"]; assert d.version > v; v = d.version
    d.update(c=3); assert d.version > v; v = d.version
    d.pop("c"); assert d.version > v; v = d.version
    d.clear(); assert d.version > v and len(d) == 0


def test_group_plan_orders_samples_by_label_then_index():
    from facerecognition_b200.recognition_engine import group_plan
    labels = np.array([2, 0, 2, 1, 0, 2])
    order, offsets = group_plan(labels, 4)
    assert order.tolist() == [1, 4, 3, 0, 2, 5] and offsets.tolist() == [0, 2, 3, 6, 6]
    with pytest.raises(IndexError):
        group_plan(np.array([0, 5]), 3)
    order, offsets = group_plan(np.zeros(0, np.int64), 2)
    assert order.size == 0 and offsets.tolist() == [0, 0, 0]


def test_threshold_sweep_matches_the_real_reference():
    """facerecognition_b200.evaluation.threshold_sweep vs the report of inference/evaluate.py:61-128 (sweep_golden.npz)."""
    import json
    from facerecognition_b200.evaluation import threshold_sweep
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sweep_golden.npz"))
    ref = json.loads(str(g["report"]))
    got = threshold_sweep(g["sim"], g["y_true"], g["y_pred"])
    assert got["best_f1_threshold"] == ref["best_f1_threshold"] and got["best_accuracy_threshold"] == ref["best_accuracy_threshold"]
    for a, b in zip(got["results"], ref["results"]):
        assert a.keys() == b.keys()
        for k in a:
            assert a[k] == pytest.approx(b[k], abs=1e-12), k


def test_lbph_cell_size_inference_from_stored_float_rows():
    """formats._infer_cell_px: a row whose smallest count does not divide the cell size (5 of 144) must still be read
    (round-1 advisor finding); a reading is valid when it reproduces the stored floats as integer counts per cell."""
    from facerecognition_b200.formats import _infer_cell_px
    rng = np.random.default_rng(0)
    for px in (144, 169, 16, 1, 2401, 7):
        for trial in range(6):
            counts = rng.multinomial(px, rng.dirichlet(np.ones(256) * rng.choice([0.05, 1, 10])), size=64).astype(np.float32)
            if trial == 0 and px >= 7:
                counts[:] = 0
                counts[:, 0], counts[:, 1] = 5, px - 5
            h = (counts * np.float32(1.0 / px)).reshape(-1)
            got = _infer_cell_px(h)
            assert got == px, (px, got, trial)
    with pytest.raises(ValueError):
        _infer_cell_px(np.full(16384, 0.3, np.float32))
    with pytest.raises(ValueError):
        _infer_cell_px(np.zeros(16384, np.float32))


def test_batched_result_formatting_fast_path_equals_the_row_by_row_form():
    """RecognitionEngine._format_db_results: the bulk form (one object-array gather of the names) must build exactly the
    tuples of the reference's per-row form (recognition_engine.py:282-289), incl. 'Unknown' below the threshold and lists
    cut short by -1 padding (galleries with fewer than 5 rows)."""
    from facerecognition_b200.recognition_engine import RecognitionEngine

    class _G:
        pass

    eng = RecognitionEngine.__new__(RecognitionEngine)
    eng.threshold = 0.85
    g = _G()
    g.names, g._names_obj = [f"id_{i:03d}" for i in range(50)], None
    rng = np.random.default_rng(0)
    s = np.sort(rng.random((9, 5)).astype(np.float32))[:, ::-1].copy()
    r = rng.integers(0, 50, (9, 5))
    eng._gallery = g
    fast = eng._format_db_results(s, r, g.names)
    eng._gallery = None
    slow = eng._format_db_results(s, r, g.names)
    assert fast == slow and g._names_obj is not None
    assert {x[0] == "Unknown" for x in fast} == {True, False}
    assert all(isinstance(x[1], float) and len(x[2]) == 5 for x in fast)
    r[3, 4] = -1
    eng._gallery = g
    padded = eng._format_db_results(s, r, g.names)
    assert len(padded[3][2]) == 4 and padded[2] == slow[2]
