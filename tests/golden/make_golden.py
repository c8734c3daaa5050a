#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.  Run in the AUTHORING container only:

    python tests/golden/make_golden.py

* cosine_golden.npz — outputs of the REAL reference, imported from /root/reference
  (inference/recognition_engine.py: cosine_similarity :41-63, RecognitionEngine.recognize_with_db
  :267-289) on seeded synthetic galleries.  These pin oracle/cosine.py and the CUDA path.
* lbph_golden.npz — seeded synthetic faces (recipe of models/lbphmodel/test_lbph_logic.py:18-33
  plus blurred / flat / saturated patches), the oracle's u16 histograms for them (PARITY
  UNPINNED: cv2.face is not installed anywhere we can run), and chi-square distances computed
  by the REAL cv2.compareHist(HISTCMP_CHISQR_ALT) of the installed OpenCV core (pinned).

* gallery_golden.npz — outputs of the REAL reference's gallery builders on seeded embeddings:
  compute_prototypes (inference/extract_embeddings.py:555-592) and RecognitionEngine.add_to_db
  (inference/recognition_engine.py:391-422, with extract_embedding stubbed to pass embeddings through).

/root/reference does not exist on the GPU box; the tests read only the .npz files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def unit(x):
    return (x / np.linalg.norm(x, axis=-1, keepdims=True)).astype(np.float32)


def make_cosine():
    sys.path.insert(0, "/root/reference")
    from inference.recognition_engine import RecognitionEngine, cosine_similarity  # the real reference

    rng = np.random.default_rng(20261018)
    D = 512
    out = {}

    # case A: unit-norm gallery (how database_builder/extract_embeddings emit it), planted + random queries
    NA, QA = 384, 48
    gal = unit(rng.standard_normal((NA, D)))
    src = rng.integers(0, NA, QA)
    q = unit(gal[src] + 0.03 * rng.standard_normal((QA, D)).astype(np.float32))
    q[::8] = unit(rng.standard_normal((len(q[::8]), D)))          # pure random -> "Unknown" at 0.5
    q[1] *= 3.5                                                    # un-normalised query -> division branch
    q[2] = 0.0                                                     # zero query -> all scores 0.0
    gal[7] = gal[3]                                                # exact duplicate rows -> tie, insertion order wins
    gal[11] = 0.0                                                  # zero gallery row -> 0.0
    gal[13] *= 1.0005                                              # inside the 1e-3 "normalised" window -> raw dot
    gal[17] *= 1.5                                                 # outside -> divided
    names = np.array([f"id_{i:05d}" for i in range(NA)])
    eng = RecognitionEngine(model_path=None, db_path=None, use_face_detection=False, threshold=0.5)
    eng.db = {n: g for n, g in zip(names, gal)}
    best_names, best_scores, top_names, top_scores = [], [], [], []
    for e in q:
        bn, bs, tk = eng.recognize_with_db(e)
        best_names.append(bn)
        best_scores.append(bs)
        top_names.append([t[0] for t in tk])
        top_scores.append([t[1] for t in tk])
    # full score matrix through the reference's scalar function
    S = np.array([[cosine_similarity(e, g) for g in gal] for e in q], np.float64)
    out.update(a_gallery=gal, a_names=names, a_queries=q, a_threshold=np.float64(0.5),
               a_best_names=np.array(best_names), a_best_scores=np.array(best_scores, np.float64),
               a_top_names=np.array(top_names), a_top_scores=np.array(top_scores, np.float64), a_scores=S)

    # case B: tiny gallery (fewer than 5 identities) + empty-db sentinel
    gal_b = unit(rng.standard_normal((3, D)))
    eng.db = {f"p{i}": g for i, g in enumerate(gal_b)}
    eng.set_threshold(0.65)
    qb = unit(gal_b[1] + 0.02 * rng.standard_normal(D).astype(np.float32))
    bn, bs, tk = eng.recognize_with_db(qb)
    out.update(b_gallery=gal_b, b_query=qb, b_best_name=np.array(bn), b_best_score=np.float64(bs),
               b_top_names=np.array([t[0] for t in tk]), b_top_scores=np.array([t[1] for t in tk], np.float64))
    eng.db = None
    sent = eng.recognize_with_db(qb)
    out.update(b_sentinel_name=np.array(sent[0]), b_sentinel_score=np.float64(sent[1]))
    np.savez_compressed(os.path.join(HERE, "cosine_golden.npz"), **out)
    print("cosine_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


def synth_faces(rng, count, h, w):
    """1/3 uniform noise + class stripe (test_lbph_logic.py:26-28), 1/3 blurred noise, 1/3 flat/saturated patches."""
    import cv2
    faces = np.zeros((count, h, w), np.uint8)
    for i in range(count):
        kind = i % 3
        if kind == 0:
            img = rng.integers(0, 255, (h, w), dtype=np.uint8)
            c = (i // 3) % max(1, h // 10)
            img[c * 10:(c + 1) * 10, :] = 255
        elif kind == 1:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
            img = cv2.GaussianBlur(img, (0, 0), 1.5 + (i % 4))
        else:
            img = np.zeros((h, w), np.uint8)
            levels = [0, 1, 2, 3, 7, 127, 128, 254, 255, 64, 200, 31]
            for _ in range(24):
                y0, x0 = rng.integers(0, h), rng.integers(0, w)
                y1, x1 = min(h, y0 + rng.integers(4, 40)), min(w, x0 + rng.integers(4, 40))
                img[y0:y1, x0:x1] = levels[rng.integers(0, len(levels))]
        faces[i] = img
    return faces


def make_lbph():
    import cv2
    from oracle import lbph as O

    rng = np.random.default_rng(2024)
    out = {}
    for tag, (n, h, w) in {"s100": (12, 100, 100), "s112": (9, 112, 112), "s57x83": (6, 57, 83)}.items():
        faces = synth_faces(rng, n, h, w)
        hist, px = O.c_lbp_hist(faces)
        codes0 = O.c_elbp(faces[0])
        hf = O.hist_to_f32(hist, px)
        # pinned anchor: the real cv2.compareHist on the float32 view, all pairs (i, j)
        d = np.array([[cv2.compareHist(hf[i], hf[j], cv2.HISTCMP_CHISQR_ALT) for j in range(n)] for i in range(n)],
                     np.float64)
        out.update({f"{tag}_faces": faces, f"{tag}_hist": hist, f"{tag}_cell_px": np.int32(px),
                    f"{tag}_codes0": codes0, f"{tag}_cv2_chisq": d})
    # flat images: gray-level dependent codes
    flat_levels = np.arange(256, dtype=np.uint8)
    flat_codes = np.array([O.c_elbp(np.full((5, 5), v, np.uint8))[1, 1] for v in flat_levels], np.int32)
    out.update(flat_levels=flat_levels, flat_codes=flat_codes)
    np.savez_compressed(os.path.join(HERE, "lbph_golden.npz"), **out)
    print("lbph_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


def make_gallery():
    sys.path.insert(0, "/root/reference")
    from inference.extract_embeddings import compute_prototypes          # the real reference
    from inference.recognition_engine import RecognitionEngine

    rng = np.random.default_rng(77)
    D, C, M = 512, 37, 411
    labels = rng.integers(0, C, M).astype(np.int64)
    labels[:C] = np.arange(C)                                             # every class present (the reference indexes by label)
    emb = unit(rng.standard_normal((M, D)) + 0.5 * rng.standard_normal((C, D))[labels])
    protos = compute_prototypes(emb, labels)
    eng = RecognitionEngine(model_path=None, db_path=None, use_face_detection=False)
    eng.extract_embedding = lambda x: x                                   # "images" are already embeddings
    groups = [emb[labels == c] for c in (0, 5, 36)]
    for c, g in zip((0, 5, 36), groups):
        assert eng.add_to_db(f"id_{c}", list(g))
    assert eng.add_to_db("nobody", []) is False
    out = dict(emb=emb, labels=labels, prototypes=protos,
               add_names=np.array(list(eng.db.keys())), add_rows=np.stack([eng.db[k] for k in eng.db]),
               add_classes=np.array([0, 5, 36]))
    np.savez_compressed(os.path.join(HERE, "gallery_golden.npz"), **out)
    print("gallery_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


def make_sweep():
    """sweep_golden.npz — the REAL inference/evaluate.py threshold_sweep (:61-128) on seeded scores.  The module's
    plotting imports (matplotlib, seaborn — not installed here) are stubbed; threshold_sweep itself is pure numpy."""
    import json
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference")
    from inference.evaluate import threshold_sweep as ref_sweep
    rng = np.random.default_rng(3)
    n = 500
    yt = rng.integers(0, 40, n)
    yp = np.where(rng.random(n) < 0.8, yt, rng.integers(0, 40, n))
    sim = np.clip(rng.normal(0.6, 0.2, n), 0, 1)
    np.savez_compressed(os.path.join(HERE, "sweep_golden.npz"), sim=sim, y_true=yt, y_pred=yp,
                        report=np.array(json.dumps(ref_sweep(sim, yt, yp))))
    print("sweep_golden.npz written")


if __name__ == "__main__":
    which = sys.argv[1:] or ["lbph", "cosine", "gallery", "sweep"]
    if "lbph" in which:
        make_lbph()
    if "cosine" in which:
        make_cosine()
    if "gallery" in which:
        make_gallery()
    if "sweep" in which:
        make_sweep()
