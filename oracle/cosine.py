"""oracle/cosine.py — NumPy restatement of the reference's cosine identification path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference lines it follows; paths are relative to /root/reference.

PINNED: tests/test_oracle_cosine.py replays tests/golden/cosine_golden.npz, which
tests/golden/make_golden.py produced by importing the real
inference/recognition_engine.py.  The FAISS restatement is PARITY UNPINNED
(faiss not installed; requirements.txt:43).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import os

import numpy as np


def cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """inference/recognition_engine.py:41-63.

    float32 flatten (:48-49); 0.0 if either norm is 0 (:55-56); raw dot when both
    norms are within 1e-3 of 1 (:59-60); else dot/(na*nb) (:63).
    """
    a = np.asarray(a).astype(np.float32).flatten()
    b = np.asarray(b).astype(np.float32).flatten()
    na = np.linalg.norm(a)
    nb = np.linalg.norm(b)
    if na == 0 or nb == 0:
        return 0.0
    if abs(na - 1.0) < 1e-3 and abs(nb - 1.0) < 1e-3:
        return float(np.dot(a, b))
    return float(np.dot(a, b) / (na * nb))


def recognize_with_db(db: Dict[str, np.ndarray], embedding: np.ndarray, threshold: float
                      ) -> Tuple[str, float, List[Tuple[str, float]]]:
    """RecognitionEngine.recognize_with_db, inference/recognition_engine.py:267-289.

    Dict insertion order, stable descending sort (ties keep insertion order, :282),
    strict '<' threshold (:286), top-5 fixed (:287,289), sentinel for no db (:274-275).
    """
    if db is None:
        return "No database", 0.0, []
    scores = [(name, cosine_similarity(embedding, vec)) for name, vec in db.items()]
    scores.sort(key=lambda x: x[1], reverse=True)
    best_name, best_score = scores[0]
    if best_score < threshold:
        return "Unknown", best_score, scores[:5]
    return best_name, best_score, scores[:5]


def build_flat_ip(embeddings: np.ndarray) -> np.ndarray:
    """build_faiss_index, inference/extract_embeddings.py:619-635: rows/(‖row‖+1e-8), IndexFlatIP.add."""
    e = np.asarray(embeddings).astype("float32")
    norms = np.linalg.norm(e, axis=1, keepdims=True)
    return (e / (norms + 1e-8)).astype(np.float32)


def flat_ip_search(index_rows: np.ndarray, queries: np.ndarray, k: int):
    """faiss.IndexFlatIP.search restated: exact fp32 inner product, scores descending,
    id -1 / score -inf padding when k > ntotal.  Tie order: lowest id first (FAISS's own
    tie order is unspecified — PARITY UNPINNED)."""
    q = np.asarray(queries, np.float32).reshape(-1, index_rows.shape[1])
    s = q @ index_rows.T
    n = index_rows.shape[0]
    order = np.lexsort((np.broadcast_to(np.arange(n), s.shape), -s), axis=1)[:, :k]
    scores = np.take_along_axis(s, order, 1)
    if k > n:
        pad = k - n
        scores = np.concatenate([scores, np.full((q.shape[0], pad), -np.inf, np.float32)], 1)
        order = np.concatenate([order, np.full((q.shape[0], pad), -1, np.int64)], 1)
    return scores.astype(np.float32), order.astype(np.int64)


def recognize_with_faiss(index_rows, id_to_label, embedding: np.ndarray, k: int, threshold: float):
    """RecognitionEngine.recognize_with_faiss, inference/recognition_engine.py:291-326.

    e/(‖e‖+1e-8) (:301-302); search (:304); skip id -1 (:310); name lookup with
    "ID_<n>" default (:312-315); empty -> ("Unknown", 0.0, []) (:318-319); strict '<' (:323).
    """
    if index_rows is None:
        return "No FAISS index", 0.0, []
    e = np.asarray(embedding).astype(np.float32).reshape(1, -1)
    e = e / (np.linalg.norm(e) + 1e-8)
    scores, indices = flat_ip_search(index_rows, e, k)
    results = []
    for idx, score in zip(indices.flatten(), scores.flatten()):
        if idx == -1:
            continue
        name = id_to_label.get(idx, f"ID_{idx}") if id_to_label else f"ID_{idx}"
        results.append((name, float(score)))
    if len(results) == 0:
        return "Unknown", 0.0, []
    best_name, best_score = results[0]
    if best_score < threshold:
        return "Unknown", best_score, results
    return best_name, best_score, results


def facenet_match(db: Dict[str, np.ndarray], embedding: np.ndarray, threshold: float):
    """Inline FaceNet matcher, web_app.py:537-562.

    e/=(‖e‖+1e-8) (:540); per row d/=(‖d‖+1e-8) (:549), score=e·d (:551),
    distance=‖e-d‖ (:552); stable sort desc (:554); strict '<' (:558); top_k[:5] (:562).
    Returns (identity, confidence, distance, top_k[(name, score, distance)]).
    """
    e = np.asarray(embedding).flatten()
    e = e / (np.linalg.norm(e) + 1e-8)
    top_k = []
    for name, db_emb in db.items():
        d = np.asarray(db_emb).flatten()
        d = d / (np.linalg.norm(d) + 1e-8)
        top_k.append((name, float(np.dot(e, d)), float(np.linalg.norm(e - d))))
    top_k.sort(key=lambda x: x[1], reverse=True)
    best_name, best_score, best_dist = top_k[0]
    if best_score < threshold:
        best_name = "Unknown"
    return best_name, best_score, best_dist, top_k[:5]


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """torch.nn.functional.normalize(x, p=2, dim=1) = x / max(‖x‖, eps):
    inference/extract_embeddings.py:381,434; models/facenet/facenet_model.py:35."""
    x = np.asarray(x, np.float32)
    n = np.sqrt(np.sum(x.astype(np.float32) ** 2, axis=-1, keepdims=True, dtype=np.float32))
    return (x / np.maximum(n, np.float32(eps))).astype(np.float32)


def mean_prototype(embeddings: Sequence[np.ndarray]) -> np.ndarray:
    """Per-identity gallery row: mean then /(‖mean‖+1e-8).
    inference/extract_embeddings.py:758-760; inference/recognition_engine.py:413-414."""
    mean_emb = np.mean(np.stack(embeddings, axis=0), axis=0)
    return mean_emb / (np.linalg.norm(mean_emb) + 1e-8)


def compute_prototypes(embeddings: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """inference/extract_embeddings.py:573-584 (prototype[label] = mean / (‖mean‖+1e-8))."""
    unique = np.unique(labels)
    out = np.zeros((len(unique), embeddings.shape[1]), np.float32)
    for lab in unique:
        p = embeddings[labels == lab].mean(axis=0)
        out[lab] = p / (np.linalg.norm(p) + 1e-8)
    return out


def batched_topk(E: np.ndarray, P: np.ndarray, k: int = 5):
    """Notebook form: S = np.dot(E, P.T); argmax; argsort[:, -k:]
    (notebooks/evaluate_arcface_kaggle.ipynb:618,713).  Returned best-first with ties
    resolved to the lowest gallery index (NumPy's quicksort order is unspecified)."""
    S = np.dot(np.asarray(E, np.float32), np.asarray(P, np.float32).T)
    n = S.shape[1]
    order = np.lexsort((np.broadcast_to(np.arange(n), S.shape), -S), axis=1)[:, :k]
    return np.take_along_axis(S, order, 1), order.astype(np.int64)


def batched_topk_fast(E: np.ndarray, P: np.ndarray, k: int = 5, threads: int = 0):
    """The same batched form with the CPU's best foot forward: np.dot (multi-threaded sgemm) + argpartition
    instead of the notebook's full argsort (notebooks/evaluate_arcface_kaggle.ipynb:713); the partition — which
    numpy runs on one thread — is spread over `threads` host threads by query row (numpy releases the GIL in it).
    Used only as the reported CPU baseline in bench.py; ties are not ordered."""
    S = np.dot(np.asarray(E, np.float32), np.asarray(P, np.float32).T)
    threads = threads or (os.cpu_count() or 1)

    def part_rows(lo_hi):
        lo, hi = lo_hi
        return np.argpartition(-S[lo:hi], k - 1, axis=1)[:, :k]

    n = S.shape[0]
    if threads > 1 and n >= 2 * threads:
        from concurrent.futures import ThreadPoolExecutor
        step = (n + threads - 1) // threads
        with ThreadPoolExecutor(threads) as pool:
            part = np.concatenate(list(pool.map(part_rows, [(lo, min(lo + step, n)) for lo in range(0, n, step)])), 0)
    else:
        part = part_rows((0, n))
    ps = np.take_along_axis(S, part, 1)
    order = np.argsort(-ps, axis=1)
    return np.take_along_axis(ps, order, 1), np.take_along_axis(part, order, 1).astype(np.int64)
