"""Evaluation consumers of the cosine path: the notebooks' batched top-1 / top-5 accuracy
(notebooks/evaluate_arcface_kaggle.ipynb:618,713) and inference/evaluate.py's threshold sweep (:61-128).

The Q x C score matrix the notebooks materialise (`np.dot(E, P.T)` then `argsort`) never exists here: the
fused kernel (frb_cosine_topk) returns the best k prototypes per embedding and everything below is bookkeeping
on [Q, k] arrays.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from . import ops


def identify_batch(embeddings: np.ndarray, prototypes: np.ndarray, k: int = 5, *, bf16: bool = False,
                   device: Optional[str] = None):
    """Best-k prototypes per embedding: (scores f32 [Q, k] descending, labels i64 [Q, k]); ties -> lowest label.
    fp32 (exact, the notebooks' arithmetic) or bf16 tensor cores (scores within 1e-3)."""
    dev = torch.device(device) if device else torch.device("cuda")
    q = torch.from_numpy(np.ascontiguousarray(embeddings, np.float32)).to(dev)
    g = torch.from_numpy(np.ascontiguousarray(prototypes, np.float32)).to(dev)
    if bf16:
        g = ops.normalize_rows(g, N.FRB_QNORM_NONE, torch.bfloat16)
    s, i = ops.cosine_topk(q, g, k, score_mode=N.FRB_SCORE_IP)
    return s.cpu().numpy(), i.cpu().numpy()


def topk_accuracy(embeddings: np.ndarray, labels: np.ndarray, prototypes: np.ndarray, ks: Sequence[int] = (1, 5), *,
                  bf16: bool = False, device: Optional[str] = None) -> Dict:
    """Top-k accuracy of nearest-prototype identification, plus the top-1 similarities / predictions the
    threshold sweep consumes."""
    kmax = max(ks)
    s, i = identify_batch(embeddings, prototypes, kmax, bf16=bf16, device=device)
    y = np.asarray(labels).reshape(-1, 1)
    hit = i == y
    out = {f"top{k}_accuracy": float(hit[:, :k].any(axis=1).mean()) if len(y) else 0.0 for k in ks}
    out.update(similarities=s[:, 0].copy(), predictions=i[:, 0].copy())
    return out


def threshold_sweep(similarities: np.ndarray, y_true: np.ndarray, y_pred_identities: np.ndarray,
                    thresholds: Optional[Sequence[float]] = None) -> Dict:
    """Same report as inference/evaluate.py:61-128: a prediction counts as known when its similarity is >= the
    threshold; accuracy and recall are correct / all samples, precision is correct / known."""
    if thresholds is None:
        thresholds = np.arange(0.3, 0.95, 0.05)
    sim, yt, yp = np.asarray(similarities), np.asarray(y_true), np.asarray(y_pred_identities)
    n = len(yt)
    results: List[Dict] = []
    for t in thresholds:
        known = (sim >= t) & (yp != -1)
        n_known = int(known.sum())
        correct = int(((yp == yt) & known).sum())
        acc = correct / n if (n and n_known) else 0.0
        prec = correct / n_known if n_known else 0.0
        rec = acc
        f1 = 2 * prec * rec / (prec + rec) if (prec + rec) > 0 else 0.0
        results.append({"threshold": float(t), "accuracy": float(acc), "precision": float(prec), "recall": float(rec),
                        "f1": float(f1), "known_ratio": float(n_known / n) if n else 0.0, "num_known": n_known,
                        "num_unknown": int(n - n_known)})
    bf = int(np.argmax([r["f1"] for r in results]))
    ba = int(np.argmax([r["accuracy"] for r in results]))
    return {"results": results, "best_f1_threshold": results[bf]["threshold"], "best_f1_score": results[bf]["f1"],
            "best_accuracy_threshold": results[ba]["threshold"], "best_accuracy_score": results[ba]["accuracy"]}
