"""B200 drop-in for the reference's cosine identification stage.

Mirrors inference/recognition_engine.py of sin0235/FaceRecognition: same class name, constructor
arguments, method names, return shapes, sentinel returns and threshold rule; the per-identity Python
loops are replaced by one fused similarity + top-k kernel call (libfrb200 frb_cosine_topk).

  cosine_similarity(a, b)                         inference/recognition_engine.py:41-63
  RecognitionEngine.recognize_with_db(emb)        :267-289   (dict DB, reference cosine rule, top-5)
  RecognitionEngine.recognize_with_faiss(emb, k)  :291-326   (flat inner-product index)
  RecognitionEngine.recognize / recognize_batch   :328-389
  RecognitionEngine.add_to_db / save_db / ...     :391-435
  match_facenet(db, embedding, threshold)         web_app.py:537-562 (inline FaceNet matcher)

Out of scope: detection, alignment and the embedding network.  `recognize()` therefore
takes its embeddings from an injected `embedder` callable (image -> float32[512] or None); with no
embedder it returns the same status='error' dict the reference returns when no checkpoint is loaded.

Batched entry points (new): recognize_embeddings(E), FlatIPIndex.search(E, k) — Q queries, one launch.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as N
from . import formats, ops


def _match_device(device: Optional[str]) -> torch.device:
    if device is not None and str(device).startswith("cuda"):
        return torch.device(device)
    return torch.device("cuda")


def _to_dev(a: np.ndarray, dev: torch.device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def _pad8(t: torch.Tensor) -> torch.Tensor:
    """Zero-pad the columns of [R, D] to a multiple of 8 (the kernels' vector width).  Zero columns change neither dot
    products nor norms, so any embedding dimension the reference accepts works here too."""
    d = t.shape[1]
    if d % 8 == 0:
        return t.contiguous()
    return torch.nn.functional.pad(t, (0, 8 - d % 8)).contiguous()


def _queries_to_dev(emb, dim: int, dev: torch.device) -> torch.Tensor:
    """Query embeddings as a padded fp32 CUDA matrix [Q, pad8(dim)]: numpy / lists are uploaded, CUDA tensors (the
    embedding network's output, inference/extract_embeddings.py:392-443) are taken where they are."""
    if isinstance(emb, torch.Tensor):
        q = emb.detach().to(device=dev, dtype=torch.float32).reshape(-1, dim)
    else:
        q = _to_dev(np.asarray(emb).astype(np.float32).reshape(-1, dim), dev)
    return _pad8(q)


def cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """inference/recognition_engine.py:41-63, evaluated by the CUDA kernel (Q = N = 1):
    0.0 if either norm is 0; raw dot if both norms are within 1e-3 of 1; else dot/(na*nb)."""
    dev = torch.device("cuda")
    qa = _pad8(_to_dev(np.asarray(a).astype(np.float32).reshape(1, -1), dev))
    gb = _pad8(_to_dev(np.asarray(b).astype(np.float32).reshape(1, -1), dev))
    s, _ = ops.cosine_topk(qa, gb, 1, score_mode=N.FRB_SCORE_REF_COSINE, q_norms=ops.row_norms(qa),
                           g_norms=ops.row_norms(gb))
    return float(s[0, 0].item())


class _GalleryDict(dict):
    """dict that counts mutations so the device copy can be refreshed lazily; the reference's callers
    read and write engine.db directly (web_app.py:545, recognition_engine.py:419)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.version = 0

    def _bump(self):
        self.version += 1

    def __setitem__(self, k, v):
        super().__setitem__(k, v); self._bump()

    def __delitem__(self, k):
        super().__delitem__(k); self._bump()

    def update(self, *a, **kw):
        super().update(*a, **kw); self._bump()

    def pop(self, *a):
        r = super().pop(*a); self._bump(); return r

    def popitem(self):
        r = super().popitem(); self._bump(); return r

    def clear(self):
        super().clear(); self._bump()

    def setdefault(self, k, d=None):
        r = super().setdefault(k, d); self._bump(); return r

    def __ior__(self, other):
        r = super().__ior__(other); self._bump(); return r

    def invalidate(self):
        """Call after editing a stored vector IN PLACE (engine.db[name][:] = v): nothing else can notice that."""
        self._bump()


class DeviceGallery:
    """Row-major fp32 [N, D] copy of a {name: vector} dict on the GPU, with per-row norms.
    Row i <-> i-th dict key (insertion order), which is what makes ties resolve as the reference's
    stable sort does."""

    def __init__(self, db: Dict[str, np.ndarray], device: torch.device, bounds: Optional[Tuple[int, int]] = None):
        """bounds = (lo, hi): keep only rows [lo, hi) on this device (one shard of an identity-sharded gallery); `names`
        always lists every identity and the searches report GLOBAL row ids (idx_base = lo)."""
        self.names: List[str] = list(db.keys())
        self._names_obj: Optional[np.ndarray] = None      # object array of the names, built on the first batched result
        self.lo, self.hi = bounds if bounds is not None else (0, len(self.names))
        mine = self.names[self.lo:self.hi]
        if self.names:
            self.dim = int(np.asarray(db[self.names[0]]).size)
        else:
            self.dim = 8
        if mine:
            mat = np.stack([np.asarray(db[n]).astype(np.float32).flatten() for n in mine], 0)
        else:
            mat = np.zeros((0, self.dim), np.float32)
        self.rows = _pad8(_to_dev(mat, device))               # [n, pad8(dim)]
        self.norms = ops.row_norms(self.rows) if len(mine) else torch.zeros(0, device=device)
        self._unit_rows = None
        self._unit_f16 = None

    def unit_rows_f16(self) -> torch.Tensor:
        """Unit-norm fp16 copy of the rows: the tensor-core first pass of ops.cosine_topk_exact."""
        if self._unit_f16 is None:
            self._unit_f16 = ops.normalize_rows(self.rows, N.FRB_QNORM_CLAMP, torch.float16)
        return self._unit_f16

    def unit_rows(self) -> torch.Tensor:
        """rows / (||row|| + 1e-8) — web_app.py:549 normalises every db row this way."""
        if self._unit_rows is None:
            self._unit_rows = ops.normalize_rows(self.rows, N.FRB_QNORM_EPS)
        return self._unit_rows


class FlatIPIndex:
    """GPU stand-in for faiss.IndexFlatIP as the reference uses it (inference/extract_embeddings.py:628-635,
    inference/recognition_engine.py:304): .d, .ntotal, .add(x), .search(x, k) -> (scores, ids)."""

    def __init__(self, d: int, device: Optional[str] = None, dtype: torch.dtype = torch.float32):
        self.d = int(d)
        self.dp = (self.d + 7) // 8 * 8            # stored row length: zero-padded to the kernels' vector width
        self.device = _match_device(device)
        self.dtype = dtype
        self.rows = torch.zeros((0, self.dp), dtype=dtype, device=self.device)

    @property
    def ntotal(self) -> int:
        return int(self.rows.shape[0])

    def add(self, x: np.ndarray) -> None:
        x = _pad8(_to_dev(np.asarray(x).reshape(-1, self.d), self.device))
        if self.dtype == torch.bfloat16:
            x = ops.normalize_rows(x, N.FRB_QNORM_NONE, torch.bfloat16)
        self.rows = torch.cat([self.rows, x], 0).contiguous()

    def search_device(self, x: torch.Tensor, k: int, qnorm_mode: int = N.FRB_QNORM_NONE):
        return ops.cosine_topk(_pad8(x), self.rows, k, score_mode=N.FRB_SCORE_IP, qnorm_mode=qnorm_mode)

    def search(self, x, k: int):
        """faiss's search(x, k) -> (scores, ids) on the host; x may also be a CUDA tensor [Q, d]."""
        s, i = self.search_device(_queries_to_dev(x, self.d, self.device), k)
        return s.cpu().numpy(), i.cpu().numpy()

    @classmethod
    def from_file(cls, path: str, device: Optional[str] = None) -> "FlatIPIndex":
        rows = formats.read_faiss_flat_ip(path)
        idx = cls(rows.shape[1], device)
        idx.add(rows)
        return idx

    def write(self, path: str) -> None:
        formats.write_faiss_flat_ip(path, self.rows[:, :self.d].float().cpu().numpy())


def group_plan(labels: np.ndarray, n_groups: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Host-side index plumbing for frb_group_mean_renorm: (order i64 [M], offsets i64 [G + 1]) with the samples of
    each label in ascending sample order (what `embeddings[labels == label]` yields in the reference)."""
    lab = np.asarray(labels).astype(np.int64).reshape(-1)
    g = int(n_groups if n_groups is not None else (lab.max() + 1 if lab.size else 0))
    if lab.size and (lab.min() < 0 or lab.max() >= g):
        raise IndexError(f"label out of range 0..{g - 1}")      # the reference's prototypes[label] = ... raises the same way
    order = np.argsort(lab, kind="stable").astype(np.int64)
    offsets = np.zeros(g + 1, np.int64)
    np.cumsum(np.bincount(lab, minlength=g), out=offsets[1:])
    return order, offsets


def compute_prototypes(embeddings: np.ndarray, labels: np.ndarray, output_path: str = None, device: Optional[str] = None
                       ) -> np.ndarray:
    """inference/extract_embeddings.py:555-592: prototype[label] = mean(embeddings[labels == label]) / (||mean|| + 1e-8),
    num_classes = len(np.unique(labels)) rows.  The means and norms run on the GPU (frb_group_mean_renorm)."""
    print("\n=== COMPUTING PROTOTYPES ===")
    emb = np.ascontiguousarray(embeddings, np.float32)
    num_classes = len(np.unique(labels))
    order, offsets = group_plan(labels, num_classes)
    dev = _match_device(device)
    out, _ = ops.group_mean_renorm(_to_dev(emb, dev), torch.from_numpy(order).to(dev), torch.from_numpy(offsets).to(dev))
    prototypes = out.cpu().numpy()
    print(f"Computed {num_classes} prototypes")
    if output_path:
        np.save(output_path, prototypes)
        print(f"Saved prototypes: {output_path}")
    return prototypes


def mean_embedding(embeddings: Sequence[np.ndarray], device: Optional[str] = None) -> np.ndarray:
    """One identity's gallery row: mean of its embeddings, / (||mean|| + 1e-8) (extract_embeddings.py:758-760,
    recognition_engine.py:413-414)."""
    emb = np.ascontiguousarray(np.stack(embeddings), np.float32)
    dev = _match_device(device)
    order = torch.arange(emb.shape[0], dtype=torch.int64, device=dev)
    offsets = torch.tensor([0, emb.shape[0]], dtype=torch.int64, device=dev)
    out, _ = ops.group_mean_renorm(_to_dev(emb, dev), order, offsets)
    return out[0].cpu().numpy()


def build_db_from_embeddings(names: Sequence[str], embeddings: np.ndarray, owner: np.ndarray, device: Optional[str] = None
                             ) -> Dict[str, np.ndarray]:
    """The gallery dict build_db emits (extract_embeddings.py:808-831), from already-extracted embeddings:
    `owner[i]` = index into `names` of sample i; identities with no sample are left out, as build_db skips them."""
    emb = np.ascontiguousarray(embeddings, np.float32)
    order, offsets = group_plan(owner, len(names))
    dev = _match_device(device)
    out, _ = ops.group_mean_renorm(_to_dev(emb, dev), torch.from_numpy(order).to(dev), torch.from_numpy(offsets).to(dev))
    rows = out.cpu().numpy()
    return {n: rows[i] for i, n in enumerate(names) if offsets[i + 1] > offsets[i]}


def build_faiss_index(embeddings: np.ndarray, output_path: str = None, use_gpu: bool = True,
                      device: Optional[str] = None) -> FlatIPIndex:
    """inference/extract_embeddings.py:595-645: rows / (||row|| + 1e-8), IndexFlatIP(dim).add, optional write."""
    dev = _match_device(device)
    e = _to_dev(np.asarray(embeddings).astype("float32"), dev)
    index = FlatIPIndex(e.shape[1], str(dev))
    index.rows = ops.normalize_rows(_pad8(e), N.FRB_QNORM_EPS)
    if output_path:
        index.write(output_path)
    return index


class RecognitionEngine:
    """Same surface as the reference's RecognitionEngine (inference/recognition_engine.py:66-435)."""

    def __init__(
        self,
        model_path: str = "models/checkpoints/arcface/arcface_best.pth",
        db_path: str = None,
        faiss_index_path: str = None,
        prototypes_path: str = None,
        label_mapping_path: str = None,
        device: str = None,
        threshold: float = 0.5,
        use_face_detection: bool = True,
        embedder: Optional[Callable[[object], Optional[np.ndarray]]] = None,
        group=None,
    ):
        """`group` (new; a torch.distributed process group, or True for the default group): the dict gallery is
        sharded by identity over the group's ranks — rank r keeps rows shard_bounds(len(db), R, r) on its GPU, every
        rank passes the same queries, each searches its shard and ONE fused NVLink exchange + merge kernel gives every
        rank the same answer as a single-GPU engine (ties -> lowest row).  All ranks must make the same calls."""
        self.group = group
        self._sharded = None
        self.device = device or "cuda"
        self.match_device = _match_device(device)
        self.threshold = threshold
        self.use_face_detection = use_face_detection
        self.model_path = model_path
        self.model = None           # the embedding network is out of scope; see `embedder`
        self.embedder = embedder
        self._db: Optional[_GalleryDict] = None
        self._gallery: Optional[DeviceGallery] = None
        self._gallery_version = -1
        self.faiss_index: Optional[FlatIPIndex] = None
        self.prototypes = None
        self.label_to_id = None
        self.id_to_label = None

        if db_path and os.path.exists(db_path):
            self.db = formats.load_embedding_db(db_path)
            print(f"Loaded database: {len(self.db)} identities")
        if faiss_index_path and os.path.exists(faiss_index_path):
            self._load_faiss(faiss_index_path, prototypes_path, label_mapping_path)

    # self.db stays a plain-looking dict for callers; assignments are wrapped so edits are noticed
    @property
    def db(self):
        return self._db

    @db.setter
    def db(self, value):
        if value is None:
            self._db = None
        elif isinstance(value, _GalleryDict):
            self._db = value
        else:
            self._db = _GalleryDict(value)
        self._gallery = None
        self._gallery_version = -1

    def _load_faiss(self, index_path: str, prototypes_path: str = None, mapping_path: str = None):
        try:
            self.faiss_index = FlatIPIndex.from_file(index_path, str(self.match_device))
            print(f"Loaded FAISS index: {self.faiss_index.ntotal} vectors")
        except Exception as e:  # same convention as the reference: report and carry on without an index
            print(f"Loi load FAISS: {e}")
            return
        if prototypes_path and os.path.exists(prototypes_path):
            self.prototypes = np.load(prototypes_path)
            print(f"Loaded prototypes: {self.prototypes.shape}")
        if mapping_path and os.path.exists(mapping_path):
            self.label_to_id, self.id_to_label = formats.load_label_mapping(mapping_path)
            print(f"Loaded label mapping: {len(self.label_to_id)} classes")

    def set_threshold(self, threshold: float):
        self.threshold = threshold

    # ---- gallery on the device ---------------------------------------------------------------
    def _dist_group(self):
        import torch.distributed as dist
        if self.group is None or not (dist.is_available() and dist.is_initialized()):
            return None
        return None if self.group is True else self.group

    def _world_rank(self) -> Tuple[int, int]:
        import torch.distributed as dist
        if self.group is None or not (dist.is_available() and dist.is_initialized()):
            return 1, 0
        g = self._dist_group()
        return dist.get_world_size(g), dist.get_rank(g)

    def gallery(self) -> DeviceGallery:
        if self._gallery is None or self._gallery_version != self._db.version:
            world, rank = self._world_rank()
            bounds = None
            if world > 1:
                from .sharded import shard_bounds
                bounds = shard_bounds(len(self._db), world, rank)
            self._gallery = DeviceGallery(self._db, self.match_device, bounds)
            self._gallery_version = self._db.version
            self._sharded = None
        return self._gallery

    def refresh_gallery(self) -> None:
        """Force the device copy to be rebuilt from self.db (after in-place edits of stored vectors)."""
        self._gallery = None

    # ---- matching -----------------------------------------------------------------------------
    def _local_topk(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(scores, global rows) of this device's rows for padded fp32 CUDA queries, the reference's cosine rule."""
        g = self.gallery()
        qn = ops.row_norms(q)
        n_local, pdim = g.rows.shape[0], g.rows.shape[1]
        if ops.refine_applicable(q.shape[0], n_local, pdim, k):
            # batches: fp16 tensor-core first pass + exact fp32 re-score of the candidates, each list proven complete;
            # the few queries without a provable margin are re-run on the fp32 kernels
            return ops.cosine_topk_exact(q, g.rows, g.unit_rows_f16(), k, q_norms=qn, g_norms=g.norms, idx_base=g.lo)
        return ops.cosine_topk(q, g.rows, k, score_mode=N.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=g.norms, idx_base=g.lo)

    def recognize_embeddings_device(self, embeddings, k: int = 5) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-level batched match: embeddings (numpy [Q, D] or a CUDA tensor, e.g. straight from the embedding
        network) -> (scores fp32 [Q, k], gallery rows int64 [Q, k]) as CUDA tensors, no host synchronisation on the
        single-kernel paths.  Row r is the r-th key of self.db."""
        g = self.gallery()
        q = _queries_to_dev(embeddings, g.dim, self.match_device)
        world, _ = self._world_rank()
        if world > 1:
            if self._sharded is None:
                self._sharded = self._make_sharded()
            return self._sharded.search(q, k)
        return self._local_topk(q, k)

    def _make_sharded(self):
        """The cross-rank step behind a sharded engine: local top-k with global ids -> fused NVLink exchange + merge."""
        from .sharded import ShardedSearch
        return ShardedSearch(self._local_topk, ops.topk_merge, True, self._dist_group(), merge_packed=ops.topk_merge_packed,
                             peer_exchange=True)

    def _db_topk(self, emb, k: int):
        """(scores [Q, k], rows [Q, k]) on the host for Q query embeddings under the reference's cosine rule."""
        s, i = self.recognize_embeddings_device(emb, k)
        return s.cpu().numpy(), i.cpu().numpy(), self.gallery().names

    def _format_db_results(self, scores: np.ndarray, rows: np.ndarray, names) -> List[Tuple[str, float, List[Tuple[str, float]]]]:
        """Result tuples of recognize_with_db for Q rows of (scores, gallery rows).  One bulk conversion to Python
        floats / ints (float64 of the fp32 value == float(np.float32)), then plain tuple building: 3x faster than
        walking numpy scalars, and the formatting is half of a 256-query call."""
        out = []
        g = self._gallery
        if rows.shape[0] > 4 and rows.size and int(rows.min()) >= 0 and g is not None and names is g.names:
            # whole lists (a gallery of at least 5 rows): the names of all rows in one object-array gather
            if g._names_obj is None:
                g._names_obj = np.array(names, dtype=object)
            for nrow, srow in zip(g._names_obj[rows].tolist(), scores.astype(np.float64).tolist()):
                out.append(("Unknown" if srow[0] < self.threshold else nrow[0], srow[0], list(zip(nrow, srow))))
            return out
        for srow, irow in zip(scores.astype(np.float64).tolist(), rows.tolist()):
            top = [(names[j], s) for s, j in zip(srow, irow) if j >= 0]
            best_name, best_score = top[0]  # IndexError on an empty dict, as in the reference (:284)
            out.append(("Unknown" if best_score < self.threshold else best_name, best_score, top))
        return out

    def recognize_with_db(self, embedding: np.ndarray) -> Tuple[str, float, List[Tuple[str, float]]]:
        """(best_name, best_score, top-5) — inference/recognition_engine.py:267-289."""
        if self.db is None:
            return "No database", 0.0, []
        s, i, names = self._db_topk(embedding, 5)
        return self._format_db_results(s[:1], i[:1], names)[0]

    def recognize_embeddings(self, embeddings) -> List[Tuple[str, float, List[Tuple[str, float]]]]:
        """Batched recognize_with_db: Q embeddings [Q, D] (numpy, or a CUDA tensor) -> Q result tuples from ONE fused
        similarity + top-k call."""
        if self.db is None:
            return [("No database", 0.0, [])] * len(embeddings)
        s, i, names = self._db_topk(embeddings, 5)
        return self._format_db_results(s, i, names)

    def recognize_with_faiss(self, embedding: np.ndarray, k: int = 5) -> Tuple[str, float, List[Tuple[str, float]]]:
        """inference/recognition_engine.py:291-326: e/(||e||+1e-8), IndexFlatIP.search, strict '<' threshold."""
        if self.faiss_index is None:
            return "No FAISS index", 0.0, []
        return self._faiss_results(np.asarray(embedding).astype(np.float32).reshape(1, -1), k)[0]

    def _faiss_results(self, embeddings, k: int) -> List[Tuple[str, float, List[Tuple[str, float]]]]:
        """recognize_with_faiss for Q embeddings in one search call."""
        q = _queries_to_dev(embeddings, self.faiss_index.d, self.match_device)
        scores, indices = self.faiss_index.search_device(q, k, qnorm_mode=N.FRB_QNORM_EPS)
        scores, indices = scores.cpu().numpy().astype(np.float64).tolist(), indices.cpu().numpy().tolist()   # bulk -> Python scalars
        return [self._format_faiss_result(irow, srow) for irow, srow in zip(indices, scores)]

    def _format_faiss_result(self, indices, scores):
        results = []
        for idx, score in zip(indices, scores):
            if idx == -1:
                continue
            if self.id_to_label:
                name = self.id_to_label.get(idx, f"ID_{idx}")
            else:
                name = f"ID_{idx}"
            results.append((name, float(score)))
        if len(results) == 0:
            return "Unknown", 0.0, []
        best_name, best_score = results[0]
        if best_score < self.threshold:
            return "Unknown", best_score, results
        return best_name, best_score, results

    # ---- image-level entry points ---------------------------------------------------------------
    def extract_embedding(self, img_input) -> Optional[np.ndarray]:
        """Detection/alignment/embedding are upstream of the hot path; delegate to the injected embedder."""
        if self.embedder is None:
            print("Model chua duoc load")
            return None
        return self.embedder(img_input)

    def recognize(self, img_input, use_faiss: bool = None, k: int = 5) -> Dict:
        """inference/recognition_engine.py:328-381 — same result dict and error statuses."""
        result = {"identity": "Unknown", "confidence": 0.0, "top_k": [], "embedding": None, "status": "success"}
        embedding = self.extract_embedding(img_input)
        if embedding is None:
            result["status"] = "error"
            result["message"] = "Cannot extract embedding (no face or invalid image)"
            return result
        result["embedding"] = embedding
        if use_faiss is None:
            use_faiss = self.faiss_index is not None
        if use_faiss and self.faiss_index is not None:
            identity, confidence, top_k = self.recognize_with_faiss(embedding, k)
        elif self.db is not None:
            identity, confidence, top_k = self.recognize_with_db(embedding)
        else:
            result["status"] = "error"
            result["message"] = "No database loaded"
            return result
        result["identity"], result["confidence"], result["top_k"] = identity, confidence, top_k
        return result

    def recognize_batch(self, img_inputs: Sequence, use_faiss: bool = None) -> List[Dict]:
        """inference/recognition_engine.py:383-389 — same list of result dicts as calling recognize() per image; the
        embeddings are extracted per image (upstream of this path) and then matched in ONE batched call."""
        results = []
        for img in img_inputs:
            r = {"identity": "Unknown", "confidence": 0.0, "top_k": [], "embedding": None, "status": "success"}
            r["embedding"] = self.extract_embedding(img)
            if r["embedding"] is None:
                r["status"] = "error"
                r["message"] = "Cannot extract embedding (no face or invalid image)"
            results.append(r)
        live = [r for r in results if r["status"] == "success"]
        if use_faiss is None:
            use_faiss = self.faiss_index is not None
        if use_faiss and self.faiss_index is not None:
            matches = self._faiss_results(np.stack([np.asarray(r["embedding"], np.float32).reshape(-1) for r in live]), 5) if live else []
        elif self.db is not None:
            matches = self.recognize_embeddings(np.stack([np.asarray(r["embedding"], np.float32).reshape(-1) for r in live])) if live else []
        else:
            for r in live:
                r["status"] = "error"
                r["message"] = "No database loaded"
            return results
        for r, (identity, confidence, top_k) in zip(live, matches):
            r["identity"], r["confidence"], r["top_k"] = identity, confidence, top_k
        return results

    def add_to_db(self, name: str, img_inputs: Sequence) -> bool:
        """inference/recognition_engine.py:391-422: mean of the embeddings, / (||mean|| + 1e-8)."""
        embeddings = [e for e in (self.extract_embedding(img) for img in img_inputs) if e is not None]
        if len(embeddings) == 0:
            print(f"Khong the extract embedding cho {name}")
            return False
        mean_emb = mean_embedding(embeddings, str(self.match_device))
        if self.db is None:
            self.db = {}
        self.db[name] = mean_emb
        print(f"Added {name} to database (from {len(embeddings)} images)")
        return True

    def save_db(self, path: str):
        if self.db:
            formats.save_embedding_db(path, dict(self.db))
            print(f"Saved database: {path}")

    def get_db_identities(self) -> List[str]:
        if self.db:
            return list(self.db.keys())
        return []


def create_engine_from_embeddings_dir(model_path: str, embeddings_dir: str, threshold: float = 0.5,
                                      device: str = None, embedder=None) -> RecognitionEngine:
    """inference/recognition_engine.py:438-464."""
    faiss_path = os.path.join(embeddings_dir, "arcface_index.faiss")
    prototypes_path = os.path.join(embeddings_dir, "arcface_prototypes.npy")
    mapping_path = os.path.join(embeddings_dir, "label_mapping.npy")
    return RecognitionEngine(
        model_path=model_path,
        faiss_index_path=faiss_path if os.path.exists(faiss_path) else None,
        prototypes_path=prototypes_path if os.path.exists(prototypes_path) else None,
        label_mapping_path=mapping_path if os.path.exists(mapping_path) else None,
        threshold=threshold, device=device, embedder=embedder)


def match_facenet(db: Dict[str, np.ndarray], embedding: np.ndarray, threshold: float = 0.5,
                  gallery: Optional[DeviceGallery] = None, device: Optional[str] = None):
    """The inline FaceNet matcher of web_app.py:537-562 as a function: e/=(||e||+1e-8); per db row
    d/=(||d||+1e-8), score=e.d, distance=||e-d||; sorted desc; strict '<' threshold; top-5 of
    (name, score, distance).
    Returns {"identity", "confidence", "distance", "top_k"}."""
    dev = _match_device(device)
    g = gallery or DeviceGallery(db, dev)
    q = _pad8(_to_dev(np.asarray(embedding).astype(np.float32).reshape(1, -1), dev))
    qn = ops.normalize_rows(q, N.FRB_QNORM_EPS)
    s, i = ops.cosine_topk(qn, g.unit_rows(), 5, score_mode=N.FRB_SCORE_IP)
    s, i = s.cpu().numpy()[0], i.cpu().numpy()[0]
    keep = [int(j) for j in i if j >= 0]
    # the 5 winning rows come back to the host so the displayed L2 distance is the reference's own
    # float32 expression ||e - d|| (web_app.py:552) rather than a cancellation-prone sqrt(2 - 2s)
    e = qn.cpu().numpy()[0]
    rows = g.unit_rows()[torch.tensor(keep, dtype=torch.int64, device=dev)].cpu().numpy() if keep else np.zeros((0, e.shape[0]))
    top_k = [(g.names[j], float(sc), float(np.linalg.norm(e - rows[r]))) for r, (sc, j) in enumerate(zip(s, keep))]
    best_name, best_score, best_distance = top_k[0]
    if best_score < threshold:
        best_name = "Unknown"
    return {"identity": best_name, "confidence": best_score, "distance": best_distance, "top_k": top_k}
