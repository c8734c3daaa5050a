"""LBPH identification on B200: a drop-in for the cv2.face.LBPHFaceRecognizer object protocol the
reference uses, plus mirrors of the reference's thin wrappers.

Reference call sites (paths relative to the reference root):
  models/lbphmodel/train_lbph.py:4-36       train_lbph_model(faces, labels, radius, neighbors, grid_x, grid_y)
  models/lbphmodel/inference_lbph.py:4-18   recognize_face(model, face_img, threshold)
  models/lbphmodel/evaluate_lbph.py:4-45    evaluate_lbph(model, faces, labels, threshold)
  models/lbphmodel/threshold_lbph.py:7-96   find_optimal_threshold(model, faces, labels, min_coverage, threshold_range)
  web_app.py:245-246, 587                   LBPHFaceRecognizer_create(); model.read(path); model.predict(img)
  models/lbphmodel/train_lbph_script.py:222 model.save(path)

The arithmetic (OpenCV-contrib lbph_faces.cpp: elbp + spatial_histogram + compareHist(CHISQR_ALT) scan)
runs in libfrb200's CUDA kernels (frb_lbp_hist_u8, frb_chisq_topk).  The gallery lives on the GPU as
integer u16 cell histograms (32 KiB per face); OpenCV's float32 view is count * float32(1/cell_px).
There is no CPU path.
"""
from __future__ import annotations

import sys
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

DBL_MAX = sys.float_info.max

try:  # raise the exception type callers of cv2.face already catch
    import cv2 as _cv2

    LBPHError = _cv2.error
except Exception:  # pragma: no cover - cv2 is part of the image
    _cv2 = None

    class LBPHError(RuntimeError):
        pass


def _as_gray_u8(img) -> np.ndarray:
    a = np.asarray(img)
    if a.ndim != 2:
        raise LBPHError(f"LBPH expects single-channel 2-D images, got shape {a.shape}")
    if a.dtype != np.uint8:
        raise LBPHError(f"LBPH on B200 supports uint8 images only (the reference always passes uint8), got {a.dtype}")
    return np.ascontiguousarray(a)


def _cat_rows(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Row-wise concatenation of two count matrices (u8, or u16 through its int16 view; a u8 / u16 mix widens to u16)."""
    if a.dtype != b.dtype:
        a, b = (t if t.dtype == torch.uint16 else t.to(torch.int16).view(torch.uint16) for t in (a, b))
    if a.dtype == torch.uint16:
        return torch.cat([a.view(torch.int16), b.view(torch.int16)], 0).view(torch.uint16)
    return torch.cat([a, b], 0)


class _Group:
    """Gallery rows that share a cell size (same image shape class)."""

    __slots__ = ("cell_px", "hist", "rows")

    def __init__(self, cell_px: int, hist: torch.Tensor, rows: torch.Tensor):
        # hist [n, L]: u8 counts when cell_px <= 255 (ops.compact_histograms: half the bytes the scan streams), else u16;
        # rows int64 [n] global row ids
        self.cell_px, self.hist, self.rows = cell_px, hist, rows


class LBPHFaceRecognizer:
    """cv2.face.LBPHFaceRecognizer protocol: train / update / predict / save / read / setThreshold ..."""

    def __init__(self, radius: int = 1, neighbors: int = 8, grid_x: int = 8, grid_y: int = 8,
                 threshold: float = DBL_MAX, device: Optional[str] = None, group=None):
        """`group` (new; a torch.distributed process group, or True for the default group): the gallery is sharded by
        training sample over the group's ranks.  Every rank makes the SAME train / update / predict calls with the same
        data; rank r keeps the histograms of its slice of each train()/update() batch on its GPU (global row ids), every
        rank extracts the query histograms, searches its shard, and ONE fused NVLink exchange + merge kernel gives all
        ranks the answer a single-GPU model gives (first row wins ties)."""
        self._radius, self._neighbors, self._grid_x, self._grid_y = int(radius), int(neighbors), int(grid_x), int(grid_y)
        self._threshold = float(threshold)
        self.device = torch.device(device or "cuda")
        if self.device.type != "cuda":
            raise ValueError("LBPHFaceRecognizer runs on CUDA only (no CPU fallback)")
        self.group = group
        self._sharded = None
        self._groups: List[_Group] = []
        self._labels = np.zeros((0,), np.int32)
        self._label_info: Dict[int, str] = {}

    # ---- parameters (cv2 getters/setters) ---------------------------------------------------
    def getRadius(self): return self._radius
    def getNeighbors(self): return self._neighbors
    def getGridX(self): return self._grid_x
    def getGridY(self): return self._grid_y
    def getThreshold(self): return self._threshold
    def setThreshold(self, val: float): self._threshold = float(val)
    def setRadius(self, v: int): self._radius = int(v)
    def setNeighbors(self, v: int): self._neighbors = int(v)
    def setGridX(self, v: int): self._grid_x = int(v)
    def setGridY(self, v: int): self._grid_y = int(v)
    def empty(self): return self.size == 0
    def setLabelInfo(self, label: int, text: str): self._label_info[int(label)] = str(text)
    def getLabelInfo(self, label: int): return self._label_info.get(int(label), "")

    @property
    def size(self) -> int:
        return int(self._labels.shape[0])

    @property
    def hist_len(self) -> int:
        return self._grid_x * self._grid_y * (1 << self._neighbors)

    # ---- feature extraction -------------------------------------------------------------------
    def compute_histograms(self, images: torch.Tensor, counts8: bool = False) -> Tuple[torch.Tensor, int]:
        """u8 CUDA tensor [B, H, W] -> (u16 [B, L], cell_px) via frb_lbp_hist_u8; counts8: the gallery form (u8 counts
        straight from the kernel when a cell has <= 255 pixels)."""
        return ops.lbp_hist(images, self._radius, self._neighbors, self._grid_x, self._grid_y, counts8=counts8)

    def _upload(self, imgs: Sequence[np.ndarray]) -> torch.Tensor:
        stack = np.stack(imgs, 0)
        return torch.from_numpy(stack).to(self.device, non_blocking=False)

    def _hist_by_shape(self, faces: Sequence[np.ndarray], counts8: bool = False):
        """Yield (positions, hist u16 (or u8 with counts8) [n, L], cell_px) per distinct image shape, positions ascending."""
        by_shape: Dict[Tuple[int, int], List[int]] = {}
        for i, f in enumerate(faces):
            by_shape.setdefault(f.shape, []).append(i)
        for shape, pos in by_shape.items():
            hist, px = self.compute_histograms(self._upload([faces[i] for i in pos]), counts8)
            yield pos, hist, px

    # ---- train / update (LBPH::train, LBPH::update) ------------------------------------------
    def train(self, src: Iterable, labels) -> None:
        self._groups = []
        self._labels = np.zeros((0,), np.int32)
        self._add(src, labels, "train")

    def update(self, src: Iterable, labels) -> None:
        self._add(src, labels, "update")

    def _add(self, src, labels, what: str) -> None:
        faces = [_as_gray_u8(f) for f in src]
        lab = np.asarray(labels).astype(np.int32).reshape(-1)
        if len(faces) == 0:
            raise LBPHError(f"Empty training data was given. You'll need more than one sample to learn a model. ({what})")
        if lab.shape[0] != len(faces):
            raise LBPHError(f"The number of samples (src) must equal the number of labels (labels). "
                            f"Was len(samples)={len(faces)}, len(labels)={lab.shape[0]}.")
        base = self.size
        world, rank = self._world_rank()
        keep = None
        if world > 1:                                   # this rank's slice of the batch; labels stay global
            from .sharded import shard_bounds
            lo, hi = shard_bounds(len(faces), world, rank)
            keep = range(lo, hi)
        mine = faces if keep is None else [faces[i] for i in keep]
        for pos, hist, px in (self._hist_by_shape(mine, counts8=True) if mine else ()):   # u8 counts straight from K2 when they fit
            if keep is not None:
                pos = [keep[i] for i in pos]
            rows = torch.tensor([base + i for i in pos], dtype=torch.int64, device=self.device)
            for g in self._groups:
                if g.cell_px == px:
                    g.hist = _cat_rows(g.hist, hist)
                    g.rows = torch.cat([g.rows, rows], 0)
                    break
            else:
                self._groups.append(_Group(px, hist, rows))
        self._labels = np.concatenate([self._labels, lab])

    def set_gallery(self, hist_u16: torch.Tensor, cell_px: int, labels, row_offset: int = 0) -> None:
        """Adopt precomputed integer histograms: the whole gallery, or — with `group` — THIS rank's shard, whose first
        row has global id `row_offset`; `labels` always lists every gallery row of the whole model."""
        assert hist_u16.dtype in (torch.uint16, torch.uint8) and hist_u16.dim() == 2 and hist_u16.shape[1] == self.hist_len
        n = hist_u16.shape[0]
        hist = hist_u16.contiguous() if hist_u16.dtype == torch.uint8 else ops.compact_histograms(hist_u16.contiguous(), int(cell_px))
        if n and hist.dtype == torch.uint8 and int(hist.max()) > int(cell_px):
            # the tensor-core filter's tables (and its error bound) cover counts 0..cell_px, all a cell can hold
            raise ValueError(f"histogram counts up to {int(hist.max())} with cell_px = {int(cell_px)}: a cell cannot hold more than cell_px codes")
        rows = torch.arange(row_offset, row_offset + n, dtype=torch.int64, device=hist_u16.device)
        self._groups = [_Group(int(cell_px), hist, rows)] if n else []
        self._labels = np.asarray(labels).astype(np.int32).reshape(-1)
        assert self._world_rank()[0] > 1 or self._labels.shape[0] == n
        self._sharded = None

    # ---- predict (LBPH::predict + StandardCollector) -----------------------------------------
    def _dist_group(self):
        return None if self.group is True else self.group

    def _world_rank(self) -> Tuple[int, int]:
        import torch.distributed as dist
        if self.group is None or not (dist.is_available() and dist.is_initialized()):
            return 1, 0
        g = self._dist_group()
        return dist.get_world_size(g), dist.get_rank(g)

    def _search(self, q_hist: torch.Tensor, q_px: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(dist fp32 [Q, k], global row int64 [Q, k]) over the whole gallery (all ranks' shards); ties -> lowest row."""
        if self._world_rank()[0] > 1:
            if self._sharded is None:
                self._sharded = self._make_sharded()
            self._sharded.local_search = lambda qh, kk: self._search_local(qh, q_px, kk)
            return self._sharded.search(q_hist, k)
        return self._search_local(q_hist, q_px, k)

    def _make_sharded(self):
        """The cross-rank step behind a sharded model: local nearest rows with global ids -> fused NVLink exchange + merge."""
        from .sharded import ShardedSearch
        return ShardedSearch(None, ops.topk_merge, False, self._dist_group(), merge_packed=ops.topk_merge_packed, peer_exchange=True)

    def _search_local(self, q_hist: torch.Tensor, q_px: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """The same over THIS device's gallery groups."""
        outs = []
        for g in self._groups:
            d, i = ops.chisq_topk(q_hist, q_px, g.hist, g.cell_px, k)
            if len(self._groups) > 1 or g.rows.shape[0] != self.size:
                ops.index_remap(i, g.rows)                               # local group rows -> global row ids
            outs.append((d, i))
        if not outs:                                                     # a rank whose shard is empty
            n = q_hist.shape[0]
            return (torch.full((n, k), float("inf"), dtype=torch.float32, device=self.device),
                    torch.full((n, k), -1, dtype=torch.int64, device=self.device))
        if len(outs) == 1:
            return outs[0]
        cd = torch.stack([o[0] for o in outs], 0).contiguous()
        ci = torch.stack([o[1] for o in outs], 0).contiguous()
        return ops.topk_merge(cd, ci, largest=False)

    def predict_device(self, images: torch.Tensor, k: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
        """u8 CUDA [Q, H, W] -> (dist fp32 [Q, k], gallery row int64 [Q, k]); no host synchronisation."""
        if self.size == 0:
            raise LBPHError("This LBPH model is not computed yet. Did you call the train method?")
        q_hist, q_px = self.compute_histograms(images)
        return self._search(q_hist, q_px, k)

    def predict_device_bgr(self, frames_bgr: torch.Tensor, k: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
        """Video-batch front end: u8 CUDA [Q, H, W, 3] BGR crops -> gray on the device (frb_bgr2gray_u8, the
        cv2.cvtColor step of web_app.py:475) -> predict_device."""
        return self.predict_device(ops.bgr_to_gray(frames_bgr), k)

    def predict_device_frames(self, frames_bgr: torch.Tensor, target_size: Tuple[int, int] = (100, 100), k: int = 1
                              ) -> Tuple[torch.Tensor, torch.Tensor]:
        """The whole no-detector front end of web_app.py:484-486 / train_lbph_script.py:69-72 on the device:
        u8 CUDA [Q, H, W, 3] BGR frames of any size -> cv2.resize(frame, target_size) + BGR2GRAY in one kernel
        (frb_resize_linear_u8, to_gray) -> predict_device.  target_size = (cols, rows) as in cv2."""
        return self.predict_device(preprocess_frames_device(frames_bgr, target_size), k)

    def predict_batch(self, images) -> Tuple[np.ndarray, np.ndarray]:
        """Batched predict: (labels int32 [Q], distances float64 [Q]); (-1, DBL_MAX) where the model
        threshold rejects the best match (dist >= threshold), as StandardCollector does."""
        if isinstance(images, torch.Tensor):
            dist, idx = self.predict_device(images, 1)
        else:
            faces = [_as_gray_u8(f) for f in images]
            if self.size == 0:
                raise LBPHError("This LBPH model is not computed yet. Did you call the train method?")
            parts = [(pos,) + self._search(hist, px, 1) for pos, hist, px in self._hist_by_shape(faces)]
            if len(parts) == 1:                      # one image shape (the usual case): results are already in order
                _, dist, idx = parts[0]
            else:
                dist = torch.empty((len(faces), 1), dtype=torch.float32, device=self.device)
                idx = torch.empty((len(faces), 1), dtype=torch.int64, device=self.device)
                for pos, d, i in parts:
                    p = torch.tensor(pos, dtype=torch.int64, device=self.device)
                    dist[p], idx[p] = d, i
        d = dist[:, 0].cpu().numpy().astype(np.float64)      # float32 -> float64 is exact wherever it is done
        i = idx[:, 0].cpu().numpy()
        ok = (i >= 0) & (d < self._threshold)
        labels = np.where(ok, self._labels[np.clip(i, 0, max(self.size - 1, 0))], -1).astype(np.int32)
        return labels, np.where(ok, d, DBL_MAX)

    def predict(self, src) -> Tuple[int, float]:
        """(label, confidence) exactly as cv2.face's predict: confidence is the chi-square distance."""
        labels, dists = self.predict_batch([src])
        return int(labels[0]), float(dists[0])

    def predict_topk(self, src, k: int = 5) -> List[Tuple[int, float]]:
        """The k nearest gallery faces as (label, distance) — a true top-k (the reference's web UI
        approximates one with repeated predicts, web_app.py:628-701)."""
        img = _as_gray_u8(src)
        dist, idx = self.predict_device(self._upload([img]), k)
        d, i = dist[0].double().cpu().numpy(), idx[0].cpu().numpy()
        return [(int(self._labels[j]), float(x)) for x, j in zip(d, i) if j >= 0 and x < self._threshold]

    # ---- model state ----------------------------------------------------------------------------
    def getLabels(self) -> np.ndarray:
        return self._labels.reshape(-1, 1).copy()

    def get_histograms_u16(self) -> Tuple[np.ndarray, np.ndarray]:
        """(u16 [N, L] counts in training order, cell_px int32 [N])."""
        if self._world_rank()[0] > 1:
            raise LBPHError("a sharded LBPH model holds only its shard of the histograms on each rank; "
                            "save / getHistograms need a single-GPU model")
        hist = np.zeros((self.size, self.hist_len), np.uint16)
        px = np.zeros((self.size,), np.int32)
        for g in self._groups:
            rows = g.rows.cpu().numpy()
            hist[rows] = g.hist.cpu().numpy().astype(np.uint16)
            px[rows] = g.cell_px
        return hist, px

    def getHistograms(self) -> List[np.ndarray]:
        """OpenCV's view: list of float32 1 x L histograms, count * float32(1/cell_px)."""
        hist, px = self.get_histograms_u16()
        return [(hist[i].astype(np.float32) * np.float32(1.0 / px[i])).reshape(1, -1) for i in range(self.size)]

    # save / read / write live in formats.py (OpenCV FileStorage layout)
    def save(self, filename: str) -> None:
        from .formats import write_lbph_model
        write_lbph_model(self, filename)

    write = save

    def read(self, filename: str) -> None:
        from .formats import read_lbph_model
        read_lbph_model(self, filename)


def LBPHFaceRecognizer_create(radius: int = 1, neighbors: int = 8, grid_x: int = 8, grid_y: int = 8,
                              threshold: float = DBL_MAX, device: Optional[str] = None, group=None) -> LBPHFaceRecognizer:
    """cv2.face.LBPHFaceRecognizer_create (models/lbphmodel/train_lbph.py:24-29; no-arg form web_app.py:245)."""
    return LBPHFaceRecognizer(radius, neighbors, grid_x, grid_y, threshold, device, group)


# ---- mirrors of the reference's wrappers ------------------------------------------------------------
def preprocess_frames_device(frames_bgr: torch.Tensor, target_size: Tuple[int, int] = (100, 100), grayscale: bool = True
                             ) -> torch.Tensor:
    """Batched `_preprocess_image_for_lbph(..., detector=None)` (train_lbph_script.py:49-76) on the device: every frame
    of u8 CUDA [B, H, W, 3] goes through cv2.resize(image, target_size) and, if `grayscale`, cv2.cvtColor(BGR2GRAY),
    bit for bit (fused: the resized colour frame is never written).  Returns u8 [B, rows, cols] or [B, rows, cols, 3]."""
    return ops.resize_linear(frames_bgr, target_size, to_gray=grayscale)


def train_lbph_model(faces, labels, radius=1, neighbors=8, grid_x=8, grid_y=8):
    """models/lbphmodel/train_lbph.py:4-36 — same signature, returns a trained recognizer."""
    model = LBPHFaceRecognizer_create(radius=radius, neighbors=neighbors, grid_x=grid_x, grid_y=grid_y)
    if not isinstance(labels, np.ndarray):
        labels = np.array(labels, dtype=np.int32)
    model.train(faces, labels)
    return model


def recognize_face(model, face_img, threshold):
    """models/lbphmodel/inference_lbph.py:4-18 — known iff confidence < threshold, else label None."""
    pred, conf = model.predict(face_img)
    if conf < threshold:
        return {"label": pred, "confidence": conf, "status": "known"}
    return {"label": None, "confidence": conf, "status": "unknown"}


def web_confidence(distance: float) -> float:
    """web_app.py:597 (and :684): the UI's confidence from a chi-square distance, max(0, min(1, (200 - d) / 200)) below 200, else 0."""
    return max(0, min(1, (200 - distance) / 200)) if distance < 200 else 0.0


def recognize_face_web(model, face_img, threshold, label_map=None):
    """The LBPH branch of the web UI (web_app.py:587-605): predict, identity from the label map (`Person_<label>` without
    one), confidence = web_confidence(distance), and "Unknown" when distance > threshold — note the strict '>' here
    against `conf < threshold` in inference_lbph.recognize_face (a distance equal to the threshold is known in both)."""
    label, distance = model.predict(face_img)
    identity = label_map.get(label, f"Person_{label}") if label_map else f"Person_{label}"
    if distance > threshold:
        identity = "Unknown"
    return {"identity": identity, "confidence": float(web_confidence(distance)), "distance": float(distance), "label": int(label)}


def evaluate_lbph(model, faces, labels, threshold):
    """models/lbphmodel/evaluate_lbph.py:4-45 — (accuracy, coverage, used, confidences); the N predicts
    of the reference's loop run as one batched kernel call."""
    preds, confs = model.predict_batch(list(faces))
    labels = np.asarray(labels)
    accepted = confs < threshold
    used = int(np.sum(accepted))
    correct = int(np.sum(preds[accepted] == labels[accepted])) if used > 0 else 0
    accuracy = (correct / used) if used > 0 else 0.0
    coverage = (used / len(labels)) if len(labels) > 0 else 0.0
    return accuracy, coverage, used, np.array(confs)


def find_optimal_threshold(model, faces, labels, min_coverage=0.3, threshold_range=None):
    """models/lbphmodel/threshold_lbph.py:7-96 — maximise accuracy*coverage subject to coverage >= min_coverage;
    returns (best_threshold, best_score, [(threshold, accuracy, coverage, score), ...])."""
    if threshold_range is None:
        threshold_range = range(40, 121, 5)
    predictions, confidences = model.predict_batch(list(faces))
    labels = np.asarray(labels)
    best_threshold, best_score, results = None, -1, []
    for threshold in threshold_range:
        accepted = confidences < threshold
        used = np.sum(accepted)
        accuracy = (np.sum(predictions[accepted] == labels[accepted]) / used) if used > 0 else 0.0
        coverage = used / len(labels) if len(labels) > 0 else 0.0
        if coverage >= min_coverage:
            score = accuracy * coverage
            results.append((threshold, accuracy, coverage, score))
            if score > best_score:
                best_score, best_threshold = score, threshold
    if best_threshold is None:
        best_threshold, best_score = max(threshold_range), 0.0
    return best_threshold, best_score, results
