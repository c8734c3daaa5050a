"""On-disk gallery formats of the reference, read and written unchanged.

* embedding DB: ``np.save(path, {name: float32[512]})`` (inference/extract_embeddings.py:831,
  inference/recognition_engine.py:135,428) — pickled dict, loaded with ``np.load(..., allow_pickle=True).item()``;
* prototypes ``float32[C, 512]`` + ``label_mapping.npy`` (inference/extract_embeddings.py:542-543,589);
* FAISS ``IndexFlatIP`` file (inference/extract_embeddings.py:628-642): parsed/emitted natively — faiss itself is
  not a dependency (PARITY UNPINNED: no faiss in this image to cross-check the bytes);
* OpenCV LBPH model XML/YAML (models/lbphmodel/train_lbph_script.py:222, web_app.py:245-246) via cv2.FileStorage;
* ``label_map.npy`` ``{int label: name}`` (models/lbphmodel/train_lbph_script.py:225-226).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Tuple

import numpy as np

# ---- embedding DB ------------------------------------------------------------------------------------


def load_embedding_db(path: str) -> Dict[str, np.ndarray]:
    """inference/recognition_engine.py:135."""
    return np.load(path, allow_pickle=True).item()


def save_embedding_db(path: str, db: Dict[str, np.ndarray]) -> None:
    """inference/recognition_engine.py:424-428 / inference/extract_embeddings.py:830-831."""
    d = os.path.dirname(path)
    os.makedirs(d if d else ".", exist_ok=True)
    np.save(path, db)


def load_label_mapping(path: str) -> Tuple[dict, dict]:
    """label_mapping.npy -> (label_to_id, id_to_label) exactly as RecognitionEngine._load_faiss reads it
    (inference/recognition_engine.py:159-162)."""
    mapping = np.load(path, allow_pickle=True).item()
    return mapping.get("label_to_id", {}), mapping.get("id_to_label", {})


def load_lbph_label_map(path: str) -> Dict[int, str]:
    """label_map.npy {int -> name} (models/lbphmodel/train_lbph_script.py:225-226)."""
    return np.load(path, allow_pickle=True).item()


# ---- FAISS IndexFlatIP -----------------------------------------------------------------------------
# faiss/impl/index_write.cpp (1.7.x): fourcc "IxFI", header {int d; int64 ntotal; int64 dummy; int64 dummy;
# uint8 is_trained; int metric_type (0 = inner product)}, then the vector as {uint64 n_floats; float32 data[]}.


def read_faiss_flat_ip(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        buf = f.read()
    fourcc = buf[:4]
    if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
        raise ValueError(f"{path}: unsupported FAISS index type {fourcc!r} (only flat indexes are handled)")
    d, ntotal, _d1, _d2, _trained, metric = struct.unpack_from("<iqqqBi", buf, 4)
    off = 4 + struct.calcsize("<iqqqBi")
    if metric > 1:
        off += 4  # metric_arg
    (count,) = struct.unpack_from("<Q", buf, off)
    off += 8
    if count != ntotal * d:
        raise ValueError(f"{path}: vector holds {count} floats, expected {ntotal}x{d}")
    return np.frombuffer(buf, dtype="<f4", count=count, offset=off).reshape(ntotal, d).copy()


def write_faiss_flat_ip(path: str, rows: np.ndarray) -> None:
    rows = np.ascontiguousarray(rows, dtype="<f4")
    ntotal, d = rows.shape
    with open(path, "wb") as f:
        f.write(b"IxFI")
        f.write(struct.pack("<iqqqBi", d, ntotal, 1 << 20, 1 << 20, 1, 0))
        f.write(struct.pack("<Q", ntotal * d))
        f.write(rows.tobytes())


# ---- OpenCV LBPH model ------------------------------------------------------------------------------


def _fits_cell_px(h: np.ndarray, n: int) -> bool:
    """True when every value of the stored row is an integer count over n pixels and each 256-bin cell holds exactly n."""
    c = h.astype(np.float64) * n
    r = np.round(c)
    if np.any(np.abs(c - r) > 1e-6 * np.maximum(r, 1.0)):        # float32(count) * float32(1/n): relative error ~2e-7
        return False
    cells = r.reshape(-1, 256).sum(1) if h.shape[0] % 256 == 0 else np.array([r.sum()])
    return bool(np.all(cells == n)) or h.shape[0] % 256 != 0


def _infer_cell_px(h: np.ndarray, hint: int = 0) -> int:
    """A stored float histogram row is count * float32(1/n) per cell and each cell's counts sum to n: recover n.  The
    smallest non-zero value is c_min / n for an unknown integer c_min, so n = round(c / v_min) is tried for c = 1, 2, ...
    against the STRUCTURE (all values integral over n, every cell summing to n) — not against a list of multiples of
    round(1 / v_min), which misses rows whose smallest count does not divide n (e.g. 5 of 144)."""
    if hint and _fits_cell_px(h, hint):
        return hint
    nz = h[h > 0]
    if nz.size == 0:
        raise ValueError("cannot infer the cell size of an all-zero histogram")
    vmin = float(nz.min())
    for c in range(1, 65536):
        n = int(round(c / vmin))
        if n > 65535:
            break
        if n >= 1 and _fits_cell_px(h, n):
            return n
    raise ValueError("histogram values are not integer counts over a common cell size; not an OpenCV LBPH histogram")


def write_lbph_model(model, filename: str) -> None:
    """LBPH::save — <opencv_lbphfaces>{threshold, radius, neighbors, grid_x, grid_y, histograms[], labels, labelsInfo[]}."""
    import cv2

    fs = cv2.FileStorage(filename, cv2.FILE_STORAGE_WRITE)
    if not fs.isOpened():
        raise model_error(f"File can't be opened for writing: {filename}")
    fs.startWriteStruct("opencv_lbphfaces", cv2.FileNode_MAP)
    fs.write("threshold", float(model.getThreshold()))
    fs.write("radius", int(model.getRadius()))
    fs.write("neighbors", int(model.getNeighbors()))
    fs.write("grid_x", int(model.getGridX()))
    fs.write("grid_y", int(model.getGridY()))
    fs.startWriteStruct("histograms", cv2.FileNode_SEQ)
    for h in model.getHistograms():
        fs.write("", h)
    fs.endWriteStruct()
    fs.write("labels", model.getLabels().astype(np.int32))
    fs.startWriteStruct("labelsInfo", cv2.FileNode_SEQ)
    for label, text in sorted(model._label_info.items()):
        fs.startWriteStruct("", cv2.FileNode_MAP)
        fs.write("label", int(label))
        fs.write("value", str(text))
        fs.endWriteStruct()
    fs.endWriteStruct()
    fs.endWriteStruct()
    fs.release()


def read_lbph_model(model, filename: str) -> None:
    """LBPH::load — accepts files written by OpenCV-contrib or by write_lbph_model."""
    import cv2
    import torch

    fs = cv2.FileStorage(filename, cv2.FILE_STORAGE_READ)
    if not fs.isOpened():
        raise model_error(f"File can't be opened for reading: {filename}")
    root = fs.getNode("opencv_lbphfaces")
    if root.empty():
        root = fs.root()
    model.setThreshold(root.getNode("threshold").real())
    model.setRadius(int(root.getNode("radius").real()))
    model.setNeighbors(int(root.getNode("neighbors").real()))
    model.setGridX(int(root.getNode("grid_x").real()))
    model.setGridY(int(root.getNode("grid_y").real()))
    hn = root.getNode("histograms")
    hists = [hn.at(i).mat().reshape(-1) for i in range(hn.size())]
    labels_node = root.getNode("labels")
    labels = labels_node.mat().reshape(-1).astype(np.int32) if not labels_node.empty() else np.zeros((0,), np.int32)
    info = root.getNode("labelsInfo")
    model._label_info = {}
    if not info.empty():
        for i in range(info.size()):
            item = info.at(i)
            model._label_info[int(item.getNode("label").real())] = item.getNode("value").string()
    fs.release()
    if len(hists) != labels.shape[0]:
        raise model_error(f"{filename}: {len(hists)} histograms but {labels.shape[0]} labels")
    model._groups = []
    model._labels = labels
    if not hists:
        return
    L = model.hist_len
    by_px: Dict[int, list] = {}
    last_px = 0
    for i, h in enumerate(hists):
        if h.shape[0] != L:
            raise model_error(f"{filename}: histogram {i} has {h.shape[0]} bins, expected {L}")
        last_px = _infer_cell_px(h, last_px)              # rows of one model nearly always share a cell size: try it first
        by_px.setdefault(last_px, []).append(i)
    from .lbph import _Group
    for px, rows in by_px.items():
        counts = np.round(np.stack([hists[i] for i in rows]).astype(np.float64) * px).astype(np.uint16)
        from . import ops
        model._groups.append(_Group(px, ops.compact_histograms(torch.from_numpy(counts).to(model.device), px),
                                    torch.tensor(rows, dtype=torch.int64, device=model.device)))


def model_error(msg: str):
    from .lbph import LBPHError
    return LBPHError(msg)
