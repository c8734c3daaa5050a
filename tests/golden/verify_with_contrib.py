#!/usr/bin/env python
"""Pins the two "parity unpinned" stages against the REAL third-party code, wherever it is installed:

    pip install opencv-contrib-python faiss-cpu        # the reference's requirements.txt:13,43
    python tests/golden/verify_with_contrib.py

* cv2.face (opencv-contrib): LBPHFaceRecognizer_create().train() on the committed fixture faces
  (tests/golden/lbph_golden.npz) must give histograms bit-equal to the fixture's u16 counts * float32(1/cell_px) — the
  counts oracle/lbph_oracle.c produced and every CUDA parity test is measured against; predict() must give the oracle's
  (label, distance); a model file written by the real save() must parse back to the same integer counts through
  facerecognition_b200.formats (cell size inferred from the stored floats).
* faiss: IndexFlatIP(d).add(rows); write_index must produce byte for byte what formats.write_faiss_flat_ip writes for the
  same rows, read_index must accept our file, and search() must return the oracle's ids (ties: which of two equal
  scores faiss lists first is recorded, the kernels and the oracle list the lower id first).

Neither module is installed in the authoring image or on the GPU box (no network), which is why the fixtures say
"parity unpinned".  The same checks run under pytest (tests/test_oracle_lbph.py::test_third_party_pins) and skip when
the modules are missing.  Needs no GPU.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def have_cv2_face():
    try:
        import cv2
        return hasattr(cv2, "face") and hasattr(cv2.face, "LBPHFaceRecognizer_create")
    except Exception:
        return False


def have_faiss():
    try:
        import faiss  # noqa: F401
        return True
    except Exception:
        return False


def check_lbph_against_contrib():
    """Returns a list of (name, ok, detail)."""
    import cv2
    from oracle import lbph as OL
    from facerecognition_b200 import formats
    g = np.load(os.path.join(HERE, "lbph_golden.npz"))
    out = []
    for tag in ("s100", "s112", "s57x83"):
        faces, hist, px = g[f"{tag}_faces"], g[f"{tag}_hist"], int(g[f"{tag}_cell_px"])
        labels = np.arange(len(faces), dtype=np.int32) + 3
        model = cv2.face.LBPHFaceRecognizer_create()            # radius 1, neighbours 8, grid 8x8: the reference's defaults
        model.train(list(faces), labels)
        real = np.stack([h.reshape(-1) for h in model.getHistograms()])
        want = hist.astype(np.float32) * np.float32(1.0 / px)
        out.append((f"{tag}: getHistograms() bit-equal to the fixture counts / cell_px", bool(np.array_equal(real, want)),
                    f"max |diff| {np.abs(real - want).max():.3g}"))
        ref = OL.OracleLBPH()
        ref.train(list(faces), labels)
        rng = np.random.default_rng(5)
        probes = [faces[0], faces[-1], np.clip(faces[1].astype(np.int16) + rng.integers(-9, 10, faces[1].shape), 0, 255).astype(np.uint8)]
        ok, detail = True, ""
        for p in probes:
            lab, dist = model.predict(p)
            rlab, rdist = ref.predict(p)
            if lab != rlab or abs(dist - rdist) > 1e-9 * max(abs(rdist), 1.0):
                ok, detail = False, f"cv2 ({lab}, {dist!r}) vs oracle ({rlab}, {rdist!r})"
        out.append((f"{tag}: predict() == oracle predict (label, float64 distance)", ok, detail))
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "lbph_model.xml")
            model.save(path)
            fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
            root = fs.getNode("opencv_lbphfaces")
            hn = root.getNode("histograms")
            rows = [hn.at(i).mat().reshape(-1) for i in range(hn.size())]
            fs.release()
            counts = np.stack([np.round(r.astype(np.float64) * formats._infer_cell_px(r)).astype(np.uint16) for r in rows])
            pxs = {formats._infer_cell_px(r) for r in rows}
        out.append((f"{tag}: a file written by the real save() parses to the fixture's integer counts", bool(np.array_equal(counts, hist)) and pxs == {px},
                    f"cell sizes read {sorted(pxs)} (fixture {px})"))
    return out


def check_faiss_against_real():
    import faiss
    from facerecognition_b200 import formats
    from oracle import cosine as OC
    g = np.load(os.path.join(HERE, "cosine_golden.npz"))
    rows = OC.build_flat_ip(g["a_gallery"])                     # rows / (||row|| + 1e-8), as extract_embeddings.py:622-623
    out = []
    index = faiss.IndexFlatIP(rows.shape[1])
    index.add(rows)
    with tempfile.TemporaryDirectory() as d:
        real_path, ours_path = os.path.join(d, "real.faiss"), os.path.join(d, "ours.faiss")
        faiss.write_index(index, real_path)
        formats.write_faiss_flat_ip(ours_path, rows)
        a, b = open(real_path, "rb").read(), open(ours_path, "rb").read()
        same = a == b
        first = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), None)
        out.append(("write_faiss_flat_ip == faiss.write_index byte for byte", same,
                    "" if same else f"sizes {len(a)} / {len(b)}, first difference at byte {first}"))
        back = faiss.read_index(ours_path)
        out.append(("faiss.read_index accepts our file", back.ntotal == rows.shape[0] and back.d == rows.shape[1], ""))
        out.append(("read_faiss_flat_ip reads the real file", bool(np.array_equal(formats.read_faiss_flat_ip(real_path), rows)), ""))
    q = OC.l2_normalize(g["a_queries"])
    s, i = index.search(q, 5)
    rs, ri = OC.flat_ip_search(rows, q, 5)
    close = np.abs(s - rs).max() <= 1e-5
    mism = np.argwhere(i != ri)
    tie_only = all(abs(rs[r, c] - s[r, c]) <= 1e-6 for r, c in mism)
    out.append(("IndexFlatIP.search ids == oracle ids (differences only between equal scores)", bool(close and tie_only),
                f"{len(mism)} positions differ; at exact ties faiss lists " +
                ("the lower id first like the oracle" if len(mism) == 0 else "a different id first than the oracle (lower id)")))
    return out


def main():
    results = []
    if have_cv2_face():
        results += check_lbph_against_contrib()
    else:
        print("SKIPPED: cv2.face is not importable (pip install opencv-contrib-python) -> LBP code / histogram stage stays parity unpinned")
    if have_faiss():
        results += check_faiss_against_real()
    else:
        print("SKIPPED: faiss is not importable (pip install faiss-cpu) -> IndexFlatIP file bytes / tie order stay parity unpinned")
    bad = 0
    for name, ok, detail in results:
        print(("PASS  " if ok else "FAIL  ") + name + (f"   [{detail}]" if detail else ""))
        bad += 0 if ok else 1
    if results:
        print(f"{len(results) - bad} of {len(results)} checks passed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
