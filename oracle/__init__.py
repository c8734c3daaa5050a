"""oracle/ — CPU restatement of the reference's identification stage.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and only as the checker or the reported CPU baseline.  The product
package ``facerecognition_b200`` never imports it and has no CPU fallback.

Parity status
-------------
* cosine path  (oracle/cosine.py): PINNED — golden vectors in
  ``tests/golden/cosine_golden.npz`` were produced by importing the real
  reference (``/root/reference/inference/recognition_engine.py``) in the
  authoring container with ``tests/golden/make_golden.py``.
* FAISS IndexFlatIP search: PARITY UNPINNED — ``faiss`` is neither vendored nor
  installed (requirements.txt:43 ``faiss-gpu==1.7.4``); restated as exact fp32
  inner-product top-k.
* LBPH LBP-code/histogram stage (oracle/lbph_oracle.c): PARITY UNPINNED —
  ``cv2.face`` (opencv-contrib) is absent; restated from lbph_faces.cpp.
* LBPH chi-square stage: PINNED against the real ``cv2.compareHist``.
"""
