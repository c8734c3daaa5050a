"""facerecognition_b200 — the identification stage of sin0235/FaceRecognition on B200 (sm_100a).

Two paths behind the reference's API, both running in hand-written CUDA kernels reached through the
C ABI in include/frb200.h (libfrb200.so, built in-tree):

* cosine gallery match: RecognitionEngine / FlatIPIndex / match_facenet  (recognition_engine.py)
* LBPH: LBPHFaceRecognizer_create().train/predict, train_lbph_model, recognize_face ...  (lbph.py)

Importing the package loads libfrb200.so and fails loudly if it is missing: there is no CPU fallback.
"""
from . import _native  # noqa: F401  (loads libfrb200.so; raises ImportError if it is not built)
from . import ops  # noqa: F401
from .lbph import (LBPHFaceRecognizer, LBPHFaceRecognizer_create, evaluate_lbph, find_optimal_threshold,  # noqa: F401
                   preprocess_frames_device, recognize_face, recognize_face_web, train_lbph_model, web_confidence)
from .recognition_engine import (DeviceGallery, FlatIPIndex, RecognitionEngine, build_db_from_embeddings,  # noqa: F401
                                 build_faiss_index, compute_prototypes, cosine_similarity,
                                 create_engine_from_embeddings_dir, group_plan, match_facenet, mean_embedding)
from .sharded import HostBatchPipeline, ShardedSearch, chisq_sharded, cosine_sharded, shard_bounds  # noqa: F401

__version__ = "0.1.0"
