"""Single-query / small-batch latency of K1 (HBM-bound regime): python profiles/run_latency.py [rows]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = torch.Generator(device='cuda').manual_seed(1)
gal32 = ops.normalize_rows(torch.randn((n, 512), generator=g, device='cuda'), NV.FRB_QNORM_CLAMP)
gal16 = ops.normalize_rows(gal32, NV.FRB_QNORM_NONE, torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for name, gal, bytes_per_row in (("bf16", gal16, 1024), ("fp32", gal32, 2048)):
    for nq in (1, 4, 8, 64):
        q = gal32[:nq].clone() + 0.01
        for _ in range(3):
            ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in ev:
            flush.zero_()
            a.record()
            s, i = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
        # the same step recorded into a CUDA graph (every entry point only enqueues on the caller's stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            gs, gi = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
        for a, b in ev:
            flush.zero_()
            a.record()
            graph.replay()
            b.record()
        torch.cuda.synchronize()
        gms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
        assert torch.equal(gi, i) and torch.equal(gs, s)
        print(f"{name} gallery {n} rows, {nq} queries: {ms:.3f} ms  ({n * bytes_per_row / ms / 1e6:.0f} GB/s of gallery)  as a graph: {gms:.3f} ms ({n * bytes_per_row / gms / 1e6:.0f} GB/s)  top1 ok={bool((i[:, 0] == torch.arange(nq, device='cuda')).all())}")

# the reference's own single-image entry (RecognitionEngine.recognize_with_db, recognition_engine.py:267-289: 133 ms at 10k
# identities on the CPU): wall clock of the whole call, host numpy in -> Python tuple out
import time
import numpy as np
from facerecognition_b200.recognition_engine import RecognitionEngine
rng = np.random.default_rng(3)
db = {f"id_{i:05d}": v for i, v in enumerate((lambda x: x / np.linalg.norm(x, axis=1, keepdims=True))(rng.standard_normal((10_000, 512)).astype(np.float32)))}
eng = RecognitionEngine(model_path=None, db_path=None, use_face_detection=False, device="cuda")
eng.db = db
emb = db["id_01234"] + 0.01 * rng.standard_normal(512).astype(np.float32)
for _ in range(20):
    name, score, top = eng.recognize_with_db(emb)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    name, score, top = eng.recognize_with_db(emb)
dt = (time.perf_counter() - t0) / 200
print(f"RecognitionEngine.recognize_with_db, 10k identities fp32, one embedding: {dt * 1e3:.3f} ms wall clock per call -> {name} {score:.4f}")
