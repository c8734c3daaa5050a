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
