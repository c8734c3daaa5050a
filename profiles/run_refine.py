"""Exact fp32 top-5 through the tensor cores (ops.cosine_topk_exact: fp16 first pass + exact re-score + proof) against the
fp32 kernels (cosine_simt / cosine_gemv) on the same inputs.  python profiles/run_refine.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerecognition_b200 import ops, _native as NV

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2], out


for n_rows in tuple(int(x) for x in os.environ.get("FRB_REFINE_ROWS", "10000,100000,1000000").split(",")):
    gen = torch.Generator(device=dev).manual_seed(n_rows)
    gal = ops.normalize_rows(torch.randn((n_rows, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP)
    g16 = ops.normalize_rows(gal, NV.FRB_QNORM_CLAMP, torch.float16)
    gn = ops.row_norms(gal)
    for nq in tuple(int(x) for x in os.environ.get("FRB_REFINE_Q", "8,64,256,4096").split(",")):
        src = torch.randint(0, n_rows, (nq,), generator=gen, device=dev)
        q = gal[src] + 0.03 * torch.randn((nq, 512), generator=gen, device=dev)
        q[: nq // 10] = torch.randn((nq // 10, 512), generator=gen, device=dev)          # 10 % without a match
        qn = ops.row_norms(q)
        t_tc, (s1, i1) = timed(lambda: ops.cosine_topk_exact(q, gal, g16, 5, q_norms=qn, g_norms=gn))
        _, _, fail, _ = ops.cosine_topk_refined(q, gal, g16, 5, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=gn)  # failure count only
        if nq * n_rows <= 4096 * 100_000 or nq <= 256:
            t_ex, (s2, i2) = timed(lambda: ops.cosine_topk(q, gal, 5, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=gn), 3)
            same = bool((((s1 - s2).abs() <= 2e-6) | (i1 == i2)).all()) and float((s1 - s2).abs().max()) <= 2e-6
        else:
            t_ex, same = float("nan"), None
        print(f"{nq:5d} q x {n_rows:8d} rows fp32 top-5: tensor-core first pass + exact re-score {t_tc:8.3f} ms "
              f"({int(fail.item())} queries re-run exactly), fp32 kernels {t_ex:8.3f} ms, identical={same}")
