"""Exact fp32 batched search on a large fp32 gallery: tiled FFMA kernel vs tensor-core first pass + exact re-score."""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
k = 5
g = torch.Generator(device='cuda').manual_seed(1)
gal = ops.normalize_rows(torch.randn((n, 512), generator=g, device='cuda'), NV.FRB_QNORM_CLAMP)
g16 = ops.normalize_rows(gal, NV.FRB_QNORM_CLAMP, torch.bfloat16)
q = gal[torch.randint(0, n, (nq,), generator=g, device='cuda')] + 0.03 * torch.randn((nq, 512), generator=g, device='cuda')
qn, gn = ops.row_norms(q), ops.row_norms(gal)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
t_ref, (s2, i2, fail) = timed(lambda: ops.cosine_topk_refined(q, gal, g16, k, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=gn))
t_ex, (s1, i1) = timed(lambda: ops.cosine_topk(q, gal, k, score_mode=NV.FRB_SCORE_REF_COSINE, q_norms=qn, g_norms=gn), reps=1)
print(f"{nq} q x {n} fp32 rows, top-{k}: exact tiled kernel {t_ex:.2f} ms, tensor-core first pass + re-score {t_ref:.2f} ms "
      f"(fail={int(fail.item())}); same ids: {bool((i1 == i2).all())}, max |ds| = {float((s1 - s2).abs().max()):.2e}")
