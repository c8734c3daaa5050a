"""K1-bf16 with a small batch (HBM-bound regime): python profiles/run_tc_small.py [queries] [rows]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
g = torch.Generator(device='cuda').manual_seed(1)
gal = ops.normalize_rows(torch.randn((n, 512), generator=g, device='cuda'), NV.FRB_QNORM_CLAMP, torch.bfloat16)
q = gal[:nq].float() + 0.01
for _ in range(3):
    ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
torch.cuda.synchronize()
NV.profile_enable(True); NV.profile_read(NV.K_COSINE_TC)
for _ in range(5):
    s, i = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
ms, k = NV.profile_read(NV.K_COSINE_TC); NV.profile_enable(False)
print(f"{nq} q x {n} rows: {ms / 5:.3f} ms in cosine_tc_kernel per call ({k // 5} launches) = {n * 1024 / (ms / 5) / 1e6:.0f} GB/s")
