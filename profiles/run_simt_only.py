"""K1-fp32 tiled kernel alone: python profiles/run_simt_only.py [queries] [rows]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
g = torch.Generator(device='cuda').manual_seed(1)
gal = ops.normalize_rows(torch.randn((n, 512), generator=g, device='cuda'), NV.FRB_QNORM_CLAMP)
q = gal[:nq].clone() + 0.01
for _ in range(3):
    s, i = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_EPS)
torch.cuda.synchronize()
NV.profile_enable(True)
NV.profile_read(NV.K_COSINE_SIMT)
for _ in range(5):
    s, i = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_EPS)
ms, k = NV.profile_read(NV.K_COSINE_SIMT)
NV.profile_enable(False)
print(f"{nq} q x {n} fp32 rows: {ms / k:.3f} ms/launch = {2 * nq * n * 512 / (ms / k) / 1e9:.1f} TFLOP/s, top1 ok={bool((i[:, 0] == torch.arange(nq, device='cuda')).all())}")
