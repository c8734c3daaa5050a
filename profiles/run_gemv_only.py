"""K1-small alone (ncu target): one query against a 12.5M-row bf16 shard (configs[3]'s per-GPU share)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
g = torch.Generator(device='cuda').manual_seed(1)
gal = torch.empty((n, 512), dtype=torch.bfloat16, device='cuda')
for r0 in range(0, n, 500_000):
    r1 = min(n, r0 + 500_000)
    gal[r0:r1] = ops.normalize_rows(torch.randn((r1 - r0, 512), generator=g, device='cuda'), NV.FRB_QNORM_CLAMP, torch.bfloat16)
q = gal[12345:12346].float() + 0.01
for _ in range(3):
    s, i = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
torch.cuda.synchronize()
NV.profile_enable(True)
NV.profile_read(NV.K_COSINE_GEMV)
for _ in range(5):
    s, i = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
ms, k = NV.profile_read(NV.K_COSINE_GEMV)
NV.profile_enable(False)
print(f"1 query x {n} bf16 rows: {ms / k:.3f} ms/launch = {n * 1024 / (ms / k) / 1e6:.0f} GB/s, top1={int(i[0, 0])}")
