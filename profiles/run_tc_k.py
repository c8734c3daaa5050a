"""K1-bf16 at different k: python profiles/run_tc_k.py [queries] [rows]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
g = torch.Generator(device='cuda').manual_seed(1)
gal = ops.normalize_rows(torch.randn((n, 512), generator=g, device='cuda'), NV.FRB_QNORM_CLAMP, torch.bfloat16)
q = gal[torch.randint(0, n, (nq,), generator=g, device='cuda')].float() + 0.03 * torch.randn((nq, 512), generator=g, device='cuda')
for k in (5, 8, 16, 32, 64):
    for _ in range(2):
        ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        s, i = ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
    b.record(); torch.cuda.synchronize()
    print(f"{nq} q x {n} rows, k={k}: {a.elapsed_time(b) / 3:.2f} ms")
