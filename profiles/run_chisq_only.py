"""K3 alone (ncu target / quick timing): python profiles/run_chisq_only.py [queries] [rows]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
g = torch.Generator(device='cuda').manual_seed(1)
faces = torch.randint(0, 256, (8192, 112, 112), generator=g, device='cuda', dtype=torch.uint8)
hist, px = ops.lbp_hist(faces)
gal = hist.view(torch.int16)[torch.randint(0, 8192, (ng,), generator=g, device='cuda')].contiguous().view(torch.uint16)
qh = hist.view(torch.int16)[torch.randint(0, 8192, (nq,), generator=g, device='cuda')].contiguous().view(torch.uint16)
for name, store in (("u16 gallery", gal), ("u8 gallery", ops.compact_histograms(gal, px))):
    for _ in range(2):
        d, i = ops.chisq_topk(qh, px, store, px, 1)
    torch.cuda.synchronize()
    NV.profile_enable(True)
    NV.profile_read(NV.K_CHISQ)
    for _ in range(3):
        d, i = ops.chisq_topk(qh, px, store, px, 1)
    ms, k = NV.profile_read(NV.K_CHISQ)
    NV.profile_enable(False)
    per = ms / k
    pairs = nq * ng
    row_bytes = store.shape[1] * store.element_size()
    print(f"{name}: {nq} q x {ng} rows: {per:.3f} ms/launch {pairs / per / 1e3:.1f} M pairs/s, {pairs * row_bytes / per / 1e6:.0f} GB/s of "
          f"{row_bytes} B rows; min d={float(d.min()):.6g} zero-dist matches={int((d[:, 0] == 0).sum())}")
