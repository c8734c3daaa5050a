"""K2 alone (ncu target / quick timing): python profiles/run_lbp_only.py [faces] [H] [W] [kind]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
h = int(sys.argv[2]) if len(sys.argv) > 2 else 112
w = int(sys.argv[3]) if len(sys.argv) > 3 else 112
kind = sys.argv[4] if len(sys.argv) > 4 else "noise"
g = torch.Generator(device='cuda').manual_seed(1)
if kind == "noise":
    faces = torch.randint(0, 256, (n, h, w), generator=g, device='cuda', dtype=torch.uint8)
elif kind == "flat":
    faces = torch.full((n, h, w), 200, device='cuda', dtype=torch.uint8)
else:  # smooth: blurred noise (many ties, skewed codes)
    x = torch.rand((n, 1, h, w), generator=g, device='cuda')
    k = torch.ones((1, 1, 5, 5), device='cuda') / 25
    faces = (torch.nn.functional.conv2d(x, k, padding=2)[:, 0] * 255).to(torch.uint8).contiguous()
for _ in range(3):
    hist, px = ops.lbp_hist(faces)
torch.cuda.synchronize()
NV.profile_enable(True)
for _ in range(5):
    hist, px = ops.lbp_hist(faces)
ms, k = NV.profile_read(NV.K_LBP_HIST)
NV.profile_enable(False)
per = ms / k
gbs = n * (h * w + 16384 * 2) / (per * 1e-3) / 1e9
print(f"{kind} {n}x{h}x{w}: cell_px={px} sum={int(hist.sum())} {per:.4f} ms/launch {n / per / 1e3:.2f} M faces/s {gbs:.0f} GB/s")
