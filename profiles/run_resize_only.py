"""Resize + gray front end alone: python profiles/run_resize_only.py [frames] [src_rows] [src_cols] [dst]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from facerecognition_b200 import ops, _native as NV
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sh = int(sys.argv[2]) if len(sys.argv) > 2 else 180
sw = int(sys.argv[3]) if len(sys.argv) > 3 else 240
d = int(sys.argv[4]) if len(sys.argv) > 4 else 112
g = torch.Generator(device='cuda').manual_seed(1)
big = torch.randint(0, 256, (n, sh, sw, 3), generator=g, device='cuda', dtype=torch.uint8)
for gray in (True, False):
    for _ in range(2):
        out = ops.resize_linear(big, (d, d), to_gray=gray)
    torch.cuda.synchronize()
    NV.profile_enable(True); NV.profile_read(NV.K_RESIZE)
    for _ in range(5):
        out = ops.resize_linear(big, (d, d), to_gray=gray)
    ms, k = NV.profile_read(NV.K_RESIZE); NV.profile_enable(False)
    by = n * (sh * sw * 3 + out[0].numel())
    print(f"{n} x {sh}x{sw}x3 -> {d}x{d} gray={gray}: {ms / k:.3f} ms = {n / (ms / k) / 1e3:.2f} M frames/s, {by / (ms / k) / 1e6:.0f} GB/s")
