"""One small program per ncu capture (round 2): python profiles/run_ncu_targets.py <target>
targets: tc | tc16 | chisq_q1_u16 | chisq_q1_u8 | chisq_b_u16 | chisq_b_u8 | filter | lbp | lbp8 | resize
Each runs the op 3 times (2 warm-up calls + the one to capture with `-s <2 x launches per call> -c <launches per call>`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerecognition_b200 import ops, _native as NV  # noqa: E402

target = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(1)


def faces_gpu(n, side, seed):
    gg = torch.Generator(device="cuda").manual_seed(seed)
    base = torch.randint(0, 256, (n, side // 4 + 2, side // 4 + 2), generator=gg, device="cuda").float()
    up = base.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :side, :side]
    return (up + 12.0 * torch.randn((n, side, side), generator=gg, device="cuda")).clamp(0, 255).to(torch.uint8)


def run(fn, label, work=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{label}: {a.elapsed_time(b):.3f} ms per call" + (f" ({work})" if work else ""))
    return out


if target == "tc16":
    # the first pass of the exact-fp32 path: fp16 operands, 16-slot lists through the CTA-pair kernel's queued admission
    gal = ops.normalize_rows(torch.randn((1_000_000, 512), generator=g, device="cuda"), NV.FRB_QNORM_CLAMP, torch.float16)
    q = gal[torch.randint(0, 1_000_000, (4096,), generator=g, device="cuda")].float() + 0.03 * torch.randn((4096, 512), generator=g, device="cuda")
    run(lambda: ops.cosine_topk(q, gal, 16, qnorm_mode=NV.FRB_QNORM_CLAMP), "cosine_tc fp16 4096 x 1M k=16")
elif target == "tc":
    gal = ops.normalize_rows(torch.randn((1_000_000, 512), generator=g, device="cuda"), NV.FRB_QNORM_CLAMP, torch.bfloat16)
    q = gal[torch.randint(0, 1_000_000, (4096,), generator=g, device="cuda")].float() + 0.03 * torch.randn((4096, 512), generator=g, device="cuda")
    run(lambda: ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP), "cosine_tc 4096 x 1M k=5")
elif target.startswith("chisq"):
    hist, px = ops.lbp_hist(faces_gpu(8192, 112, 3))
    ng = 100_000
    gal = hist.view(torch.int16)[torch.randint(0, 8192, (ng,), generator=g, device="cuda")].contiguous().view(torch.uint16)
    if target.endswith("u8"):
        gal = ops.compact_histograms(gal, px)
    nq = 1 if "_q1_" in target else 64
    qh = hist[:nq].contiguous()
    ops.FILTER_ENABLED = False
    run(lambda: ops.chisq_topk(qh, px, gal, px, 1), f"chisq exact {nq} x {ng} {gal.dtype}")
elif target == "filter":
    n_gal, nq = 148 * 256, 256
    gh, px = ops.lbp_hist(faces_gpu(n_gal, 112, 5), counts8=True)
    qh, _ = ops.lbp_hist(faces_gpu(nq, 112, 6))
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    run(lambda: ops.chisq_top1_filtered(qh, gh, px, stats=stats), f"chisq filter {nq} x {n_gal}",
        f"{2 * nq * n_gal * 16384 * 8 / 1e12:.2f} TFLOP in the filter kernel")
    print("stats", stats.tolist())
elif target in ("lbp", "lbp8"):
    faces = faces_gpu(65536, 112, 7)
    run(lambda: ops.lbp_hist(faces, counts8=target == "lbp8"), f"lbp_hist 65536 x 112x112 counts8={target == 'lbp8'}")
elif target == "resize":
    big = torch.randint(0, 256, (8192, 180, 240, 3), generator=g, device="cuda", dtype=torch.uint8)
    run(lambda: ops.resize_linear(big, (112, 112), to_gray=True), "resize 8192 x 180x240x3 -> 112x112 gray")
else:
    raise SystemExit(f"unknown target {target}")
