"""cosine_tc_kernel at the weak-scaling shapes of bench.py --gpus N, all on ONE GPU (same flops per step):
4096 x 1M (N=1), 8192 x 500k (N=2), 16384 x 250k (N=4), 32768 x 125k (N=8).  Separates the kernel's shape dependence
from anything the other ranks do."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerecognition_b200 import ops, _native as NV

dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
gal_all = ops.normalize_rows(torch.randn((1_000_000, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP, torch.bfloat16)
q_all = torch.randn((32768, 512), generator=gen, device=dev)
src = torch.randint(0, 125_000, (32768,), generator=gen, device=dev)
q_all[3000:] = gal_all[src[3000:]].float() + 0.03 * q_all[3000:]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def measure(q, gal, reps=6):
    for _ in range(3):
        ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
    torch.cuda.synchronize()
    NV.profile_enable(True); NV.profile_read(NV.K_COSINE_TC)
    ev = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    kms, kn = NV.profile_read(NV.K_COSINE_TC); NV.profile_enable(False)
    return sum(a.elapsed_time(b) for a, b in ev) / reps, kms / reps


# A/B in one process, interleaved (the power cap makes back-to-back runs drift by several per cent)
MODES = (("single CTA, rule of thumb", {"FRB_TC_PAIR": "0", "FRB_TC_BALANCE": "0"}),
         ("single CTA, balanced groups", {"FRB_TC_PAIR": "0", "FRB_TC_BALANCE": "1"}),
         ("CTA pair (cta_group::2)", {"FRB_TC_PAIR": "1", "FRB_TC_BALANCE": "1"}))
for n in (1, 2, 4, 8):
    gal = gal_all[:1_000_000 // n].contiguous()
    q = q_all[:4096 * n].contiguous()
    acc = {name: [] for name, _ in MODES}
    ref = None
    for rnd in range(5):
        for name, env in MODES:
            os.environ.update(env)
            acc[name].append(measure(q, gal))
            if rnd == 0:
                out = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
                if ref is None:
                    ref = out
                assert torch.equal(out[1], ref[1]) and torch.equal(out[0], ref[0]), name
    for name, _ in MODES:
        ks = sorted(k for _, k in acc[name])
        st = sorted(s for s, _ in acc[name])
        print(f"N={n}: {4096 * n:6d} q x {1_000_000 // n:7d} rows, {name:28s}: cosine_tc median {ks[2]:.3f} ms (min {ks[0]:.3f} max {ks[-1]:.3f}), "
              f"step median {st[2]:.3f} ms, {2 * 4096 * 1e6 * 512 / (ks[2] * 1e-3) / 1e12:.0f} TFLOP/s")
