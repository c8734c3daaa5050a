"""Times the tensor-core chi-square filter against the exact scan on configs[4]-shaped data (one GPU's share).
Usage: python profiles/run_chisq_filter.py [n_gallery=125000] [n_query=1024] [side=112] [reps=3]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from facerecognition_b200 import ops, _native as NV  # noqa: E402
from test_gpu_chisq_filter import faces_gpu, exact_top1  # noqa: E402


def main():
    n_gal = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
    n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    side = int(sys.argv[3]) if len(sys.argv) > 3 else 112
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    exact_q = int(os.environ.get("EXACT_QUERIES", n_q))
    parts = []
    px = None
    for lo in range(0, n_gal, 16384):
        n = min(16384, n_gal - lo)
        h, px = ops.lbp_hist(faces_gpu(n, side, 100 + lo))
        parts.append(ops.compact_histograms(h, px))
    g8 = torch.cat(parts, 0)
    del parts
    gen = torch.Generator(device="cuda").manual_seed(5)
    n_plant = n_q // 4
    src = torch.randint(0, n_gal, (n_plant,), generator=gen, device="cuda")
    fresh_h, _ = ops.lbp_hist(faces_gpu(n_q - n_plant, side, 7))
    # planted queries: a gallery row's counts with a few bins moved (a re-shot of the same face)
    ph = g8[src].to(torch.int16)
    bump = (torch.rand(ph.shape, generator=gen, device="cuda") < 0.02).to(torch.int16)
    ph = (ph + bump - bump.roll(1, 1)).clamp_(0, px).view(torch.uint16)
    qh = torch.cat([ph.contiguous().view(torch.int16), fresh_h.view(torch.int16)], 0).view(torch.uint16).contiguous()
    out = {"n_gallery": n_gal, "n_query": n_q, "side": side, "cell_px": px, "hist_len": int(g8.shape[1])}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    NV.profile_enable(True)
    for name in ("filtered", "exact"):
        times = []
        for r in range(reps + 1):
            torch.cuda.synchronize()
            ev[0].record()
            if name == "filtered":
                stats.zero_()
                d, i = ops.chisq_top1_filtered(qh, g8, px, stats=stats)
            else:
                de, ie = exact_top1(qh[:exact_q].contiguous(), g8, px)
            ev[1].record()
            torch.cuda.synchronize()
            if r:
                times.append(ev[0].elapsed_time(ev[1]))
        ms = sorted(times)[len(times) // 2]
        nq = n_q if name == "filtered" else exact_q
        out[name] = {"ms": round(ms, 3), "pairs_per_s": round(nq * n_gal / ms * 1e3, 1), "faces_per_s": round(nq / ms * 1e3, 1)}
        if name == "filtered":
            kms, kn = NV.profile_read(NV.K_CHISQ_FILTER)
            st = stats.cpu().tolist()
            flop = 2.0 * n_q * n_gal * g8.shape[1] * 8
            out[name].update({"filter_kernel_ms": round(kms / max(kn, 1), 3), "tflops": round(flop / (kms / max(kn, 1)) / 1e9, 1),
                              "fallback_queries": st[0], "survivors_per_query": round(st[1] / n_q, 1),
                              "raw_candidates_per_query": round(st[2] / n_q, 1)})
    out["identical"] = bool(torch.equal(i[:exact_q], ie) and torch.equal(d[:exact_q].view(torch.int32), de.view(torch.int32)))
    out["speedup"] = round(out["exact"]["faces_per_s"] and out["filtered"]["faces_per_s"] / out["exact"]["faces_per_s"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
