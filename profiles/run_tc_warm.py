"""Is the threshold warm-up pass worth its launch on small problems?  A/B with FRB_TC_WARM=0/1, interleaved."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerecognition_b200 import ops, _native as NV
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
gal_all = ops.normalize_rows(torch.randn((1_000_000, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP, torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for nq, rows in ((4096, 125_000), (4096, 250_000), (1024, 125_000), (256, 1_000_000), (4096, 1_000_000), (32768, 125_000)):
    gal = gal_all[:rows].contiguous()
    q = torch.randn((nq, 512), generator=gen, device=dev)
    src = torch.randint(0, rows, (nq,), generator=gen, device=dev)
    q[nq // 10:] = gal[src[nq // 10:]].float() + 0.03 * q[nq // 10:]
    res = {}
    ref = None
    for rnd in range(5):
        for mode in ("1", "0"):
            os.environ["FRB_TC_WARM"] = mode
            for _ in range(2):
                out = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
            if ref is None:
                ref = out
            assert torch.equal(out[1], ref[1]) and torch.equal(out[0], ref[0])
            ts = []
            for _ in range(4):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            res.setdefault(mode, []).append(sorted(ts)[1])
    print(f"{nq:6d} q x {rows:8d} rows: with warm-up pass {sorted(res['1'])[2]:.3f} ms, without {sorted(res['0'])[2]:.3f} ms (whole call, median of 5 rounds)")
