"""Long lists (9..16 slots) in the CTA-pair kernel with the per-lane admission queue + share-of-k bounds, against the
single-CTA kernel.  Interleaved A/B through the environment knobs of cosine_tc.cu; results must be bit-identical."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerecognition_b200 import ops, _native as NV
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
gal_all = ops.normalize_rows(torch.randn((1_000_000, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP, torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
MODES = {"single": dict(FRB_TC_PAIR="0"),
         "pair": dict(FRB_TC_PAIR_LONG="0"),
         "pair+long-units": dict(FRB_TC_PAIR_LONG="1"),
         "shipped": dict()}
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or \
         [(4096, 1_000_000, 16), (4096, 1_000_000, 10), (2048, 1_000_000, 16), (1024, 1_000_000, 16), (256, 1_000_000, 16),
          (8192, 500_000, 16), (32768, 125_000, 16), (4096, 125_000, 16), (4096, 1_000_000, 5)]
for nq, rows, k in shapes:
    gal = gal_all[:rows].contiguous()
    q = torch.randn((nq, 512), generator=gen, device=dev)
    src = torch.randint(0, rows, (nq,), generator=gen, device=dev)
    q[nq // 10:] = gal[src[nq // 10:]].float() + 0.03 * q[nq // 10:]
    res, ref = {}, None
    for rnd in range(3):
        for mode, env in MODES.items():
            for kk in ("FRB_TC_PAIR", "FRB_TC_PAIR_LONG"):
                os.environ.pop(kk, None)
            os.environ.update(env)
            for _ in range(2):
                out = ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
            if ref is None:
                ref = out
            assert torch.equal(out[1], ref[1]) and torch.equal(out[0], ref[0]), (mode, nq, rows, k)
            ts = []
            for _ in range(4):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            res.setdefault(mode, []).append(sorted(ts)[1])
    print(f"{nq:6d} q x {rows:8d} rows k={k:2d}: " + ", ".join(f"{m} {sorted(v)[1]:.3f} ms" for m, v in res.items()) + "  (identical results)", flush=True)
