"""Unit length (gallery tiles per work unit) of the CTA-pair kernel on a LARGE gallery: the planner's rule gives 16 units
per CTA, i.e. units of thousands of tiles at 10^7..10^8 rows.  python profiles/run_tc_unit_len.py [rows]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerecognition_b200 import ops, _native as NV
dev = torch.device("cuda")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
gal = torch.empty((rows, 512), dtype=torch.bfloat16, device=dev)
for b in range(0, rows, 1_000_000):
    gen = torch.Generator(device=dev).manual_seed(9000 + b)
    n = min(1_000_000, rows - b)
    gal[b:b + n] = ops.normalize_rows(torch.randn((n, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP, torch.bfloat16)
gen = torch.Generator(device=dev).manual_seed(5)
for nq in (4096, 256):
    q = torch.randn((nq, 512), generator=gen, device=dev)
    src = torch.randint(0, rows, (nq,), generator=gen, device=dev)
    q[nq // 10:] = gal[src[nq // 10:]].float() + 0.03 * q[nq // 10:]
    ref = None
    for ut in ("", "4096", "1024", "256", "128", "64", "32", ""):
        if ut:
            os.environ["FRB_TC_UNIT_TILES"] = ut
        else:
            os.environ.pop("FRB_TC_UNIT_TILES", None)
        out = ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP)
        if ref is None:
            ref = out
        assert torch.equal(out[1], ref[1]) and torch.equal(out[0], ref[0])
        NV.profile_enable(True); NV.profile_read(NV.K_COSINE_TC)
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.cosine_topk(q, gal, 5, qnorm_mode=NV.FRB_QNORM_CLAMP); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        kms, kn = NV.profile_read(NV.K_COSINE_TC); NV.profile_enable(False)
        tf = 2.0 * nq * rows * 512 / (kms / 3 * 1e-3) / 1e12
        print(f"{nq} q x {rows} rows, unit = {ut or 'planner'} tiles: call {sorted(ts)[1]:.2f} ms, kernel {kms / 3:.2f} ms = {tf:.0f} TFLOP/s", flush=True)
