"""Does a step launched kernel by kernel take longer after the same step has been captured into a CUDA graph?"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from facerecognition_b200 import ops, _native as NV
from facerecognition_b200.sharded import cosine_sharded

dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
gal = ops.normalize_rows(torch.randn((1_000_000, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP, torch.bfloat16)
q = torch.randn((4096, 512), generator=gen, device=dev)
search = cosine_sharded(gal, 0, qnorm_mode=NV.FRB_QNORM_CLAMP)

def timed(label, graph, prof, n=10):
    NV.profile_enable(prof)
    for _ in range(3):
        search.search(q, 5, graph=graph)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    t0 = time.perf_counter()
    for a, b in ev:
        a.record(); search.search(q, 5, graph=graph); b.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    kms, kn = NV.profile_read(NV.K_COSINE_TC) if prof else (0, 0)
    NV.profile_enable(False)
    print(f"{label:40s} device ms/step {sum(a.elapsed_time(b) for a, b in ev) / n:7.3f}  host launch ms/step {host:7.3f}  wall {wall:7.3f}  kernel ms/step {kms / n:6.3f} ({kn} launches)")

timed("eager, before any graph", False, False)
timed("eager + profile, before any graph", False, True)
timed("graph", True, False)
timed("eager, after graph", False, False)
timed("eager + profile, after graph", False, True)
timed("graph again", True, False)
