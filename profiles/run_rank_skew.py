"""Why a sharded step takes longer than the slowest rank's AVERAGE kernel time: per-step kernel durations on every rank
(event pairs inside libfrb200) next to the step time, for the configs[3] shape (100 M rows / N per GPU, 4096 queries).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29530 profiles/run_rank_skew.py [steps] [rows] [queries]
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
from facerecognition_b200 import ops, _native as NV                         # noqa: E402
from facerecognition_b200.sharded import cosine_sharded, shard_bounds        # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000          # gallery rows of the whole job
NQ = int(sys.argv[3]) if len(sys.argv) > 3 else 4096                 # queries per step
lo, hi = shard_bounds(N, world, rank)
shard = torch.empty((hi - lo, 512), dtype=torch.bfloat16, device=dev)
for b in range(lo, hi, 1_000_000):
    gen = torch.Generator(device=dev).manual_seed(9000 + b)
    n = min(1_000_000, hi - b)
    shard[b - lo:b - lo + n] = ops.normalize_rows(torch.randn((n, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP, torch.bfloat16)
gen = torch.Generator(device=dev).manual_seed(99)
q = torch.randn((NQ, 512), generator=gen, device=dev)
search = cosine_sharded(shard, lo, qnorm_mode=NV.FRB_QNORM_CLAMP)
for _ in range(3):
    search.search(q, 5)
dist.barrier()
torch.cuda.synchronize()
NV.profile_enable(True)
NV.profile_read(NV.K_COSINE_TC)
kern, step = [], []
for _ in range(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    search.search(q, 5)
    b.record()
    torch.cuda.synchronize()
    kern.append(NV.profile_read(NV.K_COSINE_TC)[0])
    step.append(a.elapsed_time(b))
NV.profile_enable(False)
allk = [torch.zeros(steps, dtype=torch.float64, device=dev) for _ in range(world)]
alls = [torch.zeros(steps, dtype=torch.float64, device=dev) for _ in range(world)]
dist.all_gather(allk, torch.tensor(kern, dtype=torch.float64, device=dev))
dist.all_gather(alls, torch.tensor(step, dtype=torch.float64, device=dev))
if rank == 0:
    K = torch.stack(allk).cpu()          # [rank, step]
    S = torch.stack(alls).cpu()
    print(f"{world} GPUs, {hi - lo} rows per GPU, {NQ} queries, {steps} steps; kernel ms per rank (mean, min..max):")
    for r in range(world):
        print(f"  rank {r}: {K[r].mean():7.2f}  {K[r].min():7.2f} .. {K[r].max():7.2f}")
    print(f"slowest rank's MEAN kernel: {K.mean(1).max():.2f} ms; mean over steps of the per-step MAX over ranks: {K.max(0).values.mean():.2f} ms; "
          f"step time (mean over steps, max over ranks): {S.max(0).values.mean():.2f} ms")
dist.destroy_process_group()
