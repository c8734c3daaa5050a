"""Soak of the sharded step (local tcgen05 search + fused NVLink exchange + merge) under torchrun: tens of thousands of
CUDA-graph replays and eager calls on small shards, so the step is exchange-dominated and the flag / epoch protocol
turns over as fast as it can; every 500th answer is compared with the unsharded one and the exchange's time-out counter
must stay 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 profiles/run_soak.py [seconds]
"""
import os
import sys
import time

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

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
gen = torch.Generator(device=dev).manual_seed(1)
N, k = 20_003, 5
gal = ops.normalize_rows(torch.randn((N, 512), generator=gen, device=dev), NV.FRB_QNORM_CLAMP, torch.bfloat16)
lo, hi = shard_bounds(N, world, rank)
search = cosine_sharded(gal[lo:hi].contiguous(), lo, qnorm_mode=NV.FRB_QNORM_CLAMP)
report = []
for Q, graph in ((1, True), (256, True), (4096, True), (64, False)):
    q = torch.randn((Q, 512), generator=gen, device=dev)
    want_s, want_i = ops.cosine_topk(q, gal, k, qnorm_mode=NV.FRB_QNORM_CLAMP)
    steps, bad = 0, 0
    t_end = time.time() + seconds / 4
    stop = torch.zeros(1, dtype=torch.int32, device=dev)
    while True:
        for _ in range(500):
            s, i = search.search(q, k, graph=graph)
        steps += 500
        if not (torch.equal(i, want_i) and torch.equal(s, want_s)):
            bad += 1
        stop.fill_(1 if time.time() > t_end else 0)                          # every rank leaves the loop at the same step
        dist.all_reduce(stop, op=dist.ReduceOp.MAX)
        if int(stop.item()):
            break
    torch.cuda.synchronize()
    timeouts, epoch = search._exchange.status() if search._exchange is not None else (-1, -1)
    report.append(f"Q={Q} graph={graph}: {steps} steps, {bad} wrong answers, exchange time-outs {timeouts}, epoch {epoch}")
    assert bad == 0 and timeouts == 0, report[-1]
if rank == 0:
    print(f"soak on {world} GPUs, {seconds:.0f} s: " + "; ".join(report))
dist.destroy_process_group()
