#!/usr/bin/env python
"""bench.py — the identification stage's headline benchmark (BASELINE.json: "queries/sec @1M-512d cosine top-5").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (configs[2]): FaceNet 512-d cosine top-5, 1 000 000-row bf16 gallery, 4096-query batch.  A step is one
pass of the hot path over one 4096-query batch: L2-normalise + bf16 tcgen05 similarity + in-TMEM top-5
(+ one NCCL all-gather and a merge when the gallery is sharded across N GPUs).  Data is synthetic
(SURVEY.md §8d): unit-norm Gaussian gallery rows, 90 % planted queries (source + 0.03 noise), 10 % random.

N > 1: the SAME 1M gallery is sharded by identity over the ranks (strong scaling); queries are replicated;
each rank reports global ids; one all-gather of [Q, 5] candidates; every rank merges.

One JSON line on stdout (rank 0).  Extra legs inside it: `e2e` (host buffers through RecognitionEngine-level
API, H2D/D2H inside the timed region), `roofline` (the tcgen05 kernel alone, event-timed per launch by
libfrb200's frb_profile_* hooks), `cpu_baseline` (oracle port of the reference's batched numpy path on the
host cores, rank 0 at N=1 only), `lbph` (K2/K3 secondary numbers with their own rooflines).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GALLERY = 1_000_000
N_QUERY = 4096
DIM = 512
TOPK = 5
BLOCK_ROWS = 65536          # generation granularity: block b is seeded with 1234 + b on every rank / world size
L2_FLUSH_BYTES = 256 << 20


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) >= 7 and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_gallery_and_queries(torch, ops, NV, device, lo, hi, n_gallery, n_query):
    """This rank's bf16 shard [lo, hi) of the logical gallery + the full replicated fp32 query batch."""
    gen_q = torch.Generator(device=device).manual_seed(4321)
    src = torch.randint(0, n_gallery, (n_query,), generator=gen_q, device=device)
    noise = torch.randn((n_query, DIM), generator=gen_q, device=device)
    queries = torch.empty((n_query, DIM), dtype=torch.float32, device=device)
    shard = torch.empty((hi - lo, DIM), dtype=torch.bfloat16, device=device)
    for b in range((n_gallery + BLOCK_ROWS - 1) // BLOCK_ROWS):
        r0, r1 = b * BLOCK_ROWS, min((b + 1) * BLOCK_ROWS, n_gallery)
        gen = torch.Generator(device=device).manual_seed(1234 + b)
        rows = ops.normalize_rows(torch.randn((r1 - r0, DIM), generator=gen, device=device), NV.FRB_QNORM_CLAMP)
        sel = (src >= r0) & (src < r1)
        if bool(sel.any()):
            queries[sel] = rows[src[sel] - r0] + 0.03 * noise[sel]
        a, e = max(lo, r0), min(hi, r1)
        if a < e:
            shard[a - lo:e - lo] = ops.normalize_rows(rows[a - r0:e - r0].contiguous(), NV.FRB_QNORM_NONE, torch.bfloat16)
    n_rand = n_query // 10
    queries[:n_rand] = torch.randn((n_rand, DIM), generator=gen_q, device=device)
    return shard, queries, src, n_rand


def cpu_topk_port(queries_f32, gallery_f32, k):
    """Oracle port of the reference's batched path (np.dot + top-k, notebooks/evaluate_arcface_kaggle.ipynb:618,713)."""
    from oracle import cosine as OC
    return OC.batched_topk_fast(OC.l2_normalize(queries_f32), gallery_f32, k)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is Python and
    cannot travel to the GPU box), all host threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    rng = np.random.default_rng(1234)
    sample_q = 64
    gal = rng.standard_normal((N_GALLERY, DIM), dtype=np.float32)
    gal /= np.linalg.norm(gal, axis=1, keepdims=True)
    q = gal[rng.integers(0, N_GALLERY, sample_q)] + 0.03 * rng.standard_normal((sample_q, DIM), dtype=np.float32)
    for _ in range(args.warmup):
        cpu_topk_port(q, gal, TOPK)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s, i = cpu_topk_port(q, gal, TOPK)
    dt = time.perf_counter() - t0
    val = sample_q * args.steps / dt
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": "queries/sec @1M-512d cosine top-5", "value": val, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2]: 1M x 512 gallery, cosine top-5 (CPU step = 64-query sample of the 4096 batch)",
                       "gallery_rows": N_GALLERY, "dim": DIM, "k": TOPK, "queries_per_step": sample_q},
            "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"{sample_q} queries x full 1M fp32 gallery per step, numpy sgemm + argpartition, {cores} threads"},
            "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def lbph_leg(torch, ops, NV, device, peaks):
    """Secondary: K2 (LBP + grid histogram) and K3 (chi-square scan) with their own rooflines."""
    out = {}
    gen = torch.Generator(device=device).manual_seed(2024)
    n_faces = 65536
    faces = torch.randint(0, 256, (n_faces, 112, 112), generator=gen, device=device, dtype=torch.uint8)
    for _ in range(3):
        hist, px = ops.lbp_hist(faces)
    torch.cuda.synchronize()
    NV.profile_enable(True)
    for _ in range(5):
        hist, px = ops.lbp_hist(faces)
    ms, n = NV.profile_read(NV.K_LBP_HIST)
    per = ms / n
    bytes_per_face = 112 * 112 + 16384 * 2
    gbs = n_faces * bytes_per_face / (per * 1e-3) / 1e9
    out["extract"] = {"faces_per_s": n_faces / (per * 1e-3), "ms_per_launch": per, "faces_per_launch": n_faces,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": gbs / peaks["hbm_gbs"], "traffic": None}}
    # K3: 64 query histograms against a 100k-row u16 gallery (3.3 GB), each query streams the gallery
    n_gal, n_q = 100_000, 64
    gal = hist.view(torch.int16)[torch.randint(0, n_faces, (n_gal,), generator=gen, device=device)].contiguous().view(torch.uint16)
    qh = hist[:n_q].contiguous()
    for _ in range(2):
        d, i = ops.chisq_topk(qh, px, gal, px, 1)
    NV.profile_read(NV.K_CHISQ)
    for _ in range(3):
        d, i = ops.chisq_topk(qh, px, gal, px, 1)
    ms, n = NV.profile_read(NV.K_CHISQ)
    NV.profile_enable(False)
    per = ms / n
    pairs = n_gal * n_q
    gbs = pairs * 32768 / (per * 1e-3) / 1e9
    out["match"] = {"pairs_per_s": pairs / (per * 1e-3), "predicts_per_s_at_100k_gallery": n_q / (per * 1e-3), "ms_per_launch": per,
                    "queries": n_q, "gallery_rows": n_gal,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                                 "note": "algorithmic bytes = 32768 B per (query, gallery row) pair: every query streams the gallery"}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gallery", type=int, default=N_GALLERY)
    ap.add_argument("--queries", type=int, default=N_QUERY)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lbph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import facerecognition_b200 as F  # loads libfrb200.so or raises
    from facerecognition_b200 import _native as NV
    from facerecognition_b200 import ops
    from facerecognition_b200.sharded import cosine_sharded, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = measured_peaks()
    n_gallery, n_query = args.gallery, args.queries
    lo, hi = shard_bounds(n_gallery, world, rank)
    shard, q_dev, src, n_rand = make_gallery_and_queries(torch, ops, NV, device, lo, hi, n_gallery, n_query)
    search = cosine_sharded(shard, lo, qnorm_mode=NV.FRB_QNORM_CLAMP)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: inputs already in HBM -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        s, i = search.search(q_dev, TOPK)
    barrier()
    # correctness of what is being timed: planted queries must come back as their source row
    ok = bool(torch.equal(i[n_rand:, 0], src[n_rand:]))
    NV.profile_enable(True)
    NV.profile_read(NV.K_COSINE_TC)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        for a, b in ev:
            flush.zero_()                      # L2 flush between timed iterations (outside the event bracket)
            a.record()
            s, i = search.search(q_dev, TOPK)
            b.record()
        barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    k_ms, k_n = NV.profile_read(NV.K_COSINE_TC)
    NV.profile_enable(False)
    value = n_query * args.steps / (total_ms * 1e-3)
    # the step launches cosine_tc_kernel twice (threshold warm-up pass over ~1/64 of the shard + main pass);
    # achieved = the step's algorithmic flops / the summed device time of those launches
    per_launch_ms = k_ms / args.steps
    flops_per_launch = 2.0 * n_query * (hi - lo) * DIM
    achieved = flops_per_launch / (per_launch_ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops"]
    launches_per_step = 2 + k_n // max(args.steps, 1) + (1 if world > 1 else 0)   # normalize_rows, cosine_tc x passes, topk_merge (+ cross-rank merge)

    # ---- end-to-end leg: host buffers, H2D + D2H inside the timed region ---------------------------
    q_host = q_dev.cpu().pin_memory()
    out_s = torch.empty((n_query, TOPK), dtype=torch.float32).pin_memory()
    out_i = torch.empty((n_query, TOPK), dtype=torch.int64).pin_memory()
    q_stage = torch.empty_like(q_dev)

    def e2e_step():
        q_stage.copy_(q_host, non_blocking=True)
        s, i = search.search(q_stage, TOPK)
        out_s.copy_(s, non_blocking=True)
        out_i.copy_(i, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = n_query * args.steps / float(e2e_s.item())
    ok = ok and bool((out_i[n_rand:, 0] == src[n_rand:].cpu()).all())

    line = {
        "metric": "queries/sec @1M-512d cosine top-5", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[2]: FaceNet 512-d cosine top-5, 1M-row bf16 gallery, 4096-query batch",
                   "gallery_rows": n_gallery, "queries_per_step": n_query, "dim": DIM, "k": TOPK,
                   "sharding": f"gallery rows by identity over {world} rank(s); 1 NCCL all-gather of [Q,5] candidates" if world > 1 else "none",
                   "l2": "flushed between timed steps (256 MiB memset outside the event bracket); shard >= L2"},
        "e2e": {"value": e2e_val, "unit": "queries/s", "h2d_bytes_per_step": n_query * DIM * 4,
                "d2h_bytes_per_step": n_query * TOPK * 12,
                "note": "host fp32 queries (pinned) -> device, search, (score, id) lists -> host; gallery resident in HBM as engine state"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "cosine_tc_kernel", "ms_per_step_in_kernel": per_launch_ms, "launches_timed": k_n, "launches_per_step": k_n // max(args.steps, 1),
                     "flops_per_step": flops_per_launch, "peak_source": peaks["source"] + ", bf16 burst"},
        "clocks": clocks.summary(),
        "planted_top1_correct": ok,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded CPU sample: 256-query chunks of the same batch against the full gallery until ~12 s
        gal_f32 = shard.float().cpu().numpy()
        qh = q_host.numpy()
        done, t0 = 0, time.perf_counter()
        while done < n_query and time.perf_counter() - t0 < 12.0:
            cpu_topk_port(qh[done:done + 256], gal_f32, TOPK)
            done += 256
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": done / dt, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{done} of the {n_query} queries x full 1M fp32 gallery, numpy sgemm + argpartition (oracle.cosine.batched_topk_fast)"}
        del gal_f32
    if rank == 0 and world == 1 and not args.no_lbph:
        try:
            line["lbph"] = lbph_leg(torch, ops, NV, device, peaks)
        except Exception as e:  # the secondary leg must never cost the headline line
            line["lbph"] = {"error": repr(e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
