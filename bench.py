#!/usr/bin/env python
"""bench.py — the identification stage's headline benchmark (BASELINE.json: "queries/sec @1M-512d cosine top-5").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (configs[2]): FaceNet 512-d cosine top-5 against a 1 000 000-row bf16 gallery.  A step is one pass of
the hot path over one query batch: L2-normalise + bf16 tcgen05 similarity + in-TMEM top-5 (+ one NCCL all-gather
of the candidates and a merge when the gallery is sharded).  Data is synthetic: unit-norm Gaussian
gallery rows, 90 % planted queries (source row + 0.03 noise), 10 % random.

N = 1: 4096 queries x 1M rows.  N > 1 (weak scaling, work per GPU fixed): the SAME 1M gallery is sharded by
identity over the ranks and the batch grows to 4096*N queries, so every rank still multiplies 4096*N queries by
1M/N rows; candidates carry global ids; one all-gather of [Q, 5] (score, id) lists; every rank merges.
`value` = queries answered per second by the whole job.

One JSON line on stdout (rank 0).  Extra legs inside it: `e2e` (host buffers: each rank uploads its 1/N slice of
the batch from pinned memory, an all-gather replicates the queries over NVLink, results come back to the host,
all inside the timed region), `roofline` (the tcgen05 kernel alone, event-timed per launch inside libfrb200),
`cpu_baseline` (oracle port of the reference's batched numpy path on the host cores, rank 0 at N=1 only),
`lbph` (N = 1: K2, K3 in predict mode (one query: HBM) and batched mode (exact: FP32 pipe; tensor-core filter: tensor
pipe), each with the roofline that actually bounds it), `strong` (N > 1: the 4096-query batch on the sharded gallery,
i.e. total work fixed), `c4` (configs[3]: a 100M-row bf16 gallery sharded over the N ranks, 4096 / 256 / 1 queries)
and `c5` (configs[4]: 1024 frames, LBPH extract + chi-square NN against 1M histograms sharded over the N ranks),
`engine_e2e` (N = 1: configs[1] through RecognitionEngine.recognize_embeddings with host numpy in and out).
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
N_QUERY = 4096              # per GPU
DIM = 512
TOPK = 5
BLOCK_ROWS = 65536          # generation granularity: block b is seeded with 1234 + b on every rank / world size
L2_FLUSH_BYTES = 256 << 20
METRIC = "queries/sec @1M-512d cosine top-5"
C4_ROWS = 100_000_000       # configs[3]
C4_BLOCK = 1 << 20          # generation granularity of the 100M gallery (global block b is seeded with 9000 + b)
C5_ROWS = 1_000_000         # configs[4]
C5_FRAMES = 1024
C5_CHUNK = 16384            # faces generated per chunk (global chunk c is seeded with 500 + c)
FP32_PEAK_TFLOPS = 72.0     # measured on this GPU by profiles/micro/fp32_peak.cu (FFMA and FFMA2 alike), DESIGN.md §3


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}
    try:
        d = json.load(open(p))
        burst = float(d["bf16_tflops"])
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": burst,
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained") or burst), "source": "measured (MEASURED_PEAKS.json)"}
    except (OSError, ValueError, KeyError, TypeError):     # absent or incomplete: the recipe's stated numbers
        return fallback


def ncu_traffic(kernel):
    """{"traffic": DRAM bytes (read + written) of `kernel` over the launches `achieved` is computed on, "traffic_detail": ...}
    from the committed `ncu --set full` capture (profiles/ncu_traffic.json); traffic None if there is no capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))[kernel]
        return {"traffic": d["bytes_per_step"], "traffic_detail": d}
    except Exception:
        return {"traffic": None}


class ClockSampler:
    """SM clock + throttle reasons sampled every 25 ms through NVML while work runs (nvidia-smi -lms 200 as fallback)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.stop, self.thread, self.proc, self.max_mhz = index, [], threading.Event(), None, None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:                                   # NVML enumerates physical GPUs: match torch's device by UUID
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def pump():
                while not self.stop.is_set():
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        why = reasons_fn(h)
                        util = pynvml.nvmlDeviceGetUtilizationRates(h).gpu
                        self.rows.append((float(mhz), int(why), int(util)))
                    except Exception:
                        pass
                    self.stop.wait(0.025)
            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
        except Exception:
            self._smi()
        return self

    def _smi(self):
        fields = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={fields}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)

            def pump():
                for line in self.proc.stdout:
                    r = [x.strip() for x in line.split(",")]
                    if len(r) >= 6 and r[0].replace(".", "").isdigit():
                        self.max_mhz = float(r[1])
                        why = sum(bit for (_, bit), v in zip(self.REASONS, r[2:6]) if v.lower().startswith("active"))
                        self.rows.append((float(r[0]), why, 100))
            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        rows = list(self.rows)
        busy = [r for r in rows if r[2] > 0] or rows        # samples taken while the GPU was running kernels
        sm = [r[0] for r in busy]
        why = 0
        for r in busy:
            why |= r[1]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in self.REASONS if why & bit], "samples": len(sm)}


def warm_in_lockstep(step, sync, min_steps, min_seconds, any_rank_wants_more):
    """Warm-up in rounds of 16 steps until >= min_steps steps have run and >= min_seconds have passed, WITH THE DECISION
    TO STOP AGREED BETWEEN THE RANKS: a sharded step ends in the exchange, which waits for every rank, so all ranks must
    run the same number of steps.  (Each rank deciding on its own clock — as this loop did before — stops the ranks in
    different rounds whenever a round boundary falls between their start times; the rank that runs one round more then
    spins in the exchange until its ~10 s time-outs, and the epochs stay out of step for the rest of the run.)
    any_rank_wants_more(flag) -> bool is the all-reduce (max) of the ranks' flags; identity on one GPU.  Returns the
    number of steps run (the same on every rank)."""
    t0, done = time.perf_counter(), 0
    while True:
        for _ in range(16):
            step()
        done += 16
        sync()
        more = done < min_steps or time.perf_counter() - t0 < min_seconds
        if not any_rank_wants_more(bool(more)):
            return done


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


def workload_config(world, n_gallery, n_query_total):
    return {"workload": "configs[2]: FaceNet 512-d cosine top-5, 1M-row bf16 gallery, 4096-query batch per GPU",
            "gallery_rows": n_gallery, "queries_per_step": n_query_total, "dim": DIM, "k": TOPK,
            "sharding": (f"gallery rows by identity over {world} ranks, batch = 4096 x {world} queries; "
                         "per-rank [Q,5] candidates exchanged once and merged on every rank") if world > 1 else "none",
            "l2": "flushed between timed steps (256 MiB memset outside the event bracket); shard >= L2"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is Python and
    cannot travel to the GPU box), all host threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is one process that should use the whole host
    cores = os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(cores)
    import numpy as np
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
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
    cfg = workload_config(world, N_GALLERY, N_QUERY * world)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": (f"each step = a {sample_q}-query sample of the batch x the full 1M fp32 gallery: numpy sgemm on {cores} threads + "
                                        f"argpartition spread over {cores} threads by query row (oracle.cosine.batched_topk_fast); the "
                                        "reference's own per-query Python loop is ~20x slower still (BASELINE.md)")},
            "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def synthetic_faces(torch, n, h, w, device, seed=2024):
    """SURVEY §8d mix: 1/3 uniform noise with a 255 stripe, 1/3 blurred noise, 1/3 piecewise-flat / saturated patches."""
    gen = torch.Generator(device=device).manual_seed(seed)
    faces = torch.randint(0, 255, (n, h, w), generator=gen, device=device, dtype=torch.uint8)
    third = n // 3
    faces[:third, 10:20, :] = 255
    x = torch.rand((third, 1, h, w), generator=gen, device=device)
    k = torch.ones((1, 1, 5, 5), device=device) / 25
    faces[third:2 * third] = (torch.nn.functional.conv2d(x, k, padding=2)[:, 0] * 255).to(torch.uint8)
    coarse = torch.randint(0, 4, (n - 2 * third, 1, (h + 15) // 16, (w + 15) // 16), generator=gen, device=device).float() * 85
    faces[2 * third:] = torch.nn.functional.interpolate(coarse, size=(h, w), mode="nearest")[:, 0].to(torch.uint8)
    return faces


def blocky_faces(torch, n, side, device, seed):
    """Gray faces with structure at two scales (a 4x-upsampled random field + pixel noise): LBP count spread closer to
    real faces than pure noise; the generator of tests/test_gpu_chisq_filter.py."""
    g = torch.Generator(device=device).manual_seed(seed)
    base = torch.randint(0, 256, (n, side // 4 + 2, side // 4 + 2), generator=g, device=device).float()
    up = base.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :side, :side]
    return (up + 12.0 * torch.randn((n, side, side), generator=g, device=device)).clamp(0, 255).to(torch.uint8)


def event_ms(torch, fn, reps, flush=None):
    """Median device time of fn() over `reps` runs (CUDA events on the current stream; optional L2 flush before each)."""
    times = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return statistics.median(times)


def lbph_leg(torch, ops, NV, device, peaks, flush):
    """Secondary (N = 1): K2 (LBP + grid histogram), K3 in its three regimes, the front end, the C5-shaped step."""
    out = {}
    gen = torch.Generator(device=device).manual_seed(2024)
    n_faces = 65536
    faces = synthetic_faces(torch, n_faces, 112, 112, device)
    for _ in range(3):
        hist, px = ops.lbp_hist(faces)
    torch.cuda.synchronize()
    NV.profile_enable(True)
    NV.profile_read(NV.K_LBP_HIST)
    for _ in range(5):
        hist, px = ops.lbp_hist(faces)
    ms, n = NV.profile_read(NV.K_LBP_HIST)
    per = ms / n
    bytes_per_face = 112 * 112 + 16384 * 2
    gbs = n_faces * bytes_per_face / (per * 1e-3) / 1e9
    out["extract"] = {"faces_per_s": n_faces / (per * 1e-3), "ms_per_launch": per, "faces_per_launch": n_faces,
                      "faces": "112x112 u8: 1/3 noise + 255 stripe, 1/3 blurred noise, 1/3 flat patches",
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": gbs / peaks["hbm_gbs"], **ncu_traffic("lbp_hist_kernel"),
                                   "note": "algorithmic bytes = 112*112 + 32768 per face; the kernel is issue-bound (DESIGN.md §3)"}}
    # ---- K3, predict mode: ONE query streams the gallery (what cv2's predict() does per call): HBM-bound --------
    n_gal = 100_000
    pick = torch.randint(0, n_faces, (n_gal,), generator=gen, device=device)
    gal16 = hist.view(torch.int16)[pick].contiguous().view(torch.uint16)
    gal8 = ops.compact_histograms(gal16, px)
    q1 = hist[:1].contiguous()
    for name, gal, row_bytes in (("predict_q1_u16_gallery", gal16, 32768), ("predict_q1_u8_gallery", gal8, 16384)):
        for _ in range(2):
            ops.chisq_topk(q1, px, gal, px, 1)
        ms1 = event_ms(torch, lambda: ops.chisq_topk(q1, px, gal, px, 1), 5, flush)
        NV.profile_read(NV.K_CHISQ)
        flush.zero_()
        ops.chisq_topk(q1, px, gal, px, 1)
        kms, kn = NV.profile_read(NV.K_CHISQ)
        gbs = n_gal * row_bytes / (kms / kn * 1e-3) / 1e9
        out[name] = {"predicts_per_s": 1e3 / ms1, "ms_per_call": ms1, "ms_in_kernel": kms / kn, "gallery_rows": n_gal,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                  **ncu_traffic("chisq_kernel_q1_u16" if row_bytes == 32768 else "chisq_kernel_q1_u8"),
                                  "note": f"physical bytes: one pass over the {n_gal} x {row_bytes} B gallery (L2 flushed before the call; "
                                          "the gallery is 13-26x the L2)"}}
    # ---- K3, batched EXACT scan (64 queries share each gallery chunk through L2): FP32-pipe bound ------------------
    n_q = 64
    qh = hist[:n_q].contiguous()
    saved = ops.FILTER_ENABLED
    ops.FILTER_ENABLED = False
    try:
        for _ in range(2):
            d, i = ops.chisq_topk(qh, px, gal8, px, 1)
        NV.profile_read(NV.K_CHISQ)
        for _ in range(3):
            d, i = ops.chisq_topk(qh, px, gal8, px, 1)
        ms, n = NV.profile_read(NV.K_CHISQ)
    finally:
        ops.FILTER_ENABLED = saved
    per = ms / n
    pairs = n_gal * n_q
    tf = pairs * 65536 / (per * 1e-3) / 1e12
    out["match_exact_batched"] = {"pairs_per_s": pairs / (per * 1e-3), "predicts_per_s_at_100k_gallery": n_q / (per * 1e-3),
                                  "ms_per_launch": per, "queries": n_q, "gallery_rows": n_gal, "gallery": "u8 counts (the form LBPHFaceRecognizer keeps)",
                                  "roofline": {"bound": "fp32", "achieved": tf, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                                               "frac": tf / FP32_PEAK_TFLOPS, **ncu_traffic("chisq_kernel"),
                                               "note": "SURVEY §8d: 65536 flop per (query, row) pair against the measured FP32 peak "
                                                       "(72 TFLOP/s, profiles/micro/fp32_peak.cu); the gallery chunk is shared through L2, "
                                                       "so DRAM traffic (`traffic`) is a small fraction of pairs x 16 KiB and HBM is not the bound. "
                                                       "The peak counts an FMA as 2 flop while the exact formula issues ~5 FP32 instructions + half a "
                                                       "reciprocal per bin, few of them FMAs: ncu has the FP32 pipe 61 % and the MUFU pipe 48 % busy "
                                                       "(profiles/r2_prof_chisq_b_u8_r2.txt). Batches of 16+ predicts do not run this kernel: they go "
                                                       "through the tensor-core filter (extract_match_c5_share, c5)"}}
    # front end: interleaved BGR video crops -> gray (3 B read + 1 B written per pixel)
    n_fr = 32768
    bgr = torch.randint(0, 256, (n_fr, 112, 112, 3), generator=gen, device=device, dtype=torch.uint8)
    for _ in range(2):
        gray = ops.bgr_to_gray(bgr)
    NV.profile_read(NV.K_BGR2GRAY)
    for _ in range(5):
        gray = ops.bgr_to_gray(bgr)
    ms, n = NV.profile_read(NV.K_BGR2GRAY)
    gbs = n_fr * 112 * 112 * 4 / (ms / n * 1e-3) / 1e9
    out["bgr2gray"] = {"frames_per_s": n_fr / (ms / n * 1e-3), "ms_per_launch": ms / n, "frames_per_launch": n_fr,
                       "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                                    "note": "algorithmic bytes = 4 per pixel (3 read, 1 written)"}}
    del bgr, gray
    # front end in front of that: camera-sized BGR frames -> cv2.resize(frame, (112, 112)) + BGR2GRAY in one kernel
    n_fr, sh, sw = 8192, 180, 240
    big = torch.randint(0, 256, (n_fr, sh, sw, 3), generator=gen, device=device, dtype=torch.uint8)
    for _ in range(2):
        small = ops.resize_linear(big, (112, 112), to_gray=True)
    NV.profile_read(NV.K_RESIZE)
    for _ in range(5):
        small = ops.resize_linear(big, (112, 112), to_gray=True)
    ms, n = NV.profile_read(NV.K_RESIZE)
    gbs = n_fr * (sh * sw * 3 + 112 * 112) / (ms / n * 1e-3) / 1e9
    out["resize_gray"] = {"frames_per_s": n_fr / (ms / n * 1e-3), "ms_per_launch": ms / n, "frames_per_launch": n_fr,
                          "frame": f"{sh}x{sw}x3 -> 112x112 gray",
                          "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                       "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                                       "note": "algorithmic bytes = the whole source frame read once + the gray crop written"}}
    del big, small, gal16, gal8, faces, hist
    NV.profile_enable(False)
    # C5 shape on one GPU's share: 1024 frames, extract + chi-square NN against 125 000 histograms (1M / 8 GPUs)
    out["extract_match_c5_share"] = c5_step(torch, ops, NV, device, peaks, 125_000, 0, None)
    return out


def c5_gallery(torch, ops, device, lo, hi):
    """u8 LBPH histograms of global gallery faces [lo, hi) (chunk c of C5_CHUNK faces is seeded with 500 + c on any rank)."""
    parts, px = [], None
    for c in range(lo // C5_CHUNK, (hi + C5_CHUNK - 1) // C5_CHUNK):
        r0 = c * C5_CHUNK
        h, px = ops.lbp_hist(blocky_faces(torch, C5_CHUNK, 112, device, 500 + c))
        a, e = max(lo, r0), min(hi, r0 + C5_CHUNK)
        parts.append(ops.compact_histograms(h[a - r0:e - r0].contiguous(), px))
    return torch.cat(parts, 0), px


def c5_step(torch, ops, NV, device, peaks, rows, lo, search_factory, reps=3):
    """configs[4]-shaped step on this rank: C5_FRAMES frames -> LBP histograms (K2) -> chi-square nearest neighbour against
    gallery rows [lo, lo + rows) through the tensor-core filter + exact re-score (-> cross-rank merge when sharded)."""
    gal8, px = c5_gallery(torch, ops, device, lo, lo + rows)
    gen = torch.Generator(device=device).manual_seed(77)
    n_plant = C5_FRAMES // 4
    src = torch.randint(0, min(C5_CHUNK, C5_ROWS), (n_plant,), generator=gen, device=device)      # global rows in chunk 0
    chunk0 = blocky_faces(torch, C5_CHUNK, 112, device, 500)
    planted = (chunk0[src].float() + 6.0 * torch.randn((n_plant, 112, 112), generator=gen, device=device)).clamp(0, 255).to(torch.uint8)
    frames = torch.cat([planted, blocky_faces(torch, C5_FRAMES - n_plant, 112, device, 31337)], 0).contiguous()
    del chunk0
    search = search_factory(gal8, px) if search_factory else None
    stats = torch.zeros(4, dtype=torch.int32, device=device)

    def step():
        qh, qpx = ops.lbp_hist(frames)
        if search is not None:
            return search.search(qh, 1)
        return ops.chisq_top1_filtered(qh, gal8, px, idx_base=lo, stats=stats)

    for _ in range(2):
        d, i = step()
    torch.cuda.synchronize()
    NV.profile_enable(True)
    NV.profile_read(NV.K_CHISQ_FILTER)
    stats.zero_()
    ms = event_ms(torch, step, reps)
    kms, kn = NV.profile_read(NV.K_CHISQ_FILTER)
    NV.profile_enable(False)
    d, i = step()
    # end to end: the frames start in pinned host memory and the (distance, row) answers end there
    frames_host = frames.cpu().pin_memory()
    out_d = torch.empty((C5_FRAMES, 1), dtype=torch.float32).pin_memory()
    out_i = torch.empty((C5_FRAMES, 1), dtype=torch.int64).pin_memory()

    def e2e_once():
        frames.copy_(frames_host, non_blocking=True)
        dd, ii = step()
        out_d.copy_(dd, non_blocking=True)
        out_i.copy_(ii, non_blocking=True)
        torch.cuda.synchronize()

    e2e_once()
    t0 = time.perf_counter()
    for _ in range(reps):
        e2e_once()
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    flop = 2.0 * C5_FRAMES * rows * gal8.shape[1] * 8
    tf = flop / (kms / max(kn, 1) * 1e-3) / 1e12 if kn else 0.0
    st = stats.cpu().tolist()
    return {"faces_per_s": C5_FRAMES / (ms * 1e-3), "ms_per_step": ms, "frames": C5_FRAMES, "gallery_rows": rows,
            "e2e": {"faces_per_s": C5_FRAMES / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "h2d_bytes_per_step": C5_FRAMES * 112 * 112,
                    "d2h_bytes_per_step": C5_FRAMES * 12, "note": "pinned host frames -> GPU, K2 + match, (distance, row) -> pinned host; wall clock on this rank"},
            "gallery": "u8 LBPH histograms (16 KiB per face) of blocky synthetic 112x112 faces; 1/4 of the frames are noisy re-shots "
                       "of gallery faces, 3/4 have no match",
            "planted_top1": (i[:n_plant, 0], src),
            "filter": {"fallback_queries_per_step": st[0] / max(reps, 1), "survivors_per_query": st[1] / max(reps, 1) / C5_FRAMES,
                       "audit_violations": st[3]} if search is None else None,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"],
                         "frac_of_sustained": tf / peaks["bf16_tflops_sustained"], "kernel": "chisq_filter_kernel",
                         "ms_in_kernel": kms / max(kn, 1), "traffic": None,
                         "traffic_note": "ncu capture at 256 queries x 37888 rows (profiles/r2_prof_filter_r2.txt): 0.69 GB read for 0.62 GB of gallery + 0.07 GB of query features",
                         "note": "algorithmic flops = 2 x frames x rows x 16384 bins x 8 fp16 features (the rank-8 feature GEMM that "
                                 "decides every pair; exact re-score of the few survivors and K2 are inside ms_per_step); peak = "
                                 "the measured cuBLAS bf16 figure (fp16 runs at the same tensor rate)"}}


def c4_leg(torch, dist, ops, NV, device, peaks, world, rank, cosine_sharded, shard_bounds, steps=3):
    """configs[3]: 100M-row bf16 gallery sharded by identity over the ranks; the SAME 4096-query batch on every N, so
    ms_per_step x N flat = linear scaling.  Also 256-query and single-query latency through the same sharded search."""
    lo, hi = shard_bounds(C4_ROWS, world, rank)
    free, _ = torch.cuda.mem_get_info(device)
    need = (hi - lo) * DIM * 2 + (6 << 30)
    if free < need:
        return {"skipped": f"shard needs {need >> 30} GiB, {free >> 30} GiB free"}
    shard = torch.empty((hi - lo, DIM), dtype=torch.bfloat16, device=device)
    gen_q = torch.Generator(device=device).manual_seed(99)
    src = torch.randint(0, C4_ROWS, (N_QUERY,), generator=gen_q, device=device)
    noise = torch.randn((N_QUERY, DIM), generator=gen_q, device=device)
    n_rand = N_QUERY // 10
    for b in range(lo // C4_BLOCK, (hi + C4_BLOCK - 1) // C4_BLOCK):
        r0, r1 = b * C4_BLOCK, min((b + 1) * C4_BLOCK, C4_ROWS)
        gen = torch.Generator(device=device).manual_seed(9000 + b)
        rows = ops.normalize_rows(torch.randn((r1 - r0, DIM), generator=gen, device=device), NV.FRB_QNORM_CLAMP, torch.bfloat16)
        a, e = max(lo, r0), min(hi, r1)
        shard[a - lo:e - lo] = rows[a - r0:e - r0]
        del rows
    # planted queries: the owner of a source row contributes (its stored bf16 row + noise), one all-reduce replicates
    queries = torch.zeros((N_QUERY, DIM), dtype=torch.float32, device=device)
    mine = (src >= lo) & (src < hi)
    mine[:n_rand] = False
    queries[mine] = shard[src[mine] - lo].float() + 0.03 * noise[mine]
    if world > 1:
        dist.all_reduce(queries)
    queries[:n_rand] = noise[:n_rand]
    search = cosine_sharded(shard, lo, qnorm_mode=NV.FRB_QNORM_CLAMP)

    def timed(q, reps, graph):
        for _ in range(2):
            s, i = search.search(q, TOPK, graph=graph)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            a.record()
            s, i = search.search(q, TOPK, graph=graph)
            b.record()
        torch.cuda.synchronize()
        t = torch.tensor([statistics.median(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), s, i

    NV.profile_enable(True)
    NV.profile_read(NV.K_COSINE_TC)
    ms, s, i = timed(queries, steps, False)
    k_ms, k_n = NV.profile_read(NV.K_COSINE_TC)
    NV.profile_enable(False)
    ok = torch.tensor([1 if bool(torch.equal(i[n_rand:, 0], src[n_rand:])) else 0], device=device)
    kernel_ms = k_ms / (steps + 2)                       # summed over the cosine_tc launch(es) of one step
    tf = 2.0 * N_QUERY * (hi - lo) * DIM / (kernel_ms * 1e-3) / 1e12
    fr = torch.tensor([tf], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        dist.all_reduce(fr, op=dist.ReduceOp.MIN)
    q256 = queries[n_rand:n_rand + 256].contiguous()
    q1 = queries[n_rand:n_rand + 1].contiguous()
    ms256, _, _ = timed(q256, 5, True)
    ms1, _, i1 = timed(q1, 9, True)
    ok1 = bool(int(i1[0, 0]) == int(src[n_rand]))
    del shard, search
    torch.cuda.empty_cache()
    return {"workload": "configs[3]: 100M x 512 bf16 gallery sharded by identity, 4096-query batch (fixed for every N), top-5",
            "rows_total": C4_ROWS, "rows_per_gpu": hi - lo, "gallery_bytes_per_gpu": (hi - lo) * DIM * 2,
            "ms_per_step": ms, "ms_per_step_x_gpus": ms * world, "queries_per_s": N_QUERY / (ms * 1e-3),
            "roofline": {"bound": "tensor", "kernel": "cosine_tc_kernel", "achieved": float(fr.item()), "unit": "TFLOP/s per GPU (slowest rank)",
                         "peak": peaks["bf16_tflops_sustained"], "frac": float(fr.item()) / peaks["bf16_tflops_sustained"],
                         "frac_of_burst": float(fr.item()) / peaks["bf16_tflops"], "ms_per_step_in_kernel": kernel_ms, "traffic": None,
                         "note": "a step keeps the tensor pipe busy for 0.04-0.3 s: the cuBLAS SUSTAINED figure is the denominator"},
            "latency_ms": {"q4096": ms, "q256": ms256, "q1": ms1,
                           "note": "whole sharded call (local search + NVLink exchange), CUDA-graph replay for 256 / 1 queries; "
                                   "one query streams the shard once: HBM-bound row streaming kernel"},
            "q1_gbs_per_gpu": (hi - lo) * DIM * 2 / (ms1 * 1e-3) / 1e9,
            "planted_top1_correct": bool(ok.item()) and ok1}


def engine_e2e_leg(torch, np, F):
    """configs[1] through the reference-named API: RecognitionEngine.recognize_embeddings, host numpy in, Python tuples out
    (ArcFace 512-d cosine, 10k-identity dict gallery, 256-query batch, fp32 exact path, top-1 + threshold)."""
    rng = np.random.default_rng(7)
    gal = rng.standard_normal((10_000, DIM)).astype(np.float32)
    gal /= np.linalg.norm(gal, axis=1, keepdims=True)
    eng = F.RecognitionEngine(model_path=None, threshold=0.5, use_face_detection=False)
    eng.db = {f"id_{i:05d}": g for i, g in enumerate(gal)}
    src = rng.integers(0, 10_000, 256)
    q = (gal[src] + 0.03 * rng.standard_normal((256, DIM))).astype(np.float32)
    for _ in range(3):
        res = eng.recognize_embeddings(q)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        res = eng.recognize_embeddings(q)
    dt = (time.perf_counter() - t0) / reps
    ok = all(r[0] == f"id_{j:05d}" for r, j in zip(res, src))
    qd = torch.from_numpy(q).cuda()
    for _ in range(3):
        eng.recognize_embeddings(qd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        res_d = eng.recognize_embeddings(qd)
    dt_d = (time.perf_counter() - t0) / reps
    return {"workload": "configs[1]: ArcFace 512-d cosine, 10k-identity dict gallery, 256-query batch, fp32, top-5 + threshold",
            "queries_per_s": 256 / dt, "ms_per_call": dt * 1e3, "h2d_bytes_per_call": 256 * DIM * 4, "d2h_bytes_per_call": 256 * 5 * 12,
            "device_tensor_in": {"queries_per_s": 256 / dt_d, "ms_per_call": dt_d * 1e3},
            "all_identities_correct": bool(ok) and all(r[0] == f"id_{j:05d}" for r, j in zip(res_d, src)),
            "note": "wall clock around RecognitionEngine.recognize_embeddings (numpy in -> list of (name, score, top-5) out), "
                    "gallery resident on the GPU as engine state; device_tensor_in = the same call with a CUDA tensor"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gallery", type=int, default=N_GALLERY)
    ap.add_argument("--queries", type=int, default=N_QUERY, help="queries per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lbph", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip configs[3] (100M-row sharded gallery)")
    ap.add_argument("--no-c5", action="store_true", help="skip configs[4] (LBPH 1024 frames vs 1M sharded histograms)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--balance", action="store_true",
                    help="gallery shards proportional to each GPU's measured rate (3 s calibration probe) instead of equal shards")
    ap.add_argument("--min-warm-seconds", type=float, default=1.0,
                    help="keep warming until this much time has passed under load (clock samples); 0 for ncu launch lists")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import facerecognition_b200 as F  # loads libfrb200.so or raises
    from facerecognition_b200 import _native as NV
    from facerecognition_b200 import ops
    from facerecognition_b200.sharded import HostBatchPipeline, chisq_sharded, cosine_sharded, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = measured_peaks()
    warmup = max(args.warmup, 3)
    n_gallery, q_per_gpu = args.gallery, args.queries
    n_query = q_per_gpu * world                      # weak scaling: the batch grows with the job
    # Several GPUs: a sharded step waits for its slowest rank, and under the power cap the GPUs of one box can differ by
    # 10-30 % (profiles/r2_rank_skew.txt).  With --balance a 3 s probe run on all ranks at once measures each GPU's rate
    # and the gallery rows are split in proportion (sharded.balanced_bounds); the answer does not depend on the split.
    # Opt-in: validated at 2 and 4 GPUs (no gain on boxes without skew); the default stays the equal split of SURVEY 8(e).
    balance = None

    def bounds_of(n_rows):
        return shard_bounds(n_rows, world, rank)

    if world > 1 and args.balance:
        from facerecognition_b200.sharded import balanced_bounds, measure_rank_weights
        gen_p = torch.Generator(device=device).manual_seed(7)
        probe_g = ops.normalize_rows(torch.randn((N_GALLERY, DIM), generator=gen_p, device=device), NV.FRB_QNORM_CLAMP, torch.bfloat16)
        probe_q = torch.randn((N_QUERY, DIM), generator=gen_p, device=device)
        weights = measure_rank_weights(lambda: ops.cosine_topk(probe_q, probe_g, TOPK, qnorm_mode=NV.FRB_QNORM_CLAMP),
                                       torch.cuda.synchronize, 3.0)
        del probe_g, probe_q

        def bounds_of(n_rows):                                              # noqa: F811
            return balanced_bounds(n_rows, weights, rank)

        balance = {"weights": [round(w, 4) for w in weights],
                   "note": "gallery rows per rank proportional to the rank's measured rate on a 3 s probe (4096 x 1M bf16 top-5, back to back, on "
                           "all ranks at once); applies to the headline, c4 and c5 shards (opt-in: --balance; the default is equal shards)"}
    lo, hi = bounds_of(n_gallery)
    shard, q_dev, src, n_rand = make_gallery_and_queries(torch, ops, NV, device, lo, hi, n_gallery, n_query)
    search = cosine_sharded(shard, lo, qnorm_mode=NV.FRB_QNORM_CLAMP)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)

    use_graph = not args.no_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(queries, steps, graph=None):
        """K steps, CUDA events per step on the launching stream, L2 flushed outside the brackets; MAX over ranks (ms)."""
        graph = (use_graph and world > 1) if graph is None else graph
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in ev:
            flush.zero_()
            a.record()
            s, i = search.search(queries, TOPK, graph=graph)
            b.record()
        barrier()
        total = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item()), s, i

    with ClockSampler(local_rank) as clocks:
        # ---- device-resident leg: inputs already in HBM ---------------------------------------------
        def any_rank_wants_more(flag):
            if world == 1:
                return flag
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return bool(int(t.item()))

        warm_out = {}

        def warm_step():
            warm_out["s"], warm_out["i"] = search.search(q_dev, TOPK, graph=use_graph and world > 1)

        # >= W steps and >= 1 s under load (clock samples); every rank runs the same number of steps
        warm_in_lockstep(warm_step, torch.cuda.synchronize, warmup, args.min_warm_seconds, any_rank_wants_more)
        s, i = warm_out["s"], warm_out["i"]
        barrier()
        if world > 1 and search._exchange is not None:
            # belt and braces: a rank that missed a step would leave the exchange epochs out of step (10 s per step from
            # then on); the status word says so, and resync() (collective) restarts the epochs
            missed = torch.tensor([search._exchange.status()[0]], dtype=torch.int32, device=device)
            dist.all_reduce(missed, op=dist.ReduceOp.MAX)
            if int(missed.item()):
                search.resync()
        # correctness of what is being timed: planted queries must come back as their source row
        ok = bool(torch.equal(i[n_rand:, 0], src[n_rand:]))
        # One GPU: the timed steps are launched kernel by kernel, so libfrb200 brackets every cosine_tc_kernel launch of the
        # timed region itself with an event pair on the launching stream (launch gaps are ~1 % of a 2.6 ms step).
        # Several GPUs: the timed steps are graph replays (launch gaps and rank skew matter there), which have no host-side
        # launch to bracket; the kernel time then comes from the same K steps repeated kernel by kernel right after the
        # timed region (a later pass under the power cap can run a few per cent slower than the timed one).
        for _ in range(3):                       # torch.cuda.graph() empties the caching allocator: re-warm the eager path
            search.search(q_dev, TOPK, graph=False)
        def measure():
            NV.profile_enable(True)
            NV.profile_read(NV.K_COSINE_TC)
            if world == 1:
                total, s_, i_ = timed_steps(q_dev, args.steps, graph=False)
                eager = total
            else:
                NV.profile_enable(False)
                total, s_, i_ = timed_steps(q_dev, args.steps)
                NV.profile_enable(True)
                eager, _, _ = timed_steps(q_dev, args.steps, graph=False)
            km, kn = NV.profile_read(NV.K_COSINE_TC)
            NV.profile_enable(False)
            return total, eager, km, kn, s_, i_

        total_ms, eager_ms, k_ms, k_n, s, i = measure()
        # a timed region that saw a hardware / thermal slowdown on any rank is measured again, once
        remeasured = None
        seen = set(clocks.summary()["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if os.environ.get("FRB_BENCH_FORCE_REMEASURE") == "1":          # exercises this path on a healthy box
            seen = seen | {"forced"}
        again = torch.tensor([1 if seen else 0], device=device)
        if world > 1:
            dist.all_reduce(again, op=dist.ReduceOp.MAX)
        if int(again.item()):
            remeasured = {"first_reasons": sorted(seen), "first_value": n_query * args.steps / (total_ms * 1e-3)}
            time.sleep(5.0)
            clocks.rows.clear()
            for _ in range(warmup):
                search.search(q_dev, TOPK, graph=use_graph and world > 1)
            total_ms, eager_ms, k_ms, k_n, s, i = measure()
        value = n_query * args.steps / (total_ms * 1e-3)
        # the step launches cosine_tc_pair_kernel once (the single-CTA kernels, <= 128 queries, add a threshold warm-up pass);
        # achieved = the step's algorithmic flops / the summed device time of its cosine_tc launches
        kernel_ms_per_step = k_ms / args.steps
        flops_per_step = 2.0 * n_query * (hi - lo) * DIM
        achieved = flops_per_step / (kernel_ms_per_step * 1e-3) / 1e12
        tc_per_step = k_n // max(args.steps, 1)
        launches_per_step = 2 + tc_per_step + (1 if world > 1 else 0)   # normalize_rows, cosine_tc passes, compact merge (+ cross-rank merge)

        # ---- end-to-end leg: host buffers, H2D + D2H inside the timed region -------------------------
        # Each rank owns the 1/N slice of the batch that "arrived" at it: pinned host -> its GPU, one all-gather
        # replicates the queries over NVLink, sharded search, each rank returns its slice's answers to the host.
        q0, q1 = rank * q_per_gpu, (rank + 1) * q_per_gpu
        q_host = q_dev[q0:q1].cpu().pin_memory()
        out_s = torch.empty((q_per_gpu, TOPK), dtype=torch.float32).pin_memory()
        out_i = torch.empty((q_per_gpu, TOPK), dtype=torch.int64).pin_memory()
        q_stage = torch.empty((q_per_gpu, DIM), dtype=torch.float32, device=device) if world > 1 else torch.empty_like(q_dev)
        q_stage16 = torch.empty((n_query, DIM), dtype=torch.bfloat16, device=device) if world > 1 else None

        def e2e_step():
            q_stage.copy_(q_host, non_blocking=True)
            if world > 1:
                # each rank normalises ITS slice, one all-gather of bf16 rows, tensor-core search on the gathered batch
                s, i = search.search(search.gather_normalized(q_stage, q_stage16, q0), TOPK, graph=use_graph)
            else:
                s, i = search.search(q_stage, TOPK, graph=use_graph)
            out_s.copy_(s[q0:q1], non_blocking=True)
            out_i.copy_(i[q0:q1], non_blocking=True)
            torch.cuda.synchronize()

        def timed_e2e(run):
            run(3)
            barrier()
            t0 = time.perf_counter()
            run(args.steps)
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return n_query * args.steps / float(t.item())

        def run_sync(n):
            for _ in range(n):
                e2e_step()

        e2e_sync = timed_e2e(run_sync)            # one blocking call per batch: copy in, search, copy out, wait
        ok = ok and bool((out_i[max(n_rand - q0, 0):, 0] == src[q0 + max(n_rand - q0, 0):q1].cpu()).all())

        # the serving form of the same call: sharded.HostBatchPipeline keeps two batches in flight so the copies of
        # one batch overlap the search of the other (every batch's H2D and D2H are still inside the timed region)
        pipe = HostBatchPipeline(search, n_query, DIM, TOPK, device, rows=(q0, q1), graph=use_graph)
        last = {}

        def run_pipelined(n):
            tickets = []
            for _ in range(n):
                tickets.append(pipe.submit(q_host))
                if len(tickets) == pipe.depth:
                    last["s"], last["i"] = pipe.result(tickets.pop(0))
            for t in tickets:
                last["s"], last["i"] = pipe.result(t)

        e2e_val = timed_e2e(run_pipelined)
        out_i = last["i"]
        ok = ok and bool((out_i[max(n_rand - q0, 0):, 0] == src[q0 + max(n_rand - q0, 0):q1].cpu()).all())

        strong = None
        if world > 1:
            # total work fixed: the 4096-query batch against the sharded gallery
            qs = q_dev[:q_per_gpu].contiguous()
            for _ in range(3):
                search.search(qs, TOPK, graph=use_graph)
            ms_s, _, _ = timed_steps(qs, args.steps)
            strong = {"value": q_per_gpu * args.steps / (ms_s * 1e-3), "unit": "queries/s", "ms_per_step": ms_s / args.steps,
                      "queries_per_step": q_per_gpu, "note": "same 1M gallery, batch NOT grown: total work fixed"}

    exchange_kind = search._exchange is not None
    okt = torch.tensor([1 if ok else 0], device=device)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    traffic = ncu_traffic("cosine_tc_kernel") if world == 1 and n_gallery == N_GALLERY and q_per_gpu == N_QUERY else {"traffic": None}
    # denominator: the timed steps follow >= 1 s of back-to-back steps; when the clock samples show the power cap
    # holding the SM clock down (median under load below 90 % of the maximum) the kernel ran in cuBLAS's "sustained"
    # regime, otherwise in its "burst" regime (B200_PROFILING.md)
    clock_summary = clocks.summary()
    sustained = ("sw_power_cap" in clock_summary["reasons"] and clock_summary["sm_mhz"] is not None
                 and clock_summary["sm_max_mhz"] and clock_summary["sm_mhz"] < 0.9 * clock_summary["sm_max_mhz"])
    peak = peaks["bf16_tflops_sustained"] if sustained else peaks["bf16_tflops"]
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world, n_gallery, n_query),
        "e2e": {"value": e2e_val, "unit": "queries/s", "h2d_bytes_per_step": n_query * DIM * 4,
                "d2h_bytes_per_step": n_query * TOPK * 12, "blocking_call_value": e2e_sync,
                "note": "pinned host fp32 queries -> GPU (each rank its 1/N slice, normalised there, bf16 rows all-gathered over "
                        "NVLink), search, (score, id) lists -> host; bytes are whole-job totals; gallery resident in HBM as engine state. "
                        "value: sharded.HostBatchPipeline, two batches in flight (copies of one overlap the search of "
                        "the other); blocking_call_value: one batch at a time, host waits for each"},
        "gpu_launches": launches_per_step * args.steps,
        "launch_mode": ("the step's kernels are replayed from one CUDA graph (ShardedSearch.search(graph=True))" if use_graph and world > 1
                        else "kernel by kernel (one GPU: every launch of the timed region is bracketed for the roofline); the e2e legs replay the step's CUDA graph"),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     **traffic, "kernel": "cosine_tc_kernel", "ms_per_step_in_kernel": kernel_ms_per_step,
                     "launches_timed": k_n, "launches_per_step": tc_per_step, "flops_per_step": flops_per_step,
                     "ms_per_step_kernel_by_kernel": eager_ms / args.steps,
                     "timed_on": ("the timed region's own launches (event pairs around each launch inside libfrb200)" if world == 1 or not use_graph
                                  else "the K steps repeated kernel by kernel right after the graph-replayed timed region"),
                     "frac_of_burst": achieved / peaks["bf16_tflops"], "frac_of_sustained": achieved / peaks["bf16_tflops_sustained"],
                     "peak_source": peaks["source"] + (", bf16 SUSTAINED figure: the timed steps ran power-capped (median SM clock < 90 % of max, see clocks) after >= 1 s of load"
                                                       if sustained else ", bf16 BURST figure: the SM clock stayed near its maximum during the run")},
        "clocks": {**clock_summary, **({"remeasured": remeasured} if remeasured else {})},
        "planted_top1_correct": bool(okt.item()),
    }
    if strong:
        line["strong"] = strong
    if balance:
        balance["rows_this_rank"] = hi - lo
        line["shard_balance"] = balance
    if world > 1:
        # (kept out of `config`, which must read the same in the reference arm)
        line["exchange"] = ("one fused kernel over NVLink peer memory (frb_exchange_topk_merge: peer stores + flags + merge)"
                            if exchange_kind else "one NCCL all-gather of packed records + frb_topk_merge_strided")

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded CPU sample: 256-query chunks of the same batch against the full gallery until ~12 s
        gal_f32 = shard.float().cpu().numpy()
        qh = q_dev.cpu().numpy()
        done, t0 = 0, time.perf_counter()
        while done < n_query and time.perf_counter() - t0 < 12.0:
            cpu_topk_port(qh[done:done + 256], gal_f32, TOPK)
            done += 256
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": done / dt, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{done} of the {n_query} queries x full 1M fp32 gallery, numpy sgemm + argpartition (oracle.cosine.batched_topk_fast)"}
        del gal_f32
    del shard, search, q_dev
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_lbph:
        try:
            line["lbph"] = lbph_leg(torch, ops, NV, device, peaks, flush)
            pl = line["lbph"]["extract_match_c5_share"].pop("planted_top1")
            line["lbph"]["extract_match_c5_share"]["planted_top1_correct"] = bool(torch.equal(pl[0], pl[1]))
        except Exception as e:  # the secondary leg must never cost the headline line
            line["lbph"] = {"error": repr(e)}
        try:
            line["engine_e2e"] = engine_e2e_leg(torch, np, F)
        except Exception as e:
            line["engine_e2e"] = {"error": repr(e)}
    if not args.no_c5:
        try:
            lo5, hi5 = bounds_of(C5_ROWS)
            c5 = c5_step(torch, ops, NV, device, peaks, hi5 - lo5, lo5,
                         (lambda g8, px: chisq_sharded(g8, px, lo5)) if world > 1 else None)
            got, want = c5.pop("planted_top1")
            ok5 = torch.tensor([1 if bool(torch.equal(got, want)) else 0], device=device)
            t5 = torch.tensor([c5["ms_per_step"], -c5["roofline"]["achieved"]], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(ok5, op=dist.ReduceOp.MIN)
                dist.all_reduce(t5, op=dist.ReduceOp.MAX)            # slowest rank's step time, lowest kernel rate
            c5.update({"workload": "configs[4]: LBPH extract + chi-square NN, 1024 frames of 112x112 vs 1M histograms sharded by identity",
                       "rows_total": C5_ROWS, "ms_per_step": float(t5[0]), "faces_per_s": C5_FRAMES / (float(t5[0]) * 1e-3),
                       "ms_per_step_x_gpus": float(t5[0]) * world, "planted_top1_correct": bool(ok5.item())})
            c5["roofline"]["achieved"] = -float(t5[1])
            c5["roofline"]["frac"] = c5["roofline"]["achieved"] / peaks["bf16_tflops"]
            c5["roofline"]["frac_of_sustained"] = c5["roofline"]["achieved"] / peaks["bf16_tflops_sustained"]
            line["c5"] = c5
        except Exception as e:
            line["c5"] = {"error": repr(e)}
        torch.cuda.empty_cache()
    if not args.no_c4:
        try:
            line["c4"] = c4_leg(torch, dist, ops, NV, device, peaks, world, rank, cosine_sharded, lambda n, w, r: bounds_of(n))
        except Exception as e:
            line["c4"] = {"error": repr(e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
