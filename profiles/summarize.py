#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep and the launch-list CSV into the small text/CSV summaries committed under profiles/.
Usage: python profiles/summarize.py <round-tag>   (reads gpurun_out/, writes profiles/)"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("FRB_SUMMARY_OUT", os.path.join(ROOT, "profiles"))   # on the GPU box: gpurun_out/summ (only gpurun_out/ comes back)
SRC = os.path.join(ROOT, "gpurun_out")
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_rep(name, tag):
    rep = os.path.join(SRC, name)
    if not os.path.exists(rep):
        return
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {name} ({tag}); one block per captured launch"]
    for r in rows[2:]:
        lines.append(f"\n## {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''}")
        for k in KEYS:
            if k in hdr:
                lines.append(f"{k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h = src[1]
        ia, isrc, iss, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        data = [r for r in src[2:] if len(r) > iss and r[iss].isdigit()]
        tot = sum(int(r[iss]) for r in data)
        lines.append(f"\n## hottest SASS by warp-state samples (total {tot}; first launch)")
        for r in sorted(data, key=lambda r: -int(r[iss]))[:24]:
            lines.append(f"{r[iss]:>8} samples  {r[iex]:>10} exec  {r[ia][-5:]}  {r[isrc][:100]}")
        ops = collections.Counter()
        for r in data:
            for m in ("UTCHMMA", "UTMALDG", "UTMAPF", "LDTM", "UTCBAR", "MUFU", "ATOMS", "HMMA"):
                if m in r[isrc]:
                    ops[m] += int(r[iex]) if r[iex].isdigit() else 0
        lines.append("\n## executed Blackwell-specific / notable opcodes: " + ", ".join(f"{k}={v}" for k, v in sorted(ops.items())))
    with open(os.path.join(OUT, f"{tag}_{name.replace('.ncu-rep', '')}.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


def summarize_launches(tag):
    p = os.path.join(SRC, f"launches_{tag}.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)   # -> microseconds
        name = r[ik].split("(")[0]
        d = per.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v * scale
    tot = sum(d[1] for d in per.values())
    with open(os.path.join(OUT, f"{tag}_launches_summary.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_us,share\n")
        for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{us:.1f},{us / tot:.4f}\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    summarize_launches(tag)
    for name in sorted(os.listdir(SRC)):
        if name.endswith(f"_{tag}.ncu-rep"):
            summarize_rep(name, tag)
