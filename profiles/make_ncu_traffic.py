"""Regenerates profiles/ncu_traffic.json from the committed round-2 ncu summaries (profiles/r2_prof_*_r2.txt):
dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel, for the shapes bench.py times."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(name):
    txt = open(os.path.join(HERE, name)).read()
    out = []
    for block in txt.split("\n## ")[1:]:
        if block.startswith("hottest") or block.startswith("executed"):
            continue
        rd = re.search(r"dram__bytes_read.sum = ([\d.]+) (\w+)", block)
        wr = re.search(r"dram__bytes_write.sum = ([\d.]+) (\w+)", block)
        t = re.search(r"gpu__time_duration.sum = ([\d.]+) (\w+)", block)
        if rd and wr:
            out.append({"kernel": block.split("(")[0].strip(), "read": float(rd.group(1)) * UNIT[rd.group(2)],
                        "write": float(wr.group(1)) * UNIT[wr.group(2)], "time": f"{t.group(1)} {t.group(2)}" if t else None})
    return out


def entry(name, algorithmic, what, pick=None):
    ls = launches(name)
    ls = ls if pick is None else [ls[i] for i in pick]
    total = sum(l["read"] + l["write"] for l in ls)
    return {"bytes_per_step": int(total), "algorithmic_bytes_per_step": int(algorithmic),
            "source": f"profiles/{name}: " + "; ".join(f"{l['kernel']} {l['read'] / 1e6:.1f} MB read + {l['write'] / 1e6:.1f} MB written in {l['time']}" for l in ls) + f" ({what})"}


doc = {
    "_doc": "dram__bytes_read.sum + dram__bytes_write.sum per bench STEP / launch of each kernel, from the round-2 ncu --set full captures "
            "summarised beside this file (profiles/ncu_round2.sh, profiles/run_ncu_targets.py); regenerate with profiles/make_ncu_traffic.py",
    "cosine_tc_kernel": entry("r2_prof_tc_r2.txt", 1_000_000 * 1024 + 4096 * 2048 + 4096 * 5 * 12, "4096 q x 1M bf16 rows, k = 5: one launch per step"),
    "lbp_hist_kernel": entry("r2_prof_lbp_r2.txt", 65536 * (112 * 112 + 32768), "65536 faces of 112x112, u16 counts out"),
    "lbp_hist_kernel_u8": entry("r2_prof_lbp8_r2.txt", 65536 * (112 * 112 + 16384), "65536 faces of 112x112, u8 counts out"),
    "chisq_kernel": entry("r2_prof_chisq_b_u8_r2.txt", 64 * 100_000 * 16384, "batched exact scan, 64 queries x 100k u8 rows: the chunk is shared through L2"),
    "chisq_kernel_q1_u16": entry("r2_prof_chisq_q1_u16_r2.txt", 100_000 * 32768, "one query x 100k u16 rows"),
    "chisq_kernel_q1_u8": entry("r2_prof_chisq_q1_u8_r2.txt", 100_000 * 16384, "one query x 100k u8 rows"),
    "chisq_filter_kernel_256x37888": entry("r2_prof_filter_r2.txt", 37888 * 16384 + 256 * 16384 * 16, "256 queries x 37888 u8 rows: gallery counts + query features"),
}
json.dump(doc, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: v["bytes_per_step"] for k, v in doc.items() if k != "_doc"}, indent=1))
