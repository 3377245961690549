#!/usr/bin/env python
"""Turn the raw ncu outputs of a gpurun call into the tracked summaries under profiles/.

    python tools/make_profiles.py --tag r01 --launches gpurun_out/launches_r1.csv --rep gpurun_out/prof_r5.ncu-rep \
        --cmd "python bench.py --steps 3 --warmup 3 --no-cpu-baseline"

Writes profiles/<tag>_launches_raw.csv (copy), <tag>_launch_shares.csv, <tag>_ncu_kernels.md and
<tag>_dram_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, read by bench.py).
"""
import argparse
import csv
import json
import os
import shutil
import subprocess
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}            # bytes / milliseconds


def short(name):
    n = name.split("(")[0].replace("void ", "").strip()
    return n.split("<")[0] if n.startswith("dm2::") or n.startswith("k_") else n.split("<")[0]


def launch_shares(path, out_csv, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name, unit, val = short(r[4]), r[13], float(r[14].replace(",", ""))
        us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    with open(out_csv, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none, cold-cache/serialised: compare SHARES)\n")
        f.write(f"# command: {cmd}   ({len(rows)} launches after the warm-up skip)\n")
        f.write("kernel,launches,total_us,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{v[0]},{v[1]:.1f},{v[1] / tot:.4f}\n")


METRICS = [
    ("time", "gpu__time_duration.sum", "ms"),
    ("regs/thread", "launch__registers_per_thread", None),
    ("warp instructions", "smsp__inst_executed.sum", None),
    ("active threads / instruction", "smsp__thread_inst_executed_per_inst_executed.ratio", None),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active", None),
    ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", None),
    ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", None),
    ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", None),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
    ("DRAM read MB", "dram__bytes_read.sum", "MB"),
    ("DRAM write MB", "dram__bytes_write.sum", "MB"),
    ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", None),
    ("global RED requests (warp level)", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", None),
    ("global ATOM requests (warp level)", "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum", None),
    ("RED sectors to L2 (32 B)", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", None),
]


def kernels(rep, out_md, out_json, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    cols, traffic, winst = [], {}, {}
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        full = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        if "<" in full:
            name += "<" + full.split("<")[1].split(">")[0] + ">"
        col = {"kernel": name}
        for label, key, want in METRICS:
            if key not in hdr:
                continue
            i = hdr.index(key)
            v = float(r[i].replace(",", ""))
            if want == "ms":
                v *= UNIT.get(units[i], 1.0)
            elif want == "MB":
                v *= UNIT.get(units[i], 1.0) / 1e6
            col[label] = v
        st = sorted(((float(r[hdr.index(h)]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stalls), reverse=True)
        col["top stalls (warps per issue)"] = ", ".join(f"{n} {v:.2f}" for v, n in st[:4])
        cols.append(col)
        traffic[name] = (col.get("DRAM read MB", 0.0) + col.get("DRAM write MB", 0.0)) * 1e6
        winst[name] = col.get("warp instructions", 0.0)
    with open(out_md, "w") as f:
        f.write("# ncu --set full --clock-control none (one launch of each kernel of one bench step; cfg2: 1M tris @1024^2)\n")
        f.write(f"# command: {cmd} ; numbers under the profiler are NOT bench values\n\n")
        f.write("| metric | " + " | ".join(c["kernel"] for c in cols) + " |\n")
        f.write("|---|" + "---|" * len(cols) + "\n")
        for label in [m[0] for m in METRICS] + ["top stalls (warps per issue)"]:
            cells = []
            for c in cols:
                v = c.get(label, "")
                cells.append(f"{v:.4g}" if isinstance(v, float) else str(v))
            f.write(f"| {label} | " + " | ".join(cells) + " |\n")
    with open(out_json, "w") as f:
        json.dump({"source": os.path.basename(out_md) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, cfg2)",
                   "bytes_per_launch": traffic,
                   "warp_instructions_per_launch": winst}, f, indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--cmd", default="python bench.py --steps 3 --warmup 3 --no-cpu-baseline")
    a = ap.parse_args()
    pdir = os.path.join(ROOT, "profiles")
    os.makedirs(pdir, exist_ok=True)
    if a.launches:
        shutil.copy(a.launches, os.path.join(pdir, f"{a.tag}_launches_raw.csv"))
        launch_shares(a.launches, os.path.join(pdir, f"{a.tag}_launch_shares.csv"), a.cmd)
    if a.rep:
        kernels(a.rep, os.path.join(pdir, f"{a.tag}_ncu_kernels.md"), os.path.join(pdir, f"{a.tag}_dram_traffic.json"), a.cmd)


if __name__ == "__main__":
    main()
