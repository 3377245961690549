#!/usr/bin/env python
"""Per-kernel one-screen summary of an ncu --set full report:  python tools/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = [("time_ms", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
        ("warp_inst", "smsp__inst_executed.sum"), ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("warps_active%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2_hit%", "lts__t_sector_hit_rate.pct")]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    out = [name]
    for label, key in want:
        if key in hdr:
            v = r[hdr.index(key)]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            out.append(f"{label}={v}")
    st = []
    for h in stalls:
        try:
            st.append((float(r[hdr.index(h)]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
    st.sort(reverse=True)
    out.append("stalls: " + " ".join(f"{n}={v:.2f}" for v, n in st[:6]))
    print("  ".join(out))
