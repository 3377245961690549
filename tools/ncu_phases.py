#!/usr/bin/env python
"""Phase breakdown of one kernel from an ncu source-page dump (per-section warp-instruction shares).

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:<k> > src.csv
    python tools/ncu_phases.py src.csv [--lines N] [--phases file:lo-hi:label ...]

The dump lists a SASS instruction under every file of its inline stack.  Instructions are
de-duplicated by address and attributed to ONE (file, line): the first match in the priority order
dm2_aa.cuh > dm2_common.cuh > dm2_raster_common.cuh > other dm2_*.cu/.cuh > toolkit headers (so a
shuffle inside fill_queue counts for fill_queue, not for sm_30_intrinsics.hpp).  Several
`file:lo-hi` specs may share a label (their sums are merged).
"""
import argparse
import csv
import os
from collections import OrderedDict, defaultdict

PRIO = ["dm2_aa_fast.cuh", "dm2_aa.cuh", "dm2_common.cuh", "dm2_raster_common.cuh"]


def rank(fname):
    if fname in PRIO:
        return PRIO.index(fname)
    return len(PRIO) if fname.startswith("dm2_") else len(PRIO) + 1


def parse(path):
    sass = {}         # address -> (rank, file, line, inst, thread_inst, samples, text)
    src = {}
    cur_file = None
    hdr = None
    cur_line = None
    with open(path, newline="") as fh:
        for r in csv.reader(fh):
            if not r:
                continue
            if r[0] in ("File Name", "File Path"):
                cur_file = os.path.basename(r[1])
                continue
            if r[0] == "Line No":
                hdr = r
                continue
            if hdr is None or len(r) < 10:
                continue
            if r[0] != "":
                cur_line = int(r[0])
                src[(cur_file, cur_line)] = r[1].strip()
                continue
            addr = r[2]
            if not addr.startswith("0x"):
                continue
            try:
                inst = int(r[hdr.index("Instructions Executed")])
                tinst = int(r[hdr.index("Predicated-On Thread Instructions Executed")])
                samples = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            rk = rank(cur_file)
            if addr not in sass or rk < sass[addr][0]:
                sass[addr] = (rk, cur_file, cur_line, inst, tinst, samples, r[3].strip())
    return sass, src


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--lines", type=int, default=0, help="print the N heaviest source lines")
    ap.add_argument("--phases", nargs="*", default=[])
    a = ap.parse_args()
    sass, src = parse(a.csv)
    per = defaultdict(lambda: [0, 0, 0])
    for rk, f, ln, i, t, s, _ in sass.values():
        p = per[(f, ln)]
        p[0] += i; p[1] += t; p[2] += s
    ti = sum(v[0] for v in per.values()); tt = sum(v[1] for v in per.values()); ts = sum(v[2] for v in per.values())
    print(f"# warp instructions {ti:.4g}, predicated-on threads / instruction {tt / max(ti, 1):.1f}, stall samples {ts}")
    if a.lines:
        for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][0])[: a.lines]:
            print(f"{f + ':' + str(ln):32s} {100 * v[0] / ti:6.2f}% inst {v[1] / max(v[0], 1):5.1f} thr "
                  f"{100 * v[2] / max(ts, 1):6.2f}% smp | {src.get((f, ln), '')[:90]}")
    if a.phases:
        groups = OrderedDict()
        for spec in a.phases:
            f, rng, label = spec.split(":", 2)
            lo, hi = (int(x) for x in rng.split("-"))
            groups.setdefault(label, []).append((f, lo, hi))
        print("| phase | where | warp instr | share | threads/instr | stall-sample share |")
        print("|---|---|---|---|---|---|")
        covered = set()
        for label, specs in groups.items():
            keys = [k for k in per if any(k[0] == f and lo <= k[1] <= hi for f, lo, hi in specs) and k not in covered]
            covered.update(keys)
            i = sum(per[k][0] for k in keys); t = sum(per[k][1] for k in keys); s = sum(per[k][2] for k in keys)
            where = ", ".join(f"{f}:{lo}-{hi}" for f, lo, hi in specs)
            print(f"| {label} | {where} | {i:.3g} | {100 * i / ti:.1f} % | {t / max(i, 1):.1f} | {100 * s / max(ts, 1):.1f} % |")
        keys = [k for k in per if k not in covered]
        i = sum(per[k][0] for k in keys); t = sum(per[k][1] for k in keys); s = sum(per[k][2] for k in keys)
        print(f"| (other) | | {i:.3g} | {100 * i / ti:.1f} % | {t / max(i, 1):.1f} | {100 * s / max(ts, 1):.1f} % |")


if __name__ == "__main__":
    main()
