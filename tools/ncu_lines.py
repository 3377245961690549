#!/usr/bin/env python
"""Aggregate an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line.

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:<k> > src.csv
    python tools/ncu_lines.py src.csv [--top 40] [--ranges file:lo-hi:label ...]

Prints, per source line, warp instructions executed, average active threads and stall samples, and
(optionally) the totals over labelled line ranges - used to attribute kernel time to phases.
"""
import argparse
import csv
import os
from collections import defaultdict


def parse(path):
    rows = []
    fname = None
    hdr = None
    cur_line = None
    with open(path, newline="") as fh:
        for r in csv.reader(fh):
            if not r:
                continue
            if r[0] in ("File Name", "File Path"):
                fname = os.path.basename(r[1])
                continue
            if r[0] == "Line No":
                hdr = r
                continue
            if hdr is None or len(r) < 10:
                continue
            if r[0] != "":
                cur_line = int(r[0])
                continue            # the source-line row itself (aggregates); SASS rows follow
            if r[2] in ("...", ""):
                continue
            try:
                inst = int(r[hdr.index("Instructions Executed")])
                tinst = int(r[hdr.index("Predicated-On Thread Instructions Executed")])
                samples = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            rows.append((fname, cur_line, r[3].strip(), inst, tinst, samples))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--ranges", nargs="*", default=[])
    a = ap.parse_args()
    rows = parse(a.csv)
    # the same SASS instruction is listed under every file it correlates to: de-duplicate by address text
    per = defaultdict(lambda: [0, 0, 0])
    for f, ln, sass, inst, tinst, smp in rows:
        p = per[(f, ln)]
        p[0] += inst; p[1] += tinst; p[2] += smp
    tot_i = sum(v[0] for v in per.values())
    tot_s = sum(v[2] for v in per.values())
    print(f"total warp inst {tot_i:.4g}  samples {tot_s}")
    print(f"{'file:line':40s} {'warp inst':>12s} {'%':>6s} {'thr/inst':>8s} {'samples':>8s} {'%':>6s}")
    for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][2])[: a.top]:
        print(f"{f + ':' + str(ln):40s} {v[0]:12d} {100 * v[0] / tot_i:6.2f} {v[1] / max(v[0], 1):8.1f} {v[2]:8d} {100 * v[2] / max(tot_s, 1):6.2f}")
    for spec in a.ranges:
        f, rng, label = spec.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        i = sum(v[0] for (ff, ln), v in per.items() if ff == f and lo <= ln <= hi)
        t = sum(v[1] for (ff, ln), v in per.items() if ff == f and lo <= ln <= hi)
        s = sum(v[2] for (ff, ln), v in per.items() if ff == f and lo <= ln <= hi)
        print(f"[{label:28s}] {f}:{lo}-{hi}: inst {100 * i / tot_i:5.1f}%  thr/inst {t / max(i, 1):5.1f}  samples {100 * s / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main()
