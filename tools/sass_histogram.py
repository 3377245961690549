#!/usr/bin/env python
"""SASS opcode histogram per kernel of libdm2_b200.so (cuobjdump -sass): profiles/r02_sass_opcodes.md.
Shows at a glance what the kernels are made of: LDGSTS (cp.async staging), REDG / RED (gradient reductions),
MATCH / REDUX / VOTE / SHFL (warp-level work), LDGMC / multimem (switch reductions), no FP64 arithmetic, no
local-memory traffic in the hot paths (the chain-gradient slow path and the recompute fallbacks use the stack).

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dmesh2_renderer_b200", "lib", "libdm2_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [o.split("(")[0].replace("void ", "").replace("dm2::", "") for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op.split(".")[0]] += 1
            if op.startswith(("LDL", "STL")):
                kernels[cur]["(local memory ld/st)"] += 1
            if re.match(r"D(ADD|MUL|FMA|SETP)", op):
                kernels[cur]["(FP64 arithmetic)"] += 1
    names = demangle(list(kernels))
    interesting = ["LDGSTS", "LDG", "STG", "LDS", "STS", "REDG", "RED", "ATOMG", "ATOMS", "MATCH", "REDUX", "VOTE", "SHFL",
                   "LDGMC", "UBLKCP", "UTMALDG", "FFMA", "FMUL", "FADD", "FSETP", "FSEL", "MUFU", "IMAD", "LOP3", "BSSY", "CALL",
                   "(local memory ld/st)", "(FP64 arithmetic)"]
    print("# SASS opcode counts per kernel (static; `cuobjdump -sass dmesh2_renderer_b200/lib/libdm2_b200.so`, sm_100a, nvcc 12.9)")
    print("# No UBLKCP / UTMALDG: the TMA staging variant (cp.async.bulk per 240-byte record) was measured 2.5 % slower than")
    print("# LDGSTS (16-byte cp.async) in round 1 (gpurun_out/ab_*.json, DESIGN.md section 3) and is not in the build.\n")
    print("| kernel | instructions | " + " | ".join(interesting) + " |")
    print("|---|---|" + "---|" * len(interesting))
    for (k, c), n in sorted(zip(kernels.items(), names), key=lambda kv: -sum(kv[0][1].values())):
        tot = sum(v for o, v in c.items() if not o.startswith("("))
        if tot < 60:
            continue
        print(f"| {n} | {tot} | " + " | ".join(str(c.get(o, 0)) for o in interesting) + " |")


if __name__ == "__main__":
    main()
