#!/usr/bin/env python
"""Dev tool for A/B kernel experiments: builds libdm2_b200 with extra -D flags into
dmesh2_renderer_b200/lib/variants/lib_<name>.so (git-ignored; travels with gpurun).  Select a variant
at run time with DM2_LIB=<path> (dmesh2_renderer_b200/_lib.py).

    python tools/build_variant.py <name> [-DFOO=1 ...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ge.LIBDIR, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    procs, objs = [], []
    for s in ge.CU_SOURCES:
        o = os.path.join(out_dir, s + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(["nvcc"] + ge.NVCC_FLAGS + flags + ["-c", os.path.join(ge.CSRC, s), "-o", o]))
    if any(p.wait() != 0 for p in procs):
        raise SystemExit("nvcc failed")
    lib = os.path.join(ge.LIBDIR, "variants", f"lib_{name}.so")
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    print(lib)


if __name__ == "__main__":
    main()
