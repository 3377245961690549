#!/usr/bin/env python
"""Step statistics of the LayeredRenderer tet walk (debug build):
    python tools/build_variant.py walkstats -DDM2_WALK_STATS
    DM2_LIB=$PWD/dmesh2_renderer_b200/lib/variants/lib_walkstats.so python tools/walk_stats.py
Prints how many walk steps took the fast path (edge-function signs) and why the others did not."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import _lib, scenes
    lib = _lib.load_library()
    fn = lib.dm2_debug_walk_stats
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    fn.restype = None
    dev = torch.device("cuda:0")
    sc = scenes.cfg3()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    lr = dm.LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    a = [t(getattr(sc, k)) for k in ("verts", "faces", "tets", "face_tets", "tet_faces", "face_existence")]
    out = (C.c_ulonglong * 8)()
    fn(out, 1)
    lr.generate([0], *a, sc.num_layers)
    torch.cuda.synchronize()
    fn(out, 1)
    names = ["steps", "no_fast_state", "margin_not_met", "not_in_front", "bad_sign_pattern", "fast_steps", "walk_ended_in_slow", "-"]
    print(json.dumps({n: int(v) for n, v in zip(names, out)}))


if __name__ == "__main__":
    main()
