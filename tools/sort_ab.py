#!/usr/bin/env python
"""Sort-only A/B on the run's own keys (verdict r1 #9): this library's tile binning (scatter into tile segments +
per-segment sort, dm2_tilesort.cu) against torch.sort(stable=True) - CUB's onesweep radix sort underneath, the
kernel the reference calls (reference cuda_impl/renderer.cu:202-207) - and against the stand-alone LSD radix sort
dm2_sort_pairs, on the emitted (tile | depth, face) pairs of cfg2 and cfg5.

    python tools/sort_ab.py [cfg2 cfg5]

torch.sort orders all 64 key bits (8 onesweep passes); the reference sorts 32 + bit(tiles) of them (6 passes at
1024^2, 6 at 2048^2): the `cub_6_of_8` column scales torch's time accordingly.  One JSON line per config.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dmesh2_renderer_b200 as dm                      # noqa: E402
from dmesh2_renderer_b200 import _C, _lib, scenes      # noqa: E402
from helpers import render_args                        # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = torch.device("cuda:0")
    for name in (sys.argv[1:] or ["cfg2", "cfg5"]):
        sc = {"cfg2": scenes.cfg2, "cfg5": scenes.cfg5}[name]()
        _r, _st, args = render_args(sc, dev)
        with torch.no_grad():
            res = _C.render_forward_cuda(*args)
        R, bb = res[0], res[8]
        taps = _C.binning_taps(bb, R)
        ku, vu = taps["keys_unsorted"].clone(), taps["vals_unsorted"].clone()
        tiles = (sc.width // 16) * (sc.height // 16)
        end_bit = 32 + int(_lib.load_library().dm2_higher_msb(tiles))
        # (a) the pipeline's own binning, from the library's per-stage CUDA events
        _lib.profile_reset(); _lib.profile_enable(True)
        n = 5
        with torch.no_grad():
            for _ in range(n):
                _C.render_forward_cuda(*args)
        st = _lib.profile_report(); _lib.profile_enable(False)
        ours = (st["scatter_instances"][0] + st["tile_sort"][0]) / n
        count_scan = (st["tile_count"][0] + st["tile_scan"][0]) / n
        # (b) CUB onesweep through torch.sort (stable) + the value gather the pair sort implies
        def cub():
            k, idx = torch.sort(ku, stable=True)
            return k, vu[idx]
        t_cub = timed(cub)
        k_ref, v_ref = cub()
        assert torch.equal(k_ref, taps["keys"]) , "sorted keys differ from a stable sort of the emitted pairs"
        # equal (tile, depth) keys: the scatter order is arbitrary, the pipeline orders them by face id
        # (c) the stand-alone LSD radix sort of this library on the significant bits only
        t_lsd = timed(lambda: _C.sort_pairs(ku, vu, end_bit), iters=5)
        print(json.dumps({"config": sc.name, "num_rendered": R, "key_bits_sorted_by_reference": end_bit,
                          "tile_binning_ms (scatter + segment sort)": ours, "tile_count_scan_ms": count_scan,
                          "torch_sort_stable_64bit_ms": t_cub, "cub_6_of_8_passes_ms": t_cub * ((end_bit + 7) // 8) / 8,
                          "dm2_sort_pairs_lsd_ms": t_lsd,
                          "keys_per_s": {"tile_binning": R / ours * 1e3, "torch_sort": R / t_cub * 1e3,
                                         "dm2_sort_pairs": R / t_lsd * 1e3}}), flush=True)


if __name__ == "__main__":
    main()
