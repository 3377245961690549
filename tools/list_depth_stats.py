#!/usr/bin/env python
"""How far down its depth-sorted list does each tile actually composite?

    python tools/list_depth_stats.py [cfg2 cfg4 cfg5]

For every tile: list length (ranges) and the deepest list position any of its pixels blended (max n_contrib, the
reference's last_contributor, forward.cu:412-431).  Prints one JSON line per configuration with percentiles of
both and the share of listed instances that lie behind the deepest visited position of their tile - the work an
ahead-of-time classification pre-pass wastes, and what running the forward pass in rounds recovers
(dm2_set_forward_rounds).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from dmesh2_renderer_b200 import _C, scenes
    from helpers import render_args
    dev = torch.device("cuda:0")
    for name in (sys.argv[1:] or ["cfg2", "cfg4", "cfg5"]):
        sc = getattr(scenes, name)()
        views = [0] if name != "cfg4" else list(range(8))
        _, _, args = render_args(sc, dev, views=views)
        with torch.no_grad():
            res = _C.render_forward_cuda(*[a.detach() if isinstance(a, torch.Tensor) else a for a in args])
        R, ib = res[0], res[9]
        taps = _C.image_taps(ib, len(views), sc.height, sc.width)
        rng = taps["ranges"].cpu().numpy().astype(np.int64)
        lens = rng[:, 1] - rng[:, 0]
        nc = taps["n_contrib"].cpu().numpy().astype(np.int64)
        B, H, W = nc.shape
        gy, gx = (H + 15) // 16, (W + 15) // 16
        pad = np.zeros((B, gy * 16, gx * 16), np.int64)
        pad[:, :H, :W] = nc
        deepest = pad.reshape(B, gy, 16, gx, 16).max(axis=(2, 4)).reshape(-1)
        # a pixel that never reaches T < 1e-4 walks its whole list: only tiles whose pixels ALL terminated stop early
        fT = np.zeros((B, gy * 16, gx * 16), np.float32)
        fT[:, :H, :W] = taps["final_T"].cpu().numpy()
        all_done = (fT.reshape(B, gy, 16, gx, 16).max(axis=(2, 4)).reshape(-1) < 1e-4)
        deepest = np.where(all_done, deepest, lens)
        out_extra = {"tiles_terminated_early_frac": float(all_done.mean())}
        pct = lambda a: [int(np.percentile(a, q)) for q in (50, 90, 99, 100)]
        out = {"config": name, "views": len(views), "num_rendered": int(R), "tiles": int(lens.size),
               "list_len_p50_p90_p99_max": pct(lens), "deepest_visited_p50_p90_p99_max": pct(deepest),
               "instances_behind_deepest_visited_frac": float((lens - deepest).sum() / max(1, lens.sum()))}
        out.update(out_extra)
        for c in (512, 1024, 2048):
            cls = np.minimum(lens, (deepest + c - 1) // c * c).sum()
            out[f"classified_frac_rounds_of_{c}"] = float(cls / max(1, lens.sum()))
        print(json.dumps(out), flush=True)
        del res, taps, ib
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
