#!/usr/bin/env python
"""Kernel-level trace of bench.py's step (cfg2, public API fwd + loss + bwd) with torch.profiler:
prints every GPU kernel / memcpy / memset of the traced steps with its summed duration, and the
GPU-idle share of the step.  Used to find what is left besides the dm2:: kernels.

    python tools/step_trace.py [--steps 3] [--faces 1000000] [--size 1024]
"""
import argparse
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--faces", type=int, default=1_000_000)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--cfg4", action="store_true", help="one multi-view step: 64 views @512^2, 8 views per Renderer call")
    a = ap.parse_args()
    import dmesh2_renderer_b200 as pkg
    from dmesh2_renderer_b200 import scenes
    pkg.load_library()
    dev = torch.device("cuda", 0)
    sc = scenes.cfg2() if (a.faces == 1_000_000 and a.size == 1024) else scenes.cfg2(n_faces=a.faces, size=a.size)
    if a.cfg4:
        sc = scenes.cfg4()
    t = torch.from_numpy
    r = pkg.Renderer(t(sc.mv).to(dev), t(sc.proj).to(dev), sc.width, sc.height, dev)
    faces, bg = t(sc.faces).to(dev), t(sc.background).to(dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    wc, wd = t(sc.w_color).to(dev), t(sc.w_depth).to(dev)
    p = {k: t(getattr(sc, k)).to(dev).requires_grad_(True)
         for k in ("verts", "verts_color", "faces_opacity", "faces_intense")}

    pm8 = torch.zeros((8, 2), dtype=torch.int32, device=dev)
    consts = {}

    def step4():
        for x in p.values():
            x.grad = None
        for b0 in range(0, sc.mv.shape[0], 8):
            vb = list(range(b0, b0 + 8))
            if b0 not in consts:         # (indexing with a Python list = a pageable H2D copy that drains the queue)
                idx = torch.tensor(vb, dtype=torch.long, device=dev)
                consts[b0] = (idx, wc.index_select(0, idx), wd.index_select(0, idx))
            idx, wcb, wdb = consts[b0]
            color, depth = r(vb, pm8, sc.width, sc.height, p["verts"], faces, p["verts_color"], p["faces_opacity"],
                             p["faces_intense"].index_select(0, idx), bg)
            ((color * wcb).sum() + (depth * wdb).sum()).backward()

    def step():
        if a.cfg4:
            return step4()
        for x in p.values():
            x.grad = None
        color, depth = r([0], pm, sc.width, sc.height, p["verts"], faces, p["verts_color"], p["faces_opacity"],
                         p["faces_intense"], bg)
        ((color * wc).sum() + (depth * wd).sum()).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    if os.environ.get("STEP_TRACE_JSON"):
        prof.export_chrome_trace(os.environ["STEP_TRACE_JSON"])
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    per = defaultdict(lambda: [0, 0.0])
    for e in ev:
        per[e.name][0] += 1
        per[e.name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    busy = sum(v[1] for v in per.values())
    t0 = min(e.time_range.start for e in ev)
    t1 = max(e.time_range.end for e in ev)
    span = t1 - t0
    print(f"steps {a.steps}  span {span / a.steps / 1e3:.3f} ms/step  busy {busy / a.steps / 1e3:.3f} ms/step  "
          f"idle {100 * (1 - busy / span):.1f} %")
    for name, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / a.steps:10.1f} us/step  x{n / a.steps:5.1f}  {name[:110]}")
    # GPU-idle gaps between consecutive activities (after -> before), summed per neighbour pair
    ev.sort(key=lambda e: e.time_range.start)
    gaps = defaultdict(lambda: [0, 0.0])
    end = ev[0].time_range.end
    prev = ev[0].name
    for e in ev[1:]:
        g = e.time_range.start - end
        if g > 0:
            k = (prev[:48], e.name[:48])
            gaps[k][0] += 1
            gaps[k][1] += g
        if e.time_range.end > end:
            end, prev = e.time_range.end, e.name
    print("idle gaps (us/step):")
    for (p, n), (c, us) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{us / a.steps:8.1f}  x{c / a.steps:4.1f}  {p}  ->  {n}")


if __name__ == "__main__":
    main()
