#!/usr/bin/env python
"""Which ray recipe (include/dm2_renderer.h, dm2_generate_rays) reproduces torch's _init_rays tensors bit for bit
on this box?  Prints, per frame size / view count, the selected recipe and how many recipes match at all."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

from dmesh2_renderer_b200 import _lib, renderer, scenes  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load_library()
    for size, nv in ((64, 1), (96, 3), (200, 2), (512, 8), (512, 64), (1024, 1), (1000, 1), (2048, 1)):
        sc = scenes.cfg4(n_views=max(nv, 2), size=size, n_faces=10)
        sc.mv, sc.proj = sc.mv[-nv:], sc.proj[-nv:]
        mv, proj = torch.from_numpy(sc.mv).to(dev), torch.from_numpy(sc.proj).to(dev)
        ro, rd, imv, ipr = renderer._unproject_pixel_rays(mv, proj, size, size, dev)
        params = torch.cat((ipr.reshape(nv, 16), imv.reshape(nv, 16)), dim=1).contiguous()
        o, d = torch.empty_like(ro), torch.empty_like(rd)
        match = []
        for recipe in range(1, _lib.RAY_RECIPES + 1):
            _lib.check(lib.dm2_generate_rays(C.c_void_p(params.data_ptr()), nv, size, size, recipe, C.c_void_p(o.data_ptr()),
                                             C.c_void_p(d.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            if torch.equal(o, ro) and torch.equal(d, rd):
                match.append(recipe)
        best = (d - rd).abs().max().item()
        print(json.dumps({"size": size, "views": nv, "matching_recipes": match, "last_recipe_max_abs_diff": best}))


if __name__ == "__main__":
    main()
