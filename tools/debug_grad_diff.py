#!/usr/bin/env python
"""Where do this implementation and the reference disagree most on dL/dverts?  Prints the worst
elements with the fp64-accumulated CPU-oracle value next to both (GPU box, needs baseline/_ref)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dmesh2_renderer_b200 as dm                      # noqa: E402
from dmesh2_renderer_b200 import scenes                # noqa: E402
from helpers import args_to_numpy, load_reference, render_args, scene_tensors   # noqa: E402
from oracle import oracle as orc                       # noqa: E402


def main():
    dev = torch.device("cuda:0")
    ref = load_reference()
    n_faces = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    sc = scenes.cfg4(n_views=1, n_faces=n_faces)
    st = scene_tensors(sc, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    grads = {}
    for name, pkg in (("b200", dm), ("ref", ref)):
        r = pkg.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
        p = {k: st[k].detach().clone().requires_grad_(True) for k in ("verts", "verts_color", "faces_opacity", "faces_intense")}
        color, depth = r([0], pm, sc.width, sc.height, p["verts"], st["faces"], p["verts_color"], p["faces_opacity"],
                         p["faces_intense"], st["background"])
        ((color * st["w_color"]).sum() + (depth * st["w_depth"]).sum()).backward()
        grads[name] = {k: v.grad.detach().cpu().numpy() for k, v in p.items()}
    # oracle at the RenderFunction boundary + torch autograd for the host prep (on the GPU tensors)
    _r, _st, args = render_args(sc, dev)
    leaves = {}
    for i, n in ((4, "verts"), (8, "verts_ndc"), (12, "aa_face_verts")):
        pass
    a = args_to_numpy(args)
    o = orc.render(a, sc.width, sc.height, 1.0, K=0, backward=True, dL_dcolor=sc.w_color, dL_ddepth=sc.w_depth)
    # chain the oracle's boundary gradients through the torch prep to get d/d verts
    v = st["verts"].detach().clone().requires_grad_(True)
    r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    a2 = r.prepare_inputs([0], pm, sc.width, sc.height, v, st["faces"], st["verts_color"], st["faces_opacity"],
                          st["faces_intense"], st["background"], 1.0)
    tot = (a2[8] * torch.from_numpy(o["dL_dverts_ndc"]).to(dev, torch.float32)).sum() + \
          (a2[12] * torch.from_numpy(o["dL_daa_face_verts"]).to(dev, torch.float32)).sum()
    tot.backward()
    gold = v.grad.cpu().numpy().astype(np.float64) + o["dL_dverts"]
    gb, gr = grads["b200"]["verts"].astype(np.float64), grads["ref"]["verts"].astype(np.float64)
    mx = np.abs(gold).max()
    print("max |gold| %.4g ; max|b200-ref| %.4g ; max|b200-gold| %.4g ; max|ref-gold| %.4g" % (
        mx, np.abs(gb - gr).max(), np.abs(gb - gold).max(), np.abs(gr - gold).max()))
    print("rel-to-max: b200-ref %.3g  b200-gold %.3g  ref-gold %.3g" % (
        np.abs(gb - gr).max() / mx, np.abs(gb - gold).max() / mx, np.abs(gr - gold).max() / mx))
    idx = np.argsort(-np.abs(gb - gr).reshape(-1))[:8]
    for i in idx:
        vi, ci = divmod(int(i), 3)
        print(f"vert {vi} comp {ci} (face {vi // 3}): b200 {gb[vi, ci]:+.6e}  ref {gr[vi, ci]:+.6e}  gold {gold[vi, ci]:+.6e}")
    for k in ("verts_color", "faces_opacity", "faces_intense"):
        d = np.abs(grads["b200"][k].astype(np.float64) - grads["ref"][k]).max() / (np.abs(grads["ref"][k]).max() + 1e-30)
        print(k, "b200 vs ref rel-to-max", d)
    # 3-D only part at the boundary
    for nm, og in (("dL_dverts(3D path)", "dL_dverts"),):
        pass


if __name__ == "__main__":
    main()
