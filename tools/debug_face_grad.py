#!/usr/bin/env python
"""Where does dL/dverts of ONE vertex differ from the fp64-accumulated oracle?  (GPU box; oracle + this library)

    python tools/debug_face_grad.py cfg2 2461333

Renders the patch around the faces of the vertex through the reference-shaped RenderFunction route (CUDA) and
through the CPU oracle, prints the six boundary gradients of those faces side by side, then repeats the
comparison pixel by pixel (1x1 patches) to find the pixels that carry the difference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dmesh2_renderer_b200 as dm                      # noqa: E402
from dmesh2_renderer_b200 import scenes                # noqa: E402
from helpers import GRAD_NAMES, args_to_numpy          # noqa: E402
from oracle import oracle as orc                       # noqa: E402


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def run_patch(r, sc, view, sel, vid, faces_sub, x0, y0, pw, ph, dev, K=20):
    pm = torch.tensor([[x0, y0]], dtype=torch.int32, device=dev)
    r.aa_grad_buffer_size = K
    args = list(r.prepare_inputs([view], pm, pw, ph, t(sc.verts[vid], dev), t(faces_sub, dev), t(sc.verts_color[vid], dev),
                                 t(sc.faces_opacity[sel], dev), t(sc.faces_intense[[view]][:, sel], dev),
                                 t(sc.background, dev), 1.0))
    wc = sc.w_color[[view], y0:y0 + ph, x0:x0 + pw]
    wd = -0.5 * sc.w_depth[[view], y0:y0 + ph, x0:x0 + pw]
    leaves = []
    for i in (4, 6, 7, 8, 10, 12):
        args[i] = args[i].detach().clone().requires_grad_(True)
        leaves.append(args[i])
    color, depth = dm.RenderFunction.apply(*args)
    ((color * t(wc, dev)).sum() + (depth * t(wd, dev)).sum()).backward()
    cuda = {n: l.grad.detach().cpu().numpy().astype(np.float64) for n, l in zip(GRAD_NAMES, leaves)}
    cuda["color"] = color.detach().cpu().numpy()
    o = orc.render(args_to_numpy(args), pw, ph, 1.0, K=0, backward=True, dL_dcolor=wc, dL_ddepth=wd)
    return cuda, o


def main():
    cfg, v = sys.argv[1], int(sys.argv[2])
    dev = torch.device("cuda:0")
    sc = {"cfg2": scenes.cfg2, "cfg5": scenes.cfg5, "cfg4": lambda: scenes.cfg4(n_views=8)}[cfg]()
    view = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    W, H = sc.width, sc.height
    r = dm.Renderer(t(sc.mv, dev), t(sc.proj, dev), W, H, dev)
    _ndc, img = r.compute_verts_ndc_image(t(sc.verts, dev), r.mv[[view]], r.proj[[view]])
    tri_img = img[0].cpu().numpy()[sc.faces]
    fmin, fmax = tri_img.min(1), tri_img.max(1)
    mine = np.nonzero((sc.faces == v).any(1))[0]
    x0 = max(int(np.floor(fmin[mine, 0].min())) - 1, 0); y0 = max(int(np.floor(fmin[mine, 1].min())) - 1, 0)
    x1 = min(int(np.ceil(fmax[mine, 0].max())) + 1, W); y1 = min(int(np.ceil(fmax[mine, 1].max())) + 1, H)
    sel = np.nonzero((fmax[:, 0] >= x0 - 1) & (fmin[:, 0] <= x1 + 1) & (fmax[:, 1] >= y0 - 1) & (fmin[:, 1] <= y1 + 1))[0]
    vid, inv = np.unique(sc.faces[sel], return_inverse=True)
    faces_sub = inv.reshape(-1, 3).astype(np.int32)
    print(f"vertex {v}: faces {mine.tolist()}, patch [{x0},{x1}) x [{y0},{y1}), {len(sel)} faces in the patch")
    print("image verts of the face:", tri_img[mine[0]].tolist())
    f_loc = int(np.searchsorted(sel, mine[0]))
    rows = [int(np.searchsorted(vid, u)) for u in sc.faces[mine[0]]]
    for K in (20, 0):
        cuda, o = run_patch(r, sc, view, sel, vid, faces_sub, x0, y0, x1 - x0, y1 - y0, dev, K=K)
        print(f"--- whole patch, K={K}: colour max diff {np.abs(cuda['color'] - o['color']).max():.3g}")
        for n in GRAD_NAMES:
            a, b = cuda[n], o[n]
            if n in ("dL_dverts", "dL_dverts_color"):
                a, b = a[rows], b[rows]
            elif n == "dL_dverts_ndc":
                a, b = a[0, rows, 2], b[0, rows, 2]
            elif n == "dL_dfaces_opacity":
                a, b = a[f_loc], b[f_loc]
            else:
                a, b = a[0, f_loc], b[0, f_loc]
            print(f"{n:20s} cuda {np.asarray(a).ravel()}\n{'':20s} gold {np.asarray(b).ravel()}")
    # pixel by pixel
    print("--- per pixel (1x1 patches): |cuda - gold| of dL_dverts rows of the face, dL_daa_face_verts")
    for py in range(y0, y1):
        for px in range(x0, x1):
            cuda, o = run_patch(r, sc, view, sel, vid, faces_sub, px, py, 1, 1, dev, K=20)
            d3 = np.abs(cuda["dL_dverts"][rows] - o["dL_dverts"][rows]).max()
            da = np.abs(cuda["dL_daa_face_verts"][0, f_loc] - o["dL_daa_face_verts"][0, f_loc]).max()
            dz = np.abs(cuda["dL_dverts_ndc"][0, rows, 2] - o["dL_dverts_ndc"][0, rows, 2]).max()
            s3 = np.abs(o["dL_dverts"][rows]).max()
            if d3 > 1e-3 * max(s3, 1.0) or da > 1e-4 or dz > 1e-4:
                print(f"pixel ({px},{py}): d3 {d3:.4g} (|gold| {s3:.4g})  daa {da:.4g}  dz {dz:.4g}")
                print("   cuda dverts", cuda["dL_dverts"][rows].ravel())
                print("   gold dverts", o["dL_dverts"][rows].ravel())
                print("   cuda daa", cuda["dL_daa_face_verts"][0, f_loc].ravel(), " gold daa", o["dL_daa_face_verts"][0, f_loc].ravel())


if __name__ == "__main__":
    main()
