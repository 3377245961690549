"""Generates tests/golden/aa_overlap_golden.npz and tests/golden/prep_golden.npz by IMPORTING the
reference's own Python (runs in the dev container only; needs /root/reference, and for the prep
golden the reference package built into baseline/_ref by baseline/build_ref.sh).

  aa_overlap_golden.npz : random (triangle, unit pixel) pairs with the overlap area and
      d(area)/d(tri verts) computed by BOTH reference implementations in
      dmesh2_renderer/pyrenderer.py - `_tri_pixel_overlap_area_analytic` (:207-425, the function
      aa.h transcribes) and `_tri_pixel_overlap_area_autograd` (:66-205, gradients by autograd).
  prep_golden.npz : outputs of the reference's host-side prep on CPU for a tiny scene:
      Renderer._init_rays, compute_verts_ndc_image, select_rays (dmesh2_renderer/__init__.py:198-302)
      and pyrenderer.Triangles (:6-30).

Usage:  python tests/golden/make_python_goldens.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_PY = "/root/reference/dmesh2_renderer/pyrenderer.py"


def load_pyrenderer():
    spec = importlib.util.spec_from_file_location("ref_pyrenderer", REF_PY)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def make_aa(n=1500, seed=1234):
    pr = load_pyrenderer()
    rng = np.random.default_rng(seed)
    rows = []
    while len(rows) < n:
        kind = rng.integers(0, 4)
        px, py = float(rng.integers(0, 6)), float(rng.integers(0, 6))
        if kind == 0:      # small triangle near the pixel
            c = np.array([px + 0.5, py + 0.5]) + rng.normal(0, 0.7, 2)
            tri = c + rng.normal(0, 0.8, (3, 2))
        elif kind == 1:    # large triangle, pixel often fully inside
            c = np.array([px + 0.5, py + 0.5]) + rng.normal(0, 1.5, 2)
            tri = c + rng.normal(0, 4.0, (3, 2))
        elif kind == 2:    # a vertex inside the pixel
            tri = np.array([[px + rng.uniform(0.1, 0.9), py + rng.uniform(0.1, 0.9)],
                            [px + rng.normal(0.5, 2.0), py + rng.normal(0.5, 2.0)],
                            [px + rng.normal(0.5, 2.0), py + rng.normal(0.5, 2.0)]])
        else:              # thin sliver
            a = np.array([px + rng.uniform(-2, 3), py + rng.uniform(-2, 3)])
            d = rng.normal(0, 3.0, 2)
            tri = np.stack([a, a + d, a + 0.5 * d + rng.normal(0, 0.15, 2)])
        tri = tri.astype(np.float32)
        res = {}
        ok = True
        for mode in ("analytic", "autograd"):
            p = [torch.tensor(tri[i:i + 1].copy(), requires_grad=True) for i in range(3)]
            tris = pr.Triangles(p[0].clone(), p[1].clone(), p[2].clone())
            pix = pr.Pixels(torch.tensor([[px, py]]), torch.tensor([[px + 1, py + 1]]))
            try:
                area, _poly = pr.tri_pixel_overlap_area(tris, pix, 0, 0, use_autograd=(mode == "autograd"))
            except ValueError:
                ok = False
                break
            if isinstance(area, torch.Tensor) and area.requires_grad:
                # TOTAL derivative w.r.t. the three input points (edges / reciprocals are functions of
                # them), re-ordered to the CCW order of Triangles.verts that the kernel sees
                area.backward()
                gl = [x.grad[0].numpy().copy() if x.grad is not None else np.zeros(2, np.float32) for x in p]
                swapped = not np.array_equal(tris.verts[0, 1].detach().numpy(), tri[1])
                g = np.stack([gl[0], gl[2], gl[1]] if swapped else gl)
                a = float(area.detach())
            else:
                g = np.zeros((3, 2), np.float32)
                a = float(area)
            res[mode] = (a, g)
            res["prep"] = (tris.verts[0].detach().numpy().copy(), tris.edges[0].detach().numpy().copy(),
                           tris.edges_iszero[0].numpy().copy(), tris.edges_recip[0].detach().numpy().copy(),
                           tris.edges_normal[0].detach().numpy().copy(), tris.edges_normal_c[0].detach().numpy().copy())
        if not ok:
            continue
        rows.append((tri, px, py, res))
    out = dict(
        tri=np.stack([r[0] for r in rows]), px=np.array([r[1] for r in rows], np.float32),
        py=np.array([r[2] for r in rows], np.float32),
        area_analytic=np.array([r[3]["analytic"][0] for r in rows], np.float32),
        grad_analytic=np.stack([r[3]["analytic"][1] for r in rows]).astype(np.float32),
        area_autograd=np.array([r[3]["autograd"][0] for r in rows], np.float32),
        grad_autograd=np.stack([r[3]["autograd"][1] for r in rows]).astype(np.float32),
        verts=np.stack([r[3]["prep"][0] for r in rows]), edges=np.stack([r[3]["prep"][1] for r in rows]),
        iszero=np.stack([r[3]["prep"][2] for r in rows]), recip=np.stack([r[3]["prep"][3] for r in rows]),
        normal=np.stack([r[3]["prep"][4] for r in rows]), normal_c=np.stack([r[3]["prep"][5] for r in rows]))
    np.savez_compressed(os.path.join(HERE, "aa_overlap_golden.npz"), **out)
    print("aa golden:", len(rows), "pairs; partial:", int(((out["area_analytic"] > 0) & (out["area_analytic"] < 1)).sum()))


def make_prep():
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import dmesh2_renderer as ref          # the UNMODIFIED reference package (needs its built _C to import)
    sys.path.insert(0, ROOT)
    from dmesh2_renderer_b200 import scenes
    sc = scenes.small(n_faces=200, size=48, n_views=3, seed=21)
    t = torch.from_numpy
    r = ref.Renderer(t(sc.mv), t(sc.proj), sc.width, sc.height, "cpu")
    idx = [2, 0]
    ndc, img = r.compute_verts_ndc_image(t(sc.verts), r.mv[idx], r.proj[idx])
    pmin = torch.tensor([[3, 5], [16, 0]], dtype=torch.int32)
    ro, rd = r.select_rays(idx, pmin, 32, 40)
    tv = img[:, t(sc.faces).long().flatten()].view(-1, 3, 2)
    tris = ref.Triangles(tv[:, 0], tv[:, 1], tv[:, 2])
    np.savez_compressed(
        os.path.join(HERE, "prep_golden.npz"), ray_o=r.ray_o.numpy(), ray_d=r.ray_d.numpy(), verts_ndc=ndc.numpy(),
        verts_image=img.numpy(), sel_ray_o=ro.numpy(), sel_ray_d=rd.numpy(), idx=np.array(idx), patch_min=pmin.numpy(),
        tri_verts=tris.verts.numpy(), tri_edges=tris.edges.numpy(), tri_iszero=tris.edges_iszero.numpy(),
        tri_recip=tris.edges_recip.numpy(), tri_normal=tris.edges_normal.numpy(), tri_normal_c=tris.edges_normal_c.numpy())
    print("prep golden written")


if __name__ == "__main__":
    make_aa()
    make_prep()
