"""Shared helpers for the parity tests (host-side only; no product logic lives here)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

ARG_NAMES = ["background", "patch_min", "patch_width", "patch_height", "verts", "faces", "verts_color",
             "faces_opacity", "verts_ndc", "verts_image", "faces_intense", "aa_temperature", "aa_face_verts",
             "aa_face_edges", "aa_face_edges_iszero", "aa_face_edges_recip", "aa_face_edges_normal",
             "aa_face_edges_normal_c", "len_oarea_buffer", "image_ray_o", "image_ray_d"]
GRAD_NAMES = ["dL_dverts", "dL_dverts_color", "dL_dfaces_opacity", "dL_dverts_ndc", "dL_dfaces_intense",
              "dL_daa_face_verts"]


def scene_tensors(sc, device):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return dict(mv=t(sc.mv), proj=t(sc.proj), verts=t(sc.verts), faces=t(sc.faces), verts_color=t(sc.verts_color),
                faces_opacity=t(sc.faces_opacity), faces_intense=t(sc.faces_intense), background=t(sc.background),
                w_color=t(sc.w_color), w_depth=t(sc.w_depth))


def render_args(sc, device, views=None, patch_min=None, patch_wh=None, tau=1.0, renderer_cls=None):
    """The 21 RenderFunction inputs for a scene, produced by the product's host-side prep."""
    from dmesh2_renderer_b200 import Renderer
    cls = renderer_cls or Renderer
    st = scene_tensors(sc, device)
    views = list(range(sc.mv.shape[0])) if views is None else list(views)
    r = cls(st["mv"], st["proj"], sc.width, sc.height, device)
    pw, ph = patch_wh if patch_wh is not None else (sc.width, sc.height)
    pm = torch.zeros((len(views), 2), dtype=torch.int32) if patch_min is None else torch.as_tensor(patch_min, dtype=torch.int32)
    pm = pm.to(device)
    args = r.prepare_inputs(views, pm, pw, ph, st["verts"], st["faces"], st["verts_color"], st["faces_opacity"],
                            st["faces_intense"][views], st["background"], tau)
    return r, st, list(args)


def args_to_numpy(args):
    out = {}
    for n, a in zip(ARG_NAMES, args):
        out[n] = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    return out


def rel_err(a, b):
    """max |a-b| / (|b|_max + tiny): scale-relative error used for float parity."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)) if a.size else 0.0


def frac_outside(a, b, rtol, atol):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    bad = np.abs(a - b) > (atol + rtol * np.abs(b))
    return float(bad.mean()) if a.size else 0.0


def load_reference():
    """The UNMODIFIED reference package built into baseline/_ref (None if absent)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    pkg = os.path.join(ref_dir, "dmesh2_renderer")
    if not os.path.isdir(pkg) or not any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(pkg)):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import dmesh2_renderer as ref
        return ref
    except Exception:
        return None
