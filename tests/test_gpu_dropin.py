"""Drop-in proof at the pybind seam: the reference's OWN, untouched Python package
(baseline/_ref/dmesh2_renderer/__init__.py - its RenderFunction, Renderer.forward and LayeredRenderer.generate)
is executed with its `_C` extension module replaced by dmesh2_renderer_b200._C (the host mirror over
libdm2_b200.so), and compared with the same package on its native `_C`.

Call sites exercised: reference dmesh2_renderer/__init__.py:4 (`from . import _C`), :78
(`_C.render_forward_cuda(*args)`, 21 positional args, 10-tuple result), :170 (`_C.render_backward_cuda(*args)`,
31 positional args, 6-tuple result), :432 (`_C.generate_render_layers_cuda(...)`, 13 positional args).
INTEGRATION.md section 2 shows the same shim for a maintainer.
"""
import importlib.util
import sys

import numpy as np
import pytest
import torch

from helpers import load_reference, scene_tensors

pytestmark = pytest.mark.gpu


def load_reference_python_on_b200_C():
    """The reference's __init__.py, loaded as package `dm2_dropin`, with `dm2_dropin._C` = our host mirror."""
    ref = load_reference()
    if ref is None:
        return None, None
    import dmesh2_renderer_b200._C as b200_C
    name = "dm2_dropin"
    if name in sys.modules:
        return sys.modules[name], ref
    spec = importlib.util.spec_from_file_location(name, ref.__file__, submodule_search_locations=[])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    sys.modules[name + "._C"] = b200_C          # what `from . import _C` resolves to
    spec.loader.exec_module(mod)
    assert mod._C is b200_C and ref._C is not b200_C
    return mod, ref


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("tau,patch", [(1.0, None), (0.5, ((9, 30), (100, 70)))])
def test_reference_renderer_runs_on_b200_extension(built, tau, patch):
    from dmesh2_renderer_b200 import scenes
    dropin, ref = load_reference_python_on_b200_C()
    if dropin is None:
        pytest.skip("baseline/_ref not present on this box")
    dev = _dev()
    sc = scenes.triangle_soup("dropin", 60_000, 256, 0.03, seed=21,
                              eyes=[tuple(e) for e in scenes.fibonacci_eyes(64)[:2]])
    st = scene_tensors(sc, dev)
    views = [1, 0]
    if patch is None:
        pm, (pw, ph) = torch.zeros((2, 2), dtype=torch.int32, device=dev), (256, 256)
    else:
        pm, (pw, ph) = torch.tensor([patch[0], (40, 3)], dtype=torch.int32, device=dev), patch[1]
    g = torch.Generator(device="cpu").manual_seed(3)
    wc = (torch.rand((2, ph, pw, 3), generator=g) * 2 - 1).to(dev)
    wd = (torch.rand((2, ph, pw), generator=g) * 2 - 1).to(dev)
    outs = []
    for pkg in (dropin, ref):
        r = pkg.Renderer(st["mv"], st["proj"], 256, 256, dev)
        p = {k: st[k].detach().clone().requires_grad_(True) for k in ("verts", "verts_color", "faces_opacity")}
        fi = st["faces_intense"][views].detach().clone().requires_grad_(True)
        color, depth = r(views, pm, pw, ph, p["verts"], st["faces"], p["verts_color"], p["faces_opacity"], fi,
                         st["background"], tau)
        ((color * wc).sum() + (depth * wd).sum()).backward()
        outs.append((color.detach(), depth.detach(), [p["verts"].grad, p["verts_color"].grad,
                                                      p["faces_opacity"].grad, fi.grad]))
    (c0, d0, g0), (c1, d1, g1) = outs
    assert float(((c0 - c1).abs() > 1e-4 * (1 + c1.abs())).float().mean()) == 0.0
    assert float(((d0 - d1).abs() > 1e-4 * (1 + d1.abs())).float().mean()) == 0.0
    for a, b, name in zip(g0, g1, ("verts", "verts_color", "faces_opacity", "faces_intense")):
        e = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)
        assert e <= 1e-3, (name, e)


def test_reference_layered_renderer_runs_on_b200_extension(built):
    from dmesh2_renderer_b200 import scenes
    dropin, ref = load_reference_python_on_b200_C()
    if dropin is None:
        pytest.skip("baseline/_ref not present on this box")
    dev = _dev()
    sc = scenes.delaunay_tets("dropin_tets", 4000, 256, seed=17, p_exist=0.2, num_layers=6)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    a = [t(getattr(sc, k)) for k in ("verts", "faces", "tets", "face_tets", "tet_faces", "face_existence")]
    res = []
    for pkg in (dropin, ref):
        lr = pkg.LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
        res.append(lr.generate([0], *a, sc.num_layers))
    assert torch.equal(res[0][0], res[1][0]), "layer ids differ"
    assert torch.equal(res[0][1], res[1][1]), "layer counts differ"
    assert res[0][0].dtype == res[1][0].dtype and res[0][0].shape == res[1][0].shape


def test_reference_error_path_is_preserved(built):
    """The reference's RenderFunction prints and re-raises what _C raises (__init__.py:77-82): same exception
    type and message through the swapped extension."""
    dropin, ref = load_reference_python_on_b200_C()
    if dropin is None:
        pytest.skip("baseline/_ref not present on this box")
    from dmesh2_renderer_b200 import scenes
    dev = _dev()
    sc = scenes.small(n_faces=50, size=32)
    st = scene_tensors(sc, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    msgs = []
    for pkg in (dropin, ref):
        r = pkg.Renderer(st["mv"], st["proj"], 32, 32, dev)
        with pytest.raises(RuntimeError) as ei:
            r([0], pm, 32, 32, st["verts"], st["faces"], st["verts_color"], st["faces_opacity"],
              st["faces_intense"], st["background"], 1.5)            # aa_temperature out of range
        msgs.append(str(ei.value).splitlines()[0])
    assert "aa_temperature must be in the range [0, 1]" in msgs[0]
    assert "aa_temperature must be in the range [0, 1]" in msgs[1]
