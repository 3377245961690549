"""CPU tests (no GPU): the oracle against the reference's own known-answer material, the host-side
prep against the reference's Python, the C-ABI surface, and size-independent properties."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN, GRAD_NAMES, ROOT, args_to_numpy, rel_err, render_args


# ----------------------------------------------------------------------------- C-ABI surface
def _header_functions():
    src = open(os.path.join(ROOT, "include", "dm2_renderer.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dm2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from dmesh2_renderer_b200 import _lib
    names = _header_functions()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dm2_renderer.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    assert _lib.load_library().dm2_abi_version() == _lib.ABI_VERSION


def test_workspace_layouts_are_aligned_and_disjoint(built):
    from dmesh2_renderer_b200 import _C
    for layered in (False, True):
        L = _C.face_layout(3, 1000, layered)
        offs = [L.depths, L.min_depths, L.max_depths, L.tiles_touched, L.rects, L.records]
        assert all(o % 256 == 0 for o in offs) and offs == sorted(offs) and len(set(offs)) == len(offs)
        assert L.total >= L.records + 3 * 1000 * L.record_stride
    L = _C.binning_layout(12345)
    assert L.vals_unsorted >= L.keys_unsorted + 8 * 12345 and L.keys >= L.vals_unsorted + 4 * 12345
    assert L.total > L.sort_temp
    I = _C.image_layout(2, 50, 70)
    assert I.final_T >= I.ranges + 2 * 4 * 5 * 8 and I.total > I.n_contrib


def test_shade_cache_sizing_policy(built, monkeypatch):
    """Host-side shading-cache policy: off for K = 0 / tau = 0 / no instances / no grad; a first call reserves a
    fixed number of bytes per instance, later calls what the previous forward pass asked for plus slack; the
    pool is a 256-byte header + a whole number of 16-byte granules, capped."""
    import torch
    from dmesh2_renderer_b200 import _C
    cpu = torch.device("cpu")
    for args in ((0, 1000, 1.0), (20, 1000, 0.0), (20, 0, 1.0)):
        assert _C.shade_cache_alloc(*args, cpu).numel() == 0
    assert _C.shade_cache_alloc(20, 1000, 1.0, cpu, want=False).numel() == 0
    for R in (1, 7, 1000, 12345):
        pool = _C.shade_cache_alloc(20, R, 0.5, cpu)
        assert pool.dtype == torch.uint8 and (pool.numel() - _C.SHADE_CACHE_HEADER) % 16 == 0
        assert pool.numel() >= _C.SHADE_CACHE_HEADER + R * _C.SHADE_CACHE_FIRST_BYTES_PER_INSTANCE
    used = 54321
    pool = _C.shade_cache_alloc(20, 1000, 1.0, cpu, used_granules=used)
    assert used * 16 * _C.SHADE_CACHE_SLACK <= pool.numel() - _C.SHADE_CACHE_HEADER <= used * 16 * 1.2 + 8192
    # a pool that overflowed under-reports its demand: the next one doubles
    pool = _C.shade_cache_alloc(20, 1000, 1.0, cpu, used_granules=1000, prev_capacity_bytes=_C.SHADE_CACHE_HEADER + 16 * 900)
    assert pool.numel() >= 2 * (_C.SHADE_CACHE_HEADER + 16 * 900)
    monkeypatch.setattr(_C, "SHADE_CACHE_FIRST_MAX_BYTES", 16 * 700)
    assert _C.shade_cache_alloc(20, 1000, 1.0, cpu).numel() == _C.SHADE_CACHE_HEADER + 16 * 700
    monkeypatch.setattr(_C, "SHADE_CACHE_MAX_BYTES", 16 * 500)
    assert _C.shade_cache_alloc(20, 1000, 1.0, cpu).numel() == _C.SHADE_CACHE_HEADER + 16 * 500


def test_higher_msb_matches_reference_formula(built):
    from dmesh2_renderer_b200 import _lib
    from oracle import oracle as orc

    def ref(n):          # reference cuda_impl/renderer.cu:396-411
        msb, step = 16, 16
        while step > 1:
            step //= 2
            msb = msb + step if (n >> msb) else msb - step
        if n >> msb:
            msb += 1
        return msb
    lib = _lib.load_library()
    for n in [1, 2, 3, 4, 255, 256, 257, 4095, 4096, 4097, 16384, 65536, 2 ** 20 + 1, 2 ** 31 - 1]:
        assert lib.dm2_higher_msb(n) == ref(n) == orc.higher_msb(n)
    assert ref(4096) == 13          # cfg2: 45 sorted key bits


def test_no_cpu_fallback(built, tmp_path):
    from dmesh2_renderer_b200 import _lib, _C
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load_library(str(tmp_path / "missing.so"))
    a = torch.zeros(3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _C.sort_pairs(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int32), 40)
    del a


# ----------------------------------------------------------------------------- oracle pinned by the reference's Python
def test_oracle_aa_overlap_matches_reference_python(built):
    """oracle AA (restating cuda_impl/aa.h) vs the reference's pyrenderer.py analytic + autograd versions."""
    from oracle import oracle as orc
    g = np.load(os.path.join(GOLDEN, "aa_overlap_golden.npz"))
    n = g["tri"].shape[0]
    worst_a, worst_g, partial = 0.0, 0.0, 0
    for i in range(n):
        err, area, grad = orc.aa_overlap(g["verts"][i], g["edges"][i], g["iszero"][i], g["recip"][i], g["normal"][i],
                                         g["normal_c"][i], float(g["px"][i]), float(g["py"][i]))
        ra, rg = float(g["area_analytic"][i]), g["grad_analytic"][i]
        if err != 0:
            assert ra == 0.0      # the reference's analytic Function swallows the error and returns 0
            continue
        assert abs(area - ra) <= 2e-6, (i, area, ra)
        assert abs(area - float(g["area_autograd"][i])) <= 2e-6
        worst_a = max(worst_a, abs(area - ra))
        scale = max(1.0, float(np.abs(rg).max()))
        worst_g = max(worst_g, float(np.abs(grad - rg).max()) / scale)
        assert np.abs(grad - rg).max() <= 2e-5 * scale, (i, grad, rg)
        assert np.abs(grad - g["grad_autograd"][i]).max() <= 2e-4 * scale, (i, grad, g["grad_autograd"][i])
        partial += 0.0 < area < 1.0
    assert partial > 300


def test_host_prep_matches_reference_python_bit_exactly(built):
    """Renderer._init_rays / compute_verts_ndc_image / select_rays / Triangles vs the reference's own code."""
    from dmesh2_renderer_b200 import Renderer, Triangles, scenes
    g = np.load(os.path.join(GOLDEN, "prep_golden.npz"))
    sc = scenes.small(n_faces=200, size=48, n_views=3, seed=21)
    t = torch.from_numpy
    r = Renderer(t(sc.mv), t(sc.proj), sc.width, sc.height, "cpu")
    assert np.array_equal(r.ray_o.numpy(), g["ray_o"]) and np.array_equal(r.ray_d.numpy(), g["ray_d"])
    idx = [int(i) for i in g["idx"]]
    ndc, img = r.compute_verts_ndc_image(t(sc.verts), r.mv[idx], r.proj[idx])
    assert np.array_equal(ndc.numpy(), g["verts_ndc"]) and np.array_equal(img.numpy(), g["verts_image"])
    ro, rd = r.select_rays(idx, t(g["patch_min"]), 32, 40)
    assert np.array_equal(ro.numpy(), g["sel_ray_o"]) and np.array_equal(rd.numpy(), g["sel_ray_d"])
    tv = img[:, t(sc.faces).long().flatten()].view(-1, 3, 2)
    tr = Triangles(tv[:, 0], tv[:, 1], tv[:, 2])
    for a, k in ((tr.verts, "tri_verts"), (tr.edges, "tri_edges"), (tr.edges_iszero, "tri_iszero"),
                 (tr.edges_recip, "tri_recip"), (tr.edges_normal, "tri_normal"), (tr.edges_normal_c, "tri_normal_c")):
        assert np.array_equal(a.numpy(), g[k], equal_nan=True), k


def test_host_prep_autograd_flows_to_verts(built):
    from dmesh2_renderer_b200 import Renderer, Triangles, scenes
    sc = scenes.small(n_faces=30, size=32)
    t = torch.from_numpy
    r = Renderer(t(sc.mv), t(sc.proj), 32, 32, "cpu")
    v = t(sc.verts).clone().requires_grad_(True)
    ndc, img = r.compute_verts_ndc_image(v, r.mv[[0]], r.proj[[0]])
    tv = img[:, t(sc.faces).long().flatten()].view(-1, 3, 2)
    tr = Triangles(tv[:, 0], tv[:, 1], tv[:, 2])
    (tr.verts.sum() + ndc[..., 2].sum()).backward()
    assert v.grad is not None and torch.isfinite(v.grad).all() and v.grad.abs().sum() > 0


# ----------------------------------------------------------------------------- oracle self-consistency
def _oracle_small(tau=1.0, K=0, **kw):
    from dmesh2_renderer_b200 import scenes
    from oracle import oracle as orc
    sc = scenes.small(**kw)
    _r, _st, args = render_args(sc, "cpu", tau=tau)
    a = args_to_numpy(args)
    out = orc.render(a, sc.width, sc.height, tau, K=K, backward=True, dL_dcolor=sc.w_color, dL_ddepth=sc.w_depth)
    return sc, a, out


def test_oracle_binning_properties(built):
    from oracle import oracle as orc
    sc, a, out = _oracle_small(n_faces=500, size=64)
    R = out["num_rendered"]
    assert R == int(out["tiles_touched"].sum()) == int(out["face_offsets"][-1])
    tiles = (sc.width // 16) * (sc.height // 16)
    bit = orc.higher_msb(tiles)
    mask = np.uint64((1 << (32 + bit)) - 1)
    k = out["keys"] & mask
    assert np.all(k[1:] >= k[:-1])
    order = np.argsort(out["keys_unsorted"] & mask, kind="stable")
    assert np.array_equal(out["vals"], out["vals_unsorted"][order])
    tile_of = (out["keys"] >> np.uint64(32)).astype(np.int64)
    for tl in range(tiles):
        lo, hi = out["ranges"][tl]
        idx = np.nonzero(tile_of == tl)[0]
        if idx.size == 0:
            assert (lo, hi) == (0, 0)
        else:
            assert (lo, hi) == (idx[0], idx[-1] + 1)
    # equal (tile, depth) keys keep ascending face id
    same = (out["keys"][1:] == out["keys"][:-1])
    assert np.all(out["vals"][1:][same] > out["vals"][:-1][same])


def test_oracle_cache_equals_recompute(built):
    """The reference's K-deep AA cache (K=20) and the K=0 recompute path give identical results."""
    _sc, _a, o0 = _oracle_small(n_faces=400, size=48, seed=3, K=0)
    _sc, _a, o20 = _oracle_small(n_faces=400, size=48, seed=3, K=20)
    _sc, _a, o2 = _oracle_small(n_faces=400, size=48, seed=3, K=2)       # overflowing cache
    for k in ("color", "depth"):
        assert np.array_equal(o0[k], o20[k]) and np.array_equal(o0[k], o2[k]), k
    for k in GRAD_NAMES:       # fp64 sums, thread-order dependent in the last bits only
        assert rel_err(o20[k], o0[k]) < 1e-12 and rel_err(o2[k], o0[k]) < 1e-12, k


@pytest.mark.parametrize("tau", [1.0, 0.5])
def test_oracle_gradients_match_finite_differences(built, tau):
    """Directional finite differences of the oracle forward (fp32) vs its analytic backward."""
    from dmesh2_renderer_b200 import Triangles
    from oracle import oracle as orc
    sc, a, out = _oracle_small(tau=tau, n_faces=120, size=32, sigma=0.25, seed=5)
    rng = np.random.default_rng(0)

    def loss(args):
        o = orc.render(args, sc.width, sc.height, tau, K=0, backward=False)
        return float((o["color"].astype(np.float64) * sc.w_color).sum() + (o["depth"].astype(np.float64) * sc.w_depth).sum())

    def rederive(av):
        tv = torch.from_numpy(av.reshape(-1, 3, 2))
        # no CCW reorder here: the perturbation is tiny and the vertices already are CCW
        e = torch.stack((tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 1], tv[:, 0] - tv[:, 2]), dim=1)
        n = torch.stack((-e[..., 1], e[..., 0]), dim=-1)
        c = (n * tv).sum(-1)
        return (e.numpy().reshape(av.shape), (e.abs() < 1e-3).numpy().reshape(av.shape),
                (1.0 / e).numpy().reshape(av.shape), n.numpy().reshape(av.shape), c.numpy().reshape(av.shape[:2] + (3,)))

    checks = [("faces_opacity", "dL_dfaces_opacity", 2e-3), ("verts_color", "dL_dverts_color", 2e-3),
              ("faces_intense", "dL_dfaces_intense", 2e-3), ("aa_face_verts", "dL_daa_face_verts", 2e-3)]
    for name, gname, h in checks:
        d = rng.standard_normal(a[name].shape).astype(np.float32)
        plus, minus = dict(a), dict(a)
        plus[name] = (a[name] + h * d).astype(np.float32)
        minus[name] = (a[name] - h * d).astype(np.float32)
        if name == "aa_face_verts":
            if tau < 1.0:
                continue      # coverage switches discontinuously with the hit test when tau < 1
            for q in (plus, minus):
                (q["aa_face_edges"], q["aa_face_edges_iszero"], q["aa_face_edges_recip"], q["aa_face_edges_normal"],
                 q["aa_face_edges_normal_c"]) = rederive(q[name])
        fd = (loss(plus) - loss(minus)) / (2 * h)
        an = float((out[gname] * d).sum())
        assert abs(fd - an) <= 0.03 * max(abs(an), abs(fd)) + 0.05, (name, fd, an)
    # (verts_ndc.z is not finite-differenced: it also drives the depth sort, which is discontinuous)


def test_oracle_forward_invariants(built):
    sc, a, out = _oracle_small(n_faces=300, size=48, seed=2)
    assert np.all(out["final_T"] >= 0) and np.all(out["final_T"] <= 1)
    assert np.all(out["final_prev_T"] >= out["final_T"])
    untouched = out["n_contrib"] == 0
    assert np.all(out["color"][untouched] == 1.0) and np.all(out["depth"][untouched] == 1.0)
    # linearity of the backward in dL/dout
    from oracle import oracle as orc
    o2 = orc.render(a, sc.width, sc.height, 1.0, K=0, backward=True, dL_dcolor=2 * sc.w_color, dL_ddepth=2 * sc.w_depth)
    for k in GRAD_NAMES:
        assert rel_err(o2[k], 2 * out[k]) < 1e-5, k


def test_oracle_layers_properties(built):
    from dmesh2_renderer_b200 import LayeredRenderer, scenes
    from oracle import oracle as orc
    sc = scenes.small_tets(n_points=250, size=48)
    t = torch.from_numpy
    lr = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, "cpu")
    ndc, img = lr.compute_verts_ndc_image(t(sc.verts), lr.mv[[0]], lr.proj[[0]])
    out = orc.layers(sc.verts, sc.faces, sc.tets, sc.face_tets, sc.tet_faces, sc.face_existence, ndc.numpy(), img.numpy(),
                     lr.ray_o[[0]].numpy(), lr.ray_d[[0]].numpy(), sc.width, sc.height, sc.num_layers)
    L, c = out["render_layers"], out["render_layers_cnt"]
    assert c.max() > 0 and c.max() <= sc.num_layers
    k = np.arange(L.shape[-1])[None, None, None, :]
    valid = k < c[..., None]
    assert np.all(L[~valid] == -1) and np.all(sc.face_existence[L[valid]] != 0)
    # with every face existing, the first layer is the first hit
    out_all = orc.layers(sc.verts, sc.faces, sc.tets, sc.face_tets, sc.tet_faces, np.ones_like(sc.face_existence),
                         ndc.numpy(), img.numpy(), lr.ray_o[[0]].numpy(), lr.ray_d[[0]].numpy(), sc.width, sc.height, 3)
    walked = out_all["render_layers_cnt"] > 0
    assert np.array_equal(out_all["render_layers"][..., 0][walked], out_all["first_face"][walked])
    # a ray that enters the hull must hit something: the central pixel has layers
    assert out_all["render_layers_cnt"][0, sc.height // 2, sc.width // 2] > 0


# ----------------------------------------------------------------------------- oracle pinned by the compiled reference (B200 goldens)
def _golden(name):
    p = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(p):
        pytest.skip(f"{name}.npz not generated yet (tests/golden/make_ref_goldens.py on a GPU box)")
    return np.load(p)


@pytest.mark.parametrize("name,tau", [("ref_small_tau1", 1.0), ("ref_small_tau05", 0.5), ("ref_small_tau0", 0.0)])
def test_oracle_matches_reference_cuda_goldens(built, name, tau):
    g = _golden(name)
    _sc, _a, out = _oracle_small(tau=tau, K=20, n_faces=600, size=64)
    live = out["tiles_touched"] > 0
    assert np.array_equal(out["tiles_touched"], g["tiles_touched"])
    assert np.array_equal(out["depths"].view(np.uint32)[live], g["depths"].view(np.uint32)[live])
    assert out["num_rendered"] == int(g["num_rendered"])
    assert np.array_equal(out["keys"], g["keys"]) and np.array_equal(out["vals"], g["face_list"])
    assert np.array_equal(out["ranges"], g["ranges"])
    B, H, W = out["n_contrib"].shape
    mism = np.mean(out["n_contrib"].reshape(-1) != g["n_contrib"])
    assert mism <= 2e-3, mism
    assert rel_err(out["depth"], g["color_ndcdepth"]) <= 1e-4 or np.mean(np.abs(out["depth"] - g["color_ndcdepth"]) > 1e-4) <= 2e-3
    bad = np.mean(np.abs(out["color"] - g["color"]) > 1e-4)
    assert bad <= 2e-3, bad
    for k in GRAD_NAMES:
        if k in g.files:
            e = np.abs(out[k] - g[k]).max() / (np.abs(g[k]).max() + 1e-30)
            assert e <= 5e-3, (k, e)


@pytest.mark.parametrize("name,n_points,size", [("ref_tets_small", 300, 64)])
def test_oracle_layers_match_reference_cuda_goldens(built, name, n_points, size):
    from dmesh2_renderer_b200 import LayeredRenderer, scenes
    from oracle import oracle as orc
    g = _golden(name)
    sc = scenes.small_tets(n_points=n_points, size=size)
    t = torch.from_numpy
    lr = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, "cpu")
    ndc, img = lr.compute_verts_ndc_image(t(sc.verts), lr.mv[[0]], lr.proj[[0]])
    out = orc.layers(sc.verts, sc.faces, sc.tets, sc.face_tets, sc.tet_faces, sc.face_existence, ndc.numpy(), img.numpy(),
                     lr.ray_o[[0]].numpy(), lr.ray_d[[0]].numpy(), sc.width, sc.height, sc.num_layers)
    # NOTE: inputs are produced on the CPU here and on the GPU for the golden (matmul rounding may
    # differ in the last bit), so a handful of rays may legitimately differ.
    same = np.all(out["render_layers"] == g["render_layers"], axis=-1) & (out["render_layers_cnt"] == g["render_layers_cnt"])
    assert same.mean() >= 0.995, same.mean()


# ------------------------------------------------------------------------------------------------
# bench.py contract (no GPU in this container: the reference arm falls back to the CPU oracle port)
def test_bench_reference_arm_prints_one_json_line_without_a_gpu(built):
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU box: the reference arm runs the reference's CUDA extension there")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--faces", "3000",
                          "--size", "128"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # the CPU oracle is also timed on cfg1, for real
    assert d["cpu_baseline"]["cfg1"]["value"] > 0 and "cfg1" in d["cpu_baseline"]["cfg1"]["sample"]


def test_bench_algorithmic_bytes_follow_survey_8d():
    """A_fwd = G + 28 BHW + 28 R split over the stages, A_bwd = G + 28 BHW + 4 R + grads."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    B, F, H, W, R = 1, 1_000_000, 1024, 1024, 2_258_496
    P = 3 * F
    G = 12 * F + 24 * P + 4 * F + B * (20 * P + 4 * F + 24 * F) + 24 * B * H * W
    grads = 24 * P + 4 * F + B * (12 * P + 4 * F + 24 * F)
    assert bench.algorithmic_bytes("raster_forward", B, P, F, H, W, R) == G + 28 * B * H * W + 4 * R
    assert bench.algorithmic_bytes("raster_backward", B, P, F, H, W, R) == G + 28 * B * H * W + 4 * R + grads
    assert bench.algorithmic_bytes("tile_sort", B, P, F, H, W, R) == 24 * R
    # the forward stages together move the 28 R instance-list bytes of A_fwd (12 w + 12 r in the sort, 4 r in the raster)
    assert bench.algorithmic_bytes("tile_sort", B, P, F, H, W, R) + 4 * R == 28 * R
