"""GPU parity tests: the CUDA path (through the C-ABI, via the host mirror of the reference API)
against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact for integer work (tiles_touched, keys, sorted list,
tile ranges, n_contrib, first_face/first_tet, layer ids + counts); forward colour/depth within
rel 1e-4; gradients within rel 1e-3 (atomic-order nondeterminism), both measured against the
tensor's max magnitude, plus an element-wise outlier budget.
"""
import numpy as np
import pytest
import torch

from helpers import GRAD_NAMES, args_to_numpy, frac_outside, rel_err, render_args, scene_tensors

pytestmark = pytest.mark.gpu

FWD_RTOL = 1e-4
BWD_RTOL = 1e-3


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _oracle_and_cuda(sc, tau=1.0, views=None, patch_min=None, patch_wh=None, K=0, mutate_args=None):
    from dmesh2_renderer_b200 import _C, RenderFunction
    from oracle import oracle as orc
    dev = _dev()
    r, st, args = render_args(sc, dev, views=views, patch_min=patch_min, patch_wh=patch_wh, tau=tau)
    if mutate_args is not None:
        mutate_args(args)
    views = list(range(sc.mv.shape[0])) if views is None else list(views)
    B = len(views)
    pw, ph = args[2], args[3]
    leaf_idx = [4, 6, 7, 8, 10, 12]
    leaves = []
    for i in leaf_idx:
        args[i] = args[i].detach().clone().requires_grad_(True)
        leaves.append(args[i])
    color, depth = RenderFunction.apply(*args)
    if patch_min is None and patch_wh is None:
        wc, wd = st["w_color"][views], st["w_depth"][views]
    else:
        g = torch.Generator(device="cpu").manual_seed(5)
        wc = (torch.rand((B, ph, pw, 3), generator=g) * 2 - 1).to(dev)
        wd = (torch.rand((B, ph, pw), generator=g) * 2 - 1).to(dev)
    loss = (color * wc).sum() + (depth * wd).sum()
    loss.backward()
    cuda = dict(color=color.detach().cpu().numpy(), depth=depth.detach().cpu().numpy())
    for n, l in zip(GRAD_NAMES, leaves):
        cuda[n] = l.grad.detach().cpu().numpy()
    # taps: run the raw entry point once more to get at the scratch buffers
    with torch.no_grad():
        res = _C.render_forward_cuda(*[a.detach() if isinstance(a, torch.Tensor) else a for a in args])
    R, _c, _d, _, _, _, _, fb, bb, ib = res
    F = args[5].shape[0]
    cuda["num_rendered"] = R
    cuda.update({k: v.cpu().numpy() for k, v in _C.face_taps(fb, B, F).items()})
    cuda.update({k: v.cpu().numpy() for k, v in _C.binning_taps(bb, R).items()})
    cuda.update({k: v.cpu().numpy() for k, v in _C.image_taps(ib, B, ph, pw).items()})
    assert np.array_equal(_c.cpu().numpy(), cuda["color"]), "forward is not run-to-run deterministic"
    ref = orc.render(args_to_numpy(args), pw, ph, tau, K=K, backward=True, dL_dcolor=wc.cpu().numpy(),
                     dL_ddepth=wd.cpu().numpy())
    return cuda, ref


def _check_ints(cuda, ref):
    live = ref["tiles_touched"] > 0
    assert np.array_equal(cuda["tiles_touched"].view(np.uint32), ref["tiles_touched"])
    assert cuda["num_rendered"] == ref["num_rendered"]
    for k in ("depths", "min_depths", "max_depths"):
        assert np.array_equal(cuda[k].view(np.uint32)[live], ref[k].view(np.uint32)[live]), k
    # the unsorted buffers hold the emitted pairs grouped by tile in scatter order: same multiset
    assert np.array_equal(np.sort(cuda["keys_unsorted"].view(np.uint64)), np.sort(ref["keys_unsorted"]))
    assert np.array_equal(cuda["keys"].view(np.uint64), ref["keys"]), "sorted keys differ"
    assert np.array_equal(cuda["vals"].view(np.uint32), ref["vals"]), "sorted face list differs"
    assert np.array_equal(cuda["ranges"].view(np.uint32), ref["ranges"]), "tile ranges differ"
    assert np.array_equal(cuda["n_contrib"].view(np.uint32), ref["n_contrib"]), "n_contrib differs"


def _check_floats(cuda, ref, skip_grads=()):
    for k in ("color", "depth", "final_T", "final_prev_T"):
        assert rel_err(cuda[k], ref[k]) <= FWD_RTOL, (k, rel_err(cuda[k], ref[k]))
        assert frac_outside(cuda[k], ref[k], FWD_RTOL, 1e-6) == 0.0, k
    for k in GRAD_NAMES:
        if k in skip_grads:
            continue
        e = rel_err(cuda[k], ref[k])
        assert e <= BWD_RTOL, (k, e)
        assert frac_outside(cuda[k], ref[k], BWD_RTOL, BWD_RTOL * np.abs(ref[k]).max() + 1e-12) == 0.0, k


@pytest.mark.parametrize("tau", [1.0, 0.5, 0.0])
def test_render_small_scene(built, tau):
    from dmesh2_renderer_b200 import scenes
    cuda, ref = _oracle_and_cuda(scenes.small(n_faces=600, size=64), tau=tau)
    _check_ints(cuda, ref)
    _check_floats(cuda, ref)


def test_render_multiview_ragged(built):
    """3 views, image size not a multiple of 16, large triangles (long lists, big tile bboxes)."""
    from dmesh2_renderer_b200 import scenes
    sc = scenes.triangle_soup("ragged", 300, 72, 0.35, seed=9, eyes=[tuple(e) for e in scenes.fibonacci_eyes(64)[[3, 17, 40]]])
    sc.height = 56
    sc.w_color = sc.w_color[:, :56]
    sc.w_depth = sc.w_depth[:, :56]
    cuda, ref = _oracle_and_cuda(sc, views=[2, 0, 1])
    _check_ints(cuda, ref)
    _check_floats(cuda, ref)


def test_render_patch_mode(built):
    """Random sub-windows (patch_min != 0, patch smaller than the frame, not a multiple of 16)."""
    from dmesh2_renderer_b200 import scenes
    sc = scenes.small(n_faces=500, size=96, sigma=0.15, seed=13, n_views=2)
    cuda, ref = _oracle_and_cuda(sc, views=[1, 0], patch_min=[[7, 20], [40, 3]], patch_wh=(50, 41))
    _check_ints(cuda, ref)
    _check_floats(cuda, ref)


def test_render_matches_reference_cache_semantics(built):
    """The oracle run WITH the reference's K=20 AA cache gives the same answer as the cache-free CUDA path."""
    from dmesh2_renderer_b200 import scenes
    cuda, ref = _oracle_and_cuda(scenes.small(n_faces=400, size=48, seed=3), K=20)
    _check_ints(cuda, ref)
    _check_floats(cuda, ref)


@pytest.mark.parametrize("fused", [False, True])
def test_shade_cache_on_off_overflow_agree(built, fused, monkeypatch):
    """Backward with the forward->backward shading cache (K > 0), without it (K = 0: hit point, clamp and clip
    are recomputed) and with a pool far too small (most blocks fall back to recomputation) must agree: the
    cached values ARE the recomputed values, so only the atomic order differs."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import _C, scenes
    dev = _dev()
    sc = scenes.small(n_faces=1500, size=96, sigma=0.08, seed=21)
    st = scene_tensors(sc, dev)

    def run(K, bytes_per_instance):
        from dmesh2_renderer_b200.fused import FusedRenderFunction
        FusedRenderFunction._size_hint.clear()          # first-call sizing policy for every run
        monkeypatch.setattr(_C, "SHADE_CACHE_FIRST_BYTES_PER_INSTANCE", bytes_per_instance)
        r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev, aa_grad_buffer_size=K)
        r.fused = fused
        leaves = {k: st[k].detach().clone().requires_grad_(True)
                  for k in ("verts", "verts_color", "faces_opacity", "faces_intense")}
        pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        color, depth = r([0], pm, sc.width, sc.height, leaves["verts"], st["faces"], leaves["verts_color"],
                         leaves["faces_opacity"], leaves["faces_intense"], st["background"], 1.0)
        ((color * st["w_color"][:1]).sum() + (depth * st["w_depth"][:1]).sum()).backward()
        return color.detach(), {k: v.grad.detach().cpu().numpy() for k, v in leaves.items()}

    c_on, g_on = run(20, 4096)
    c_off, g_off = run(0, 4096)
    c_ovf, g_ovf = run(20, 24)
    assert torch.equal(c_on, c_off) and torch.equal(c_on, c_ovf), "the cache must not change the forward result"
    for k in g_on:
        assert np.abs(g_on[k]).max() > 0
        for other in (g_off, g_ovf):
            assert rel_err(other[k], g_on[k]) <= 1e-5, (k, rel_err(other[k], g_on[k]))


@pytest.mark.parametrize("n_faces,longest", [(12_000, 2048), (45_000, 8192), (130_000, 24576)])
def test_long_tile_lists_sort_paths(built, n_faces, longest):
    """4 tiles with > 2 048, > 8 192 and > 24 576 instances each (the three longer tiers of the tile-segmented
    sort: depth-binned in 64 KB / 192 KB of shared memory, block-wide network in global memory): sorted keys,
    values and ranges must still equal the stable sort of the oracle."""
    from dmesh2_renderer_b200 import scenes
    sc = scenes.triangle_soup("long_lists", n_faces, 32, 0.01, seed=31)
    cuda, ref = _oracle_and_cuda(sc)
    lens = ref["ranges"][:, 1].astype(np.int64) - ref["ranges"][:, 0]
    assert lens.max() > longest
    _check_ints(cuda, ref)
    _check_floats(cuda, ref)


def _snap_screen_geometry(step):
    """args mutator: snaps the projected vertices to a grid of `step` pixels and rebuilds the AA tensors from
    them exactly as Renderer.prepare_inputs does - vertices ON pixel corners / edges, exactly horizontal and
    vertical edges (edges_iszero, infinite reciprocals), zero-area triangles: the reference's AA error codes
    1-6 (aa.h:265-437, "skip the face for this pixel") fire all over the image."""
    def mutate(args):
        from dmesh2_renderer_b200.triangles import Triangles
        img = torch.round(args[9] / step) * step
        faces = args[5].long()
        B, F = img.shape[0], faces.shape[0]
        corners = img[:, faces.flatten()].view(-1, 3, 2)
        tri = Triangles(corners[:, 0], corners[:, 1], corners[:, 2])
        args[9] = img
        args[12] = tri.verts.reshape(B, F, 3, 2).contiguous()
        args[13] = tri.edges.reshape(B, F, 3, 2).contiguous()
        args[14] = tri.edges_iszero.reshape(B, F, 3, 2).contiguous()
        args[15] = tri.edges_recip.reshape(B, F, 3, 2).contiguous()
        args[16] = tri.edges_normal.reshape(B, F, 3, 2).contiguous()
        args[17] = tri.edges_normal_c.reshape(B, F, 3).contiguous()
    return mutate


@pytest.mark.parametrize("step,tau", [(1.0, 1.0), (0.5, 1.0), (4.0, 0.5)])
def test_pixel_aligned_and_degenerate_screen_geometry(built, step, tau):
    from dmesh2_renderer_b200 import scenes
    sc = scenes.small(n_faces=900, size=64, sigma=0.1, seed=17)
    cuda, ref = _oracle_and_cuda(sc, tau=tau, mutate_args=_snap_screen_geometry(step))
    _check_ints(cuda, ref)
    # step 0.5 puts vertices exactly ON pixel edges (not corners: those raise error 1 and the face is skipped):
    # the overlap area has a kink there and d(area)/d(vertex) is one-sided.  The reference differentiates
    # "the intersection point slides along the pixel edge" (Jacobian chain, aa.h:280-294); the clip detects
    # these pairs and evaluates the reference's chain for them (aa_grad_reference), so ALL gradients agree.
    _check_floats(cuda, ref)


def test_inconsistent_polygon_near_a_pixel_corner(built):
    """A triangle edge that passes within an ulp of a pixel corner at coordinates ~900: the reference's
    inside-corner flags and its edge intersections disagree, it appends two spurious corners and its chained
    d(area)/d(verts) is O(1) off the true derivative (found at cfg2 by the full-size parity test: face 820444,
    pixel (179, 910)).  The forward result already matched bit for bit; the gradient must follow the reference."""
    from dmesh2_renderer_b200 import scenes
    from dmesh2_renderer_b200.triangles import Triangles
    tri = torch.tensor([[180.46987915039062, 902.59228515625], [184.44287109375, 902.9342651367188],
                        [178.21258544921875, 914.2448120117188]], dtype=torch.float32)

    def mutate(args, K=20):
        args[18] = K                                            # shading cache on (forward clip) / off (backward clip)
        img = args[9].clone()
        faces = args[5].long()
        img[0, faces[0]] = tri.to(img.device)                   # face 0 lands on the problematic spot
        B, F = img.shape[0], faces.shape[0]
        corners = img[:, faces.flatten()].view(-1, 3, 2)
        t = Triangles(corners[:, 0], corners[:, 1], corners[:, 2])
        args[9] = img
        args[12] = t.verts.reshape(B, F, 3, 2).contiguous()
        args[13] = t.edges.reshape(B, F, 3, 2).contiguous()
        args[14] = t.edges_iszero.reshape(B, F, 3, 2).contiguous()
        args[15] = t.edges_recip.reshape(B, F, 3, 2).contiguous()
        args[16] = t.edges_normal.reshape(B, F, 3, 2).contiguous()
        args[17] = t.edges_normal_c.reshape(B, F, 3).contiguous()

    sc = scenes.triangle_soup("corner_case", 3000, 1024, 0.02, seed=77)
    for K in (20, 0):
        cuda, ref = _oracle_and_cuda(sc, patch_min=[[170, 896]], patch_wh=(24, 24),
                                     mutate_args=lambda a, K=K: mutate(a, K), K=0)
        _check_ints(cuda, ref)
        _check_floats(cuda, ref)
        g = ref["dL_daa_face_verts"][0, 0]
        assert np.abs(g).max() > 0, "the special face must contribute"
        assert np.abs(cuda["dL_daa_face_verts"][0, 0] - g).max() <= BWD_RTOL * np.abs(ref["dL_daa_face_verts"]).max()


def test_large_triangles_cover_many_tiles(built):
    """A few huge triangles: hundreds of tiles per face (warp-cooperative count / scatter), mostly fully covered
    pixels, long identical lists in every tile."""
    from dmesh2_renderer_b200 import scenes
    sc = scenes.triangle_soup("huge", 60, 160, 0.9, seed=23)
    cuda, ref = _oracle_and_cuda(sc)
    assert cuda["tiles_touched"].max() > 64
    _check_ints(cuda, ref)
    _check_floats(cuda, ref)


def test_faces_around_and_behind_the_camera(built):
    """Vertices beside, at and behind the eye (|w| small or negative: the +-1e-4 clamp of
    compute_verts_ndc_image, huge screen coordinates, no t > 0 test in the compositor)."""
    from dmesh2_renderer_b200 import scenes
    sc = scenes.triangle_soup("around_eye", 500, 64, 0.25, seed=29)
    sc.verts[:, 2] = sc.verts[:, 2] * 2.5 + 1.5          # z in [-1, 4]; the eye sits at z = 3
    cuda, ref = _oracle_and_cuda(sc)
    _check_ints(cuda, ref)
    _check_floats(cuda, ref)


def test_cfg1_vs_oracle(built):
    """BASELINE config 0: 10k triangles at 256x256 (the oracle takes a few seconds)."""
    from dmesh2_renderer_b200 import scenes
    cuda, ref = _oracle_and_cuda(scenes.cfg1())
    _check_ints(cuda, ref)
    _check_floats(cuda, ref)


def test_empty_and_degenerate(built):
    from dmesh2_renderer_b200 import Renderer, scenes
    dev = _dev()
    sc = scenes.small(n_faces=50, size=32)
    st = scene_tensors(sc, dev)
    r = Renderer(st["mv"], st["proj"], 32, 32, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    # everything behind the far plane: nothing rendered -> background, depth_ndc = 1 -> depth 0
    far = st["verts"] * 0.01 + torch.tensor([0.0, 0.0, -50.0], device=dev)
    color, depth = r([0], pm, 32, 32, far, st["faces"], st["verts_color"], st["faces_opacity"], st["faces_intense"],
                     st["background"])
    assert torch.all(color == 1.0) and torch.all(depth == 0.0)
    # zero faces
    f0 = st["faces"][:0]
    color, depth = r([0], pm, 32, 32, st["verts"], f0, st["verts_color"], st["faces_opacity"][:0],
                     st["faces_intense"][:, :0], st["background"])
    assert torch.all(color == 1.0)
    # degenerate (zero-area) triangles must not produce NaNs
    deg = st["verts"].clone()
    deg[1::3] = deg[0::3]
    color, depth = r([0], pm, 32, 32, deg, st["faces"], st["verts_color"], st["faces_opacity"], st["faces_intense"],
                     st["background"])
    assert torch.isfinite(color).all() and torch.isfinite(depth).all()


def test_error_behaviour(built):
    from dmesh2_renderer_b200 import _C, scenes
    dev = _dev()
    _, _, args = render_args(scenes.small(n_faces=20, size=32), dev)
    bad = list(args)
    bad[11] = 1.5
    with pytest.raises(RuntimeError, match="aa_temperature"):
        _C.render_forward_cuda(*bad)
    bad = list(args)
    bad[0] = torch.ones(4, device=dev)
    with pytest.raises(RuntimeError, match="background"):
        _C.render_forward_cuda(*bad)
    bad = list(args)
    bad[7] = bad[7][:-1]
    with pytest.raises(RuntimeError, match="face opacity"):
        _C.render_forward_cuda(*bad)


def test_patch_bounds_are_checked_without_a_host_sync(built):
    """The reference asserts patch_min + patch size <= frame on the host (a device sync per call for CUDA
    tensors); the fused path checks inside the preprocess kernel and raises with the same message."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    dev = _dev()
    sc = scenes.small(n_faces=50, size=64)
    st = scene_tensors(sc, dev)
    r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    args = (st["verts"], st["faces"], st["verts_color"], st["faces_opacity"], st["faces_intense"][:1], st["background"])
    ok = torch.tensor([[24, 16]], dtype=torch.int32, device=dev)
    r([0], ok, 40, 48, *args)
    with pytest.raises(RuntimeError, match="b_patch_max_x"):
        r([0], torch.tensor([[25, 16]], dtype=torch.int32, device=dev), 40, 48, *args)
    with pytest.raises(RuntimeError, match="b_patch_max_y"):
        r([0], torch.tensor([[24, 17]], dtype=torch.int32, device=dev), 40, 48, *args)
    with pytest.raises(AssertionError, match="b_patch_max_x"):
        r([0], torch.tensor([[25, 16]], dtype=torch.int32), 40, 48, *args)


def test_sort_pairs_stable_and_bit_range(built):
    from dmesh2_renderer_b200 import _C
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(0)
    for n, end_bit in ((1, 40), (31, 33), (4096, 45), (4097, 45), (100_003, 47), (1_000_000, 45)):
        keys = torch.randint(0, 2 ** 62, (n,), generator=g, dtype=torch.int64)
        keys[::3] = keys[::3] & 0xFFFF        # lots of duplicates
        vals = torch.arange(n, dtype=torch.int32)
        kout, vout = _C.sort_pairs(keys.to(dev), vals.to(dev), end_bit)
        mask = (1 << end_bit) - 1
        order = np.argsort(keys.numpy() & mask, kind="stable")
        assert np.array_equal(vout.cpu().numpy(), vals.numpy()[order]), (n, end_bit)
        assert np.array_equal(kout.cpu().numpy(), keys.numpy()[order]), (n, end_bit)


def test_inclusive_scan(built):
    from dmesh2_renderer_b200 import _C
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(1)
    for n in (1, 255, 2048, 2049, 1_000_003):
        x = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32)
        out = _C.inclusive_scan_u32(x.to(dev))
        assert np.array_equal(out.cpu().numpy().astype(np.int64), np.cumsum(x.numpy().astype(np.int64)))


def _layers_both(sc, views=None):
    from dmesh2_renderer_b200 import LayeredRenderer, _C
    from oracle import oracle as orc
    dev = _dev()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    lr = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    views = list(range(sc.mv.shape[0])) if views is None else views
    layers, cnt = lr.generate(views, t(sc.verts), t(sc.faces), t(sc.tets), t(sc.face_tets), t(sc.tet_faces),
                              t(sc.face_existence), sc.num_layers)
    ndc, img = lr.compute_verts_ndc_image(t(sc.verts), lr.mv[views], lr.proj[views])
    ro, rd = lr.ray_o[views], lr.ray_d[views]
    _l, _c, taps = _C.generate_render_layers_cuda(sc.width, sc.height, t(sc.verts), t(sc.faces), t(sc.tets),
                                                  t(sc.face_tets), t(sc.tet_faces), t(sc.face_existence), ndc, img,
                                                  ro, rd, sc.num_layers, return_taps=True)
    assert torch.equal(_l, layers) and torch.equal(_c, cnt)
    # the plain C-ABI pair (geometry records built after the header read-back instead of during it): same result
    _l2, _c2 = _C.generate_render_layers_cuda(sc.width, sc.height, t(sc.verts), t(sc.faces), t(sc.tets),
                                              t(sc.face_tets), t(sc.tet_faces), t(sc.face_existence), ndc, img,
                                              ro, rd, sc.num_layers, overlap_geometry=False)
    assert torch.equal(_l2, layers) and torch.equal(_c2, cnt)
    B = len(views)
    img_t = _C.image_taps(taps["image_buffer"], B, sc.height, sc.width, layered=True)
    bin_t = _C.binning_taps(taps["binning_buffer"], taps["num_rendered"])
    ref = orc.layers(sc.verts, sc.faces, sc.tets, sc.face_tets, sc.tet_faces, sc.face_existence, ndc.cpu().numpy(),
                     img.cpu().numpy(), ro.cpu().numpy(), rd.cpu().numpy(), sc.width, sc.height, sc.num_layers)
    assert taps["num_rendered"] == ref["num_rendered"]
    assert np.array_equal(bin_t["keys"].cpu().numpy().view(np.uint64), ref["keys"])
    assert np.array_equal(bin_t["vals"].cpu().numpy().view(np.uint32), ref["vals"])
    assert np.array_equal(img_t["ranges"].cpu().numpy().view(np.uint32), ref["ranges"])
    assert np.array_equal(img_t["first_face"].cpu().numpy(), ref["first_face"])
    assert np.array_equal(img_t["first_tet"].cpu().numpy(), ref["first_tet"])
    assert np.array_equal(layers.cpu().numpy(), ref["render_layers"]), "layer ids differ"
    assert np.array_equal(cnt.cpu().numpy(), ref["render_layers_cnt"]), "layer counts differ"
    return layers, cnt


def test_layers_small(built):
    from dmesh2_renderer_b200 import scenes
    layers, cnt = _layers_both(scenes.small_tets(n_points=300, size=64))
    assert int(cnt.max()) > 0


def test_layers_ragged_multiview(built):
    from dmesh2_renderer_b200 import scenes
    sc = scenes.delaunay_tets("tets_mv", 500, 80, seed=5, p_exist=0.3, num_layers=6,
                              eyes=[tuple(e) for e in scenes.fibonacci_eyes(64)[[1, 30]]])
    sc.height = 50
    _layers_both(sc, views=[1, 0])


@pytest.mark.parametrize("n,size,eye", [(5, 64, (0.0, 0.0, 3.0)), (6, 96, (0.0, 0.0, 3.0)), (5, 80, (3.0, 0.0, 0.0)),
                                        (4, 64, (1.7, 1.7, 1.7))])
def test_layers_lattice_degenerate_rays(built, n, size, eye):
    """Points on a regular lattice seen along an axis / the space diagonal: cospherical points (flat slivers in the
    joggled Delaunay mesh), many rays through vertices, along edges and inside faces.  The walk's fast path (signs of
    edge functions with a robustness margin) must hand every such step to the reference's three hit tests: ids and
    counts equal the oracle's on every ray."""
    from dmesh2_renderer_b200 import scenes
    g = np.linspace(-1.0, 1.0, n)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    sc = scenes.delaunay_tets(f"lattice_{n}", 0, size, seed=2, p_exist=0.35, num_layers=6, eyes=[eye], points=pts,
                              qhull_options="QJ Qbb Qc")
    layers, cnt = _layers_both(sc)
    assert int(cnt.max()) > 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_layers_inconsistent_adjacency(built, seed):
    """Adjacency tensors that contradict the geometry - a face id repeated inside a tet, a face listed by a tet it
    does not belong to, wrong / missing face_tets entries, a tet with a repeated vertex: the reference's walk stops
    (or wanders) in a defined way on such input (forward.cu:880-893, 985-992); tets that are not regular never take
    the fast path, so the result is the oracle's on every ray."""
    from dmesh2_renderer_b200 import scenes
    sc = scenes.small_tets(n_points=250, size=64, seed=21 + seed, num_layers=5)
    rng = np.random.default_rng(100 + seed)
    T, F = sc.tets.shape[0], sc.faces.shape[0]
    tf, ft, tets = sc.tet_faces.copy(), sc.face_tets.copy(), sc.tets.copy()
    for t in rng.choice(T, 25, replace=False):
        kind = rng.integers(0, 4)
        if kind == 0:
            tf[t, rng.integers(0, 4)] = tf[t, rng.integers(0, 4)]            # repeated face id
        elif kind == 1:
            tf[t, rng.integers(0, 4)] = rng.integers(0, F)                    # a face of some other tet
        elif kind == 2:
            tets[t, rng.integers(0, 4)] = tets[t, rng.integers(0, 4)]        # repeated vertex
        else:
            tets[t, rng.integers(0, 4)] = rng.integers(0, sc.verts.shape[0])  # a vertex that is not on its faces
    for f in rng.choice(F, 40, replace=False):
        kind = rng.integers(0, 3)
        if kind == 0:
            ft[f, rng.integers(0, 2)] = rng.integers(0, T)                    # wrong neighbour
        elif kind == 1:
            ft[f, rng.integers(0, 2)] = -1                                    # missing neighbour
        else:
            ft[f] = ft[f, ::-1]                                               # swapped
    sc.tet_faces, sc.face_tets, sc.tets = tf, ft, tets
    _layers_both(sc)


def test_layers_incremental_regenerate(built):
    """Same geometry, new face_existence: the walk-only path must give what a full generate gives, and
    an in-place geometry update must invalidate the cache."""
    from dmesh2_renderer_b200 import LayeredRenderer, scenes
    dev = _dev()
    sc = scenes.small_tets(n_points=400, size=80, seed=5, num_layers=5)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    lr = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    full = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    full.incremental = False
    geo = [t(getattr(sc, k)) for k in ("verts", "faces", "tets", "face_tets", "tet_faces")]
    rng = np.random.default_rng(0)
    for step in range(4):
        ex = t((rng.random(sc.faces.shape[0]) < 0.1 + 0.2 * step).astype(np.int32))
        a = lr.generate([0], *geo, ex, sc.num_layers)
        b = full.generate([0], *geo, ex, sc.num_layers)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), step
        if step == 0:
            assert "_layer_cache" in lr.__dict__
        if step == 2:                      # move the points in place: version bump -> full path again
            geo[0].mul_(0.97)
    taps_before = lr.__dict__["_layer_cache"]["taps"]
    lr.generate([0], *geo, ex, 3)          # a different layer count reuses the same cache entry (walk only)
    assert lr.__dict__["_layer_cache"]["taps"] is taps_before


def test_layers_properties(built):
    """Size-independent properties: every recorded id exists, no id repeats along a ray, entries beyond cnt are -1."""
    from dmesh2_renderer_b200 import scenes
    sc = scenes.delaunay_tets("tets_prop", 3000, 256, seed=8, p_exist=0.2, num_layers=8)
    layers, cnt = _layers_both(sc)
    L = layers.cpu().numpy()
    c = cnt.cpu().numpy()
    k = np.arange(L.shape[-1])[None, None, None, :]
    valid = k < np.minimum(c, L.shape[-1])[..., None]
    assert np.all(L[~valid] == -1)
    assert np.all(sc.face_existence[L[valid]] != 0)
    Ls = np.sort(np.where(valid, L, -np.arange(1, L.shape[-1] + 1)[None, None, None, :]), axis=-1)
    assert np.all(Ls[..., 1:] != Ls[..., :-1])


# ------------------------------------------------------------------------------------------------
# against the compiled reference: committed goldens (tests/golden/ref_*.npz, produced on a B200 by
# tests/golden/make_ref_goldens.py) and, when baseline/_ref travelled to the box, the live extension
import os

from helpers import GOLDEN, load_reference


def _golden(name):
    p = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(p):
        pytest.skip(f"{name}.npz missing")
    return np.load(p)


def _check_against_reference(cuda, g, has_ndc=True):
    live = g["tiles_touched"] > 0
    assert np.array_equal(cuda["tiles_touched"].view(np.uint32), g["tiles_touched"])
    assert np.array_equal(cuda["depths"].view(np.uint32)[live], g["depths"].view(np.uint32)[live])
    assert cuda["num_rendered"] == int(g["num_rendered"])
    assert np.array_equal(cuda["keys"].view(np.uint64), g["keys"]), "sorted keys differ from the reference"
    assert np.array_equal(cuda["vals"].view(np.uint32), g["face_list"]), "sorted face list differs from the reference"
    assert np.array_equal(cuda["ranges"].view(np.uint32), g["ranges"]), "tile ranges differ from the reference"
    # float outputs: rel 1e-4 forward on every pixel / 1e-3 backward on every element (north_star); n_contrib
    # is a discrete per-pixel decision and must match on every pixel
    assert np.array_equal(cuda["n_contrib"].reshape(-1).view(np.uint32), g["n_contrib"].reshape(-1)), "n_contrib differs"
    assert np.mean(np.abs(cuda["color"] - g["color"]) > 1e-4 * (1 + np.abs(g["color"]))) == 0.0
    assert np.mean(np.abs(cuda["depth"] - g["color_ndcdepth"]) > 1e-4 * (1 + np.abs(g["color_ndcdepth"]))) == 0.0
    for k in GRAD_NAMES:
        if k in g.files:
            e = np.abs(cuda[k] - g[k]).max() / (np.abs(g[k]).max() + 1e-30)
            assert e <= BWD_RTOL, (k, e)
            assert rel_err(cuda[k].sum(), g[k].astype(np.float64).sum()) <= 1e-2 or abs(g[k].sum()) < 1e-3


@pytest.mark.parametrize("name,tau", [("ref_small_tau1", 1.0), ("ref_small_tau05", 0.5), ("ref_small_tau0", 0.0)])
def test_cuda_matches_reference_goldens_small(built, name, tau):
    from dmesh2_renderer_b200 import scenes
    cuda, _ref = _oracle_and_cuda(scenes.small(n_faces=600, size=64), tau=tau)
    _check_against_reference(cuda, _golden(name))


def test_cuda_matches_reference_golden_cfg1(built):
    from dmesh2_renderer_b200 import scenes
    cuda, _ref = _oracle_and_cuda(scenes.cfg1())
    _check_against_reference(cuda, _golden("ref_cfg1"))


@pytest.mark.parametrize("name", ["ref_tets_small", "ref_tets_3000"])
def test_layers_match_reference_goldens(built, name):
    from dmesh2_renderer_b200 import scenes
    g = _golden(name)
    sc = scenes.small_tets(n_points=300, size=64) if name == "ref_tets_small" else \
        scenes.delaunay_tets("tets_prop", 3000, 256, seed=8, p_exist=0.2, num_layers=8)
    layers, cnt = _layers_both(sc)
    assert np.array_equal(layers.cpu().numpy(), g["render_layers"]), "layer ids differ from the reference"
    assert np.array_equal(cnt.cpu().numpy(), g["render_layers_cnt"])


@pytest.mark.parametrize("exact_clip", [False, True])
def test_live_reference_mid_size(built, exact_clip):
    """200k triangles at 512x512 through BOTH public APIs on the same GPU (skipped without baseline/_ref).  With the
    exact polygon clip for every pair (dm2_set_exact_clip) the forward images are the reference's bit for bit; with the
    polygon-free overlap area (the default) they are within a few ulp."""
    ref = load_reference()
    if ref is None:
        pytest.skip("baseline/_ref not present on this box")
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    dev = _dev()
    prev = dm.set_exact_clip(exact_clip)
    try:
        _live_reference_mid_size(dm, ref, scenes, dev, exact_clip)
    finally:
        dm.set_exact_clip(prev)


def _live_reference_mid_size(dm, ref, scenes, dev, exact_clip):
    sc = scenes.triangle_soup("mid", 200_000, 512, 0.02, seed=12)
    st = scene_tensors(sc, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    outs = []
    for pkg in (dm, ref):
        r = pkg.Renderer(st["mv"], st["proj"], 512, 512, dev)
        p = {k: st[k].detach().clone().requires_grad_(True) for k in ("verts", "verts_color", "faces_opacity", "faces_intense")}
        color, depth = r([0], pm, 512, 512, p["verts"], st["faces"], p["verts_color"], p["faces_opacity"],
                         p["faces_intense"], st["background"])
        ((color * st["w_color"]).sum() + (depth * st["w_depth"]).sum()).backward()
        outs.append((color.detach().cpu().numpy(), depth.detach().cpu().numpy(), {k: v.grad.cpu().numpy() for k, v in p.items()}))
    (c0, d0, g0), (c1, d1, g1) = outs
    assert np.mean(np.abs(c0 - c1) > 1e-4 * (1 + np.abs(c1))) == 0.0
    assert np.mean(np.abs(d0 - d1) > 1e-4 * (1 + np.abs(d1))) == 0.0
    if exact_clip:
        assert np.array_equal(c0, c1) and np.array_equal(d0, d1), float(np.abs(c0 - c1).max())
    else:
        assert float(np.abs(c0 - c1).max()) <= 2e-6 and float(np.abs(d0 - d1).max()) <= 2e-6
    for k in g0:
        e = np.abs(g0[k] - g1[k]).max() / (np.abs(g1[k]).max() + 1e-30)
        assert e <= BWD_RTOL, (k, e)


# ------------------------------------------------------------------------------------------------
# fused host prep (Renderer.forward default) vs the reference-shaped RenderFunction path
@pytest.mark.parametrize("tau,patch", [(1.0, None), (0.5, None), (1.0, ([[7, 20], [40, 3]], (50, 41)))])
def test_fused_path_equals_reference_shaped_path(built, tau, patch):
    from dmesh2_renderer_b200 import Renderer, scenes
    dev = _dev()
    sc = scenes.small(n_faces=700, size=96, sigma=0.15, seed=31, n_views=2)
    st = scene_tensors(sc, dev)
    views = [1, 0]
    if patch is None:
        pm, (pw, ph) = torch.zeros((2, 2), dtype=torch.int32, device=dev), (sc.width, sc.height)
    else:
        pm, (pw, ph) = torch.tensor(patch[0], dtype=torch.int32, device=dev), patch[1]
    g = torch.Generator(device="cpu").manual_seed(3)
    wc = (torch.rand((2, ph, pw, 3), generator=g) * 2 - 1).to(dev)
    wd = (torch.rand((2, ph, pw), generator=g) * 2 - 1).to(dev)
    res = []
    for fused in (True, False):
        r = Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
        r.fused = fused
        p = {k: st[k].detach().clone().requires_grad_(True) for k in ("verts", "verts_color", "faces_opacity")}
        fi = st["faces_intense"][views].detach().clone().requires_grad_(True)
        color, depth = r(views, pm, pw, ph, p["verts"], st["faces"], p["verts_color"], p["faces_opacity"], fi,
                         st["background"], tau)
        ((color * wc).sum() + (depth * wd).sum()).backward()
        res.append((color.detach(), depth.detach(), {**{k: v.grad for k, v in p.items()}, "faces_intense": fi.grad}))
    (c0, d0, g0), (c1, d1, g1) = res
    assert torch.equal(c0, c1) and torch.equal(d0, d1), "fused forward is not bit-identical to the unfused path"
    for k in g0:
        e = (g0[k] - g1[k]).abs().max().item() / (g1[k].abs().max().item() + 1e-30)
        assert e <= BWD_RTOL, (k, e)


# ------------------------------------------------------------------------------------------------
# multi-GPU gradient exchange (csrc/dm2_peer.cu): needs >= 2 GPUs on the box, skipped otherwise
def test_peer_allreduce_degenerate_arguments(built):
    import ctypes as C
    from dmesh2_renderer_b200 import _lib
    _dev()
    lib = _lib.load_library()
    ptrs = (C.c_void_p * 1)(0)
    # world 1 / empty bucket: nothing to do; bad arguments are rejected before any launch
    assert lib.dm2_peer_allreduce_sum(ptrs, None, 1, 0, 1024, None) == 0
    assert lib.dm2_peer_allreduce_sum(ptrs, None, 2, 0, 0, None) == 0
    assert lib.dm2_peer_allreduce_sum(ptrs, None, 2, 2, 1024, None) != 0
    assert lib.dm2_peer_allreduce_sum(ptrs, None, 2, 0, 1023, None) != 0
    assert lib.dm2_peer_allreduce_sum(None, None, 2, 0, 1024, None) != 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_allreduce_matches_nccl_two_gpus(built):
    import json
    import subprocess
    import sys
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533",
           os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "check_peer_allreduce.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("small_p2p_fused", "tiny_p2p_fused", "cfg2_p2p_fused", "small_p2p_unfused", "cfg2_p2p_unfused"):
        assert res[k]["max_rel_err_vs_nccl"] < 1e-6 and res[k]["same_on_all_ranks"], (k, res[k])


def test_fused_workspace_hint_survives_a_much_larger_frame(built):
    """The fused route sizes its binning workspace / AA cache from the previous call's num_rendered of the same
    problem shape; a following call with the same shapes but far more instances must re-allocate, not overflow."""
    from dmesh2_renderer_b200 import Renderer, scenes
    dev = _dev()
    res = {}
    for fused in (True, False):
        out = []
        for sigma in (0.01, 0.25, 0.01):            # same (B, P, F, H, W); num_rendered differs by > 10x
            sc = scenes.small(n_faces=500, size=96, sigma=sigma, seed=5)
            st = scene_tensors(sc, dev)
            r = Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
            r.fused = fused
            v = st["verts"].detach().clone().requires_grad_(True)
            pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
            color, depth = r([0], pm, sc.width, sc.height, v, st["faces"], st["verts_color"], st["faces_opacity"],
                             st["faces_intense"][:1], st["background"])
            (color.sum() + depth.sum()).backward()
            out.append((color.detach(), depth.detach(), v.grad.detach()))
        res[fused] = out
    for (c0, d0, g0), (c1, d1, g1) in zip(res[True], res[False]):
        assert torch.equal(c0, c1) and torch.equal(d0, d1)
        assert (g0 - g1).abs().max().item() <= BWD_RTOL * (g1.abs().max().item() + 1e-30)


# ------------------------------------------------------------------------------------------------
# BASELINE.json's headline size (cfg2: 1 M triangles at 1024^2), where the oracle is too slow to be the
# checker: size-independent properties of the path
@pytest.fixture(scope="module")
def cfg2_state(built):
    from dmesh2_renderer_b200 import scenes
    dev = _dev()
    sc = scenes.cfg2()
    r, st, args = render_args(sc, dev)
    return sc, r, st, args


def test_full_size_binning_properties(cfg2_state):
    """sortedness, the range table as a partition of the sorted list, instance counts per face = tiles_touched,
    a checksum of the emitted pairs - on all 2.26 M instances of the headline frame."""
    from dmesh2_renderer_b200 import _C
    sc, r, st, args = cfg2_state
    with torch.no_grad():
        R, color, depth, _, _, _, _, fb, bb, ib = _C.render_forward_cuda(*args)
    F, W, H = args[5].shape[0], sc.width, sc.height
    ft, bt, it = _C.face_taps(fb, 1, F), _C.binning_taps(bb, R), _C.image_taps(ib, 1, H, W)
    tt = ft["tiles_touched"].to(torch.int64)
    assert int(tt.sum()) == R and R > 2_000_000
    keys, vals = bt["keys"].to(torch.int64), bt["vals"].to(torch.int64)        # 45 significant bits: no sign issue
    assert bool((keys[1:] >= keys[:-1]).all()), "sorted keys are not non-decreasing"
    same = keys[1:] == keys[:-1]
    assert bool((vals[1:][same] > vals[:-1][same]).all()), "equal (tile, depth) keys must be in ascending face order"
    # every face appears exactly tiles_touched times; its key carries its own depth bits
    assert torch.equal(torch.bincount(vals, minlength=F), tt)
    dbits = ft["depths"].view(torch.int32).to(torch.int64)
    assert torch.equal(keys & 0xFFFFFFFF, dbits[vals])
    # the unsorted buffer holds the same multiset of pairs (checksums of keys and of key * (face + 1))
    ku, vu = bt["keys_unsorted"].to(torch.int64), bt["vals_unsorted"].to(torch.int64)
    assert int(ku.sum()) == int(keys.sum()) and int((ku % 1000003 * (vu + 1)).sum()) == int((keys % 1000003 * (vals + 1)).sum())
    # ranges: a partition of [0, R) in tile order, entries of range t carry tile id t
    rg = it["ranges"].to(torch.int64).view(-1, 2)
    tiles = (W // 16) * (H // 16)
    assert rg.shape[0] == tiles
    lens = rg[:, 1] - rg[:, 0]
    nz = lens > 0
    assert int(lens.sum()) == R and bool((lens >= 0).all())
    starts = torch.cumsum(lens, 0) - lens
    assert torch.equal(rg[nz, 0], starts[nz])
    tile_of = keys >> 32
    assert torch.equal(torch.bincount(tile_of, minlength=tiles), lens)
    # n_contrib never points past the tile's list
    nc = it["n_contrib"].to(torch.int64).view(H // 16, 16, W // 16, 16).permute(0, 2, 1, 3).reshape(tiles, 256)
    assert bool((nc.max(dim=1).values <= lens).all())


def test_full_size_forward_properties(cfg2_state):
    """run-to-run determinism, background linearity (colour = C + T * background), value ranges, fused route =
    reference-shaped route bit for bit - at the headline size."""
    from dmesh2_renderer_b200 import _C
    sc, r, st, args = cfg2_state
    with torch.no_grad():
        a1 = list(args)
        res1 = _C.render_forward_cuda(*a1)
        res2 = _C.render_forward_cuda(*a1)
        assert torch.equal(res1[1], res2[1]) and torch.equal(res1[2], res2[2]), "forward is not deterministic"
        a0 = list(args)
        a0[0] = torch.zeros_like(args[0])                       # black background
        res0 = _C.render_forward_cuda(*a0)
        T = _C.image_taps(res1[9], 1, sc.height, sc.width)["final_T"].view(1, sc.height, sc.width)
        diff = res1[1] - res0[1]                                 # = T * (1, 1, 1)
        assert bool((T >= 0).all()) and bool((T <= 1).all())
        assert float((diff - T.unsqueeze(-1)).abs().max()) <= 1e-6
        assert torch.equal(res1[2], res0[2]), "depth must not depend on the background"
        assert bool(torch.isfinite(res1[1]).all()) and bool(torch.isfinite(res1[2]).all())
        # public fused route (depth already mapped to [0, 1]) against the reference-shaped route
        pm = torch.zeros((1, 2), dtype=torch.int32, device=args[4].device)
        color_f, depth_f = r([0], pm, sc.width, sc.height, st["verts"], st["faces"], st["verts_color"],
                             st["faces_opacity"], st["faces_intense"][:1], st["background"])
        assert torch.equal(color_f, res1[1])
        assert torch.equal(depth_f, 1.0 - (res1[2] + 1.0) / 2.0)


def test_full_size_backward_is_linear_in_the_output_gradient(cfg2_state):
    """gradients(2 * dL/dout) = 2 * gradients(dL/dout) and gradients(dL1 + dL2) = gradients(dL1) + gradients(dL2)
    within the backward tolerance (atomic-order nondeterminism) - all six outputs, headline size."""
    from dmesh2_renderer_b200 import RenderFunction
    sc, r, st, args = cfg2_state
    dev = args[4].device
    leaf_idx = [4, 6, 7, 8, 10, 12]
    g = torch.Generator(device="cpu").manual_seed(11)
    w = [((torch.rand((1, sc.height, sc.width, 3), generator=g) * 2 - 1).to(dev),
          (torch.rand((1, sc.height, sc.width), generator=g) * 2 - 1).to(dev)) for _ in range(2)]

    def grads(wc, wd):
        a = list(args)
        for i in leaf_idx:
            a[i] = args[i].detach().clone().requires_grad_(True)
        color, depth = RenderFunction.apply(*a)
        ((color * wc).sum() + (depth * wd).sum()).backward()
        return [a[i].grad for i in leaf_idx]

    g1, g2 = grads(*w[0]), grads(*w[1])
    g12 = grads(w[0][0] + w[1][0], w[0][1] + w[1][1])
    g1x2 = grads(2 * w[0][0], 2 * w[0][1])
    for a_, b_, ab, a2, name in zip(g1, g2, g12, g1x2, GRAD_NAMES):
        scale = float(torch.maximum(a_.abs().max(), b_.abs().max())) + 1e-30
        assert float((a2 - 2 * a_).abs().max()) <= BWD_RTOL * 2 * scale, name
        assert float((ab - (a_ + b_)).abs().max()) <= BWD_RTOL * 2 * scale, name


def test_full_size_layers_properties(built):
    """cfg3 of BASELINE.json (Delaunay mesh of 37 300 points, ~500 k faces, 1024^2, 8 layers): every recorded id
    exists, no id repeats along a ray, entries beyond cnt are -1, hit distances increase along each ray,
    two runs agree, the incremental walk-only path equals a full generate, and - with every face existing - the
    first layer is the first hit of the ray."""
    from dmesh2_renderer_b200 import LayeredRenderer, scenes
    dev = _dev()
    sc = scenes.cfg3()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    lr = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    geo = [t(getattr(sc, k)) for k in ("verts", "faces", "tets", "face_tets", "tet_faces")]
    ex = t(sc.face_existence)
    layers, cnt = lr.generate([0], *geo, ex, sc.num_layers)
    again = lr.generate([0], *geo, ex, sc.num_layers)                  # walk-only (incremental) path
    full = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    full.incremental = False
    ref = full.generate([0], *geo, ex, sc.num_layers)
    assert torch.equal(layers, again[0]) and torch.equal(cnt, again[1])
    assert torch.equal(layers, ref[0]) and torch.equal(cnt, ref[1])
    L = sc.num_layers
    k = torch.arange(L, device=dev).view(1, 1, 1, L)
    valid = k < cnt.clamp(max=L).unsqueeze(-1)
    assert bool((layers[~valid] == -1).all()) and bool((layers[valid] >= 0).all())
    assert int(cnt.max()) == L and float((cnt > 0).float().mean()) > 0.5
    assert bool((ex[layers[valid].long()] != 0).all())
    srt = torch.sort(torch.where(valid, layers, -k - 1), dim=-1).values
    assert bool((srt[..., 1:] != srt[..., :-1]).all()), "a face id repeats along a ray"
    # hit distance of every recorded face along its ray (plain fp32 Moller-Trumbore in torch): non-decreasing
    ids = layers.clamp(min=0).long()
    tri = geo[0][geo[1].long()[ids]]                                    # [1, H, W, L, 3, 3]
    o, d = lr.ray_o[[0]].unsqueeze(-2), lr.ray_d[[0]].unsqueeze(-2)     # [1, H, W, 1, 3]
    e1, e2 = tri[..., 1, :] - tri[..., 0, :], tri[..., 2, :] - tri[..., 0, :]
    pv = torch.cross(d.expand_as(e2), e2, dim=-1)
    den = (pv * e1).sum(-1)
    tt = (torch.cross((o - tri[..., 0, :]), e1, dim=-1) * e2).sum(-1) / den
    ok = valid[..., 1:] & valid[..., :-1]
    assert bool((tt[..., 1:][ok] >= tt[..., :-1][ok] - 1e-4).all()), "layers are not in depth order along the ray"
    assert bool((tt[valid] >= -1e-4).all())
    # all faces exist: layer 0 is the ray's first hit and the walk fills all 8 layers wherever it enters the mesh
    all_layers, all_cnt = full.generate([0], *geo, torch.ones_like(ex), sc.num_layers)
    hit = all_cnt > 0
    assert bool((all_cnt[hit] >= 1).all()) and float(hit.float().mean()) > 0.5
    first = all_layers[..., 0]
    sub = layers[..., 0]
    both = hit & (cnt > 0)
    # the first EXISTING face along a ray can only be at or behind the first face of the full walk
    t_first = tt[..., 0]
    tri0 = geo[0][geo[1].long()[first.clamp(min=0).long()]]
    e1f, e2f = tri0[..., 1, :] - tri0[..., 0, :], tri0[..., 2, :] - tri0[..., 0, :]
    of, df = lr.ray_o[[0]], lr.ray_d[[0]]
    pvf = torch.cross(df, e2f, dim=-1)
    t0 = (torch.cross(of - tri0[..., 0, :], e1f, dim=-1) * e2f).sum(-1) / (pvf * e1f).sum(-1)
    assert bool((t_first[both] >= t0[both] - 1e-4).all())
    assert bool((sub[both & (ex[first.clamp(min=0).long()] != 0)] == first[both & (ex[first.clamp(min=0).long()] != 0)]).all())


def test_runs_on_the_callers_stream(built):
    """Everything is enqueued on torch's CURRENT stream (the reference uses the legacy default stream with device
    syncs): a frame rendered on a side stream while the default stream is kept busy must equal the
    frame rendered on the default stream, forward bit for bit."""
    from dmesh2_renderer_b200 import LayeredRenderer, Renderer, scenes
    dev = _dev()
    sc = scenes.small(n_faces=3000, size=128, sigma=0.1, seed=17)
    st = scene_tensors(sc, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)

    def frame():
        r = Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
        v = st["verts"].detach().clone().requires_grad_(True)
        color, depth = r([0], pm, sc.width, sc.height, v, st["faces"], st["verts_color"], st["faces_opacity"],
                         st["faces_intense"][:1], st["background"])
        ((color * st["w_color"][:1]).sum() + (depth * st["w_depth"][:1]).sum()).backward()
        return color.detach(), depth.detach(), v.grad.detach()

    base = frame()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)
    busy = torch.randn(4096, 4096, device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    for _ in range(20):                                   # ~tens of ms of work queued on the default stream
        busy = busy @ busy * 1e-4
    with torch.cuda.stream(side):
        other = frame()
    side.synchronize()
    torch.cuda.synchronize()
    assert torch.equal(base[0], other[0]) and torch.equal(base[1], other[1])
    assert float((base[2] - other[2]).abs().max()) <= BWD_RTOL * float(base[2].abs().max())

    tets = scenes.small_tets(n_points=300, size=64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    geo = [t(getattr(tets, k)) for k in ("verts", "faces", "tets", "face_tets", "tet_faces", "face_existence")]
    lr = LayeredRenderer(t(tets.mv), t(tets.proj), tets.width, tets.height, dev)
    a = lr.generate([0], *geo, tets.num_layers)
    torch.cuda.synchronize()
    for _ in range(10):
        busy = busy @ busy * 1e-4
    with torch.cuda.stream(side):
        lr2 = LayeredRenderer(t(tets.mv), t(tets.proj), tets.width, tets.height, dev)
        b = lr2.generate([0], *geo, tets.num_layers)
    side.synchronize()
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_non_contiguous_and_int64_inputs(built):
    """The reference reads its inputs through stride-aware accessors and converts dtypes in Renderer.forward;
    strided views (a column slice of a wider tensor) and int64 faces must give the same frame and gradients."""
    from dmesh2_renderer_b200 import Renderer, scenes
    dev = _dev()
    sc = scenes.small(n_faces=800, size=96, sigma=0.12, seed=23)
    st = scene_tensors(sc, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    r = Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)

    def run(verts, colours, faces):
        color, depth = r([0], pm, sc.width, sc.height, verts, faces, colours, st["faces_opacity"],
                         st["faces_intense"][:1], st["background"])
        ((color * st["w_color"][:1]).sum() + (depth * st["w_depth"][:1]).sum()).backward()
        return color.detach(), depth.detach()

    v0 = st["verts"].detach().clone().requires_grad_(True)
    c0 = st["verts_color"].detach().clone().requires_grad_(True)
    base = run(v0, c0, st["faces"])
    wide_v = torch.cat((st["verts"], torch.zeros_like(st["verts"][:, :1])), dim=1).detach().clone().requires_grad_(True)
    wide_c = torch.cat((torch.ones_like(st["verts_color"][:, :2]), st["verts_color"]), dim=1).detach().clone().requires_grad_(True)
    v1, c1 = wide_v[:, :3], wide_c[:, 2:]
    assert not v1.is_contiguous() and not c1.is_contiguous()
    other = run(v1, c1, st["faces"].to(torch.int64))
    assert torch.equal(base[0], other[0]) and torch.equal(base[1], other[1])
    assert float((wide_v.grad[:, :3] - v0.grad).abs().max()) <= BWD_RTOL * float(v0.grad.abs().max())
    assert float((wide_c.grad[:, 2:] - c0.grad).abs().max()) <= BWD_RTOL * float(c0.grad.abs().max())
    assert float(wide_v.grad[:, 3].abs().max()) == 0.0 and float(wide_c.grad[:, :2].abs().max()) == 0.0


def test_double_backward_call_and_no_grad_forward(built):
    """backward through the same graph twice (retain_graph) accumulates exactly twice the gradient - the per-face
    gradient records are re-zeroed per call and the AA cache is read-only in the backward pass; a forward under
    no_grad (no AA cache is allocated) returns the same frame."""
    from dmesh2_renderer_b200 import Renderer, scenes
    dev = _dev()
    sc = scenes.small(n_faces=1500, size=96, sigma=0.1, seed=29)
    st = scene_tensors(sc, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    r = Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    p = {k: st[k].detach().clone().requires_grad_(True) for k in ("verts", "verts_color", "faces_opacity")}
    fi = st["faces_intense"][:1].detach().clone().requires_grad_(True)
    color, depth = r([0], pm, sc.width, sc.height, p["verts"], st["faces"], p["verts_color"], p["faces_opacity"], fi,
                     st["background"])
    loss = (color * st["w_color"][:1]).sum() + (depth * st["w_depth"][:1]).sum()
    loss.backward(retain_graph=True)
    once = {k: v.grad.detach().clone() for k, v in p.items()}
    once["fi"] = fi.grad.detach().clone()
    loss.backward()
    twice = {k: v.grad.detach() for k, v in p.items()}
    twice["fi"] = fi.grad.detach()
    for k in once:
        scale = float(once[k].abs().max()) + 1e-30
        assert float((twice[k] - 2 * once[k]).abs().max()) <= BWD_RTOL * 2 * scale, k
    with torch.no_grad():
        c2, d2 = r([0], pm, sc.width, sc.height, p["verts"], st["faces"], p["verts_color"], p["faces_opacity"], fi,
                   st["background"])
    assert torch.equal(color.detach(), c2) and torch.equal(depth.detach(), d2)


# ------------------------------------------------------------------------------------------------
# round 2: pipelined forward (no host wait before the raster kernels), input flags, gradient sink, layer cache
def _public_step(r, st, sc, dev, views=(0,), sink=None):
    leaves = {k: st[k].detach().clone().requires_grad_(True) for k in ("verts", "verts_color", "faces_opacity")}
    fi = st["faces_intense"][list(views)].detach().clone().requires_grad_(True)
    pm = torch.zeros((len(views), 2), dtype=torch.int32, device=dev)
    if sink is not None:
        r.grad_sink = sink
    color, depth = r(list(views), pm, sc.width, sc.height, leaves["verts"], st["faces"], leaves["verts_color"],
                     leaves["faces_opacity"], fi, st["background"])
    ((color * st["w_color"][list(views)]).sum() + (depth * st["w_depth"][list(views)]).sum()).backward()
    return color.detach(), depth.detach(), leaves, fi


def test_pipelined_forward_equals_first_call(built):
    """Call 1 of a problem shape goes preprocess -> host wait -> forward; call 2 is enqueued in one piece on a
    workspace sized from call 1 (dm2_fused_forward_pipelined).  Same image bit for bit, same gradients."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    from dmesh2_renderer_b200.fused import FusedRenderFunction
    dev = _dev()
    sc = scenes.small(n_faces=3000, size=128, sigma=0.06, seed=41)
    st = scene_tensors(sc, dev)
    r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    FusedRenderFunction._size_hint.clear()
    c1, d1, l1, f1 = _public_step(r, st, sc, dev)
    hint = next(iter(FusedRenderFunction._size_hint.values()))
    assert hint.num_rendered > 0 and hint.longest > 0
    torch.cuda.synchronize()
    c2, d2, l2, f2 = _public_step(r, st, sc, dev)
    hint.poll()
    assert hint.cache_granules > 0, "the shading-cache demand of the previous call must have been read back"
    c3, d3, l3, f3 = _public_step(r, st, sc, dev)                       # pool now sized from the measured demand
    for c, d in ((c2, d2), (c3, d3)):
        assert torch.equal(c1, c) and torch.equal(d1, d)
    for k in l1:
        for other in (l2, l3):
            assert rel_err(other[k].grad.cpu().numpy(), l1[k].grad.cpu().numpy()) <= 1e-5, k
    assert rel_err(f3.grad.cpu().numpy(), f1.grad.cpu().numpy()) <= 1e-5


def test_pipelined_forward_repairs_a_longer_segment_than_hinted(built):
    """The hinted longest tile segment selects the sort tiers that are enqueued speculatively; when the frame
    has a longer segment the missing tiers run afterwards and classification + compositing are redone."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    from dmesh2_renderer_b200.fused import FusedRenderFunction
    dev = _dev()
    sc = scenes.triangle_soup("long_lists", 45_000, 32, 0.01, seed=31)            # segments of > 8192 instances
    st = scene_tensors(sc, dev)
    r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    FusedRenderFunction._size_hint.clear()
    c1, d1, l1, _ = _public_step(r, st, sc, dev)
    hint = next(iter(FusedRenderFunction._size_hint.values()))
    assert hint.longest > 8192
    hint.longest = 100                                   # pretend the previous frame had short lists only
    c2, d2, l2, _ = _public_step(r, st, sc, dev)
    assert hint.longest > 8192
    assert torch.equal(c1, c2) and torch.equal(d1, d2)
    assert rel_err(l2["verts"].grad.cpu().numpy(), l1["verts"].grad.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize("round_len,pool_bytes,opacity_scale", [(256, None, 1.0), (512, None, 1.0), (1024, 1 << 20, 1.0),
                                                                 (256, None, 0.02), (1024, None, 0.02)])
def test_forward_in_rounds_equals_single_pass(built, round_len, pool_bytes, opacity_scale):
    """Deep scenes composite in rounds of `round_len` list entries per tile with classification only for the tiles
    that are still running (dm2_set_forward_rounds): images bit-identical to the single pass, same gradients -
    synchronous first call, pipelined second call, repaired call (hint shorter than the rounds needed), and with
    a shading cache too small for the frame (blocks that do not fit fall back to recomputation).  In rounds the tile
    segments are sorted lazily (the first two rounds' worth at once, the rest of a tile's list before the round that
    reaches it): opacity_scale 0.02 keeps every pixel running to the end of its list, so every long segment goes
    through the second phase."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes, _lib
    from dmesh2_renderer_b200 import fused as _C
    from dmesh2_renderer_b200.fused import FusedRenderFunction
    dev = _dev()
    sc = scenes.triangle_soup("deep", 80_000, 64, 0.02, seed=77)
    st = scene_tensors(sc, dev)
    st["faces_opacity"] = st["faces_opacity"] * opacity_scale
    r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    lib = _lib.load_library()
    real_alloc = _C.shade_cache_alloc
    try:
        if pool_bytes is not None:
            def tiny(*a, **k):
                t = real_alloc(*a, **k)
                return t[: pool_bytes] if t is not None and t.numel() > pool_bytes else t
            _C.shade_cache_alloc = tiny
        lib.dm2_set_forward_rounds(1024, 1 << 40)            # never in rounds
        FusedRenderFunction._size_hint.clear()
        c0, d0, l0, f0 = _public_step(r, st, sc, dev)
        hint = next(iter(FusedRenderFunction._size_hint.values()))
        assert hint.longest > 4096, hint.longest             # (also above the default threshold)
        lib.dm2_set_forward_rounds(round_len, 0)             # always in rounds
        FusedRenderFunction._size_hint.clear()
        runs = [_public_step(r, st, sc, dev), _public_step(r, st, sc, dev)]
        hint = next(iter(FusedRenderFunction._size_hint.values()))
        hint.longest = round_len                             # the speculative launch covers one round only -> repair
        runs.append(_public_step(r, st, sc, dev))
        for c, d, l, f in runs:
            assert torch.equal(c0, c) and torch.equal(d0, d)
            for k in l0:
                assert rel_err(l[k].grad.cpu().numpy(), l0[k].grad.cpu().numpy()) <= 1e-5, k
            assert rel_err(f.grad.cpu().numpy(), f0.grad.cpu().numpy()) <= 1e-5
    finally:
        _C.shade_cache_alloc = real_alloc
        lib.dm2_set_forward_rounds(512, 6144)
        FusedRenderFunction._size_hint.clear()


def test_input_flags_vertex_ids_and_negative_patch(built):
    """Out-of-range vertex ids raise (torch indexing raises for them in the reference's host prep; raw pointers
    would read out of bounds) - on the first (synchronous) call and on a pipelined one; a negative patch_min is
    rejected by the fused route and handled by the reference-shaped route as the reference does (wrap-around)."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    from dmesh2_renderer_b200.fused import FusedRenderFunction
    dev = _dev()
    sc = scenes.small(n_faces=200, size=64)
    st = scene_tensors(sc, dev)
    r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    rest = (st["verts_color"], st["faces_opacity"], st["faces_intense"][:1], st["background"])
    bad = st["faces"].clone()
    bad[17, 1] = st["verts"].shape[0] + 5
    FusedRenderFunction._size_hint.clear()
    with pytest.raises(RuntimeError, match="vertex id"):
        r([0], pm, 64, 64, st["verts"], bad, *rest)
    r([0], pm, 64, 64, st["verts"], st["faces"], *rest)                 # establishes the hint -> pipelined next
    with pytest.raises(RuntimeError, match="vertex id"):
        r([0], pm, 64, 64, st["verts"], bad, *rest)
    color, _ = r([0], pm, 64, 64, st["verts"], st["faces"], *rest)      # and the renderer still works afterwards
    assert bool(torch.isfinite(color).all())
    neg = torch.tensor([[-3, 0]], dtype=torch.int32, device=dev)
    with pytest.raises(RuntimeError, match="negative patch_min"):
        r([0], neg, 32, 32, st["verts"], st["faces"], *rest)
    # the reference-shaped entry point checks too (its torch prep would already have raised in the reference)
    _r, _st, args = render_args(sc, dev)
    args[5] = bad
    with pytest.raises(RuntimeError, match="vertex id"):
        dm.RenderFunction.apply(*args)


def test_grad_sink_accumulates_in_place(built):
    """Gradients of the replicated parameters land directly in caller-provided tensors (the views of a gradient
    bucket) and accumulate over calls; autograd gets None for them."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    from dmesh2_renderer_b200.fused import GradSink
    from dmesh2_renderer_b200.multiview import GradBucket
    dev = _dev()
    sc = scenes.small(n_faces=900, size=96, sigma=0.1, seed=13, n_views=2)
    st = scene_tensors(sc, dev)
    r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    _, _, la, fa = _public_step(r, st, sc, dev, views=(0,))
    _, _, lb, fb = _public_step(r, st, sc, dev, views=(1,))
    bucket = GradBucket(st["verts"].shape[0], st["faces"].shape[0], dev)
    sink = GradSink(bucket.dverts, bucket.dverts_color, bucket.dfaces_opacity)
    r2 = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    _, _, l0, f0 = _public_step(r2, st, sc, dev, views=(0,), sink=sink)
    _, _, l1, f1 = _public_step(r2, st, sc, dev, views=(1,), sink=sink)
    assert all(l0[k].grad is None and l1[k].grad is None for k in l0)
    assert rel_err(f0.grad.cpu().numpy(), fa.grad.cpu().numpy()) <= 1e-5
    for name, view in (("verts", bucket.dverts), ("verts_color", bucket.dverts_color), ("faces_opacity", bucket.dfaces_opacity)):
        want = (la[name].grad + lb[name].grad).cpu().numpy()
        assert rel_err(view.cpu().numpy(), want) <= 1e-5, name


def test_layer_cache_is_not_fooled_by_recycled_temporaries(built):
    """geometry passed as a fresh temporary each call (same shape, freed after the call - the caching
    allocator hands the next temporary the same address at version 0) must never hit the incremental cache."""
    from dmesh2_renderer_b200 import LayeredRenderer, scenes
    dev = _dev()
    sc = scenes.small_tets(n_points=400, size=64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    lr = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    rest = [t(getattr(sc, k)) for k in ("faces", "tets", "face_tets", "tet_faces", "face_existence")]
    base = t(sc.verts)
    outs = []
    for scale in (1.0, 0.8, 1.0):
        layers, cnt = lr.generate([0], base * scale, *rest, sc.num_layers)       # temporary, freed right after
        outs.append((layers.clone(), cnt.clone()))
    full = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    full.incremental = False
    for scale, (layers, cnt) in zip((1.0, 0.8, 1.0), outs):
        want = full.generate([0], (base * scale).clone(), *rest, sc.num_layers)
        assert torch.equal(layers, want[0]) and torch.equal(cnt, want[1]), scale
    assert not torch.equal(outs[0][0], outs[1][0])
    # same objects, same versions: the walk-only path is taken and agrees
    kept = base * 0.9
    a = lr.generate([0], kept, *rest, sc.num_layers)
    b = lr.generate([0], kept, *rest, sc.num_layers)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_camera_gradients_use_the_reference_shaped_route(built):
    """cameras that require grad get their gradient (through the torch prep, as in the reference);
    negative view ids are accepted like `self.mv[idx]` does."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    dev = _dev()
    sc = scenes.small(n_faces=300, size=48, n_views=3)
    st = scene_tensors(sc, dev)
    mv = st["mv"].clone().requires_grad_(True)
    r = dm.Renderer(mv, st["proj"], sc.width, sc.height, dev)
    pm = torch.zeros((1, 2), dtype=torch.int32, device=dev)
    rest = (st["verts"], st["faces"], st["verts_color"], st["faces_opacity"])
    color, depth = r([1], pm, 48, 48, *rest, st["faces_intense"][[1]], st["background"])
    (color.sum() + depth.sum()).backward()
    assert mv.grad is not None and float(mv.grad[1].abs().max()) > 0 and float(mv.grad[0].abs().max()) == 0
    r2 = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    a = r2([-1], pm, 48, 48, *rest, st["faces_intense"][[2]], st["background"])
    b = r2([2], pm, 48, 48, *rest, st["faces_intense"][[2]], st["background"])
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    with pytest.raises(IndexError):
        r2([3], pm, 48, 48, *rest, st["faces_intense"][[2]], st["background"])


# ------------------------------------------------------------------------------------------------
# in-kernel ray generation
@pytest.mark.parametrize("size,n_views", [(64, 3), (96, 2), (200, 1)])
def test_kernel_rays_are_bit_identical_to_the_torch_ray_tensors(built, size, n_views):
    """A ray recipe is only selected when it reproduces torch's ray_o / ray_d bit for bit (power-of-two and other
    frame sizes: torch's CUDA division by the frame size is a multiplication by its reciprocal); with it the
    kernels generate the rays themselves: same image, same layer ids, same gradients - also after the ray
    tensors have been released."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import _lib, scenes
    dev = _dev()
    sc = scenes.small(n_faces=500, size=size, sigma=0.12, seed=9, n_views=n_views)
    st = scene_tensors(sc, dev)
    r = dm.Renderer(st["mv"], st["proj"], size, size, dev)
    assert r._ray_recipe == 0, "in-kernel rays are opt-in"
    assert r.enable_kernel_rays() and 1 <= r._ray_recipe <= _lib.RAY_RECIPES, "no ray recipe reproduces torch's rays"
    off = dm.Renderer(st["mv"], st["proj"], size, size, dev)      # the renderer that reads the tensors
    views = list(range(n_views))
    a = _public_step(r, st, sc, dev, views=views)
    b = _public_step(off, st, sc, dev, views=views)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[2]:
        assert rel_err(a[2][k].grad.cpu().numpy(), b[2][k].grad.cpu().numpy()) <= 1e-5, k
    # patches (rays addressed through patch_min) and released tensors
    pm = torch.tensor([[5, 9]] * n_views, dtype=torch.int32, device=dev)
    args = (st["verts"], st["faces"], st["verts_color"], st["faces_opacity"], st["faces_intense"][views], st["background"])
    p0 = off(views, pm, 40, 33, *args)
    r.release_ray_buffers()
    assert r.ray_o is None
    p1 = r(views, pm, 40, 33, *args)
    assert torch.equal(p0[0], p1[0]) and torch.equal(p0[1], p1[1])
    # a renderer whose rays no recipe reproduces keeps its tensors
    off.ray_d = off.ray_d * 1.0000001
    with pytest.raises(RuntimeError):
        off.release_ray_buffers()
    assert off.ray_o is not None


def test_kernel_rays_layered(built):
    from dmesh2_renderer_b200 import LayeredRenderer, scenes
    dev = _dev()
    sc = scenes.small_tets(n_points=400, size=80)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    a = [t(getattr(sc, k)) for k in ("verts", "faces", "tets", "face_tets", "tet_faces", "face_existence")]
    lr = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    assert lr.enable_kernel_rays()
    off = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    x, y = lr.generate([0], *a, sc.num_layers), off.generate([0], *a, sc.num_layers)
    assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])
    lr.release_ray_buffers()
    lr.incremental = False
    z = lr.generate([0], *a, sc.num_layers)
    assert torch.equal(z[0], y[0]) and torch.equal(z[1], y[1])


# ------------------------------------------------------------------------------------------------
# layer consumer: fused gather + front-to-back composite of the generated layers
@pytest.mark.parametrize("with_intense", [False, True])
def test_composite_layers_matches_torch(built, with_intense):
    """Forward within 1e-5 and all gradients within 1e-4 (atomic order) of the plain-torch statement of the op,
    on real layers of a tet scene (ragged counts, -1 padding, two views)."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    dev = _dev()
    sc = scenes.delaunay_tets("comp", 1500, 96, seed=19, p_exist=0.3, num_layers=6,
                              eyes=[tuple(e) for e in scenes.fibonacci_eyes(64)[:2]])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    lr = dm.LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    layers, cnt = lr.generate([0, 1], t(sc.verts), t(sc.faces), t(sc.tets), t(sc.face_tets), t(sc.tet_faces),
                              t(sc.face_existence), sc.num_layers)
    assert int(cnt.max()) == sc.num_layers and int(cnt.min()) == 0
    F = sc.faces.shape[0]
    g = torch.Generator(device="cpu").manual_seed(2)
    mk = lambda *s_: torch.rand(*s_, generator=g).to(dev)
    base = dict(color=mk(F, 3), opacity=mk(F) * 0.9 + 0.05, intense=mk(2, F) * 0.5 + 0.5)
    bg = torch.tensor([0.2, 0.5, 0.9], device=dev)
    wc, wa = mk(2, sc.height, sc.width, 3) - 0.5, mk(2, sc.height, sc.width) - 0.5
    res = []
    for fn in (dm.composite_layers, dm.composite_layers_torch):
        p = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        color, alpha = fn(layers, cnt, p["color"], p["opacity"], bg, p["intense"] if with_intense else None)
        ((color * wc).sum() + (alpha * wa).sum()).backward()
        res.append((color.detach(), alpha.detach(), p))
    (c0, a0, p0), (c1, a1, p1) = res
    assert float((c0 - c1).abs().max()) <= 1e-5 and float((a0 - a1).abs().max()) <= 1e-5
    for k in ("color", "opacity") + (("intense",) if with_intense else ()):
        e = float((p0[k].grad - p1[k].grad).abs().max()) / (float(p1[k].grad.abs().max()) + 1e-30)
        assert e <= 1e-4, (k, e)
    assert with_intense or p0["intense"].grad is None


def test_composite_layers_edge_cases(built):
    import dmesh2_renderer_b200 as dm
    dev = _dev()
    bg = torch.tensor([1.0, 0.0, 0.5], device=dev)
    col = torch.rand(4, 3, device=dev)
    op = torch.tensor([0.5, 0.25, 0.0, 0.9], device=dev)
    layers = torch.tensor([[[[0, 1, -1], [-1, -1, -1], [3, 3, 3]]]], dtype=torch.int32, device=dev)      # (1,1,3,3)
    cnt = torch.tensor([[[2, 0, 5]]], dtype=torch.int32, device=dev)                                       # cnt > L is clamped
    color, alpha = dm.composite_layers(layers, cnt, col, op, bg)
    want_c, want_a = dm.composite_layers_torch(layers, cnt, col, op, bg)
    assert torch.allclose(color, want_c, atol=1e-6) and torch.allclose(alpha, want_a, atol=1e-6)
    assert torch.allclose(color[0, 0, 1], bg) and float(alpha[0, 0, 1]) == 0.0
    empty = dm.composite_layers(layers[:, :, :0], cnt[:, :, :0], col, op, bg)
    assert empty[0].shape == (1, 1, 0, 3)
    with pytest.raises(RuntimeError):
        dm.composite_layers(layers, cnt, col[:3], op, bg)
    with pytest.raises(RuntimeError):
        dm.composite_layers(layers.cpu(), cnt.cpu(), col.cpu(), op.cpu(), bg.cpu())


@pytest.mark.parametrize("n_views,n_faces", [(1, 700), (8, 700), (1, 1_000_000), (8, 1_000_000)])
def test_fused_projection_is_bit_identical_to_torch_matmul(built, n_views, n_faces):
    """The fused preprocess kernel reproduces the two K = 4 matmuls of compute_verts_ndc_image as the fma chain
    cuBLAS runs for them (dm2_common.cuh: clip_vertex) - for B in {1, 8} and P in {2 k, 3 M} (different cuBLAS
    kernels: gemv / batched).  Checked on what the projection feeds: tiles_touched, the three depth keys, and the
    packed records (NDC z, screen-space AA vertices / edges / reciprocals / normals) must be bit-identical between
    the fused preprocess and the reference-shaped one, which consumes torch's tensors."""
    import ctypes as C
    from dmesh2_renderer_b200 import _C, _lib, scenes
    from dmesh2_renderer_b200.fused import _fused_struct
    dev = _dev()
    sc = scenes.cfg4(n_views=n_views, size=256, n_faces=n_faces)
    r, st, args = render_args(sc, dev)
    with torch.no_grad():
        res = _C.render_forward_cuda(*args)
    R, fb = res[0], res[7]
    B, F = n_views, n_faces
    lib = _lib.load_library()
    key = tuple(range(n_views))
    _k, ray_view, view_idx = r._view_ids(key)
    pm = torch.zeros((B, 2), dtype=torch.int32, device=dev)
    s, keep = _fused_struct(st["background"], pm, sc.width, sc.height, st["verts"], st["faces"], st["verts_color"],
                            st["faces_opacity"], r.mv.index_select(0, view_idx), r.proj.index_select(0, view_idx),
                            st["faces_intense"], 1.0, r.ray_o, r.ray_d, ray_view, False)
    fb2 = torch.empty_like(fb)
    ib2 = torch.empty((_C.image_layout(B, sc.height, sc.width).total,), dtype=torch.uint8, device=dev)
    nr, longest = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.dm2_fused_preprocess(C.byref(s), _C._ptr(fb2), _C._ptr(ib2), C.byref(nr), C.byref(longest),
                                        _C._stream(dev)))
    assert int(nr.value) == R
    a, b = _C.face_taps(fb, B, F), _C.face_taps(fb2, B, F)
    live = a["tiles_touched"] > 0
    assert torch.equal(a["tiles_touched"], b["tiles_touched"]) and int(live.sum()) > 0
    for k in ("depths", "min_depths", "max_depths"):
        assert torch.equal(a[k].view(torch.int32)[live], b[k].view(torch.int32)[live]), k
    L = _C.face_layout(B, F)
    ra = fb[L.records:L.records + B * F * L.record_stride].view(torch.int32).view(B * F, -1)[live]
    rb = fb2[L.records:L.records + B * F * L.record_stride].view(torch.int32).view(B * F, -1)[live]
    # words 0..58: bbox, normals, offsets, opacity / intensity, 3-D vertices + NDC z, colours, vertex ids, AA verts /
    # edges / reciprocals, face id; word 15 (flags) differs by design: the fused kernel records the CCW flip in bit 8
    same = (ra[:, :59] == rb[:, :59])
    same[:, 15] |= ((ra[:, 15] ^ rb[:, 15]) & ~0x100) == 0
    assert bool(same.all()), f"{int((~same).sum())} record words differ"


def test_layered_index_validation_debug_switch(built, monkeypatch):
    """DM2_VALIDATE_INDICES=1: out-of-range tet / face ids raise instead of being followed through raw pointers."""
    from dmesh2_renderer_b200 import LayeredRenderer, scenes
    dev = _dev()
    sc = scenes.small_tets(n_points=200, size=48)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    lr = LayeredRenderer(t(sc.mv), t(sc.proj), sc.width, sc.height, dev)
    a = {k: t(getattr(sc, k)) for k in ("verts", "faces", "tets", "face_tets", "tet_faces", "face_existence")}
    monkeypatch.setenv("DM2_VALIDATE_INDICES", "1")
    lr.generate([0], *a.values(), sc.num_layers)
    bad = dict(a)
    bad["tet_faces"] = a["tet_faces"].clone()
    bad["tet_faces"][3, 2] = a["faces"].shape[0] + 7
    with pytest.raises(RuntimeError, match="tet_faces"):
        lr.generate([0], *bad.values(), sc.num_layers)
    bad = dict(a)
    bad["face_tets"] = a["face_tets"].clone()
    bad["face_tets"][5, 0] = -2
    with pytest.raises(RuntimeError, match="face_tets"):
        lr.generate([0], *bad.values(), sc.num_layers)


@pytest.mark.parametrize("n_faces,size,sigma", [(3000, 128, 0.06), (40_000, 256, 0.02)])
def test_fast_overlap_area_agrees_with_the_exact_clip(built, n_faces, size, sigma):
    """dm2_set_exact_clip: the polygon-free overlap area (csrc/dm2_aa_fast.cuh, the default) against the reference's
    polygon construction for every pair - images equal to 2e-6, the same contributors, gradients equal to 1e-4 of
    their largest entry (the contract itself is checked pair by pair on the CPU, tests/test_aa_fast.py)."""
    import dmesh2_renderer_b200 as dm
    from dmesh2_renderer_b200 import scenes
    dev = _dev()
    sc = scenes.small(n_faces=n_faces, size=size, sigma=sigma, seed=43)
    st = scene_tensors(sc, dev)
    r = dm.Renderer(st["mv"], st["proj"], sc.width, sc.height, dev)
    prev = dm.set_exact_clip(True)
    try:
        c0, d0, l0, f0 = _public_step(r, st, sc, dev)
        torch.cuda.synchronize()
        assert dm.set_exact_clip(False) is True
        c1, d1, l1, f1 = _public_step(r, st, sc, dev)
        torch.cuda.synchronize()
    finally:
        dm.set_exact_clip(prev)
    assert float((c0 - c1).abs().max()) <= 2e-6 and float((d0 - d1).abs().max()) <= 2e-6
    assert float((c0 - c1).abs().max()) > 0.0 or n_faces < 10_000, "the two settings must run different code"
    for k in l0:
        assert rel_err(l1[k].grad.cpu().numpy(), l0[k].grad.cpu().numpy()) <= 1e-4, k
    assert rel_err(f1.grad.cpu().numpy(), f0.grad.cpu().numpy()) <= 1e-4
