"""Host-side mirror of the reference's pybind module ``dmesh2_renderer._C`` (reference ext.cpp:6-9).

Same three function names, same positional arguments, same tuple shapes in the results, same
error behaviour (RuntimeError with the reference's messages) - but the work is done by the
hand-written sm_100a kernels in libdm2_b200.so through the C-ABI of include/dm2_renderer.h.

Differences that are private to this implementation (allowed by SURVEY.md section 8b):
  * the reference's four AA cache tensors are ONE compact "shading cache" pool (16 bytes per touched and
    24 bytes per partially covered (pixel, face) pair of the frame) returned in the ``oarea`` slot of the
    reference's 10-tuple; the other three cache tensors come back empty.  ``len_oarea_buffer == 0``
    disables it (the backward pass then recomputes, as it does for blocks beyond the pool capacity);
  * the three scratch buffers are laid out per DM2FaceLayout / DM2ImageLayout / DM2BinningLayout.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import DM2BinningLayout, DM2FaceLayout, DM2ImageLayout, DM2LayerInputs, DM2RenderGrads, DM2RenderInputs


# Shading-cache sizing (include/dm2_renderer.h: 16 bytes per touched (pixel, face) pair + 24 bytes per
# partially covered one, handed from the forward to the backward pass).  The pool is sized from what the
# previous forward call of the same problem shape asked for (+ SHADE_CACHE_SLACK); a first call reserves
# SHADE_CACHE_FIRST_BYTES_PER_INSTANCE per sorted (face, tile) instance, at most SHADE_CACHE_FIRST_MAX_BYTES
# (deep scenes terminate early and stage a fraction of their instances: 10 M triangles @2048^2 need 100 bytes per
# instance, 1 M @1024^2 360).  Blocks that do not fit are simply not cached (the backward pass recomputes them),
# so a pool that is too small costs time, never results; a pool that overflowed under-reports its demand (warps
# stop asking once it is full), so the next one is twice as large.
SHADE_CACHE_FIRST_BYTES_PER_INSTANCE = 512
SHADE_CACHE_FIRST_MAX_BYTES = 2 << 30
SHADE_CACHE_SLACK = 1.08
SHADE_CACHE_MAX_BYTES = 48 << 30
SHADE_CACHE_HEADER = 256


def shade_cache_alloc(len_oarea_buffer, num_rendered, aa_temperature, device, want=True, used_granules=None,
                      prev_capacity_bytes=0):
    """Returns the pool tensor (uint8; empty when the cache is off: len_oarea_buffer == 0 - the reference's K -,
    aa_temperature == 0, nothing to render, or no gradient wanted).  used_granules / prev_capacity_bytes: demand
    counter and pool size of the previous forward call of this problem shape."""
    if not (want and len_oarea_buffer > 0 and aa_temperature > 0 and num_rendered > 0):
        return torch.empty((0,), dtype=torch.uint8, device=device)
    if used_granules is not None and used_granules > 0:
        nbytes = int(used_granules * 16 * SHADE_CACHE_SLACK) + 4096
        if prev_capacity_bytes and used_granules * 16 > prev_capacity_bytes - SHADE_CACHE_HEADER:
            nbytes = max(nbytes, 2 * int(prev_capacity_bytes))            # overflowed: the demand is a lower bound
    else:
        nbytes = min(int(num_rendered) * SHADE_CACHE_FIRST_BYTES_PER_INSTANCE, SHADE_CACHE_FIRST_MAX_BYTES)
    nbytes = min(nbytes, SHADE_CACHE_MAX_BYTES)
    nbytes = SHADE_CACHE_HEADER + (nbytes + 15) // 16 * 16
    return torch.empty((nbytes,), dtype=torch.uint8, device=device)


def shade_cache_used_granules(pool):
    """16-byte granules the forward pass asked for (device -> host read: synchronises; tests / sizing only)."""
    return int(pool[:4].view(torch.int32).item()) & 0xFFFFFFFF if pool.numel() >= 4 else 0


def _err(msg):
    raise RuntimeError(msg)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _prep(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        _err(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        _err(f"{name} must be a CUDA tensor (dmesh2_renderer_b200 has no CPU path)")
    if t.dtype != dtype:
        _err(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t.contiguous()


def _check_render_shapes(background, patch_min, verts, faces, verts_color, faces_opacity, verts_ndc, verts_image,
                         faces_intense, aa_temperature, aa_face_verts, aa_face_edges, aa_face_edges_iszero,
                         aa_face_edges_recip, aa_face_edges_normal, aa_face_edges_normal_c, len_oarea_buffer,
                         image_ray_o, image_ray_d):
    # reference render.cu:62-118 (and :247-303 for backward), same messages
    if background.dim() != 1 or background.size(0) != 3:
        _err("background must have dimensions (3,)")
    if patch_min.dim() != 2 or patch_min.size(1) != 2:
        _err("patch_min must have dimensions (B, 2)")
    if verts.dim() != 2 or verts.size(1) != 3:
        _err("verts must have dimensions (P, 3)")
    if faces.dim() != 2 or faces.size(1) != 3:
        _err("faces must have dimensions (F, 3)")
    if verts_color.dim() != 2 or verts_color.size(1) != 3:
        _err("vert color must have dimensions (P, 3)")
    if faces_opacity.dim() != 1 or faces_opacity.size(0) != faces.size(0):
        _err("face opacity must have dimensions (F,)")
    if verts_ndc.dim() != 3 or verts_ndc.size(2) != 3:
        _err("verts_ndc must have dimensions (B, P, 3)")
    if verts_image.dim() != 3 or verts_image.size(2) != 2:
        _err("verts_image must have dimensions (B, P, 2)")
    if faces_intense.dim() != 2 or faces_intense.size(1) != faces.size(0):
        _err("faces_intense must have dimensions (B, F,)")
    for name, t in (("aa_face_verts", aa_face_verts), ("aa_face_edges", aa_face_edges),
                    ("aa_face_edges_iszero", aa_face_edges_iszero), ("aa_face_edges_recip", aa_face_edges_recip),
                    ("aa_face_edges_normal", aa_face_edges_normal)):
        if t.dim() != 4 or t.size(2) != 3 or t.size(3) != 2:
            _err(f"{name} must have dimensions (B, F, 3, 2)")
    if aa_face_edges_normal_c.dim() != 3 or aa_face_edges_normal_c.size(2) != 3:
        _err("aa_face_edges_normal_c must have dimensions (B, F, 3)")
    if image_ray_o.dim() != 4 or image_ray_o.size(3) != 3:
        _err("image_ray_o must have dimensions (B, H, W, 3)")
    if image_ray_d.dim() != 4 or image_ray_d.size(3) != 3:
        _err("image_ray_d must have dimensions (B, H, W, 3)")
    if aa_temperature < 0 or aa_temperature > 1:
        _err("aa_temperature must be in the range [0, 1]")
    if len_oarea_buffer < 0:
        _err("len_oarea_buffer must be non-negative")


def _check_consistency(B, P, F, H, W, verts_color, patch_min, verts_ndc, verts_image, faces_intense, aa_list,
                       aa_face_edges_normal_c, image_ray_o, image_ray_d):
    # The reference indexes through stride-aware accessors and trusts these; raw pointers cannot.
    if verts_color.size(0) != P:
        _err("vert color must have dimensions (P, 3)")
    if patch_min.size(0) != B:
        _err("patch_min must have dimensions (B, 2)")
    if verts_ndc.size(1) != P or verts_image.size(0) != B or verts_image.size(1) != P:
        _err("verts_ndc / verts_image must have dimensions (B, P, .)")
    if faces_intense.size(0) != B:
        _err("faces_intense must have dimensions (B, F,)")
    for t in aa_list:
        if t.size(0) != B or t.size(1) != F:
            _err("aa tensors must have dimensions (B, F, 3, 2)")
    if aa_face_edges_normal_c.size(0) != B or aa_face_edges_normal_c.size(1) != F:
        _err("aa_face_edges_normal_c must have dimensions (B, F, 3)")
    for t in (image_ray_o, image_ray_d):
        if t.size(0) != B or t.size(1) != H or t.size(2) != W:
            _err("image rays must have dimensions (B, patch_height, patch_width, 3)")


def _render_inputs(background, patch_min, patch_width, patch_height, verts, faces, verts_color, faces_opacity,
                   verts_ndc, verts_image, faces_intense, aa_temperature, aa_face_verts, aa_face_edges,
                   aa_face_edges_iszero, aa_face_edges_recip, aa_face_edges_normal, aa_face_edges_normal_c,
                   len_oarea_buffer, image_ray_o, image_ray_d):
    _check_render_shapes(background, patch_min, verts, faces, verts_color, faces_opacity, verts_ndc, verts_image,
                         faces_intense, aa_temperature, aa_face_verts, aa_face_edges, aa_face_edges_iszero,
                         aa_face_edges_recip, aa_face_edges_normal, aa_face_edges_normal_c, len_oarea_buffer,
                         image_ray_o, image_ray_d)
    f32, i32 = torch.float32, torch.int32
    keep = dict(
        background=_prep(background, f32, "background"), patch_min=_prep(patch_min, i32, "patch_min"),
        verts=_prep(verts, f32, "verts"), faces=_prep(faces, i32, "faces"),
        verts_color=_prep(verts_color, f32, "verts_color"), faces_opacity=_prep(faces_opacity, f32, "faces_opacity"),
        verts_ndc=_prep(verts_ndc, f32, "verts_ndc"), verts_image=_prep(verts_image, f32, "verts_image"),
        faces_intense=_prep(faces_intense, f32, "faces_intense"),
        aa_face_verts=_prep(aa_face_verts, f32, "aa_face_verts"),
        aa_face_edges=_prep(aa_face_edges, f32, "aa_face_edges"),
        aa_face_edges_iszero=_prep(aa_face_edges_iszero, torch.bool, "aa_face_edges_iszero"),
        aa_face_edges_recip=_prep(aa_face_edges_recip, f32, "aa_face_edges_recip"),
        aa_face_edges_normal=_prep(aa_face_edges_normal, f32, "aa_face_edges_normal"),
        aa_face_edges_normal_c=_prep(aa_face_edges_normal_c, f32, "aa_face_edges_normal_c"),
        image_ray_o=_prep(image_ray_o, f32, "image_ray_o"), image_ray_d=_prep(image_ray_d, f32, "image_ray_d"),
    )
    B, P, F = int(verts_ndc.size(0)), int(verts.size(0)), int(faces.size(0))
    H, W = int(patch_height), int(patch_width)
    _check_consistency(B, P, F, H, W, keep["verts_color"], keep["patch_min"], keep["verts_ndc"], keep["verts_image"],
                       keep["faces_intense"],
                       [keep[k] for k in ("aa_face_verts", "aa_face_edges", "aa_face_edges_iszero",
                                          "aa_face_edges_recip", "aa_face_edges_normal")],
                       keep["aa_face_edges_normal_c"], keep["image_ray_o"], keep["image_ray_d"])
    s = DM2RenderInputs()
    s.B, s.P, s.F, s.H, s.W = B, P, F, H, W
    s.aa_temperature = float(aa_temperature)
    s.len_oarea_buffer = int(len_oarea_buffer)
    for k, t in keep.items():
        setattr(s, k, t.data_ptr() if t.numel() > 0 else 0)
    return s, keep


def face_layout(B, F, layered=False):
    L = DM2FaceLayout()
    _lib.check(_lib.load_library().dm2_face_layout(B, F, 1 if layered else 0, C.byref(L)))
    return L


def image_layout(B, H, W, layered=False):
    L = DM2ImageLayout()
    _lib.check(_lib.load_library().dm2_image_layout(B, H, W, 1 if layered else 0, C.byref(L)))
    return L


def binning_layout(R):
    L = DM2BinningLayout()
    _lib.check(_lib.load_library().dm2_binning_layout(int(R), C.byref(L)))
    return L


def render_forward_cuda(background, patch_min, patch_width, patch_height, verts, faces, verts_color, faces_opacity,
                        verts_ndc, verts_image, faces_intense, aa_temperature, aa_face_verts, aa_face_edges,
                        aa_face_edges_iszero, aa_face_edges_recip, aa_face_edges_normal, aa_face_edges_normal_c,
                        len_oarea_buffer, image_ray_o, image_ray_d, want_cache=True):
    """reference render.h:12-45 / render.cu:28-195.  Returns the reference's 10-tuple."""
    lib = _lib.load_library()
    s, keep = _render_inputs(background, patch_min, patch_width, patch_height, verts, faces, verts_color,
                             faces_opacity, verts_ndc, verts_image, faces_intense, aa_temperature, aa_face_verts,
                             aa_face_edges, aa_face_edges_iszero, aa_face_edges_recip, aa_face_edges_normal,
                             aa_face_edges_normal_c, len_oarea_buffer, image_ray_o, image_ray_d)
    dev = keep["verts"].device
    B, H, W = s.B, s.H, s.W
    with torch.cuda.device(dev):
        stream = _stream(dev)
        f_opts = dict(dtype=torch.float32, device=dev)
        out_color = torch.empty((B, H, W, 3), **f_opts)
        out_depth = torch.empty((B, H, W), **f_opts)
        face_buffer = torch.empty((face_layout(B, s.F).total,), dtype=torch.uint8, device=dev)
        image_buffer = torch.empty((image_layout(B, H, W).total,), dtype=torch.uint8, device=dev)
        num_rendered, longest = C.c_int64(0), C.c_int64(0)
        _lib.check(lib.dm2_render_preprocess(C.byref(s), _ptr(face_buffer), _ptr(image_buffer),
                                             C.byref(num_rendered), C.byref(longest), stream))
        R = int(num_rendered.value)
        binning_buffer = torch.empty((binning_layout(R).total,), dtype=torch.uint8, device=dev)
        oarea = shade_cache_alloc(int(len_oarea_buffer), R, float(aa_temperature), dev, want_cache)
        _lib.check(lib.dm2_render_forward(C.byref(s), _ptr(face_buffer), _ptr(binning_buffer), _ptr(image_buffer),
                                          R, int(longest.value), _ptr(out_color), _ptr(out_depth), _ptr(oarea),
                                          int(oarea.numel()), stream))
        # the other AA cache tensors are placeholders, kept so the tuple has the reference's arity
        tri_id = torch.empty((B, H, W, 0), dtype=torch.int32, device=dev)
        tri_cnt = torch.empty((0,), dtype=torch.int32, device=dev)
        doarea = torch.empty((B, H, W, 0, 3, 2), **f_opts)
    return R, out_color, out_depth, oarea, tri_id, tri_cnt, doarea, face_buffer, binning_buffer, image_buffer


def render_backward_cuda(num_rendered, background, patch_min, patch_width, patch_height, verts, faces, verts_color,
                         faces_opacity, verts_ndc, verts_image, faces_intense, aa_temperature, aa_face_verts,
                         aa_face_edges, aa_face_edges_iszero, aa_face_edges_recip, aa_face_edges_normal,
                         aa_face_edges_normal_c, len_oarea_buffer, image_ray_o, image_ray_d, dL_dout_color,
                         dL_dout_depth, face_buffer, binning_buffer, image_buffer, oarea_buffer_oarea=None,
                         oarea_buffer_tri_id=None, oarea_buffer_tri_cnt=None, oarea_buffer_doarea_dtri_verts=None):
    """reference render.h:47-94 / render.cu:198-373.  Returns the six gradient tensors."""
    lib = _lib.load_library()
    s, keep = _render_inputs(background, patch_min, patch_width, patch_height, verts, faces, verts_color,
                             faces_opacity, verts_ndc, verts_image, faces_intense, aa_temperature, aa_face_verts,
                             aa_face_edges, aa_face_edges_iszero, aa_face_edges_recip, aa_face_edges_normal,
                             aa_face_edges_normal_c, len_oarea_buffer, image_ray_o, image_ray_d)
    dev = keep["verts"].device
    gc = _prep(dL_dout_color, torch.float32, "dL_dout_color")
    gd = _prep(dL_dout_depth, torch.float32, "dL_dout_depth")
    if tuple(gc.shape) != (s.B, s.H, s.W, 3) or tuple(gd.shape) != (s.B, s.H, s.W):
        _err("dL_dout_color / dL_dout_depth must have dimensions (B, H, W, 3) / (B, H, W)")
    with torch.cuda.device(dev):
        dL_dverts = torch.zeros_like(keep["verts"])
        dL_dverts_color = torch.zeros_like(keep["verts_color"])
        dL_dfaces_opacity = torch.zeros_like(keep["faces_opacity"])
        dL_dverts_ndc = torch.zeros_like(keep["verts_ndc"])
        dL_dfaces_intense = torch.zeros_like(keep["faces_intense"])
        dL_daa_face_verts = torch.zeros_like(keep["aa_face_verts"])
        if s.F != 0:
            g = DM2RenderGrads(dL_dverts.data_ptr(), dL_dverts_color.data_ptr(), dL_dfaces_opacity.data_ptr(),
                               dL_dverts_ndc.data_ptr(), dL_dfaces_intense.data_ptr(), dL_daa_face_verts.data_ptr())
            nb = C.c_size_t(0)
            _lib.check(lib.dm2_backward_workspace_bytes(s.B, s.F, C.byref(nb)))
            grad_ws = torch.empty((nb.value,), dtype=torch.uint8, device=dev)
            pool = oarea_buffer_oarea
            nb_pool = int(pool.numel()) * pool.element_size() if pool is not None else 0
            _lib.check(lib.dm2_render_backward(C.byref(s), _ptr(face_buffer), _ptr(binning_buffer),
                                               _ptr(image_buffer), int(num_rendered), _ptr(gc), _ptr(gd),
                                               C.byref(g), _ptr(grad_ws),
                                               _ptr(pool) if nb_pool else C.c_void_p(0), nb_pool, _stream(dev)))
    return dL_dverts, dL_dverts_color, dL_dfaces_opacity, dL_dverts_ndc, dL_dfaces_intense, dL_daa_face_verts


def generate_render_layers_cuda(width, height, verts, faces, tets, face_tets, tet_faces, face_existence, verts_ndc,
                                verts_image, image_ray_o, image_ray_d, num_layers, return_taps=False, fused=None,
                                reuse=None, ray_spec=None, overlap_geometry=True):
    """reference render.h:101-119 / render.cu:378-476.  Returns (render_layers, render_layers_cnt).

    fused = (mv [B,4,4], proj [B,4,4], ray_view [B] int32): the fused host prep - the preprocess kernel
    projects the vertices itself (verts_ndc / verts_image are then ignored and may be None) and
    image_ray_o / image_ray_d are the renderer's full (V,H,W,3) buffers indexed through ray_view.

    reuse = the `taps` dict of an earlier call with IDENTICAL geometry, views and rays (return_taps=True):
    only face_existence may differ - binning and first hits are reused, the tet walk alone is redone
    (dm2_regenerate_render_layers).  The caller vouches for the condition.

    ray_spec = (recipe, params [V,32], (V, H, W)) (fused only): the kernels generate the rays themselves;
    image_ray_o / image_ray_d may then be None.

    overlap_geometry: build the geometry / link records while the host waits for num_rendered
    (dm2_layers_preprocess_with_geometry + dm2_generate_render_layers_prepared) instead of after it
    (dm2_layers_preprocess + dm2_generate_render_layers); same results."""
    lib = _lib.load_library()
    use_ray_spec = ray_spec is not None and fused is not None
    if use_ray_spec and image_ray_o is None:        # shape-only placeholders (never dereferenced)
        image_ray_o = image_ray_d = verts.new_empty((0, int(height), int(width), 3))
    if fused is not None:       # shape-only placeholders (never dereferenced: the kernel projects the vertices itself)
        nb = int(fused[0].size(0))
        verts_ndc = verts.new_empty((nb, 0, 3))
        verts_image = verts.new_empty((nb, 0, 2))
    if verts.dim() != 2 or verts.size(1) != 3:
        _err("verts must have dimensions (P, 3)")
    if faces.dim() != 2 or faces.size(1) != 3:
        _err("faces must have dimensions (F, 3)")
    if tets.dim() != 2 or tets.size(1) != 4:
        _err("tets must have dimensions (T, 4)")
    if face_tets.dim() != 2 or face_tets.size(1) != 2:
        _err("face_tets must have dimensions (F, 2)")
    if tet_faces.dim() != 2 or tet_faces.size(1) != 4:
        _err("tet_faces must have dimensions (T, 4)")
    if face_existence.dim() != 1 or face_existence.size(0) != faces.size(0):
        _err("face_existence must have dimensions (F,)")
    if verts_ndc.dim() != 3 or verts_ndc.size(2) != 3:
        _err("verts_ndc must have dimensions (B, P, 3)")
    if verts_image.dim() != 3 or verts_image.size(2) != 2:
        _err("verts_image must have dimensions (B, P, 2)")
    if image_ray_o.dim() != 4 or image_ray_o.size(3) != 3:
        _err("image_ray_o must have dimensions (B, H, W, 3)")
    if image_ray_d.dim() != 4 or image_ray_d.size(3) != 3:
        _err("image_ray_d must have dimensions (B, H, W, 3)")
    if num_layers < 0:
        _err("num_layers must be non-negative")
    f32, i32 = torch.float32, torch.int32
    if os.environ.get("DM2_VALIDATE_INDICES", "0") != "0":
        # debug aid (synchronises): the kernels follow tets / tet_faces / face_tets through raw pointers, where
        # torch indexing in the reference would raise for out-of-range ids
        P_, F_, T_ = int(verts.size(0)), int(faces.size(0)), int(tets.size(0))
        for name, t, lo, hi in (("faces", faces, 0, P_), ("tets", tets, 0, P_), ("tet_faces", tet_faces, 0, F_),
                                ("face_tets", face_tets, -1, T_)):
            if t.numel() and (int(t.min()) < lo or int(t.max()) >= hi):
                _err(f"{name} holds an index outside [{lo}, {hi})")
    keep = dict(verts=_prep(verts, f32, "verts"), faces=_prep(faces, i32, "faces"), tets=_prep(tets, i32, "tets"),
                face_tets=_prep(face_tets, i32, "face_tets"), tet_faces=_prep(tet_faces, i32, "tet_faces"),
                face_existence=_prep(face_existence, i32, "face_existence"),
                verts_ndc=_prep(verts_ndc, f32, "verts_ndc"), verts_image=_prep(verts_image, f32, "verts_image"),
                image_ray_o=image_ray_o if use_ray_spec else _prep(image_ray_o, f32, "image_ray_o"),
                image_ray_d=image_ray_d if use_ray_spec else _prep(image_ray_d, f32, "image_ray_d"))
    B, P, F, T = int(verts_ndc.size(0)), int(verts.size(0)), int(faces.size(0)), int(tets.size(0))
    H, W, L = int(height), int(width), int(num_layers)
    if face_tets.size(0) != F or tet_faces.size(0) != T:
        _err("face_tets / tet_faces must have dimensions (F, 2) / (T, 4)")
    if fused is None and (verts_ndc.size(1) != P or tuple(verts_image.shape[:2]) != (B, P)):
        _err("verts_ndc / verts_image must have dimensions (B, P, .)")
    for t in (keep["image_ray_o"], keep["image_ray_d"]):
        if tuple(t.shape[1:3]) != (H, W) or (fused is None and t.size(0) != B):
            _err("image rays must have dimensions (B, height, width, 3)")
    s = DM2LayerInputs()
    s.B, s.P, s.F, s.T, s.H, s.W, s.num_layers = B, P, F, T, H, W, L
    for k, t in keep.items():
        setattr(s, k, t.data_ptr() if t.numel() > 0 else 0)
    if fused is not None:
        keep["mv"], keep["proj"] = _prep(fused[0], f32, "mv"), _prep(fused[1], f32, "proj")
        keep["ray_view"] = _prep(fused[2], i32, "ray_view")
        if tuple(keep["mv"].shape) != (B, 4, 4) or tuple(keep["proj"].shape) != (B, 4, 4) or keep["ray_view"].numel() != B:
            _err("mv / proj must have dimensions (B, 4, 4)")
        s.mv, s.proj, s.ray_view = keep["mv"].data_ptr(), keep["proj"].data_ptr(), keep["ray_view"].data_ptr()
        s.verts_ndc = s.verts_image = 0
        if use_ray_spec:
            keep["ray_params"] = _prep(ray_spec[1], f32, "ray_params")
            s.ray_params, s.ray_recipe, s.ray_v = keep["ray_params"].data_ptr(), int(ray_spec[0]), int(ray_spec[2][0])
            s.image_ray_o = s.image_ray_d = 0
    dev = keep["verts"].device
    with torch.cuda.device(dev):
        stream = _stream(dev)
        render_layers_cnt = torch.zeros((B, H, W), dtype=i32, device=dev)
        render_layers = torch.full((B, H, W, L), -1, dtype=i32, device=dev)
        if reuse is not None:
            _lib.check(lib.dm2_regenerate_render_layers(C.byref(s), _ptr(reuse["image_buffer"]), _ptr(reuse["geo_buffer"]),
                                                        _ptr(render_layers), _ptr(render_layers_cnt), stream))
            return (render_layers, render_layers_cnt, reuse) if return_taps else (render_layers, render_layers_cnt)
        face_buffer = torch.empty((face_layout(B, F, True).total,), dtype=torch.uint8, device=dev)
        image_buffer = torch.empty((image_layout(B, H, W, True).total,), dtype=torch.uint8, device=dev)
        geo_bytes = C.c_size_t(0)
        _lib.check(lib.dm2_layers_geo_bytes(F, T, C.byref(geo_bytes)))
        geo_buffer = torch.empty((geo_bytes.value,), dtype=torch.uint8, device=dev)
        num_rendered, longest = C.c_int64(0), C.c_int64(0)
        if overlap_geometry:
            # the geometry / link records are enqueued behind the header read-back: the GPU builds them while the
            # host waits for num_rendered and allocates the binning workspace
            _lib.check(lib.dm2_layers_preprocess_with_geometry(C.byref(s), _ptr(face_buffer), _ptr(image_buffer),
                                                               _ptr(geo_buffer), C.byref(num_rendered),
                                                               C.byref(longest), stream))
        else:
            _lib.check(lib.dm2_layers_preprocess(C.byref(s), _ptr(face_buffer), _ptr(image_buffer),
                                                 C.byref(num_rendered), C.byref(longest), stream))
        R = int(num_rendered.value)
        binning_buffer = torch.empty((binning_layout(R).total,), dtype=torch.uint8, device=dev)
        gen = lib.dm2_generate_render_layers_prepared if overlap_geometry else lib.dm2_generate_render_layers
        _lib.check(gen(C.byref(s), _ptr(face_buffer), _ptr(binning_buffer), _ptr(image_buffer), _ptr(geo_buffer), R,
                       int(longest.value), _ptr(render_layers), _ptr(render_layers_cnt), stream))
    if return_taps:
        return render_layers, render_layers_cnt, dict(num_rendered=R, face_buffer=face_buffer,
                                                      binning_buffer=binning_buffer, image_buffer=image_buffer,
                                                      geo_buffer=geo_buffer)
    return render_layers, render_layers_cnt


# ------------------------------------------------------------------------------------------------
# debug taps: typed views into the scratch buffers, used by the parity tests
def _view(buf, off, count, dtype):
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    return buf[off:off + nbytes].view(dtype)


def face_taps(face_buffer, B, F, layered=False):
    L = face_layout(B, F, layered)
    n = B * F
    return dict(depths=_view(face_buffer, L.depths, n, torch.float32),
                min_depths=_view(face_buffer, L.min_depths, n, torch.float32),
                max_depths=_view(face_buffer, L.max_depths, n, torch.float32),
                tiles_touched=_view(face_buffer, L.tiles_touched, n, torch.int32))


def binning_taps(binning_buffer, R):
    L = binning_layout(R)
    return dict(keys_unsorted=_view(binning_buffer, L.keys_unsorted, R, torch.int64),
                vals_unsorted=_view(binning_buffer, L.vals_unsorted, R, torch.int32),
                keys=_view(binning_buffer, L.keys, R, torch.int64),
                vals=_view(binning_buffer, L.vals, R, torch.int32))


def image_taps(image_buffer, B, H, W, layered=False):
    L = image_layout(B, H, W, layered)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out = dict(ranges=_view(image_buffer, L.ranges, B * tiles * 2, torch.int32).view(B * tiles, 2))
    n = B * H * W
    if layered:
        out["first_face"] = _view(image_buffer, L.first_face, n, torch.int32).view(B, H, W)
        out["first_tet"] = _view(image_buffer, L.first_tet, n, torch.int32).view(B, H, W)
    else:
        out["final_T"] = _view(image_buffer, L.final_T, n, torch.float32).view(B, H, W)
        out["final_prev_T"] = _view(image_buffer, L.final_prev_T, n, torch.float32).view(B, H, W)
        out["n_contrib"] = _view(image_buffer, L.n_contrib, n, torch.int32).view(B, H, W)
    return out


def sort_pairs(keys, vals, end_bit):
    """Stable radix sort of (int64 keys, int32 vals) on key bits [0, end_bit); inputs are not modified."""
    lib = _lib.load_library()
    keys = _prep(keys, torch.int64, "keys").clone()
    vals = _prep(vals, torch.int32, "vals").clone()
    n = keys.numel()
    if vals.numel() != n:
        _err("keys and vals must have the same length")
    dev = keys.device
    with torch.cuda.device(dev):
        kout, vout = torch.empty_like(keys), torch.empty_like(vals)
        tb = C.c_size_t(0)
        _lib.check(lib.dm2_sort_temp_bytes(n, C.byref(tb)))
        temp = torch.empty((max(tb.value, 1),), dtype=torch.uint8, device=dev)
        _lib.check(lib.dm2_sort_pairs(_ptr(keys), _ptr(vals), _ptr(kout), _ptr(vout), n, int(end_bit), _ptr(temp),
                                      tb.value, _stream(dev)))
    return kout, vout


def inclusive_scan_u32(x):
    lib = _lib.load_library()
    x = _prep(x, torch.int32, "x")
    n = x.numel()
    dev = x.device
    with torch.cuda.device(dev):
        out = torch.empty_like(x)
        cnt = C.c_int64(0)
        _lib.check(lib.dm2_scan_partials_count(n, C.byref(cnt)))
        partials = torch.empty((max(int(cnt.value), 1),), dtype=torch.int32, device=dev)
        _lib.check(lib.dm2_inclusive_scan_u32(_ptr(x), _ptr(out), n, _ptr(partials), _stream(dev)))
    return out
