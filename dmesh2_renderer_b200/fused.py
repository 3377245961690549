"""Fused host prep path of ``Renderer.forward``.

The reference's ``Renderer.forward`` runs ~25 torch kernels of host-side prep before its single
autograd Function (dmesh2_renderer/__init__.py:325-347, pyrenderer.py:6-30) and autograd replays
them backwards.  ``FusedRenderFunction`` takes the vertices and the selected mv / proj matrices and
lets the preprocess kernel do everything: the two K=4 matmuls (as the fma chain cuBLAS runs for
them, bit-for-bit), the w clamp, the perspective divide, the image transform, the per-face gather,
the CCW re-ordering and the derived AA terms, with the same IEEE operations in the same order; the
backward scatter kernel applies the matching chain rule down to d/d verts.  Colour/depth, sort
order and tile ranges are bit-identical to the unfused ``RenderFunction`` path
(tests/test_gpu_parity.py::test_fused_path_equals_reference_shaped_path).
"""
import ctypes as C

import torch

from . import _lib
from ._C import (_err, _prep, _ptr, _stream, binning_layout, face_layout, image_layout, shade_cache_alloc)
from ._lib import DM2FusedGrads, DM2FusedInputs


def _fused_struct(background, patch_min, patch_width, patch_height, verts, faces, verts_color, faces_opacity,
                  mv, proj, faces_intense, aa_temperature, ray_o, ray_d, ray_view, depth_unit=False, ray_spec=None):
    f32, i32 = torch.float32, torch.int32
    keep = dict(background=_prep(background, f32, "background"), patch_min=_prep(patch_min, i32, "patch_min"),
                verts=_prep(verts, f32, "verts"), faces=_prep(faces, i32, "faces"),
                verts_color=_prep(verts_color, f32, "verts_color"),
                faces_opacity=_prep(faces_opacity, f32, "faces_opacity"),
                mv=_prep(mv, f32, "mv"), proj=_prep(proj, f32, "proj"),
                faces_intense=_prep(faces_intense, f32, "faces_intense"),
                ray_view=_prep(ray_view, i32, "ray_view"))
    if ray_spec is None:
        keep["ray_o"], keep["ray_d"] = _prep(ray_o, f32, "ray_o"), _prep(ray_d, f32, "ray_d")
    else:
        # the kernels generate the rays (recipe verified bit-identical by Renderer._init_rays): no ray tensors
        keep["ray_params"] = _prep(ray_spec[1], f32, "ray_params")
    if keep["background"].dim() != 1 or keep["background"].size(0) != 3:
        _err("background must have dimensions (3,)")
    if keep["verts"].dim() != 2 or keep["verts"].size(1) != 3:
        _err("verts must have dimensions (P, 3)")
    if keep["faces"].dim() != 2 or keep["faces"].size(1) != 3:
        _err("faces must have dimensions (F, 3)")
    P, F = int(keep["verts"].size(0)), int(keep["faces"].size(0))
    if keep["mv"].dim() != 3 or tuple(keep["mv"].shape[1:]) != (4, 4) or keep["mv"].shape != keep["proj"].shape:
        _err("mv / proj must have dimensions (B, 4, 4)")
    B = int(keep["mv"].size(0))
    if tuple(keep["verts_color"].shape) != (P, 3):
        _err("vert color must have dimensions (P, 3)")
    if tuple(keep["faces_opacity"].shape) != (F,):
        _err("face opacity must have dimensions (F,)")
    if tuple(keep["faces_intense"].shape) != (B, F):
        _err("faces_intense must have dimensions (B, F,)")
    if tuple(keep["patch_min"].shape) != (B, 2) or keep["ray_view"].numel() != B:
        _err("patch_min must have dimensions (B, 2)")
    if ray_spec is None:
        ro, rd = keep["ray_o"], keep["ray_d"]
        if ro.dim() != 4 or ro.size(3) != 3 or ro.shape != rd.shape:
            _err("image_ray_o must have dimensions (B, H, W, 3)")
        frame = (int(ro.size(1)), int(ro.size(2)))
    else:
        if tuple(keep["ray_params"].shape) != (ray_spec[2][0], 32):
            _err("ray_params must have dimensions (V, 32)")
        frame = (int(ray_spec[2][1]), int(ray_spec[2][2]))
    if aa_temperature < 0 or aa_temperature > 1:
        _err("aa_temperature must be in the range [0, 1]")
    s = DM2FusedInputs()
    s.B, s.P, s.F, s.H, s.W = B, P, F, int(patch_height), int(patch_width)
    s.image_h, s.image_w = frame
    s.aa_temperature = float(aa_temperature)
    s.depth_unit = 1 if depth_unit else 0
    s.ray_recipe = int(ray_spec[0]) if ray_spec is not None else 0
    s.ray_v = int(ray_spec[2][0]) if ray_spec is not None else 0
    for k, t in keep.items():
        setattr(s, k, t.data_ptr() if t.numel() > 0 else 0)
    return s, keep


class _FrameHint:
    """What the last forward call of one problem shape needed: num_rendered, the longest tile segment and the
    shading-cache demand (read back asynchronously into pinned memory; consumed once its event has completed)."""

    def __init__(self):
        self.num_rendered = 0
        self.longest = -1
        self.cache_granules = 0
        self.cache_bytes = 0          # size of the pool the demand counter above was measured on
        self._pending_bytes = 0
        self._pinned = None
        self._event = None

    def note_cache_counter(self, pool, stream):
        if pool.numel() < 4:
            return
        if self._pinned is None:
            self._pinned = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self._pinned.copy_(pool[:4].view(torch.int32), non_blocking=True)
        self._pending_bytes = int(pool.numel())
        self._event = torch.cuda.Event()
        self._event.record(stream)

    def poll(self):
        if self._event is not None and self._event.query():
            self.cache_granules = int(self._pinned[0]) & 0xFFFFFFFF
            self.cache_bytes = self._pending_bytes
            self._event = None


class GradSink:
    """Optional destination of the replicated-parameter gradients: when a sink is passed to
    FusedRenderFunction, its backward ADDS dL/dverts, dL/dverts_color and dL/dfaces_opacity into these
    (caller-zeroed, fp32, contiguous) tensors with the kernels' own atomics and returns None for the three
    inputs - e.g. the views of a multiview.GradBucket, so that a multi-view step needs no gradient copies."""

    def __init__(self, dverts, dverts_color, dfaces_opacity):
        self.dverts, self.dverts_color, self.dfaces_opacity = dverts, dverts_color, dfaces_opacity


class FusedRenderFunction(torch.autograd.Function):
    """(background, patch_min, pw, ph, verts, faces, verts_color, faces_opacity, mv [B,4,4], proj [B,4,4],
    faces_intense, aa_temperature, ray_o, ray_d, ray_view, len_oarea_buffer, depth_unit, grad_sink, ray_spec) -> (color
    [B,H,W,3], depth [B,H,W]).  Differentiable w.r.t. verts, verts_color, faces_opacity, faces_intense.
    len_oarea_buffer > 0 turns the forward -> backward shading cache on (see _C.py).  depth_unit = False: depth is
    the NDC depth render_forward_cuda returns; True: the kernels apply Renderer.forward's remap
    1 - (depth_ndc + 1) / 2 (reference __init__.py:377-378) themselves, bit-identically, forward and backward.
    grad_sink: see GradSink.  ray_spec: (recipe, params [V,32], (V, H, W)) from Renderer._ray_spec() - the kernels
    then generate the rays themselves and ray_o / ray_d are ignored (may be None)."""

    # Per problem shape: what the last call needed.  The binning workspace and the shading cache are allocated
    # from it BEFORE anything is launched and the whole forward pass is enqueued in one library call
    # (dm2_fused_forward_pipelined): the kernels check on the device that the frame fits; the host only waits
    # for the 12-byte header after everything is in the queue and re-runs the tail in the rare case it does not.
    _size_hint = {}

    @staticmethod
    def forward(ctx, background, patch_min, patch_width, patch_height, verts, faces, verts_color, faces_opacity,
                mv, proj, faces_intense, aa_temperature, ray_o, ray_d, ray_view, len_oarea_buffer=20,
                depth_unit=False, grad_sink=None, ray_spec=None):
        lib = _lib.load_library()
        s, keep = _fused_struct(background, patch_min, patch_width, patch_height, verts, faces, verts_color,
                                faces_opacity, mv, proj, faces_intense, aa_temperature, ray_o, ray_d, ray_view,
                                depth_unit, ray_spec)
        dev = keep["verts"].device
        B, H, W = s.B, s.H, s.W
        with torch.cuda.device(dev):
            stream = _stream(dev)
            color = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
            depth = torch.empty((B, H, W), dtype=torch.float32, device=dev)
            face_buffer = torch.empty((face_layout(B, s.F).total,), dtype=torch.uint8, device=dev)
            image_buffer = torch.empty((image_layout(B, H, W).total,), dtype=torch.uint8, device=dev)
            want_cache = any(ctx.needs_input_grad)
            hint_key = (B, s.P, s.F, H, W, dev.index)
            hint = FusedRenderFunction._size_hint.setdefault(hint_key, _FrameHint())
            hint.poll()
            nr, longest, fits = C.c_int64(0), C.c_int64(-1), C.c_int32(0)
            cap = 0
            if hint.num_rendered > 0:
                cap = hint.num_rendered + hint.num_rendered // 16 + 1024
                binning_buffer = torch.empty((binning_layout(cap).total,), dtype=torch.uint8, device=dev)
                pool = shade_cache_alloc(int(len_oarea_buffer), cap, float(aa_temperature), dev, want_cache,
                                         hint.cache_granules, hint.cache_bytes)
                lhint = -1 if hint.longest < 0 else hint.longest       # (the library adds its own slack)
                _lib.check(lib.dm2_fused_forward_pipelined(
                    C.byref(s), _ptr(face_buffer), _ptr(binning_buffer), _ptr(image_buffer), cap, lhint, _ptr(color),
                    _ptr(depth), _ptr(pool), int(pool.numel()), C.byref(nr), C.byref(longest), C.byref(fits), stream))
            else:
                _lib.check(lib.dm2_fused_preprocess(C.byref(s), _ptr(face_buffer), _ptr(image_buffer), C.byref(nr),
                                                    C.byref(longest), stream))
            R = int(nr.value)
            if not fits.value:
                # first call of this shape, or the frame outgrew the previous one by more than the slack
                # (the layout of R instances is what the kernels index with: it travels to the backward call)
                cap = R
                binning_buffer = torch.empty((binning_layout(cap).total,), dtype=torch.uint8, device=dev)
                pool = shade_cache_alloc(int(len_oarea_buffer), cap, float(aa_temperature), dev, want_cache,
                                         hint.cache_granules, hint.cache_bytes)
                _lib.check(lib.dm2_fused_forward(C.byref(s), _ptr(face_buffer), _ptr(binning_buffer), _ptr(image_buffer),
                                                 cap, int(longest.value), _ptr(color), _ptr(depth), _ptr(pool),
                                                 int(pool.numel()), stream))
            hint.num_rendered, hint.longest = R, int(longest.value)
            hint.note_cache_counter(pool, torch.cuda.current_stream(dev))
            if depth_unit and s.P == 0:
                depth.fill_(0.5)          # the library's P == 0 shortcut writes depth_ndc = 0
        ctx.save_for_backward(*keep.values(), face_buffer, binning_buffer, image_buffer, pool)
        ctx.keys = list(keep.keys())
        ctx.num_rendered = R
        ctx.capacity = cap
        ctx.patch = (patch_width, patch_height)
        ctx.aa_temperature = aa_temperature
        ctx.depth_unit = bool(depth_unit)
        ctx.grad_sink = grad_sink
        ctx.ray_spec = None if ray_spec is None else (ray_spec[0], None, ray_spec[2])
        return color, depth

    @staticmethod
    def backward(ctx, grad_color, grad_depth):
        lib = _lib.load_library()
        saved = ctx.saved_tensors
        keep = dict(zip(ctx.keys, saved[:len(ctx.keys)]))
        face_buffer, binning_buffer, image_buffer, pool = saved[len(ctx.keys):]
        ray_spec = None if ctx.ray_spec is None else (ctx.ray_spec[0], keep["ray_params"], ctx.ray_spec[2])
        s, keep = _fused_struct(keep["background"], keep["patch_min"], ctx.patch[0], ctx.patch[1], keep["verts"],
                                keep["faces"], keep["verts_color"], keep["faces_opacity"], keep["mv"], keep["proj"],
                                keep["faces_intense"], ctx.aa_temperature, keep.get("ray_o"), keep.get("ray_d"),
                                keep["ray_view"], ctx.depth_unit, ray_spec)
        dev = keep["verts"].device
        gc = _prep(grad_color, torch.float32, "dL_dout_color")
        gd = _prep(grad_depth, torch.float32, "dL_dout_depth")
        sink = ctx.grad_sink
        with torch.cuda.device(dev):
            if sink is not None:
                g_verts, g_color, g_opacity = sink.dverts, sink.dverts_color, sink.dfaces_opacity
                for t, ref in ((g_verts, keep["verts"]), (g_color, keep["verts_color"]), (g_opacity, keep["faces_opacity"])):
                    if not (t.is_cuda and t.device == dev and t.dtype == torch.float32 and t.is_contiguous()
                            and t.numel() == ref.numel()):
                        _err("grad_sink tensors must be contiguous fp32 CUDA tensors of the parameters' sizes")
            else:
                g_verts = torch.zeros_like(keep["verts"])
                g_color = torch.zeros_like(keep["verts_color"])
                g_opacity = torch.zeros_like(keep["faces_opacity"])
            g_intense = torch.zeros_like(keep["faces_intense"])
            if s.F != 0:
                g = DM2FusedGrads(g_verts.data_ptr(), g_color.data_ptr(), g_opacity.data_ptr(), g_intense.data_ptr())
                nb = C.c_size_t(0)
                _lib.check(lib.dm2_backward_workspace_bytes(s.B, s.F, C.byref(nb)))
                grad_ws = torch.empty((nb.value,), dtype=torch.uint8, device=dev)
                _lib.check(lib.dm2_fused_backward(C.byref(s), _ptr(face_buffer), _ptr(binning_buffer),
                                                  _ptr(image_buffer), ctx.capacity, _ptr(gc), _ptr(gd),
                                                  C.byref(g), _ptr(grad_ws), _ptr(pool), int(pool.numel()),
                                                  _stream(dev)))
        if sink is not None:
            g_verts = g_color = g_opacity = None
        return (None, None, None, None, g_verts, None, g_color, g_opacity, None, None, g_intense, None, None, None,
                None, None, None, None, None)
