"""Layer consumer: what DMesh++ does with the output of ``LayeredRenderer.generate``.

The reference returns per-ray face ids only (dmesh2_renderer/__init__.py:404-451); the caller then gathers the
attributes of the recorded faces and composites them front to back with torch ops (L gathers, a cumulative
product, a sum - and their autograd mirrors).  ``composite_layers`` does both in one kernel per direction
(csrc/dm2_layers.cu, C-ABI dm2_layers_composite_forward / _backward):

    C = sum_l T_l a_l c_l + T_L background,   T_0 = 1,  T_{l+1} = T_l (1 - a_l),   alpha = 1 - T_L

over the first min(cnt, L) layers of every ray, with a_l = faces_opacity[f_l] and c_l = faces_color[f_l]
(times faces_intense[b, f_l] when given).  Differentiable w.r.t. faces_color, faces_opacity and faces_intense;
the layer ids are a discrete input.  ``composite_layers_torch`` is the plain-torch statement of the same op
(the test oracle).
"""
import ctypes as C

import torch

from . import _lib
from ._C import _err, _prep, _ptr, _stream


def composite_layers_torch(render_layers, render_layers_cnt, faces_color, faces_opacity, background, faces_intense=None):
    """Reference semantics in plain torch (any device / dtype)."""
    B, H, W, L = render_layers.shape
    ids = render_layers.clamp(min=0).long()
    valid = (torch.arange(L, device=ids.device).view(1, 1, 1, L) < render_layers_cnt.clamp(max=L).unsqueeze(-1)) & \
        (render_layers >= 0)
    a = faces_opacity[ids] * valid
    c = faces_color[ids]
    if faces_intense is not None:
        c = c * torch.gather(faces_intense, 1, ids.view(B, -1)).view(B, H, W, L, 1)
    one_minus = 1.0 - a
    T = torch.cumprod(torch.cat((torch.ones_like(a[..., :1]), one_minus[..., :-1]), dim=-1), dim=-1)
    T_end = T[..., -1] * one_minus[..., -1] if L > 0 else torch.ones((B, H, W), dtype=faces_color.dtype, device=ids.device)
    color = ((T * a).unsqueeze(-1) * c).sum(-2) + T_end.unsqueeze(-1) * background
    return color, 1.0 - T_end


class LayerCompositeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render_layers, render_layers_cnt, faces_color, faces_opacity, background, faces_intense):
        lib = _lib.load_library()
        f32, i32 = torch.float32, torch.int32
        layers = _prep(render_layers, i32, "render_layers")
        cnt = _prep(render_layers_cnt, i32, "render_layers_cnt")
        color = _prep(faces_color, f32, "faces_color")
        opacity = _prep(faces_opacity, f32, "faces_opacity")
        bg = _prep(background, f32, "background")
        intense = _prep(faces_intense, f32, "faces_intense") if faces_intense is not None else None
        if layers.dim() != 4 or tuple(cnt.shape) != tuple(layers.shape[:3]):
            _err("render_layers / render_layers_cnt must have dimensions (B, H, W, L) / (B, H, W)")
        B, H, W, L = (int(x) for x in layers.shape)
        F = int(opacity.shape[0])
        if color.dim() != 2 or tuple(color.shape) != (F, 3) or opacity.dim() != 1:
            _err("faces_color / faces_opacity must have dimensions (F, 3) / (F,)")
        if bg.dim() != 1 or bg.shape[0] != 3:
            _err("background must have dimensions (3,)")
        if intense is not None and tuple(intense.shape) != (B, F):
            _err("faces_intense must have dimensions (B, F,)")
        dev = layers.device
        with torch.cuda.device(dev):
            out_color = torch.empty((B, H, W, 3), dtype=f32, device=dev)
            out_alpha = torch.empty((B, H, W), dtype=f32, device=dev)
            _lib.check(lib.dm2_layers_composite_forward(B, H, W, L, F, _ptr(layers), _ptr(cnt), _ptr(color), _ptr(opacity),
                                                        _ptr(intense), _ptr(bg), _ptr(out_color), _ptr(out_alpha),
                                                        _stream(dev)))
        ctx.save_for_backward(layers, cnt, color, opacity, bg, *((intense,) if intense is not None else ()))
        ctx.has_intense = intense is not None
        ctx.dims = (B, H, W, L, F)
        return out_color, out_alpha

    @staticmethod
    def backward(ctx, g_color, g_alpha):
        lib = _lib.load_library()
        saved = ctx.saved_tensors
        layers, cnt, color, opacity, bg = saved[:5]
        intense = saved[5] if ctx.has_intense else None
        B, H, W, L, F = ctx.dims
        dev = layers.device
        gc = _prep(g_color, torch.float32, "dL_dout_color")
        ga = _prep(g_alpha, torch.float32, "dL_dout_alpha")
        with torch.cuda.device(dev):
            d_color = torch.zeros_like(color)
            d_opacity = torch.zeros_like(opacity)
            d_intense = torch.zeros_like(intense) if (intense is not None and ctx.needs_input_grad[5]) else None
            if F > 0:
                ws = torch.empty((max(F, 1) * 16,), dtype=torch.uint8, device=dev)
                _lib.check(lib.dm2_layers_composite_backward(B, H, W, L, F, _ptr(layers), _ptr(cnt), _ptr(color),
                                                             _ptr(opacity), _ptr(intense), _ptr(bg), _ptr(gc), _ptr(ga),
                                                             _ptr(ws), _ptr(d_color), _ptr(d_opacity), _ptr(d_intense),
                                                             _stream(dev)))
        return None, None, d_color, d_opacity, None, d_intense


def composite_layers(render_layers, render_layers_cnt, faces_color, faces_opacity, background, faces_intense=None):
    """(render_layers [B,H,W,L] int32, render_layers_cnt [B,H,W] int32, faces_color [F,3], faces_opacity [F],
    background [3], faces_intense [B,F] or None) -> (color [B,H,W,3], alpha [B,H,W]).  CUDA tensors only."""
    return LayerCompositeFunction.apply(render_layers, render_layers_cnt, faces_color, faces_opacity, background,
                                        faces_intense)
