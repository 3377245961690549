"""``RenderFunction``: the single autograd boundary of the renderer.

Signature-compatible with the reference's ``RenderFunction`` (dmesh2_renderer/__init__.py:11-177):
21 positional inputs, outputs ``(color [B,H,W,3], depth [B,H,W])`` (depth in NDC, before the
caller's 1 - (d+1)/2 remap), and a 21-tuple of gradients that is non-None at positions
4 (verts), 6 (verts_color), 7 (faces_opacity), 8 (verts_ndc: z channel only), 10 (faces_intense),
12 (aa_face_verts).  What is stashed on ``ctx`` is private: the scratch buffers and the compact AA
cache pool (``len_oarea_buffer > 0``; see _C.py).
"""
import torch

from . import _C


class RenderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx,
                background,               # (3)            no grad
                patch_min,                # (B, 2) int32   no grad
                patch_width, patch_height,
                verts,                    # (P, 3)         grad
                faces,                    # (F, 3) int32   no grad
                verts_color,              # (P, 3)         grad
                faces_opacity,            # (F,)           grad
                verts_ndc,                # (B, P, 3)      grad (z channel only)
                verts_image,              # (B, P, 2)      no grad
                faces_intense,            # (B, F)         grad
                aa_temperature,
                aa_face_verts,            # (B, F, 3, 2)   grad
                aa_face_edges,            # (B, F, 3, 2)   no grad
                aa_face_edges_iszero,     # (B, F, 3, 2)   no grad (bool)
                aa_face_edges_recip,      # (B, F, 3, 2)   no grad
                aa_face_edges_normal,     # (B, F, 3, 2)   no grad
                aa_face_edges_normal_c,   # (B, F, 3)      no grad
                len_oarea_buffer,         # int
                image_ray_o,              # (B, H, W, 3)   no grad
                image_ray_d):             # (B, H, W, 3)   no grad
        tensors = (background, patch_min, verts, faces, verts_color, faces_opacity, verts_ndc, verts_image,
                   faces_intense, aa_face_verts, aa_face_edges, aa_face_edges_iszero, aa_face_edges_recip,
                   aa_face_edges_normal, aa_face_edges_normal_c, image_ray_o, image_ray_d)
        try:
            (num_rendered, color, depth, oarea, _tri_id, _tri_cnt, _doarea,
             face_buffer, binning_buffer, image_buffer) = _C.render_forward_cuda(
                background, patch_min, patch_width, patch_height, verts, faces, verts_color, faces_opacity,
                verts_ndc, verts_image, faces_intense, aa_temperature, aa_face_verts, aa_face_edges,
                aa_face_edges_iszero, aa_face_edges_recip, aa_face_edges_normal, aa_face_edges_normal_c,
                len_oarea_buffer, image_ray_o, image_ray_d, want_cache=any(ctx.needs_input_grad))
        except Exception as ex:
            print("\nAn error occured in renderer forward.")
            print(ex)
            raise ex
        ctx.save_for_backward(*tensors, face_buffer, binning_buffer, image_buffer, oarea)
        ctx.num_rendered = num_rendered
        ctx.patch_width = patch_width
        ctx.patch_height = patch_height
        ctx.aa_temperature = aa_temperature
        ctx.len_oarea_buffer = len_oarea_buffer
        return color, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_depth):
        (background, patch_min, verts, faces, verts_color, faces_opacity, verts_ndc, verts_image, faces_intense,
         aa_face_verts, aa_face_edges, aa_face_edges_iszero, aa_face_edges_recip, aa_face_edges_normal,
         aa_face_edges_normal_c, image_ray_o, image_ray_d, face_buffer, binning_buffer, image_buffer,
         oarea) = ctx.saved_tensors
        try:
            (g_verts, g_color, g_opacity, g_ndc, g_intense, g_aa) = _C.render_backward_cuda(
                ctx.num_rendered, background, patch_min, ctx.patch_width, ctx.patch_height, verts, faces,
                verts_color, faces_opacity, verts_ndc, verts_image, faces_intense, ctx.aa_temperature,
                aa_face_verts, aa_face_edges, aa_face_edges_iszero, aa_face_edges_recip, aa_face_edges_normal,
                aa_face_edges_normal_c, ctx.len_oarea_buffer, image_ray_o, image_ray_d,
                grad_out_color, grad_out_depth, face_buffer, binning_buffer, image_buffer, oarea)
        except Exception as ex:
            print("\nAn error occured in renderer backward.")
            print(ex)
            raise ex
        grads = [None] * 21
        grads[4], grads[6], grads[7], grads[8], grads[10], grads[12] = g_verts, g_color, g_opacity, g_ndc, g_intense, g_aa
        return tuple(grads)
