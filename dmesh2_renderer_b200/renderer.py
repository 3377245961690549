"""``Renderer`` and ``LayeredRenderer``: drop-in replacements for the two classes of
SonSang/dmesh2_renderer (reference dmesh2_renderer/__init__.py:179-380, 388-451).

Constructor arguments, method names, argument order, defaults, tensor layouts and dtypes are the
reference's.  The per-call host work (projection, CCW triangle set-up, ray selection) issues the
same fp32 torch ops in the same order, so the tensors handed to ``RenderFunction`` are bit-identical
to the reference's; everything below ``RenderFunction`` is libdm2_b200.so.
"""
from typing import List

import torch

from . import _C
from .function import RenderFunction
from .triangles import Triangles

W_EPS = 1e-4      # |w| clamp of the projection (reference __init__.py:253-255)


def _unproject_pixel_rays(mv, proj, width, height, device):
    """One ray per pixel centre for every camera (reference __init__.py:198-237).

    Origin: camera position (translation column of mv^-1).  Direction: normalised vector to the
    pixel centre un-projected from NDC z = -1 *without* the homogeneous divide (upstream behaviour:
    exact only when the near plane is at distance 1 or the projection is orthographic).
    """
    n_view = mv.shape[0]
    inv_mv = torch.inverse(mv)
    inv_proj = torch.inverse(proj)
    origin = torch.zeros((n_view, height, width, 3), device=device)
    for k in range(3):
        origin[:, :, :, k] = inv_mv[:, k, 3].reshape(-1, 1, 1)

    xs = torch.arange(width, device=device).float()
    ys = torch.arange(height, device=device).float()
    centre = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1).float() + 0.5      # [W, H, 2]
    centre = centre.transpose(0, 1).unsqueeze(0).expand(n_view, -1, -1, -1)                  # [B, H, W, 2]
    ndc = torch.zeros_like(centre)
    ndc[:, :, :, 0] = (centre[:, :, :, 0] / width * 2) - 1
    ndc[:, :, :, 1] = (centre[:, :, :, 1] / height * 2) - 1
    near = -torch.ones_like(ndc[:, :, :, 0:1])
    clip = torch.cat((ndc, near), dim=-1)
    clip = torch.cat((clip, torch.ones_like(clip[:, :, :, 0:1])), dim=-1).unsqueeze(-2)      # [B, H, W, 1, 4]
    view = torch.matmul(clip, inv_proj.transpose(1, 2).unsqueeze(1).unsqueeze(1))
    world = torch.matmul(view, inv_mv.transpose(1, 2).unsqueeze(1).unsqueeze(1))
    target = world[..., :3].squeeze(-2)                                                      # [B, H, W, 3]
    direction = target - origin
    length = torch.norm(direction, dim=-1, keepdim=True) + 1e-6
    return origin, direction / length, inv_mv, inv_proj


def _find_ray_recipe(ray_o, ray_d, inv_mv, inv_proj):
    """In-kernel ray generation (include/dm2_renderer.h `dm2_generate_rays`).

    The rays are a function of 32 floats per view, but the rounding of the torch ops that produced ray_o /
    ray_d depends on library choices (division by a host scalar, cuBLAS accumulation order, the norm's reduction
    order).  The library implements a small family of arithmetic "recipes"; this probe generates ALL rays with
    each of them once and returns (recipe, params [V, 32]) for the first one whose output equals the torch
    tensors BIT FOR BIT - only then do the kernels compute rays instead of reading them.  (0, None) otherwise:
    the kernels keep reading the tensors, which is always correct."""
    import ctypes as C

    from . import _lib
    if not (ray_o.is_cuda and ray_o.dtype == torch.float32 and inv_mv.dtype == torch.float32 and ray_o.numel() > 0):
        return 0, None
    V, H, W = int(ray_o.shape[0]), int(ray_o.shape[1]), int(ray_o.shape[2])
    if V * H * W >= 2 ** 31:
        return 0, None
    lib = _lib.load_library()
    params = torch.cat((inv_proj.reshape(V, 16), inv_mv.reshape(V, 16)), dim=1).contiguous()
    dev = ray_o.device
    with torch.cuda.device(dev):
        o = torch.empty_like(ray_o)
        d = torch.empty_like(ray_d)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ro, rd = ray_o.contiguous(), ray_d.contiguous()
        # (the combinations measured on B200 first: D = 1, M = 0, N = 2, with and without the small-last-chunk rule)
        first = [18 * 1 + 9 * 1 + 0 + 2 + 1, 9 * 1 + 0 + 2 + 1]
        for recipe in first + [q for q in range(1, _lib.RAY_RECIPES + 1) if q not in first]:
            _lib.check(lib.dm2_generate_rays(C.c_void_p(params.data_ptr()), V, H, W, recipe, C.c_void_p(o.data_ptr()),
                                             C.c_void_p(d.data_ptr()), stream))
            if torch.equal(d, rd) and torch.equal(o, ro):
                return recipe, params
    return 0, None


class Renderer(torch.nn.Module):
    """Differentiable renderer with global (per-tile) depth ordering and anti-aliased coverage."""

    def __init__(self, mv, proj, width, height, device, aa_grad_buffer_size=20):
        super().__init__()
        self._setup(mv, proj, width, height, device)
        # > 0: the forward pass hands its per-contribution shading state (clamped barycentrics, coverage,
        # d(area)/d(aa verts)) to the backward pass through a compact pool (_C.shade_cache_alloc - the role of
        # the reference's K-deep per-pixel AA buffers); 0: the backward pass recomputes it
        # self.grad_sink (optional, fused.GradSink): gradients of verts / verts_color / faces_opacity are added
        # into the sink's tensors instead of being returned to autograd (multi-view steps: no gradient copies)
        self.aa_grad_buffer_size = aa_grad_buffer_size

    def _setup(self, mv, proj, width, height, device):
        self.mv = mv                  # [B, 4, 4]
        self.proj = proj              # [B, 4, 4]
        self.width = width
        self.height = height
        self.device = device
        self.num_batch = self.mv.shape[0]
        self.ray_o = None
        self.ray_d = None
        self._init_rays()

    def _init_rays(self):
        self.ray_o, self.ray_d, self._inv_mv, self._inv_proj = _unproject_pixel_rays(self.mv, self.proj, self.width,
                                                                                     self.height, self.device)
        self._ray_recipe, self._ray_params = 0, None

    def enable_kernel_rays(self):
        """In-kernel ray generation: looks for a ray recipe that reproduces this renderer's
        ray_o / ray_d tensors bit for bit (_find_ray_recipe); if there is one, the fused routes compute the rays in
        the kernels' prologues from 32 floats per view instead of reading 24 bytes per pixel.  Returns whether it
        is active.  OFF by default: measured on B200 at 1 M triangles @1024^2 it costs 1.2 % of the frame (the
        kernels are instruction bound, HBM is idle) - its use is the memory release_ray_buffers() frees."""
        if not self._ray_recipe and self.ray_o is not None and self.ray_o.is_cuda:
            try:
                self._ray_recipe, self._ray_params = _find_ray_recipe(self.ray_o, self.ray_d, self._inv_mv.detach(),
                                                                      self._inv_proj.detach())
            except RuntimeError:
                self._ray_recipe, self._ray_params = 0, None
        return bool(self._ray_recipe)

    def release_ray_buffers(self):
        """Drops the resident (V, H, W, 3) ray tensors (24 bytes per pixel and view) and lets the kernels generate
        the rays themselves.  Only the fused routes keep working afterwards; raises if no ray recipe reproduces
        the tensors (then they are needed)."""
        if not self.enable_kernel_rays():
            raise RuntimeError("in-kernel ray generation is not available for this renderer: the ray tensors are needed")
        self.ray_o = self.ray_d = None

    def _view_ids(self, batch_mvp_idx):
        """(key, ray_view int32 [B], view_idx int64 [B]) for the fused routes.  Python-style negative view ids
        are normalised (the reference's `self.mv[idx]` accepts them); out-of-range ids raise IndexError as
        torch indexing does."""
        n = self.num_batch
        key = []
        for i in batch_mvp_idx:
            i = int(i)
            if not -n <= i < n:
                raise IndexError(f"index {i} is out of bounds for dimension 0 with size {n}")
            key.append(i + n if i < 0 else i)
        key = tuple(key)
        cache = self.__dict__.setdefault("_view_cache", {})
        if key not in cache:
            idx = torch.tensor(list(key), dtype=torch.int64, device=self.mv.device)
            cache[key] = (idx.to(torch.int32), idx)
        return (key,) + cache[key]

    def _cameras_need_grad(self):
        return torch.is_grad_enabled() and (self.mv.requires_grad or self.proj.requires_grad)

    def _ray_spec(self):
        """(recipe, params, (V, H, W)) for the kernels' own ray generation, or None (read the ray tensors)."""
        if not getattr(self, "_ray_recipe", 0):
            return None
        return self._ray_recipe, self._ray_params, (self.num_batch, self.height, self.width)

    def compute_verts_ndc_image(self, verts, mv, proj):
        """verts [P,3], mv/proj [B,4,4] -> (verts_ndc [B,P,3], verts_image [B,P,2]) (reference :239-262)."""
        hom = torch.cat((verts, torch.ones_like(verts[:, [0]])), dim=-1)           # [P, 4]
        clip = torch.matmul(torch.matmul(hom, mv.transpose(1, 2)), proj.transpose(1, 2))   # [B, P, 4]
        w = clip[..., [3]]
        w = torch.where(torch.logical_and(w >= 0.0, w < W_EPS), torch.full_like(w, W_EPS), w)
        w = torch.where(torch.logical_and(w < 0.0, w > -W_EPS), torch.full_like(w, -W_EPS), w)
        ndc = clip[..., :3] / w
        half = (ndc[..., :2] + 1) * 0.5
        image = torch.stack((half[..., 0] * self.width, half[..., 1] * self.height), dim=-1)
        return ndc, image

    def select_rays(self, batch_mvp_idx, batch_patch_min, patch_width, patch_height):
        """Rays of a [patch_height, patch_width] window per selected view (reference :264-302)."""
        all_o = self.ray_o[batch_mvp_idx]
        all_d = self.ray_d[batch_mvp_idx]
        x0 = batch_patch_min[:, 0]
        y0 = batch_patch_min[:, 1]
        assert (x0 + patch_width <= self.width).all(), "Some b_patch_max_x exceed self.width"
        assert (y0 + patch_height <= self.height).all(), "Some b_patch_max_y exceed self.height"
        n = all_o.shape[0]
        dev = all_o.device
        rows = (y0.view(n, 1, 1).to(dev) + torch.arange(patch_height, device=dev).view(1, patch_height, 1)).long()
        cols = (x0.view(n, 1, 1).to(dev) + torch.arange(patch_width, device=dev).view(1, 1, patch_width)).long()
        view = torch.arange(n, device=dev).view(n, 1, 1)
        return all_o[view, rows, cols], all_d[view, rows, cols]

    def forward(self,
                batch_mvp_idx: List[int],           # [B,]
                batch_patch_min: torch.Tensor,      # [B, 2]
                patch_width: int,
                patch_height: int,
                verts: torch.Tensor,                # [P, 3]
                faces: torch.Tensor,                # [F, 3]
                verts_color: torch.Tensor,          # [P, 3]
                faces_opacity: torch.Tensor,        # [F,]
                faces_intense: torch.Tensor,        # [B, F]
                background: torch.Tensor,           # [3]
                aa_temperature: float = 1.0):
        # The fused route does not differentiate w.r.t. the cameras: when mv / proj require grad the
        # reference-shaped route below is used (autograd then reaches them through the torch matmuls of
        # compute_verts_ndc_image, exactly as in the reference).
        if getattr(self, "fused", True) and verts.is_cuda and not self._cameras_need_grad():
            # (the NDC z -> [0, 1] remap below happens inside the kernels, bit-identically: depth_unit)
            return self._forward_fused(batch_mvp_idx, batch_patch_min, patch_width, patch_height, verts,
                                       faces, verts_color, faces_opacity, faces_intense, background,
                                       aa_temperature)
        else:
            args = self.prepare_inputs(batch_mvp_idx, batch_patch_min, patch_width, patch_height, verts, faces,
                                       verts_color, faces_opacity, faces_intense, background, aa_temperature)
            color, depth = RenderFunction.apply(*args)
        depth = 1.0 - (depth + 1.0) / 2.0        # NDC z -> [0, 1], near = 1 (reference :377-378)
        return color, depth

    def _forward_fused(self, batch_mvp_idx, batch_patch_min, patch_width, patch_height, verts, faces, verts_color,
                       faces_opacity, faces_intense, background, aa_temperature):
        """Same result as the reference-shaped path below, with the host prep done inside the kernels
        (fused.py), including the two clip-space matmuls (same fma chain as cuBLAS, bit for bit)."""
        from .fused import FusedRenderFunction
        _key, ray_view, view_idx = self._view_ids(batch_mvp_idx)
        if not batch_patch_min.is_cuda:
            # the reference's asserts (__init__.py:283-284); for a CUDA tensor they would force a device sync
            # per call - there the preprocess kernel checks the bounds and the verdict comes back with
            # num_rendered (same messages, raised as RuntimeError)
            assert (batch_patch_min[:, 0] + patch_width <= self.width).all(), "Some b_patch_max_x exceed self.width"
            assert (batch_patch_min[:, 1] + patch_height <= self.height).all(), "Some b_patch_max_y exceed self.height"
        f32 = torch.float32
        dev = verts.device
        return FusedRenderFunction.apply(
            background.to(dtype=f32), batch_patch_min.to(device=dev, dtype=torch.int32), patch_width, patch_height,
            verts.to(dtype=f32), faces.to(dtype=torch.int32), verts_color.to(dtype=f32), faces_opacity.to(dtype=f32),
            self.mv.index_select(0, view_idx).to(dtype=f32), self.proj.index_select(0, view_idx).to(dtype=f32),
            faces_intense.to(dtype=f32), aa_temperature, self.ray_o, self.ray_d, ray_view, self.aa_grad_buffer_size,
            True, getattr(self, "grad_sink", None), self._ray_spec())

    def prepare_inputs(self, batch_mvp_idx, batch_patch_min, patch_width, patch_height, verts, faces, verts_color,
                       faces_opacity, faces_intense, background, aa_temperature=1.0):
        """The 21 positional inputs of ``RenderFunction`` for this call (reference :325-376)."""
        n_view = len(batch_mvp_idx)
        n_face = faces.shape[0]
        verts_ndc, verts_image = self.compute_verts_ndc_image(verts, self.mv[batch_mvp_idx], self.proj[batch_mvp_idx])

        corners = verts_image[:, faces.flatten()].view(-1, 3, 2)                  # [B*F, 3, 2]
        tri = Triangles(corners[:, 0], corners[:, 1], corners[:, 2])
        shape4 = (n_view, n_face, 3, 2)

        ray_o, ray_d = self.select_rays(batch_mvp_idx, batch_patch_min, patch_width, patch_height)

        f32 = torch.float32
        return (
            background.to(dtype=f32),
            batch_patch_min.to(dtype=torch.int32), patch_width, patch_height,
            verts.to(dtype=f32), faces.to(dtype=torch.int32), verts_color.to(dtype=f32), faces_opacity.to(dtype=f32),
            verts_ndc.to(dtype=f32), verts_image.to(dtype=f32), faces_intense.to(dtype=f32),
            aa_temperature,
            tri.verts.reshape(shape4).to(dtype=f32),
            tri.edges.reshape(shape4).to(dtype=f32),
            tri.edges_iszero.reshape(shape4).to(dtype=torch.bool),
            tri.edges_recip.reshape(shape4).to(dtype=f32),
            tri.edges_normal.reshape(shape4).to(dtype=f32),
            tri.edges_normal_c.reshape((n_view, n_face, 3)).to(dtype=f32),
            self.aa_grad_buffer_size,
            ray_o.to(dtype=f32), ray_d.to(dtype=f32))


class LayeredRenderer(Renderer):
    """Per-ray, strictly depth-ordered layers of existing faces through a tetrahedral mesh.

    Like the reference (dmesh2_renderer/__init__.py:388-402) the constructor does NOT initialise
    ``torch.nn.Module``; the object only carries attributes and ``generate`` is not differentiable.
    """

    def __init__(self, mv, proj, width, height, device):
        self._setup(mv, proj, width, height, device)

    def generate(self,
                 batch_mvp_idx: List[int],          # [B,]
                 verts: torch.Tensor,               # [P, 3]
                 faces: torch.Tensor,               # [F, 3]
                 tets: torch.Tensor,                # [T, 4]
                 face_tets: torch.Tensor,           # [F, 2]  (-1 = no tet on that side)
                 tet_faces: torch.Tensor,           # [T, 4]
                 faces_existence: torch.Tensor,     # [F,]
                 num_layers: int):
        f32, i32 = torch.float32, torch.int32
        if getattr(self, "fused", True) and verts.is_cuda:
            # fused host prep (as Renderer._forward_fused; generate() is not differentiable, so cameras that
            # require grad change nothing here): projection inside the preprocess kernel, rays read
            # in place from the full-resolution buffers - bit-identical layer ids, no torch prep kernels
            key, ray_view, view_idx = self._view_ids(batch_mvp_idx)
            # Incremental re-generation ("fixed point positions" in DMesh++'s layered stage): when
            # the geometry tensors, cameras and views are the SAME OBJECTS at the SAME VERSION as in the previous
            # call, only face_existence can have changed - binning and first hits are reused and the tet walk
            # alone is redone.  In-place updates bump torch's version counter and invalidate the entry.  The
            # cache entry holds STRONG references to the geometry tensors and compares with `is`: a freed
            # temporary's address (and version 0) can be handed to the next temporary by the caching allocator,
            # an address of a tensor that is still alive cannot.
            geo = (verts, faces, tets, face_tets, tet_faces, self.mv, self.proj) + \
                ((self.ray_o, self.ray_d) if self.ray_o is not None else ())
            sig = (key, self.width, self.height)
            vers = tuple(t._version for t in geo)
            prev = self.__dict__.get("_layer_cache") if getattr(self, "incremental", True) else None
            reuse = None
            if (prev is not None and prev["sig"] == sig and prev["versions"] == vers
                    and all(a is b for a, b in zip(prev["geo"], geo))):
                reuse = prev["taps"]
            layers, cnt, taps = _C.generate_render_layers_cuda(
                self.width, self.height,
                verts.to(dtype=f32), faces.to(dtype=i32), tets.to(dtype=i32), face_tets.to(dtype=i32),
                tet_faces.to(dtype=i32), faces_existence.to(dtype=i32), None, None, self.ray_o, self.ray_d, num_layers,
                fused=(self.mv.index_select(0, view_idx).to(dtype=f32), self.proj.index_select(0, view_idx).to(dtype=f32),
                       ray_view), return_taps=True, reuse=reuse, ray_spec=self._ray_spec(),
                overlap_geometry=getattr(self, "overlap_geometry", True))
            self.__dict__["_layer_cache"] = dict(sig=sig, versions=vers, geo=geo, taps=taps)
            return layers, cnt
        verts_ndc, verts_image = self.compute_verts_ndc_image(verts, self.mv[batch_mvp_idx], self.proj[batch_mvp_idx])
        ray_o, ray_d = self.ray_o[batch_mvp_idx], self.ray_d[batch_mvp_idx]
        return _C.generate_render_layers_cuda(
            self.width, self.height,
            verts.to(dtype=f32), faces.to(dtype=i32), tets.to(dtype=i32), face_tets.to(dtype=i32),
            tet_faces.to(dtype=i32), faces_existence.to(dtype=i32),
            verts_ndc.to(dtype=f32), verts_image.to(dtype=f32), ray_o.to(dtype=f32), ray_d.to(dtype=f32),
            num_layers)
