/*
 * dm2_renderer.h — C-ABI of libdm2_b200.so, the B200-native (sm_100a) replacement for the
 * `dmesh2_renderer._C` torch extension of SonSang/dmesh2_renderer.
 *
 * Drop-in boundary.  The reference binds three pybind functions (reference ext.cpp:6-9):
 *     render_forward_cuda          (reference render.h:12-45,  render.cu:28-195)
 *     render_backward_cuda         (reference render.h:47-94,  render.cu:198-373)
 *     generate_render_layers_cuda  (reference render.h:101-119, render.cu:378-476)
 * called only from RenderFunction.forward/backward and LayeredRenderer.generate
 * (reference dmesh2_renderer/__init__.py:78, :170, :432).  This header declares the plain-C
 * entry points that replace them: raw device pointers, explicit sizes, an explicit CUDA
 * stream, no torch types.  The host side (dmesh2_renderer_b200/) binds them with ctypes
 * and keeps the reference's Python signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a densely packed (contiguous) array unless the
 *     name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); nothing in the
 *     library calls cudaDeviceSynchronize.  dm2_*_preprocess waits on an event behind the 12-byte
 *     read-back of (num_rendered, input flags, longest tile segment) - the reference performs the
 *     num_rendered one too, with a blocking copy (reference cuda_impl/renderer.cu:174-179);
 *     dm2_fused_forward_pipelined enqueues the WHOLE forward pass before it waits for that event,
 *     so the GPU never idles on the host;
 *   - workspaces are caller-allocated byte buffers (torch caching allocator on the host
 *     side); their sizes come from the dm2_*_bytes functions and their internal layout from
 *     DM2FaceLayout / DM2ImageLayout / DM2BinningLayout (used by the parity tests as the
 *     "debug taps" of SURVEY.md §8b.5);
 *   - every function returns 0 on success; on failure it returns non-zero and
 *     dm2_last_error() returns a thread-local message (the Python wrapper raises
 *     RuntimeError, mirroring AT_ERROR / CHECK_CUDA in the reference);
 *   - gradient outputs of dm2_render_backward must be ZEROED by the caller (the reference
 *     uses torch::zeros_like, render.cu:313-318); the kernels accumulate with atomics.
 */
#ifndef DM2_RENDERER_H_INCLUDED
#define DM2_RENDERER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM2_TILE 16              /* BLOCK_X = BLOCK_Y of reference cuda_impl/config.h:4-5 */
#define DM2_ABI_VERSION 17

/* Inputs of RenderFunction (reference dmesh2_renderer/__init__.py:13-47). */
typedef struct DM2RenderInputs {
    int32_t B, P, F;             /* views in this call, vertices, faces                     */
    int32_t H, W;                /* patch_height, patch_width                               */
    float   aa_temperature;      /* in [0,1]                                                */
    int32_t len_oarea_buffer;    /* K; accepted for parity (the host side sizes the AA cache)  */
    const float*   background;               /* (3)            */
    const int32_t* patch_min;                /* (B,2)          */
    const float*   verts;                    /* (P,3)          */
    const int32_t* faces;                    /* (F,3)          */
    const float*   verts_color;              /* (P,3)          */
    const float*   faces_opacity;            /* (F)            */
    const float*   verts_ndc;                /* (B,P,3)        */
    const float*   verts_image;              /* (B,P,2)        */
    const float*   faces_intense;            /* (B,F)          */
    const float*   aa_face_verts;            /* (B,F,3,2)      */
    const float*   aa_face_edges;            /* (B,F,3,2)      */
    const uint8_t* aa_face_edges_iszero;     /* (B,F,3,2) bool */
    const float*   aa_face_edges_recip;      /* (B,F,3,2)      */
    const float*   aa_face_edges_normal;     /* (B,F,3,2)      */
    const float*   aa_face_edges_normal_c;   /* (B,F,3)        */
    const float*   image_ray_o;              /* (B,H,W,3)      */
    const float*   image_ray_d;              /* (B,H,W,3)      */
    /* Optional ray indirection (NULL/0 for the plain RenderFunction path): when ray_view != NULL the
     * ray of patch pixel (x,y) of batch entry b is  image_ray_*[ray_view[b]][y + patch_min.y][x + patch_min.x]
     * inside a (V, ray_h, ray_w, 3) buffer.                                                      */
    const int32_t* ray_view;
    int32_t ray_h, ray_w;
    /* != 0: out_depth receives the depth the reference's Renderer.forward returns,
     * 1 - (depth_ndc + 1) / 2 (reference __init__.py:377-378; same IEEE operations as the three torch
     * ops, so bit-identical), and dL_dout_depth is the gradient w.r.t. THAT value (the backward pass
     * applies the exact factor -1/2).  0: out_depth = depth_ndc, as render_forward_cuda returns it.   */
    int32_t depth_unit;
    /* In-kernel ray generation: ray_recipe > 0 and ray_params != NULL - the kernels
     * compute the ray of frame pixel (x, y) of view ray_view[b] from ray_params[view] = [inv(proj) | inv(mv)]
     * (2 x 16 floats, row-major) with the arithmetic of recipe `ray_recipe` (dm2_generate_rays) instead of
     * reading image_ray_o / image_ray_d (which may then be NULL).  Needs ray_view / ray_h / ray_w.              */
    int32_t ray_recipe;
    const float* ray_params;
    int32_t ray_v;               /* number of views V behind ray_params / the ray buffers */
    int32_t reserved_;
} DM2RenderInputs;

/* Gradient outputs of render_backward_cuda (reference render.cu:313-318, 372). */
typedef struct DM2RenderGrads {
    float* dL_dverts;            /* (P,3)      */
    float* dL_dverts_color;      /* (P,3)      */
    float* dL_dfaces_opacity;    /* (F)        */
    float* dL_dverts_ndc;        /* (B,P,3) — only channel 2 is written */
    float* dL_dfaces_intense;    /* (B,F)      */
    float* dL_daa_face_verts;    /* (B,F,3,2)  */
} DM2RenderGrads;

/* Inputs of generate_render_layers_cuda (reference render.h:101-119). */
typedef struct DM2LayerInputs {
    int32_t B, P, F, T;
    int32_t H, W;                /* full frame; patch_min is 0 (reference renderer.cu:557-558) */
    int32_t num_layers;          /* L */
    const float*   verts;                    /* (P,3)   */
    const int32_t* faces;                    /* (F,3)   */
    const int32_t* tets;                     /* (T,4)   */
    const int32_t* face_tets;                /* (F,2), -1 = none */
    const int32_t* tet_faces;                /* (T,4)   */
    const int32_t* face_existence;           /* (F)     */
    const float*   verts_ndc;                /* (B,P,3) */
    const float*   verts_image;              /* (B,P,2) */
    const float*   image_ray_o;              /* (B,H,W,3) */
    const float*   image_ray_d;              /* (B,H,W,3) */
    /* Optional fused host prep (NULL for the plain generate_render_layers_cuda path), as for the
     * Renderer: with mv / proj != NULL the preprocess kernel projects the vertices itself (same
     * fma chain and IEEE operations as compute_verts_ndc_image, reference __init__.py:239-262;
     * verts_ndc / verts_image may then be NULL), and with ray_view != NULL image_ray_* are the
     * renderer's full (V,H,W,3) buffers and batch entry b uses row ray_view[b].               */
    const float*   mv;                       /* (B,4,4) */
    const float*   proj;                     /* (B,4,4) */
    const int32_t* ray_view;                 /* (B)     */
    const float*   ray_params;               /* (V,32), as DM2RenderInputs.ray_params (needs ray_view)   */
    int32_t ray_recipe;                      /* as DM2RenderInputs.ray_recipe                            */
    int32_t ray_v;                           /* V                                                        */
} DM2LayerInputs;

/* Byte offsets inside the three workspaces (all 256-byte aligned).                         */
typedef struct DM2FaceLayout {        /* per (view,face) state; reference FaceState, state.h:19-30 */
    size_t depths;               /* f32[B*F]  clamp((mean z+1)/2,0,1); uninitialised if culled  */
    size_t min_depths;           /* f32[B*F]                                                    */
    size_t max_depths;           /* f32[B*F]                                                    */
    size_t tiles_touched;        /* u32[B*F]                                                    */
    size_t rects;                /* uint2[B*F] tile bbox: x = x0 | y0<<16, y = x1 | y1<<16      */
    size_t records;              /* packed per-(view,face) records staged by the tile kernels   */
    size_t record_stride;        /* bytes per record                                            */
    size_t total;
} DM2FaceLayout;

typedef struct DM2ImageLayout {       /* reference ImageState / ImageRenderLayerState, state.h:32-61 */
    size_t ranges;               /* uint2[B*tiles] (x = first, y = last+1; (0,0) if empty)      */
    size_t tile_cursor;          /* u32[B*tiles+3] per-tile instance counts, then scatter cursors;
                                  *               then num_rendered, the patch-bound flags and the
                                  *               longest tile segment                           */
    size_t final_T;              /* f32[B*H*W]   (render)                                       */
    size_t final_prev_T;         /* f32[B*H*W]   (render)                                       */
    size_t n_contrib;            /* u32[B*H*W]   (render)                                       */
    size_t first_face;           /* i32[B*H*W]   (layers)                                       */
    size_t first_tet;            /* i32[B*H*W]   (layers)                                       */
    size_t total;
} DM2ImageLayout;

typedef struct DM2BinningLayout {     /* reference BinningState, state.h:41-51 */
    size_t keys_unsorted;        /* u64[R]  (tile + tiles*view) << 32 | float bits of depth, grouped
                                  *         by tile, unsorted inside a tile (scatter order)      */
    size_t vals_unsorted;        /* u32[R]  face id within the view, same order                 */
    size_t keys;                 /* u64[R]  sorted                                              */
    size_t vals;                 /* u32[R]  sorted  (= reference face_list)                     */
    size_t sort_temp;            /* zero-sized (the tile-segmented sort needs no scratch; kept for layout stability) */
    size_t imask;                /* u8[R]   render path: bit w = the face touches a pixel of the
                                  *         8x4-pixel footprint w of its tile (x = (w&1)*8, y = (w>>1)*4) */
    size_t meta;                 /* uint4[8][R] render path: per (footprint, instance) touched-pixel mask
                                  *         and 9x5 inside-corner grid (reference aa.h:96-149 evaluated once) */
    size_t total;
} DM2BinningLayout;

const char* dm2_last_error(void);
int dm2_abi_version(void);

/* ---- workspace sizing ------------------------------------------------------------------ */
int dm2_face_layout   (int32_t B, int32_t F, int32_t layered, DM2FaceLayout* out);
int dm2_image_layout  (int32_t B, int32_t H, int32_t W, int32_t layered, DM2ImageLayout* out);
int dm2_binning_layout(int64_t num_rendered, DM2BinningLayout* out);

/* ---- Renderer: replaces render_forward_cuda / render_backward_cuda ---------------------- */
/* Stage 1 (reference renderer.cu:150-179): per-face depth / tile bbox / tiles_touched, the
 * packed face records, the per-tile instance counts and their scan (= the tile ranges), and
 * the read-back of num_rendered and of the longest tile segment (waits on an event recorded on
 * `stream`; longest_host may be NULL).  Fails (message in dm2_last_error) when a face references a
 * vertex id outside [0, P) - the reference's torch indexing raises for that, raw pointers cannot. */
int dm2_render_preprocess(const DM2RenderInputs* in, void* face_ws, void* image_ws,
                          int64_t* num_rendered_host, int64_t* longest_host, void* stream);
/* Stage 2 (reference renderer.cu:181-266): key emission straight into the tile segments, per-
 * segment sort by (depth bits, face id) - together the reference's stable radix sort on the low
 * 32+bit key bits -, the per-instance pixel classification, then the forward tile compositor.
 * out_color (B,H,W,3) and out_depth (B,H,W) are fully overwritten.
 * capacity: the instance count binning_ws was laid out for by dm2_binning_layout (>= num_rendered; the
 * actual count is read on the device); longest: what stage 1 returned, or -1 (= run every sort tier). */
int dm2_render_forward(const DM2RenderInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                       int64_t capacity, int64_t longest, float* out_color, float* out_depth,
                       void* shade_cache, int64_t shade_cache_bytes, void* stream);
/* Optional shading cache (the role of the reference's four oarea buffers, render.cu:143-146 - a zero-
 * filled B*H*W*K*32-byte array there - as a compact pool): when shade_cache != NULL (shade_cache_bytes
 * >= 512, a multiple of 16) the forward pass leaves, for every (instance, 8x4-pixel footprint) it stages,
 * one 16-byte record (clamped barycentrics, coverage ratio, clamp code) per touched pixel and
 * d(area)/d(aa verts) (24 bytes) per partially covered pixel; the backward pass reads them back instead of
 * re-running the hit point, the clamp and the polygon clip, and falls back to recomputation for blocks
 * that did not fit.  The first 4 bytes hold the number of 16-byte granules the forward pass asked for
 * (may exceed the capacity: size the next call's pool from it).  The same buffer and size must be passed
 * to the matching backward call.  No initialisation needed; other contents are private.               */
/* reference renderer.cu:271-392 / backward.cu:17-532.  grad_ws: scratch of
 * dm2_backward_workspace_bytes(B, F) bytes (per-(view,face) gradient records; zeroed inside).
 * capacity: as passed to the forward call.                                                     */
int dm2_backward_workspace_bytes(int32_t B, int32_t F, size_t* bytes);
int dm2_render_backward(const DM2RenderInputs* in, const void* face_ws, const void* binning_ws,
                        const void* image_ws, int64_t capacity,
                        const float* dL_dout_color, const float* dL_dout_depth,
                        const DM2RenderGrads* grads, void* grad_ws, const void* shade_cache,
                        int64_t shade_cache_bytes, void* stream);

/* ---- LayeredRenderer: replaces generate_render_layers_cuda ------------------------------- */
int dm2_layers_preprocess(const DM2LayerInputs* in, void* face_ws, void* image_ws,
                          int64_t* num_rendered_host, int64_t* longest_host, void* stream);
/* render_layers (B,H,W,L) must be pre-filled with -1 and render_layers_cnt (B,H,W) with 0 by
 * the caller (reference render.cu:437-438); geo_ws needs dm2_layers_geo_bytes(F,T) bytes (64 bytes per face +
 * 256 bytes per tet: the ray-independent geometry and adjacency records of the walk, built once per geometry). */
int dm2_layers_geo_bytes(int32_t F, int32_t T, size_t* bytes);
int dm2_generate_render_layers(const DM2LayerInputs* in, void* face_ws, void* binning_ws,
                               void* image_ws, void* geo_ws, int64_t num_rendered, int64_t longest,
                               int32_t* render_layers, int32_t* render_layers_cnt, void* stream);
/* The same pair with the geometry / link records of geo_ws built INSIDE the preprocess call, behind the header
 * read-back: they depend on the geometry tensors only, so the GPU computes them while the host waits for
 * num_rendered (the reference idles there, renderer.cu:571-580).  dm2_generate_render_layers_prepared then runs the
 * binning, the first hit and the walk on the records it finds. */
int dm2_layers_preprocess_with_geometry(const DM2LayerInputs* in, void* face_ws, void* image_ws, void* geo_ws,
                                        int64_t* num_rendered_host, int64_t* longest_host, void* stream);
int dm2_generate_render_layers_prepared(const DM2LayerInputs* in, void* face_ws, void* binning_ws,
                                        void* image_ws, void* geo_ws, int64_t num_rendered, int64_t longest,
                                        int32_t* render_layers, int32_t* render_layers_cnt, void* stream);

/* Incremental re-generation: when ONLY face_existence changed since a
 * dm2_generate_render_layers call on the same image_ws / geo_ws (same geometry, views and rays),
 * the binning, every pixel's first hit and entry tet are still valid: this entry point refreshes the
 * per-face existence flags and redoes the tet walk alone.  render_layers / render_layers_cnt must be
 * pre-filled as above.  The caller is responsible for the "only face_existence changed" condition. */
int dm2_regenerate_render_layers(const DM2LayerInputs* in, void* image_ws, void* geo_ws,
                                 int32_t* render_layers, int32_t* render_layers_cnt, void* stream);

/* ---- layer consumer ----------------------------------------------------------------
 * The reference's LayeredRenderer returns face ids only (reference dmesh2_renderer/__init__.py:404-451); DMesh++
 * then gathers per-layer face attributes and composites them in torch (L gathers + a cumulative product).  One
 * kernel per direction: front-to-back compositing of the first min(cnt, L) recorded faces of every ray,
 *     C = sum_l T_l a_l c_l + T_L background,  T_{l+1} = T_l (1 - a_l),  alpha = 1 - T_L,
 * a_l = faces_opacity[f_l], c_l = faces_color[f_l] (* faces_intense[b, f_l] if faces_intense != NULL).
 * Backward: dL/dfaces_color (F,3), dL/dfaces_opacity (F) are fully overwritten; dL/dfaces_intense (B,F) is
 * accumulated into a caller-zeroed buffer (NULL: not wanted); grad_ws: 16 * F bytes of scratch; L <= 32.       */
int dm2_layers_composite_forward(int32_t B, int32_t H, int32_t W, int32_t L, int32_t F, const int32_t* render_layers,
                                 const int32_t* render_layers_cnt, const float* faces_color, const float* faces_opacity,
                                 const float* faces_intense, const float* background, float* out_color, float* out_alpha,
                                 void* stream);
int dm2_layers_composite_backward(int32_t B, int32_t H, int32_t W, int32_t L, int32_t F, const int32_t* render_layers,
                                  const int32_t* render_layers_cnt, const float* faces_color, const float* faces_opacity,
                                  const float* faces_intense, const float* background, const float* dL_dout_color,
                                  const float* dL_dout_alpha, void* grad_ws, float* dL_dfaces_color,
                                  float* dL_dfaces_opacity, float* dL_dfaces_intense, void* stream);

/* ---- Renderer, fused host prep ----------------------------------------
 * Renderer.forward of the reference spends ~25 torch kernels per call on host-side prep
 * (reference dmesh2_renderer/__init__.py:239-262, 335-347 and pyrenderer.py:6-30): the w clamp and
 * perspective divide, the image transform, the per-face gather, the CCW re-ordering and the five
 * derived AA tensors, plus the ray patch gather - and autograd replays all of it backwards.  The
 * fused entry points take verts and the selected mv / proj matrices and perform every step inside
 * the preprocess kernel with the same IEEE operations in the same order: the two K=4 matmuls as the
 * fma chain cuBLAS runs for them on sm_100 (k = 0..3, first term a plain product; verified
 * bit-for-bit by tools/probe_matmul_rounding.py), then clamp / divide / image transform / CCW /
 * AA terms.  The packed records, tile bboxes and depth keys are bit-identical to the unfused path;
 * the backward scatter kernel applies the matching chain rule down to d/d verts.
 * Rays are read straight from the renderer's full-resolution ray buffers.                          */
typedef struct DM2FusedInputs {
    int32_t B, P, F;             /* views in this call, vertices, faces                            */
    int32_t H, W;                /* patch_height, patch_width                                      */
    int32_t image_h, image_w;    /* full frame size used by the NDC -> image transform             */
    float   aa_temperature;
    const float*   background;   /* (3)                                                            */
    const int32_t* patch_min;    /* (B,2)                                                          */
    const float*   verts;        /* (P,3)                                                          */
    const int32_t* faces;        /* (F,3)                                                          */
    const float*   verts_color;  /* (P,3)                                                          */
    const float*   faces_opacity;/* (F)                                                            */
    const float*   mv;           /* (B,4,4) model-view of each batch entry, row-major              */
    const float*   proj;         /* (B,4,4) projection of each batch entry, row-major              */
    const float*   faces_intense;/* (B,F)                                                          */
    const float*   ray_o;        /* (V,image_h,image_w,3) full-resolution rays of ALL views        */
    const float*   ray_d;
    const int32_t* ray_view;     /* (B) row of ray_o/ray_d used by batch entry b                   */
    int32_t depth_unit;          /* as DM2RenderInputs.depth_unit                                  */
    int32_t ray_recipe;          /* as DM2RenderInputs.ray_recipe (0: read ray_o / ray_d)          */
    const float* ray_params;     /* (V,32) [inv(proj) | inv(mv)] per view, or NULL                 */
    int32_t ray_v;               /* V                                                              */
    int32_t reserved_;
} DM2FusedInputs;

typedef struct DM2FusedGrads {   /* caller-zeroed */
    float* dL_dverts;            /* (P,3)   hit point + depth (z/w) + AA coverage (x/w, y/w) paths */
    float* dL_dverts_color;      /* (P,3)                                                          */
    float* dL_dfaces_opacity;    /* (F)                                                            */
    float* dL_dfaces_intense;    /* (B,F)                                                          */
} DM2FusedGrads;

int dm2_fused_preprocess(const DM2FusedInputs* in, void* face_ws, void* image_ws, int64_t* num_rendered_host,
                         int64_t* longest_host, void* stream);
int dm2_fused_forward(const DM2FusedInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                      int64_t capacity, int64_t longest, float* out_color, float* out_depth, void* shade_cache,
                      int64_t shade_cache_bytes, void* stream);
/* dm2_fused_preprocess + dm2_fused_forward in ONE call without a host wait between them (the reference
 * blocks on a device->host copy of num_rendered before it can size the binning scratch, reference
 * cuda_impl/renderer.cu:174-183).  The caller sizes binning_ws for `capacity` instances from the previous
 * frame of the same problem (and passes that frame's longest tile segment, or -1, as longest_hint - the
 * library covers 25 % more than the hint); all
 * kernels are enqueued at once - those behind the tile scan compare num_rendered with `capacity` ON THE
 * DEVICE and return immediately if the frame does not fit - and only then the call waits for the 12-byte
 * header.  *fits_host = 1: the outputs are valid.  *fits_host = 0: *num_rendered_host > capacity; the caller
 * allocates a workspace for *num_rendered_host instances and calls dm2_fused_forward (the preprocess results
 * in face_ws / image_ws stay valid).  A longest segment beyond the hinted sort tier is repaired inside.    */
int dm2_fused_forward_pipelined(const DM2FusedInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                                int64_t capacity, int64_t longest_hint, float* out_color, float* out_depth,
                                void* shade_cache, int64_t shade_cache_bytes, int64_t* num_rendered_host,
                                int64_t* longest_host, int32_t* fits_host, void* stream);
int dm2_fused_backward(const DM2FusedInputs* in, const void* face_ws, const void* binning_ws,
                       const void* image_ws, int64_t capacity, const float* dL_dout_color,
                       const float* dL_dout_depth, const DM2FusedGrads* grads, void* grad_ws,
                       const void* shade_cache, int64_t shade_cache_bytes, void* stream);

/* ---- pixel rays -------------------------------------------------------------
 * The reference keeps one origin and one direction per pixel of EVERY view resident (Renderer._init_rays,
 * reference dmesh2_renderer/__init__.py:198-237: 24 bytes per pixel and view) and gathers patches of them per
 * call (:264-302).  A ray is a function of 32 floats per view: dm2_generate_rays evaluates it - pixel centre ->
 * NDC -> inv(proj) -> inv(mv) -> minus the camera position -> normalise with the reference's `+ 1e-6` - and the
 * tile kernels can do the same in their prologue.  The torch ops of _init_rays fix the ROUNDING only up to
 * choices the libraries make (division by the frame size as a multiplication by its reciprocal on CUDA, the
 * accumulation order of the two K = 4 products in cuBLAS - which differs for the last (V H W) mod 65535 pixels,
 * a separate small batch -, of the squared norm in the reduction kernel), so the
 * arithmetic is parameterised by a recipe id in [1, DM2_RAY_RECIPES]; the host generates the rays with every
 * recipe once at construction time and only uses in-kernel rays with a recipe whose output is BIT-IDENTICAL to
 * the tensors torch produced (dmesh2_renderer_b200/renderer.py).  out_o / out_d: (V, H, W, 3).              */
#define DM2_RAY_RECIPES 36
int dm2_generate_rays(const float* ray_params, int32_t V, int32_t H, int32_t W, int32_t recipe, float* out_o,
                      float* out_d, void* stream);

/* ---- multi-view optimisation step: the gradient exchange ----------------
 * The reference is single-GPU; DMesh++'s multi-view step shards views over the GPUs of one box and
 * sums the replicated-parameter gradients [dverts | dverts_color | dfaces_opacity] once per step.
 * In-place sum over `world` ranks of an n_floats-long fp32 bucket (n_floats % 4 == 0) that every rank
 * holds at the same offset of a symmetric peer-mapped allocation: peer_ptrs[q] = rank q's bucket as
 * mapped into THIS process (host array of `world` device pointers), multicast_ptr = the bucket through
 * the NVSwitch multicast mapping, or NULL.  Rank r reduces slice r and writes the sum into every rank's
 * bucket: with a multicast mapping through multimem.ld_reduce / multimem.st (the add happens inside the
 * switch), else with peer loads in rank order + peer stores.  The CALLER provides the two cross-GPU
 * barriers around the call (all buckets complete before; all stores landed after) - see
 * dmesh2_renderer_b200/multiview.py:PeerGradBucket.                                                   */
int dm2_peer_allreduce_sum(const void* const* peer_ptrs, void* multicast_ptr, int32_t world, int32_t rank,
                           int64_t n_floats, void* stream);
/* The same exchange with BOTH cross-GPU barriers inside the kernel (one launch per optimisation step instead of
 * three): peer_flag_ptrs[q] = rank q's flag area as mapped into this process - DM2_PEER_FLAG_WORDS u32 words of
 * the same symmetric allocation, zeroed once before the first call (and made visible to all ranks by one
 * barrier) -, epoch = 1, 2, 3, ... the same on every rank for the same step.  Ranks signal arrival / departure
 * with st.release.sys into their peers' flag areas and spin on their own with ld.acquire.sys (bounded: a peer
 * that never arrives makes the kernel trap after ~20 s instead of hanging the job).                          */
#define DM2_PEER_FLAG_WORDS 64
int dm2_peer_allreduce_sum_fused(const void* const* peer_ptrs, void* multicast_ptr, const void* const* peer_flag_ptrs,
                                 int32_t world, int32_t rank, int64_t n_floats, uint32_t epoch, void* stream);

/* ---- binning primitives exposed on their own (sort keys/s metric, parity tests) ---------- */
/* Stable LSD radix sort of (u64 key, u32 value) pairs on key bits [0, end_bit): a stand-alone
 * equivalent of cub::DeviceRadixSort::SortPairs (reference renderer.cu:202-207), kept as a
 * primitive and as the cross-check of the pipeline's tile-segmented sort.  temp: sort_temp area. */
int dm2_sort_pairs(const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                   uint32_t* vals_out, int64_t n, int32_t end_bit, void* temp, size_t temp_bytes,
                   void* stream);
int dm2_sort_temp_bytes(int64_t n, size_t* bytes);
/* Inclusive u32 prefix sum; replaces cub::DeviceScan::InclusiveSum (renderer.cu:165-171).   */
int dm2_inclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* partials,
                           void* stream);
int dm2_scan_partials_count(int64_t n, int64_t* count);
/* getHigherMsb of reference renderer.cu:396-411.                                            */
uint32_t dm2_higher_msb(uint32_t n);

/* number of kernels launched by this library on this thread since the last reset
 * (bench.py's `gpu_launches`).                                                              */
int64_t dm2_launch_count(int32_t reset);

/* Optional per-stage device timing: CUDA events recorded on the launching stream around each
 * stage (preprocess_pack, tile_count, tile_scan, scatter_instances, tile_sort, classify, raster_forward,
 * raster_backward, ...).  dm2_profile_report fills buf with "name total_ms count" lines
 * (it waits for the recorded events) and returns the buffer size needed.                      */
void dm2_profile_enable(int32_t on);
/* Deep scenes: when the longest tile list exceeds `threshold` instances, classification + forward compositing run
 * in rounds - list entries [0, round_len), [round_len, 2 round_len), [2 round_len, 4 round_len), ... of every tile
 * (round_len a multiple of 256) - and tiles whose pixels have all terminated are not classified any further.  The
 * reference classifies nothing ahead of time (forward.cu:139-432); this is where its early termination is recovered
 * for the classification pre-pass.  Results do not depend on these values.  Defaults 512 / 6144; process-wide; not
 * thread safe against running calls.
 * A frame that runs in rounds also sorts its tile segments lazily: list positions [0, 2 round_len) of every tile at once,
 * the rest of a tile's segment only if a round reaches it - behind a tile's sorted prefix the `keys` / `vals` arrays of the
 * binning workspace are binned by depth but not sorted (the reference sorts everything, renderer.cu:202-207; no pixel
 * reads those entries).  DM2_LAZY_SORT=0 in the environment sorts everything. */
void dm2_set_forward_rounds(int64_t round_len, int64_t threshold);
/* Overlap area of a partially covered (pixel, face) pair (reference aa.h:151-441).  Default (0): the polygon-free
 * evaluation of dm2_aa_fast.cuh wherever the pair is far - by thresholds measured against the oracle, tests/test_aa_fast.py - from every
 * rounding-dependent decision of the reference, so that it yields the reference's value (same rounded polygon vertices, area
 * equal to ~1e-7), the reference's polygon construction for the other ~0.5 % of the pairs.  1: the polygon construction
 * for every pair - forward images then are bit-identical to the reference's (tests), at +0.46 ms per 1 M-triangle frame.
 * Process-wide; DM2_EXACT_CLIP=1 in the environment sets the initial value. */
void dm2_set_exact_clip(int32_t on);
int32_t dm2_get_exact_clip(void);
void dm2_profile_reset(void);
int64_t dm2_profile_report(char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DM2_RENDERER_H_INCLUDED */
