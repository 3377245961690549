// dm2_binning.cu — per-face preprocess + record packing, the per-instance pixel classification
// pre-pass, and the two stand-alone primitives (inclusive scan, stable LSD radix sort) that replace
// the reference's CUB calls.  Replaces reference cuda_impl/forward.cu:16-137 (preprocessFaceCUDA);
// the tile binning itself (reference renderer.cu:165-218, 415-492) lives in dm2_tilesort.cu.
#include "dm2_common.cuh"
#include "dm2_aa_fast.cuh"

namespace dm2 {

// torch.min / torch.max over the 3 triangle vertices propagate NaN (reference forward.cu:480-481)
__device__ __forceinline__ float nanmin(float a, float b) { return (a != a || a < b) ? a : b; }
__device__ __forceinline__ float nanmax(float a, float b) { return (a != a || a > b) ? a : b; }

// ------------------------------------------------------------------ preprocess (shared part)
struct PreOut { bool live; int x0, y0, x1, y1; float depth, dmin, dmax; };

__device__ __forceinline__ PreOut preprocess_face(int b, int P, int v0, int v1, int v2,
                                                  const float* __restrict__ verts_ndc,
                                                  const float* __restrict__ verts_image,
                                                  uint32_t pmx, uint32_t pmy, int gx, int gy) {
    // reference forward.cu:41-107.  SASS of the reference shows no FMA in this chain.
    PreOut o; o.live = false; o.x0 = o.y0 = o.x1 = o.y1 = 0; o.depth = o.dmin = o.dmax = 0.f;
    const size_t vb = (size_t)b * P;
    const float z0 = verts_ndc[(vb + v0) * 3 + 2];
    const float z1 = verts_ndc[(vb + v1) * 3 + 2];
    const float z2 = verts_ndc[(vb + v2) * 3 + 2];
    const float2 i0 = *reinterpret_cast<const float2*>(verts_image + (vb + v0) * 2);
    const float2 i1 = *reinterpret_cast<const float2*>(verts_image + (vb + v1) * 2);
    const float2 i2 = *reinterpret_cast<const float2*>(verts_image + (vb + v2) * 2);
    float max_z = fmaxf(fmaxf(z0, z1), z2);
    float min_z = fminf(fminf(z0, z1), z2);
    float depth = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, z0), z1), z2);
    depth = __fdiv_rn(depth, 3.0f);
    if (max_z < -1.0f || min_z > 1.0f) return o;
    patch_rect(pmx, pmy, i0.x, i0.y, i1.x, i1.y, i2.x, i2.y, gx, gy, o.x0, o.y0, o.x1, o.y1);
    if ((uint32_t)(o.x1 - o.x0) * (uint32_t)(o.y1 - o.y0) == 0u) return o;
    o.live = true;
    o.depth = clamp01_depth(depth);
    o.dmin = clamp01_depth(min_z);
    o.dmax = clamp01_depth(max_z);
    return o;
}

// ------------------------------------------------------------------ coalesced record write-out
// Every thread builds its 240-byte record in shared memory (272-byte stride: conflict-free 16-byte
// accesses); the warp then writes its 32 records as contiguous 16-byte chunks - 512 bytes per store
// instruction instead of 32 chunks scattered at a 256-byte stride.
constexpr int kPreThreads = 256;
constexpr int kPreSlot = 17;                                    // float4 per staged record (15 used)
constexpr size_t kPreSmem = (size_t)kPreThreads * kPreSlot * sizeof(float4);

__device__ __forceinline__ void write_records_coalesced(const float4* __restrict__ s_warp, bool live, long long idx,
                                                        long long BF, float* __restrict__ records, int lane) {
    __syncwarp();
    const unsigned live_mask = __ballot_sync(0xffffffffu, live);
    const long long base = idx - lane;
    float4* out = reinterpret_cast<float4*>(records);
#pragma unroll 4
    for (int c = lane; c < 32 * 16; c += 32) {
        const int slot = c >> 4, chunk = c & 15;
        if (chunk < 15 && ((live_mask >> slot) & 1u) && base + slot < BF)
            out[(size_t)(base + slot) * (kRecWords / 4) + chunk] = s_warp[slot * kPreSlot + chunk];
    }
}

__device__ __forceinline__ bool render_preprocess_one(const DM2RenderInputs& in, int gx, int gy, long long idx,
                                                      float* __restrict__ depths, float* __restrict__ min_depths,
                                                      float* __restrict__ max_depths, uint32_t* __restrict__ tiles_touched,
                                                      uint2* __restrict__ rects, float4* __restrict__ rec,
                                                      uint32_t* __restrict__ in_flags);
__device__ __forceinline__ bool fused_preprocess_one(const DM2FusedInputs& in, int gx, int gy, long long idx,
                                                     float* __restrict__ depths, float* __restrict__ min_depths,
                                                     float* __restrict__ max_depths, uint32_t* __restrict__ tiles_touched,
                                                     uint2* __restrict__ rects, float4* __restrict__ rec,
                                                     uint32_t* __restrict__ patch_flags);

// ------------------------------------------------------------------ render preprocess + pack
__global__ void __launch_bounds__(kPreThreads)
k_render_preprocess(DM2RenderInputs in, int gx, int gy, float* __restrict__ depths,
                    float* __restrict__ min_depths, float* __restrict__ max_depths,
                    uint32_t* __restrict__ tiles_touched, uint2* __restrict__ rects,
                    float* __restrict__ records, uint32_t* __restrict__ flags) {
    extern __shared__ __align__(16) float4 s_pre[];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long BF = (long long)in.B * in.F;
    const int lane = threadIdx.x & 31;
    float4* const s_warp = s_pre + (threadIdx.x - lane) * kPreSlot;
    bool live = false;
    if (idx < BF) live = render_preprocess_one(in, gx, gy, idx, depths, min_depths, max_depths, tiles_touched, rects,
                                               s_warp + lane * kPreSlot, flags);
    write_records_coalesced(s_warp, live, idx, BF, records, lane);
}

__device__ __forceinline__ bool render_preprocess_one(const DM2RenderInputs& in, int gx, int gy, long long idx,
                                                      float* __restrict__ depths, float* __restrict__ min_depths,
                                                      float* __restrict__ max_depths, uint32_t* __restrict__ tiles_touched,
                                                      uint2* __restrict__ rects, float4* __restrict__ rec,
                                                      uint32_t* __restrict__ in_flags) {
    const int b = (int)(idx / in.F), f = (int)(idx % in.F);
    const int v0 = in.faces[3 * (size_t)f], v1 = in.faces[3 * (size_t)f + 1], v2 = in.faces[3 * (size_t)f + 2];
    // torch indexing raises for out-of-range vertex ids (the reference gathers through it); raw pointers cannot
    if ((unsigned)v0 >= (unsigned)in.P || (unsigned)v1 >= (unsigned)in.P || (unsigned)v2 >= (unsigned)in.P) {
        atomicOr(in_flags, kFlagVertexId);
        tiles_touched[idx] = 0; rects[idx] = make_uint2(0, 0);
        return false;
    }
    const uint32_t pmx = (uint32_t)in.patch_min[2 * b], pmy = (uint32_t)in.patch_min[2 * b + 1];
    const PreOut o = preprocess_face(b, in.P, v0, v1, v2, in.verts_ndc, in.verts_image, pmx, pmy, gx, gy);
    if (!o.live) { tiles_touched[idx] = 0; rects[idx] = make_uint2(0, 0); return false; }
    tiles_touched[idx] = (uint32_t)(o.y1 - o.y0) * (uint32_t)(o.x1 - o.x0);
    rects[idx] = make_uint2((uint32_t)o.x0 | ((uint32_t)o.y0 << 16), (uint32_t)o.x1 | ((uint32_t)o.y1 << 16));
    depths[idx] = o.depth; min_depths[idx] = o.dmin; max_depths[idx] = o.dmax;

    // ---- packed record (layout: dm2_common.cuh) ----
    const float* av = in.aa_face_verts + (size_t)idx * 6;
    const float* ae = in.aa_face_edges + (size_t)idx * 6;
    const float* ar = in.aa_face_edges_recip + (size_t)idx * 6;
    const float* an = in.aa_face_edges_normal + (size_t)idx * 6;
    const float* ac = in.aa_face_edges_normal_c + (size_t)idx * 3;
    const uint8_t* az = in.aa_face_edges_iszero + (size_t)idx * 6;
    float a[6], e[6], r[6], n[6];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float2 t;
        t = *reinterpret_cast<const float2*>(av + 2 * k); a[2 * k] = t.x; a[2 * k + 1] = t.y;
        t = *reinterpret_cast<const float2*>(ae + 2 * k); e[2 * k] = t.x; e[2 * k + 1] = t.y;
        t = *reinterpret_cast<const float2*>(ar + 2 * k); r[2 * k] = t.x; r[2 * k + 1] = t.y;
        t = *reinterpret_cast<const float2*>(an + 2 * k); n[2 * k] = t.x; n[2 * k + 1] = t.y;
    }
    const float c0 = ac[0], c1 = ac[1], c2 = ac[2];
    uint32_t flags = 0;
    bool on_grid = false;         // some screen coordinate is an integer: possible kink of the overlap area (dm2_aa.cuh)
#pragma unroll
    for (int k = 0; k < 6; k++) { flags |= (az[k] ? 1u : 0u) << k; on_grid |= a[k] == floorf(a[k]); }
    flags |= on_grid ? 0x200u : 0u;
    uint32_t safe_word = 0u;
    flags |= aaf_face_safety(a, e, r, flags, &safe_word);
    const float txmin = nanmin(nanmin(a[0], a[2]), a[4]), txmax = nanmax(nanmax(a[0], a[2]), a[4]);
    const float tymin = nanmin(nanmin(a[1], a[3]), a[5]), tymax = nanmax(nanmax(a[1], a[3]), a[5]);
    const float* p0 = in.verts + 3 * (size_t)v0;
    const float* p1 = in.verts + 3 * (size_t)v1;
    const float* p2 = in.verts + 3 * (size_t)v2;
    const float* q0 = in.verts_color + 3 * (size_t)v0;
    const float* q1 = in.verts_color + 3 * (size_t)v1;
    const float* q2 = in.verts_color + 3 * (size_t)v2;
    const size_t vb = (size_t)b * in.P;
    const float z0 = in.verts_ndc[(vb + v0) * 3 + 2];
    const float z1 = in.verts_ndc[(vb + v1) * 3 + 2];
    const float z2 = in.verts_ndc[(vb + v2) * 3 + 2];
    rec[0]  = make_float4(txmin, txmax, tymin, tymax);
    rec[1]  = make_float4(n[0], n[1], n[2], n[3]);
    rec[2]  = make_float4(n[4], n[5], c0, c1);
    rec[3]  = make_float4(c2, in.faces_opacity[f], in.faces_intense[idx], __uint_as_float(flags));
    rec[4]  = make_float4(p0[0], p0[1], p0[2], z0);
    rec[5]  = make_float4(p1[0], p1[1], p1[2], z1);
    rec[6]  = make_float4(p2[0], p2[1], p2[2], z2);
    rec[7]  = make_float4(q0[0], q0[1], q0[2], q1[0]);
    rec[8]  = make_float4(q1[1], q1[2], q2[0], q2[1]);
    rec[9]  = make_float4(q2[2], __int_as_float(v0), __int_as_float(v1), __int_as_float(v2));
    rec[10] = make_float4(a[0], a[1], a[2], a[3]);
    rec[11] = make_float4(a[4], a[5], e[0], e[1]);
    rec[12] = make_float4(e[2], e[3], e[4], e[5]);
    rec[13] = make_float4(r[0], r[1], r[2], r[3]);
    rec[14] = make_float4(r[4], r[5], __int_as_float(f), __uint_as_float(safe_word));
    return true;
}

int launch_render_preprocess(const DM2RenderInputs& in, const FaceWs& fw, uint32_t* flags, cudaStream_t s) {
    const long long BF = (long long)in.B * in.F;
    if (BF == 0) return 0;
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        DM2_CHECK(cudaFuncSetAttribute(k_render_preprocess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPreSmem));
        attr_set.done();
    }
    ProfScope _ps("preprocess_pack", s);
    k_render_preprocess<<<cdiv(BF, kPreThreads), kPreThreads, kPreSmem, s>>>(in, gx, gy, fw.depths, fw.min_depths,
                                                                             fw.max_depths, fw.tiles_touched, fw.rects,
                                                                             fw.records, flags);
    DM2_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ fused host prep + preprocess + pack
// Per-vertex part of Renderer.compute_verts_ndc_image (reference __init__.py:251-260) on one
// clip-space vertex: w clamp to +-1e-4, perspective divide, NDC -> image.  Separately rounded IEEE
// ops in torch's order, so the results are bit-identical to the torch tensors of the unfused path.
struct ProjVert { float nx, ny, nz, ix, iy, w; bool clamped; };
__device__ __forceinline__ ProjVert project_vertex(const float4 c, float img_w, float img_h) {
    ProjVert o;
    float w = c.w;
    o.clamped = false;
    if (w >= 0.0f && w < 1e-4f) { w = 1e-4f; o.clamped = true; }
    if (w < 0.0f && w > -1e-4f) { w = -1e-4f; o.clamped = true; }
    o.w = w;
    o.nx = __fdiv_rn(c.x, w); o.ny = __fdiv_rn(c.y, w); o.nz = __fdiv_rn(c.z, w);
    o.ix = __fmul_rn(__fmul_rn(__fadd_rn(o.nx, 1.0f), 0.5f), img_w);
    o.iy = __fmul_rn(__fmul_rn(__fadd_rn(o.ny, 1.0f), 0.5f), img_h);
    return o;
}

#ifndef DM2_PRE_MINB
#define DM2_PRE_MINB 1
#endif
__global__ void __launch_bounds__(kPreThreads, DM2_PRE_MINB)
k_fused_preprocess(DM2FusedInputs in, int gx, int gy, float* __restrict__ depths, float* __restrict__ min_depths,
                   float* __restrict__ max_depths, uint32_t* __restrict__ tiles_touched, uint2* __restrict__ rects,
                   float* __restrict__ records, uint32_t* __restrict__ patch_flags) {
    extern __shared__ __align__(16) float4 s_pre[];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long BF = (long long)in.B * in.F;
    const int lane = threadIdx.x & 31;
    float4* const s_warp = s_pre + (threadIdx.x - lane) * kPreSlot;
    bool live = false;
    if (idx < BF) live = fused_preprocess_one(in, gx, gy, idx, depths, min_depths, max_depths, tiles_touched, rects,
                                              s_warp + lane * kPreSlot, patch_flags);
    write_records_coalesced(s_warp, live, idx, BF, records, lane);
}

__device__ __forceinline__ bool fused_preprocess_one(const DM2FusedInputs& in, int gx, int gy, long long idx,
                                                     float* __restrict__ depths, float* __restrict__ min_depths,
                                                     float* __restrict__ max_depths, uint32_t* __restrict__ tiles_touched,
                                                     uint2* __restrict__ rects, float4* __restrict__ rec,
                                                     uint32_t* __restrict__ patch_flags) {
    if (idx < in.B) {
        // the reference asserts these on the host with a device sync (__init__.py:283-284); here the
        // verdict travels back with num_rendered
        if (in.patch_min[2 * idx] + in.W > in.image_w) atomicOr(patch_flags, kFlagPatchX);
        if (in.patch_min[2 * idx + 1] + in.H > in.image_h) atomicOr(patch_flags, kFlagPatchY);
        // the reference's select_rays wraps negative indices (torch indexing); rays are read in place here
        if (in.patch_min[2 * idx] < 0 || in.patch_min[2 * idx + 1] < 0) atomicOr(patch_flags, kFlagPatchNeg);
    }
    const int b = (int)(idx / in.F), f = (int)(idx % in.F);
    const int v0 = in.faces[3 * (size_t)f], v1 = in.faces[3 * (size_t)f + 1], v2 = in.faces[3 * (size_t)f + 2];
    if ((unsigned)v0 >= (unsigned)in.P || (unsigned)v1 >= (unsigned)in.P || (unsigned)v2 >= (unsigned)in.P) {
        atomicOr(patch_flags, kFlagVertexId);
        tiles_touched[idx] = 0; rects[idx] = make_uint2(0, 0);
        return false;
    }
    const uint32_t pmx = (uint32_t)in.patch_min[2 * b], pmy = (uint32_t)in.patch_min[2 * b + 1];
    const float* mvb = in.mv + 16 * (size_t)b;
    const float* prb = in.proj + 16 * (size_t)b;
    const float fw = (float)in.image_w, fh = (float)in.image_h;
    const ProjVert a0 = project_vertex(clip_vertex(in.verts + 3 * (size_t)v0, mvb, prb), fw, fh);
    const ProjVert a1 = project_vertex(clip_vertex(in.verts + 3 * (size_t)v1, mvb, prb), fw, fh);
    const ProjVert a2 = project_vertex(clip_vertex(in.verts + 3 * (size_t)v2, mvb, prb), fw, fh);

    // ---- preprocessFaceCUDA (reference forward.cu:41-107), same chain as preprocess_face()
    const float max_z = fmaxf(fmaxf(a0.nz, a1.nz), a2.nz);
    const float min_z = fminf(fminf(a0.nz, a1.nz), a2.nz);
    float depth = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, a0.nz), a1.nz), a2.nz);
    depth = __fdiv_rn(depth, 3.0f);
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    bool live = !(max_z < -1.0f || min_z > 1.0f);
    if (live) {
        patch_rect(pmx, pmy, a0.ix, a0.iy, a1.ix, a1.iy, a2.ix, a2.iy, gx, gy, x0, y0, x1, y1);
        live = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0) != 0u;
    }
    if (!live) { tiles_touched[idx] = 0; rects[idx] = make_uint2(0, 0); return false; }
    tiles_touched[idx] = (uint32_t)(y1 - y0) * (uint32_t)(x1 - x0);
    rects[idx] = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));
    depths[idx] = clamp01_depth(depth);
    min_depths[idx] = clamp01_depth(min_z);
    max_depths[idx] = clamp01_depth(max_z);

    // ---- Triangles (reference pyrenderer.py:6-30, 521-535): CCW order, edges, flags, reciprocals, normals
    const float p0x = a0.ix, p0y = a0.iy;
    const float sarea = __fmul_rn(0.5f, __fsub_rn(__fmul_rn(__fsub_rn(a1.ix, p0x), __fsub_rn(a2.iy, p0y)),
                                                  __fmul_rn(__fsub_rn(a2.ix, p0x), __fsub_rn(a1.iy, p0y))));
    const bool flip = sarea < 0.0f;
    const float p1x = flip ? a2.ix : a1.ix, p1y = flip ? a2.iy : a1.iy;
    const float p2x = flip ? a1.ix : a2.ix, p2y = flip ? a1.iy : a2.iy;
    const float e[6] = {__fsub_rn(p1x, p0x), __fsub_rn(p1y, p0y), __fsub_rn(p2x, p1x), __fsub_rn(p2y, p1y),
                        __fsub_rn(p0x, p2x), __fsub_rn(p0y, p2y)};
    float r[6];
    uint32_t flags = flip ? 0x100u : 0u;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        r[k] = __fdiv_rn(1.0f, e[k]);
        flags |= (fabsf(e[k]) < 1e-3f ? 1u : 0u) << k;
    }
    // some screen coordinate is an integer: possible kink of the overlap area (dm2_aa.cuh)
    if (p0x == floorf(p0x) || p0y == floorf(p0y) || p1x == floorf(p1x) || p1y == floorf(p1y) || p2x == floorf(p2x) ||
        p2y == floorf(p2y)) flags |= 0x200u;
    uint32_t safe_word = 0u;
    {
        const float av[6] = {p0x, p0y, p1x, p1y, p2x, p2y};
        flags |= aaf_face_safety(av, e, r, flags, &safe_word);
    }
    // n_k = (-e_k.y, e_k.x), c_k = n_k . anchor_k  (two rounded products, one add)
    const float n[6] = {-e[1], e[0], -e[3], e[2], -e[5], e[4]};
    const float c0 = __fadd_rn(__fmul_rn(n[0], p0x), __fmul_rn(n[1], p0y));
    const float c1 = __fadd_rn(__fmul_rn(n[2], p1x), __fmul_rn(n[3], p1y));
    const float c2 = __fadd_rn(__fmul_rn(n[4], p2x), __fmul_rn(n[5], p2y));
    const float txmin = nanmin(nanmin(p0x, p1x), p2x), txmax = nanmax(nanmax(p0x, p1x), p2x);
    const float tymin = nanmin(nanmin(p0y, p1y), p2y), tymax = nanmax(nanmax(p0y, p1y), p2y);

    const float* w0 = in.verts + 3 * (size_t)v0;
    const float* w1 = in.verts + 3 * (size_t)v1;
    const float* w2 = in.verts + 3 * (size_t)v2;
    const float* q0 = in.verts_color + 3 * (size_t)v0;
    const float* q1 = in.verts_color + 3 * (size_t)v1;
    const float* q2 = in.verts_color + 3 * (size_t)v2;
    rec[0]  = make_float4(txmin, txmax, tymin, tymax);
    rec[1]  = make_float4(n[0], n[1], n[2], n[3]);
    rec[2]  = make_float4(n[4], n[5], c0, c1);
    rec[3]  = make_float4(c2, in.faces_opacity[f], in.faces_intense[idx], __uint_as_float(flags));
    rec[4]  = make_float4(w0[0], w0[1], w0[2], a0.nz);
    rec[5]  = make_float4(w1[0], w1[1], w1[2], a1.nz);
    rec[6]  = make_float4(w2[0], w2[1], w2[2], a2.nz);
    rec[7]  = make_float4(q0[0], q0[1], q0[2], q1[0]);
    rec[8]  = make_float4(q1[1], q1[2], q2[0], q2[1]);
    rec[9]  = make_float4(q2[2], __int_as_float(v0), __int_as_float(v1), __int_as_float(v2));
    rec[10] = make_float4(p0x, p0y, p1x, p1y);
    rec[11] = make_float4(p2x, p2y, e[0], e[1]);
    rec[12] = make_float4(e[2], e[3], e[4], e[5]);
    rec[13] = make_float4(r[0], r[1], r[2], r[3]);
    rec[14] = make_float4(r[4], r[5], __int_as_float(f), __uint_as_float(safe_word));
    return true;
}

int launch_fused_preprocess(const DM2FusedInputs& in, const FaceWs& fw, uint32_t* patch_flags, cudaStream_t s) {
    const long long BF = (long long)in.B * in.F;
    if (BF == 0) return 0;
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        DM2_CHECK(cudaFuncSetAttribute(k_fused_preprocess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPreSmem));
        attr_set.done();
    }
    ProfScope _ps("preprocess_pack", s);
    k_fused_preprocess<<<cdiv(BF, kPreThreads), kPreThreads, kPreSmem, s>>>(in, gx, gy, fw.depths, fw.min_depths,
                                                                            fw.max_depths, fw.tiles_touched, fw.rects,
                                                                            fw.records, patch_flags);
    DM2_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ layered preprocess + pack
__global__ void __launch_bounds__(256)
k_layers_preprocess(DM2LayerInputs in, int gx, int gy, float* __restrict__ depths,
                    float* __restrict__ min_depths, float* __restrict__ max_depths,
                    uint32_t* __restrict__ tiles_touched, uint2* __restrict__ rects,
                    float* __restrict__ records, uint32_t* __restrict__ flags) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long BF = (long long)in.B * in.F;
    if (idx >= BF) return;
    const int b = (int)(idx / in.F), f = (int)(idx % in.F);
    const int v0 = in.faces[3 * (size_t)f], v1 = in.faces[3 * (size_t)f + 1], v2 = in.faces[3 * (size_t)f + 2];
    if ((unsigned)v0 >= (unsigned)in.P || (unsigned)v1 >= (unsigned)in.P || (unsigned)v2 >= (unsigned)in.P) {
        atomicOr(flags, kFlagVertexId);
        tiles_touched[idx] = 0; rects[idx] = make_uint2(0, 0);
        return;
    }
    PreOut o;
    if (in.mv != nullptr) {
        // fused host prep: same projection chain as k_fused_preprocess, then preprocess_face()'s arithmetic
        const float* mvb = in.mv + 16 * (size_t)b;
        const float* prb = in.proj + 16 * (size_t)b;
        const float fw = (float)in.W, fh = (float)in.H;
        const ProjVert a0 = project_vertex(clip_vertex(in.verts + 3 * (size_t)v0, mvb, prb), fw, fh);
        const ProjVert a1 = project_vertex(clip_vertex(in.verts + 3 * (size_t)v1, mvb, prb), fw, fh);
        const ProjVert a2 = project_vertex(clip_vertex(in.verts + 3 * (size_t)v2, mvb, prb), fw, fh);
        o.live = false; o.x0 = o.y0 = o.x1 = o.y1 = 0; o.depth = o.dmin = o.dmax = 0.f;
        const float max_z = fmaxf(fmaxf(a0.nz, a1.nz), a2.nz);
        const float min_z = fminf(fminf(a0.nz, a1.nz), a2.nz);
        float depth = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, a0.nz), a1.nz), a2.nz);
        depth = __fdiv_rn(depth, 3.0f);
        if (!(max_z < -1.0f || min_z > 1.0f)) {
            patch_rect(0u, 0u, a0.ix, a0.iy, a1.ix, a1.iy, a2.ix, a2.iy, gx, gy, o.x0, o.y0, o.x1, o.y1);
            if ((uint32_t)(o.x1 - o.x0) * (uint32_t)(o.y1 - o.y0) != 0u) {
                o.live = true;
                o.depth = clamp01_depth(depth);
                o.dmin = clamp01_depth(min_z);
                o.dmax = clamp01_depth(max_z);
            }
        }
    } else {
        o = preprocess_face(b, in.P, v0, v1, v2, in.verts_ndc, in.verts_image, 0u, 0u, gx, gy);
    }
    if (!o.live) { tiles_touched[idx] = 0; rects[idx] = make_uint2(0, 0); return; }
    tiles_touched[idx] = (uint32_t)(o.y1 - o.y0) * (uint32_t)(o.x1 - o.x0);
    rects[idx] = make_uint2((uint32_t)o.x0 | ((uint32_t)o.y0 << 16), (uint32_t)o.x1 | ((uint32_t)o.y1 << 16));
    depths[idx] = o.depth; min_depths[idx] = o.dmin; max_depths[idx] = o.dmax;
    float4* rec = reinterpret_cast<float4*>(records + (size_t)idx * kLRecWords);
    const float* p0 = in.verts + 3 * (size_t)v0;
    const float* p1 = in.verts + 3 * (size_t)v1;
    const float* p2 = in.verts + 3 * (size_t)v2;
    rec[0] = make_float4(p0[0], p0[1], p0[2], p1[0]);
    rec[1] = make_float4(p1[1], p1[2], p2[0], p2[1]);
    rec[2] = make_float4(p2[2], o.dmin, o.dmax, __int_as_float(f));
}

int launch_layers_preprocess(const DM2LayerInputs& in, const FaceWs& fw, uint32_t* flags, cudaStream_t s) {
    const long long BF = (long long)in.B * in.F;
    if (BF == 0) return 0;
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    ProfScope _ps("layers_preprocess", s);
    k_layers_preprocess<<<cdiv(BF, 256), 256, 0, s>>>(in, gx, gy, fw.depths, fw.min_depths, fw.max_depths,
                                                      fw.tiles_touched, fw.rects, fw.records, flags);
    DM2_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ inclusive u32 scan
// reduce-then-scan: (1) per-block sums, (2) one block scans the sums, (3) per-block scan + base.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

long long scan_partials_count(long long n) { return (n + kScanTile - 1) / kScanTile + 1; }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix,
// total in `total`.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t& total) {
    __shared__ uint32_t wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { uint32_t s = wsum[i]; if (i < w) base += s; tot += s; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_reduce(const uint32_t* __restrict__ in, long long n, uint32_t* __restrict__ partials) {
    const long long base = (long long)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        long long j = base + (long long)i * kScanThreads + threadIdx.x;
        if (j < n) s += in[j];
    }
    uint32_t tot;
    block_excl_scan(s, tot);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_partials(uint32_t* __restrict__ partials, int nblocks) {
    // exclusive scan in place; partials[nblocks] receives the grand total
    uint32_t carry = 0;
    for (int base = 0; base < nblocks; base += kScanThreads) {
        int j = base + threadIdx.x;
        uint32_t v = (j < nblocks) ? partials[j] : 0u;
        uint32_t tot;
        uint32_t ex = block_excl_scan(v, tot);
        if (j < nblocks) partials[j] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) partials[nblocks] = carry;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long n,
             const uint32_t* __restrict__ partials) {
    // each thread owns kScanItems consecutive elements
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { v[i] = (base + i < n) ? in[base + i] : 0u; s += v[i]; }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, tot) + partials[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { ex += v[i]; if (base + i < n) out[base + i] = ex; }
}

int launch_scan(const uint32_t* in, uint32_t* out, long long n, uint32_t* partials, cudaStream_t s) {
    if (n <= 0) return 0;
    const int nb = cdiv(n, kScanTile);
    ProfScope _ps("scan", s);
    k_scan_reduce<<<nb, kScanThreads, 0, s>>>(in, n, partials);
    DM2_LAUNCH_CHECK();
    k_scan_partials<<<1, kScanThreads, 0, s>>>(partials, nb);
    DM2_LAUNCH_CHECK();
    k_scan_apply<<<nb, kScanThreads, 0, s>>>(in, out, n, partials);
    DM2_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------ stable LSD radix sort
// 8-bit digits, 4096 keys per block.  Per pass: per-block digit histogram -> digit-major
// exclusive scan -> stable scatter (warp-level match_any ranking).  Replaces
// cub::DeviceRadixSort::SortPairs (reference renderer.cu:202-207): sorts on key bits
// [0, end_bit) only and keeps the input order of equal keys.
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;   // 4096
constexpr int kRadix = 256;

// temp = [counts (nblk*256) | inclusive scan of counts (nblk*256) | scan partials]
static size_t sort_counts_bytes(long long n) {
    const long long nblk = (n + kSortTile - 1) / kSortTile;
    return align_up((size_t)(nblk * kRadix + 1) * sizeof(uint32_t), 256);
}
size_t sort_temp_bytes(long long n) {
    const long long nblk = (n + kSortTile - 1) / kSortTile;
    return 2 * sort_counts_bytes(n) + align_up((size_t)scan_partials_count(nblk * kRadix) * sizeof(uint32_t), 256);
}

__device__ __forceinline__ void warp_rank_digit(uint32_t* wc /*this warp's 256 counters*/, bool valid,
                                                uint32_t d, int lane, uint32_t& rank_out) {
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        const unsigned peers = __match_any_sync(vm, d);
        const int leader = __ffs(peers) - 1;
        const uint32_t below = __popc(peers & ((1u << lane) - 1u));
        uint32_t prev = 0;
        if (lane == leader) { prev = wc[d]; wc[d] = prev + __popc(peers); }
        prev = __shfl_sync(peers, prev, leader);
        rank_out = prev + below;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kSortThreads)
k_radix_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t mask,
             uint32_t* __restrict__ counts, int nblk) {
    __shared__ uint32_t wc[8][kRadix];
    for (int i = threadIdx.x; i < 8 * kRadix; i += kSortThreads) (&wc[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * kSortTile + w * (kSortItems * 32);
#pragma unroll 4
    for (int it = 0; it < kSortItems; it++) {
        const uint32_t i = base + it * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? (uint32_t)(keys[i] >> shift) & mask : 0u;
        uint32_t r;
        warp_rank_digit(wc[w], valid, d, lane, r);
    }
    __syncthreads();
    const int d = threadIdx.x;
    uint32_t c = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ww++) c += wc[ww][d];
    counts[(size_t)d * nblk + blockIdx.x] = c;
}

// exclusive scan over the digit-major count matrix (single block, 1024 threads)
__global__ void __launch_bounds__(1024)
k_radix_scan(uint32_t* __restrict__ counts, uint32_t total) {
    __shared__ uint32_t wsum[32];
    const uint32_t per = (total + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, total), hi = min(lo + per, total);
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; i++) s += counts[i];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(s, lane);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) { uint32_t v = wsum[lane]; uint32_t iv = warp_incl_scan(v, lane); wsum[lane] = iv - v; }
    __syncthreads();
    uint32_t run = wsum[w] + inc - s;
    for (uint32_t i = lo; i < hi; i++) { uint32_t c = counts[i]; counts[i] = run; run += c; }
}

__global__ void __launch_bounds__(kSortThreads)
k_radix_scatter(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                uint64_t* __restrict__ kout, uint32_t* __restrict__ vout, uint32_t n, int shift,
                uint32_t mask, const uint32_t* __restrict__ gbase, int nblk) {
    __shared__ uint32_t wc[8][kRadix];
    for (int i = threadIdx.x; i < 8 * kRadix; i += kSortThreads) (&wc[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * kSortTile + w * (kSortItems * 32);
    uint64_t k[kSortItems];
    uint32_t r[kSortItems];
#pragma unroll
    for (int it = 0; it < kSortItems; it++) {
        const uint32_t i = base + it * 32 + lane;
        k[it] = (i < n) ? kin[i] : 0ull;
    }
#pragma unroll
    for (int it = 0; it < kSortItems; it++) {
        const uint32_t i = base + it * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = (uint32_t)(k[it] >> shift) & mask;
        r[it] = 0;
        warp_rank_digit(wc[w], valid, d, lane, r[it]);
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        // gbase holds the INCLUSIVE scan of the digit-major count matrix; exclusive = previous entry
        const size_t gi = (size_t)d * nblk + blockIdx.x;
        uint32_t run = (gi == 0) ? 0u : gbase[gi - 1];
#pragma unroll
        for (int ww = 0; ww < 8; ww++) { uint32_t c = wc[ww][d]; wc[ww][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSortItems; it++) {
        const uint32_t i = base + it * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(k[it] >> shift) & mask;
            const uint32_t dst = wc[w][d] + r[it];
            kout[dst] = k[it];
            vout[dst] = vin[i];
        }
    }
}

int launch_sort_pairs(const uint64_t* kin, const uint32_t* vin, uint64_t* kout, uint32_t* vout,
                      long long n, int end_bit, void* temp, size_t temp_bytes, cudaStream_t s) {
    if (n <= 0) return 0;
    DM2_REQUIRE(n < (1ll << 31), "dm2_sort_pairs: n must be < 2^31");
    DM2_REQUIRE(end_bit >= 0 && end_bit <= 64, "dm2_sort_pairs: end_bit out of range");
    DM2_REQUIRE(temp_bytes >= sort_temp_bytes(n), "dm2_sort_pairs: temp buffer too small");
    const int nblk = cdiv(n, kSortTile);
    const int passes = (end_bit + 7) / 8;
    ProfScope _ps("radix_sort", s);
    uint32_t* counts = reinterpret_cast<uint32_t*>(temp);
    uint32_t* incl = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(temp) + sort_counts_bytes(n));
    uint32_t* partials = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(temp) + 2 * sort_counts_bytes(n));
    if (passes == 0) {
        DM2_CHECK(cudaMemcpyAsync(kout, kin, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        DM2_CHECK(cudaMemcpyAsync(vout, vin, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    // ping-pong between (kin,vin) and (kout,vout); the input buffers are clobbered when passes > 1.
    // With an even number of passes the first hop goes through a copy so that the result lands in
    // (kout, vout).
    uint64_t* ka = const_cast<uint64_t*>(kin); uint32_t* va = const_cast<uint32_t*>(vin);
    uint64_t* kb = kout;                       uint32_t* vb = vout;
    if (passes % 2 == 0) {
        DM2_CHECK(cudaMemcpyAsync(kout, kin, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        DM2_CHECK(cudaMemcpyAsync(vout, vin, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        ka = kout; va = vout; kb = const_cast<uint64_t*>(kin); vb = const_cast<uint32_t*>(vin);
    }
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        const int bits = min(8, end_bit - shift);
        const uint32_t mask = (1u << bits) - 1u;
        k_radix_hist<<<nblk, kSortThreads, 0, s>>>(ka, (uint32_t)n, shift, mask, counts, nblk);
        DM2_LAUNCH_CHECK();
        if (launch_scan(counts, incl, (long long)nblk * kRadix, partials, s)) return 1;
        k_radix_scatter<<<nblk, kSortThreads, 0, s>>>(ka, va, kb, vb, (uint32_t)n, shift, mask, incl, nblk);
        DM2_LAUNCH_CHECK();
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    return 0;
}

// ------------------------------------------------------------------ per-instance pixel classification
// Steps 1-3 of the reference's per-(pixel, face) AA routine (aa.h:96-149, 446-496: triangle bbox
// reject, 3 edges x 4 pixel corners half-plane tests -> outside / fully inside / partial) do not
// depend on the compositing order, and the reference evaluates them per pixel per listed face inside
// both raster kernels.  Here they run ONCE per sorted instance and 8x4-pixel footprint, bit-parallel:
// the corner predicate  fma(cx, n.x, cy * n.y) - c >= 0  is monotone in cx (one rounding of an
// affine function), so its value over the 9 corner columns of a footprint row is a run of bits whose
// end is found with 5 exact evaluations; the 32 per-pixel outcomes then follow from shifts and ANDs.
// Output per (footprint w, instance i), 16 bytes at meta[w * R + i]:
//   x: bit l = pixel l of the footprint (x = l & 7, y = l >> 3) is touched (not AA_OUT)
//   y,z: 45-bit corner grid, bit 9 * r + c = corner (c, r) lies inside all three edge half-planes
//   w: AA-cache entry offset, written by the forward tile kernel (0xFFFFFFFF = not cached)
// and imask[i] bit w = "x != 0": the tile kernels build their per-warp candidate lists from it with
// one coalesced byte load per instance.  With aa_temperature == 0 the reference runs the ray test on
// every listed face for every pixel: imask = 0xFF and no meta.

// X0f = (float)X0 (exact); corner columns are X0f + 0 .. X0f + 8, exact in fp32 for coordinates < 2^24.
// The bisection runs on float column offsets (midpoints of small integers are exact) so that no
// int -> float conversion sits inside the loop.
__device__ __forceinline__ uint32_t corner_row_mask(float nx, float ylo, float c, float X0f) {
    auto pred = [&](float cx) { return __fsub_rn(__fmaf_rn(__fadd_rn(X0f, cx), nx, ylo), c) >= 0.0f; };
    const bool p0 = pred(0.0f), p8 = pred(8.0f);
    float lo = 0.0f, hi = 8.0f;
#pragma unroll
    for (int it = 0; it < 3; it++) {
        const float mid = 0.5f * (lo + hi);
        const bool same = pred(mid) == p0;
        lo = same ? mid : lo;
        hi = same ? hi : mid;
    }
    // columns 0..lo carry p0, columns hi..8 carry p8 (hi == lo + 1)
    const uint32_t below = (1u << (int)hi) - 1u;                 // bits 0 .. lo
    return (p0 ? below : 0u) | (p8 ? (0x1FFu & ~below) : 0u);
}

constexpr int kClsThreads = 256;
constexpr uint32_t kClsRoundBlocks = 4;     // blocks per tile of a classification round

// ROUND: kClsRoundBlocks blocks per tile classify entries [round_begin, round_begin + 256 chunks) of the tile's list,
// 256 at a time, only for the footprints of the tile that have not finished compositing (warp_done).  work == nullptr:
// every tile (grid = kClsRoundBlocks x tiles, the first round); else a grid-stride loop over the list of tiles that
// are still running (k_round_worklist) - a round in which (nearly) every tile has finished must not cost a launch of
// tiles x 4 blocks (92 us at 2048^2).  The flat variant covers every instance.
template <bool ROUND>
__global__ void __launch_bounds__(kClsThreads)
k_classify_instances(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t cap, BinGuard guard,
                     const float* __restrict__ records, const int32_t* __restrict__ patch_min, int F, int gx, int gy,
                     int use_bbox, uint8_t* __restrict__ imask, uint4* __restrict__ meta,
                     const uint2* __restrict__ ranges, const uint8_t* __restrict__ warp_done, int round_begin, int chunks,
                     const uint32_t* __restrict__ work, const uint32_t* __restrict__ work_n) {
    __shared__ float s_f[13][kClsThreads];
    __shared__ uint32_t s_o[2][kClsThreads];
    __shared__ uint32_t s_im[kClsThreads];
    __shared__ uint16_t s_pairs[kClsThreads * 8];
    const int t = threadIdx.x;
    if (bin_invalid(guard)) return;
    const uint32_t n_items = !ROUND ? blockIdx.x + 1u : work ? *work_n * kClsRoundBlocks : gridDim.x;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    uint32_t R = __ldg(guard.hdr);                  // num_rendered (<= cap, the stride of meta)
    uint32_t base = blockIdx.x * kClsThreads, active = 0xFFu;
    if (ROUND) {
        const uint32_t tile = work ? work[item / kClsRoundBlocks] : item / kClsRoundBlocks;
        const uint2 r = ranges[tile];
        base = r.x + (uint32_t)round_begin + (item % kClsRoundBlocks) * kClsThreads;
        R = min(r.y, r.x + (uint32_t)round_begin + (uint32_t)chunks * kClsThreads);
        const uint2 wd = *reinterpret_cast<const uint2*>(warp_done + (size_t)tile * 8);
#pragma unroll
        for (int w = 0; w < 4; w++) {
            if ((wd.x >> (8 * w)) & 0xFFu) active &= ~(1u << w);
            if ((wd.y >> (8 * w)) & 0xFFu) active &= ~(16u << w);
        }
        if (active == 0u) continue;
    }
    for (; base < R; base += kClsRoundBlocks * kClsThreads) {           // (one trip in the flat variant)
    const uint32_t i = base + t;
    if (!use_bbox) { if (i < R) imask[i] = (uint8_t)active; if (ROUND) continue; return; }
    unsigned bm = 0u;
    if (i < R) {
        const uint32_t cur = (uint32_t)(keys[i] >> 32);
        const uint32_t tiles = (uint32_t)(gx * gy);
        const uint32_t b = cur / tiles, t_in = cur % tiles;
        const uint32_t tx = t_in % (uint32_t)gx, ty = t_in / (uint32_t)gx;
        const float* rec = records + ((size_t)b * F + vals[i]) * kRecWords;
        const float4 bb = *reinterpret_cast<const float4*>(rec + R_BBOX);
        const float4 n01 = *reinterpret_cast<const float4*>(rec + R_N01);
        const float4 n2c = *reinterpret_cast<const float4*>(rec + R_N2C);
        const float c2 = rec[R_MISC];
        const uint32_t ox = tx * kTile + (uint32_t)patch_min[2 * b], oy = ty * kTile + (uint32_t)patch_min[2 * b + 1];
        s_f[0][t] = n01.x; s_f[1][t] = n01.y; s_f[2][t] = n01.z; s_f[3][t] = n01.w; s_f[4][t] = n2c.x; s_f[5][t] = n2c.y;
        s_f[6][t] = n2c.z; s_f[7][t] = n2c.w; s_f[8][t] = c2;
        s_f[9][t] = bb.x; s_f[10][t] = bb.y; s_f[11][t] = bb.z; s_f[12][t] = bb.w;
        s_o[0][t] = ox; s_o[1][t] = oy;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const float fx0 = (float)(ox + (w & 1) * 8), fy0 = (float)(oy + (w >> 1) * 4);
            const bool hit = !((fx0 + 8.0f < bb.x) || (fx0 > bb.y) || (fy0 + 4.0f < bb.z) || (fy0 > bb.w));
            bm |= (hit ? 1u : 0u) << w;
        }
        bm &= active;
    }
    s_im[t] = 0u;
    uint32_t total;
    uint32_t off = block_excl_scan(__popc(bm), total);
    for (unsigned m = bm; m; m &= m - 1u) s_pairs[off++] = (uint16_t)((t << 3) | (__ffs(m) - 1));
    __syncthreads();

    for (uint32_t p = t; p < total; p += kClsThreads) {
        const int e = s_pairs[p];
        const int it = e >> 3, w = e & 7;
        const float nxs[3] = {s_f[0][it], s_f[2][it], s_f[4][it]};
        const float nys[3] = {s_f[1][it], s_f[3][it], s_f[5][it]};
        const float ccs[3] = {s_f[6][it], s_f[7][it], s_f[8][it]};
        const float bx0 = s_f[9][it], bx1 = s_f[10][it], by0 = s_f[11][it], by1 = s_f[12][it];
        const uint32_t X0 = s_o[0][it] + (w & 1) * 8, Y0 = s_o[1][it] + (w >> 1) * 4;
        const float X0f = (float)X0;
        // per-pixel triangle-bbox test (aa.h:96-101), negated so that NaN boxes reject nothing
        uint32_t colmask = 0u, rowmask = 0u;
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const float fx = (float)(X0 + x);
            colmask |= (!((__fadd_rn(fx, 1.0f) < bx0) || (fx > bx1)) ? 1u : 0u) << x;
        }
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const float fy = (float)(Y0 + y);
            rowmask |= (!((__fadd_rn(fy, 1.0f) < by0) || (fy > by1)) ? 1u : 0u) << y;
        }
        uint32_t anyr[4] = {0xFFu, 0xFFu, 0xFFu, 0xFFu};      // per pixel row: every edge has >= 1 corner inside
        uint32_t allc[5] = {0x1FFu, 0x1FFu, 0x1FFu, 0x1FFu, 0x1FFu};   // per corner row: inside all edges
#pragma unroll
        for (int ed = 0; ed < 3; ed++) {
            uint32_t m[5];
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const float ylo = __fmul_rn((float)(Y0 + r), nys[ed]);
                m[r] = corner_row_mask(nxs[ed], ylo, ccs[ed], X0f);
                allc[r] &= m[r];
            }
#pragma unroll
            for (int y = 0; y < 4; y++) anyr[y] &= (m[y] | (m[y] >> 1) | m[y + 1] | (m[y + 1] >> 1)) & 0xFFu;
        }
        uint32_t hit = 0u;
#pragma unroll
        for (int y = 0; y < 4; y++) hit |= (((rowmask >> y) & 1u) ? (anyr[y] & colmask) : 0u) << (8 * y);
        const unsigned long long corners = (unsigned long long)allc[0] | ((unsigned long long)allc[1] << 9) |
                                           ((unsigned long long)allc[2] << 18) | ((unsigned long long)allc[3] << 27) |
                                           ((unsigned long long)allc[4] << 36);
        const uint32_t gi = base + it;
        meta[(size_t)w * cap + gi] = make_uint4(hit, (uint32_t)corners, (uint32_t)(corners >> 32), 0xFFFFFFFFu);
        if (hit) atomicOr(&s_im[it], 1u << w);
    }
    __syncthreads();
    if (i < R) imask[i] = (uint8_t)s_im[t];
    if (!ROUND) return;
    __syncthreads();                      // s_f / s_pairs are rewritten by the next chunk
    }
    }
}

// Tiles that still have an unfinished footprint and list entries at / behind `next_begin`: their ids go to
// cls_list (counter n[0]); their unfinished half tiles (tile * 2 + half) go to fwd_list (counter n[1]).
__global__ void __launch_bounds__(256)
k_round_worklist(const uint2* __restrict__ ranges, const uint8_t* __restrict__ warp_done, int num_tiles, BinGuard guard,
                 uint32_t next_begin, uint32_t* __restrict__ n, uint32_t* __restrict__ cls_list,
                 uint32_t* __restrict__ fwd_list) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= num_tiles || bin_invalid(guard)) return;
    const uint2 r = ranges[tile];
    if (r.y - r.x <= next_begin) return;
    const uint2 wd = *reinterpret_cast<const uint2*>(warp_done + (size_t)tile * 8);
    const bool lo = (wd.x & 0x01010101u) != 0x01010101u, hi = (wd.y & 0x01010101u) != 0x01010101u;
    if (!lo && !hi) return;
    cls_list[atomicAdd(n, 1u)] = (uint32_t)tile;
    const uint32_t at = atomicAdd(n + 1, (lo ? 1u : 0u) + (hi ? 1u : 0u));
    if (lo) fwd_list[at] = 2u * tile;
    if (hi) fwd_list[at + (lo ? 1u : 0u)] = 2u * tile + 1u;
}

int launch_classify(const uint64_t* keys, long long capacity, long long longest_cap, const uint32_t* hdr, const TileMaskArgs& tm,
                    cudaStream_t s) {
    if (capacity <= 0) return 0;
    ProfScope _ps("classify", s);
    // (the grid covers the workspace capacity; blocks past the actual num_rendered exit at once)
    k_classify_instances<false><<<cdiv(capacity, kClsThreads), kClsThreads, 0, s>>>(
        keys, tm.vals, (uint32_t)capacity, BinGuard{hdr, (uint32_t)capacity, (uint32_t)longest_cap}, tm.records, tm.patch_min, tm.F, tm.gx, tm.gy,
        tm.use_bbox, tm.imask, reinterpret_cast<uint4*>(tm.meta), nullptr, nullptr, 0, 1, nullptr, nullptr);
    DM2_LAUNCH_CHECK();
    return 0;
}

int launch_classify_round(const uint64_t* keys, long long capacity, long long longest_cap, const uint32_t* hdr,
                          const TileMaskArgs& tm, const uint2* ranges, const uint8_t* warp_done, int num_tiles,
                          int round_begin, int round_len, const RoundWork& work, const RoundSort& sort, cudaStream_t s) {
    if (capacity <= 0 || num_tiles <= 0) return 0;
    const BinGuard guard{hdr, (uint32_t)capacity, (uint32_t)longest_cap};
    const bool listed = round_begin > 0;
    if (listed) {
        // the tiles / half tiles still running after the rounds so far
        DM2_CHECK(cudaMemsetAsync(work.n, 0, 4 * sizeof(uint32_t), s));
        k_round_worklist<<<cdiv(num_tiles, 256), 256, 0, s>>>(ranges, warp_done, num_tiles, guard, (uint32_t)round_begin,
                                                              work.n, work.cls_list, work.fwd_list);
        DM2_LAUNCH_CHECK();
        // lazy sort: the listed tiles whose sorted prefix ends inside this round's window are sorted to the end
        if (launch_tile_sort_finish(num_tiles, ranges, sort.cursor, sort.keys_unsorted, sort.vals_unsorted, sort.keys, sort.vals,
                                    capacity, sort.tiers, work.cls_list, work.n, (long long)round_begin + round_len, sort.lazy, s)) return 1;
    }
    ProfScope _ps("classify", s);
    const int chunks = cdiv(round_len, kClsThreads);
    const long long grid = listed ? std::min<long long>((long long)num_tiles * kClsRoundBlocks, 4ll * sm_count())
                                  : (long long)num_tiles * kClsRoundBlocks;
    k_classify_instances<true><<<(unsigned)grid, kClsThreads, 0, s>>>(
        keys, tm.vals, (uint32_t)capacity, guard, tm.records, tm.patch_min, tm.F, tm.gx, tm.gy,
        tm.use_bbox, tm.imask, reinterpret_cast<uint4*>(tm.meta), ranges, warp_done, round_begin, chunks,
        listed ? work.cls_list : nullptr, listed ? work.n : nullptr);
    DM2_LAUNCH_CHECK();
    return 0;
}

}  // namespace dm2
