// dm2_raster_common.cuh — device code shared by the forward (dm2_raster_fwd.cu) and backward
// (dm2_raster_bwd.cu) tile kernels: record staging, candidate lists, per-(pixel,face) shading,
// sub-chunk classification + dense polygon clip.
#pragma once
#include "dm2_common.cuh"
#include "dm2_aa.cuh"

namespace dm2 {

constexpr int kRecStride = 68;              // words between staged records (272 B, 16-byte aligned)
constexpr int kRasterThreads = 256;
constexpr int kGradWords = 32;              // per-(view,face) gradient record (29 used)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Stage `cnt` records.  Slot i receives the record of face list[first + i*step] (step = +1 forward,
// -1 backward).  16 threads cooperate on one record (15 x 16-byte chunks carry data).
__device__ __forceinline__ void stage_records(float* __restrict__ smem, const float* __restrict__ records,
                                              const uint32_t* __restrict__ list, long long first, int step,
                                              int cnt, size_t view_base) {
    for (int q = threadIdx.x; q < cnt * 16; q += kRasterThreads) {
        const int slot = q >> 4, chunk = q & 15;
        if (chunk == 15) continue;
        const uint32_t fid = list[first + (long long)slot * step];
        const float* src = records + (view_base + fid) * kRecWords + chunk * 4;
        cp_async16(smem + slot * kRecStride + chunk * 4, src);
    }
    cp_async_wait_all();
}

// Warp-level candidate list: slots whose bbox overlaps the warp's 8x4 pixel footprint (tau > 0),
// or every slot (tau == 0: the reference runs the ray test on every listed face).
__device__ __forceinline__ int build_candidates(const float* __restrict__ srec, int cnt, bool use_bbox, float fx0,
                                                float fy0, uint8_t* __restrict__ cand, int lane) {
    int n = 0;
    for (int c = 0; c < cnt; c += 32) {
        const int j = c + lane;
        bool hit = j < cnt;
        if (hit && use_bbox) {
            const float4 bb = *reinterpret_cast<const float4*>(srec + j * kRecStride + R_BBOX);
            hit = !((fx0 + 8.0f < bb.x) || (fx0 > bb.y) || (fy0 + 4.0f < bb.z) || (fy0 > bb.w));
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) cand[n + __popc(m & ((1u << lane) - 1u))] = (uint8_t)j;
        n += __popc(m);
    }
    __syncwarp();
    return n;
}

struct PixelGeom {
    float pxmin, pymin;     // image-space lower-left corner (patch offset added)
    f3 o, d;
};

// Everything the forward and the backward pass must agree on bit-for-bit for one (pixel, face):
// coverage ratio, clamped barycentrics, interpolated colour/depth (forward.cu:341-395).
struct Shade {
    float area, ratio, w0, w1, w2, iC[3], iD;
    int code;
};

__device__ __forceinline__ bool shade_tail(const float* __restrict__ rec, const PixelGeom& pg, float tau, Shade& s);

// is_part: the pixel is only partially covered (aa_classify == AA_PART); ignored when tau == 0.
template <bool GRAD>
__device__ __forceinline__ bool shade(const float* __restrict__ rec, const PixelGeom& pg, float tau, bool is_part,
                                      Shade& s, float* __restrict__ g) {
    s.area = 0.0f;
    if (GRAD) {
#pragma unroll
        for (int k = 0; k < 6; k++) g[k] = 0.0f;
    }
    if (tau > 0.0f) {
        s.area = 1.0f;
        if (is_part) {
            const unsigned inmask = aa_inmask(rec, pg.pxmin, pg.pymin);
            if (!aa_clip<GRAD>(rec, pg.pxmin, pg.pymin, inmask, s.area, g)) return false;
        }
    }
    return shade_tail(rec, pg, tau, s);
}

// Everything after the coverage area (s.area must be set): hit point, clamped barycentrics,
// coverage ratio, interpolated colour and depth (forward.cu:354-395).
__device__ __forceinline__ bool shade_tail(const float* __restrict__ rec, const PixelGeom& pg, float tau, Shade& s) {
    const float4 a0 = *reinterpret_cast<const float4*>(rec + R_P0);
    const float4 a1 = *reinterpret_cast<const float4*>(rec + R_P1);
    const float4 a2 = *reinterpret_cast<const float4*>(rec + R_P2);
    float t, u, v;
    if (!ray_tri(pg.o, pg.d, f3{a0.x, a0.y, a0.z}, f3{a1.x, a1.y, a1.z}, f3{a2.x, a2.y, a2.z}, t, u, v))
        return false;
    float uc, vc;
    clamp_bary(u, v, uc, vc, s.code);
    s.w0 = __fsub_rn(__fsub_rn(1.0f, uc), vc);
    s.w1 = uc;
    s.w2 = vc;
    // forward.cu:375-378: 1.0*(1-tau) + ratio*tau  |  0.0*(1-tau) + ratio*tau   (double literals)
    const float rt = __fmul_rn(s.area, tau);
    s.ratio = (s.code == 0) ? (float)((double)__fsub_rn(1.0f, tau) + (double)rt) : rt;
    if (s.ratio == 0.0f) return false;
    const float4 c0 = *reinterpret_cast<const float4*>(rec + R_COL);        // c0.rgb c1.r
    const float4 c1 = *reinterpret_cast<const float4*>(rec + R_COL + 4);    // c1.gb c2.rg
    const float c2b = rec[R_COL + 8];
    const float intense = rec[R_MISC + 2];
    const float col0[3] = {c0.x, c0.y, c0.z}, col1[3] = {c0.w, c1.x, c1.y}, col2[3] = {c1.z, c1.w, c2b};
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float c = __fmaf_rn(s.w2, col2[ch], __fmaf_rn(s.w0, col0[ch], __fmul_rn(s.w1, col1[ch])));
        s.iC[ch] = __fmul_rn(c, intense);
    }
    s.iD = __fmaf_rn(s.w2, a2.w, __fmaf_rn(s.w0, a0.w, __fmul_rn(s.w1, a1.w)));
    return true;
}

// ------------------------------------------------------------------------------------------------
// Sub-chunk processing (16 candidates at a time), shared by the forward walk and backward phase A:
//   1. lock-step classification -> per-lane hit / partial bit masks (+ the 4-bit corner mask);
//   2. the warp compacts all (lane, candidate) pairs that need the polygon clip into a pair list;
//   3. the clip runs DENSE over that list - any lane clips any pixel's pair (the area does not depend
//      on the compositing order), so all 32 lanes are busy instead of the ~12 that own a partially
//      covered pixel at a given step; areas land in a per-warp shared-memory table;
//   4. the sequential per-pixel walk then only looks its areas up.
constexpr int kSub = 16;
constexpr int kPairCap = 32 * kSub;
struct WarpClip {
    uint16_t plist[kPairCap];          // k | lane << 4 | inmask << 9
    float parea[kPairCap];             // overlap area, or -1 when the reference would skip the pair
};

struct SubMasks {
    uint32_t hit, part;                // bit k: candidate k of the sub-chunk touches / partially covers my pixel
    int base_reg;                      // lane k holds the pair-list base of candidate k ...
    uint32_t mask_reg;                 // ... and the ballot of lanes with a partial pixel for it
};

template <typename SlotOk>
__device__ __forceinline__ void subchunk_prepare(const float* __restrict__ srec, const uint8_t* __restrict__ cand,
                                                 int ncs, bool use_aa, float pxmin, float pymin, bool lane_on,
                                                 float fx0, float fy0, WarpClip& wc, int lane, SubMasks& sm_,
                                                 SlotOk slot_ok) {
    uint32_t hit = 0u, part = 0u;
    unsigned long long im = 0ull;
    for (int k = 0; k < ncs; k++) {                            // lock-step, broadcast reads
        const int j = cand[k];
        int cls = AA_FULL;
        unsigned inmask = 0xF;
        if (use_aa) cls = aa_classify(srec + j * kRecStride, pxmin, pymin, inmask);
        const bool on = lane_on && (cls != AA_OUT) && slot_ok(j);
        const bool p = on && (cls == AA_PART);
        hit |= (on ? 1u : 0u) << k;
        part |= (p ? 1u : 0u) << k;
        if (p) im |= (unsigned long long)inmask << (4 * k);
    }
    int total = 0, base_reg = 0;
    uint32_t mask_reg = 0u;
    if (__any_sync(0xffffffffu, part != 0u)) {
        for (int k = 0; k < ncs; k++) {
            const bool p = (part >> k) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, p);
            if (lane == k) { base_reg = total; mask_reg = m; }
            if (p) wc.plist[total + __popc(m & ((1u << lane) - 1u))] =
                       (uint16_t)(k | (lane << 4) | ((unsigned)((im >> (4 * k)) & 15ull) << 9));
            total += __popc(m);
        }
        __syncwarp();
        for (int pid = lane; pid < total; pid += 32) {         // dense clip
            const unsigned e = wc.plist[pid];
            const int k = e & 15, l = (e >> 4) & 31;
            float area;
            const bool ok = aa_clip<false>(srec + cand[k] * kRecStride, fx0 + (float)(l & 7), fy0 + (float)(l >> 3),
                                           e >> 9, area, nullptr);
            wc.parea[pid] = ok ? area : -1.0f;
        }
        __syncwarp();
    }
    sm_.hit = hit; sm_.part = part; sm_.base_reg = base_reg; sm_.mask_reg = mask_reg;
}

// Pops this lane's next hit of the sub-chunk (-1: none) and fetches its coverage area.
// Must be called by all 32 lanes.  ok == false: the reference skips this (pixel, face) pair.
__device__ __forceinline__ int subchunk_pop(SubMasks& sm_, bool allowed, const WarpClip& wc, int lane, bool use_aa,
                                            float& area, bool& ok, bool& is_part) {
    int k = -1;
    if (allowed && sm_.hit) { k = __ffs(sm_.hit) - 1; sm_.hit &= sm_.hit - 1u; }
    const int kk = k < 0 ? 0 : k;
    const int b = __shfl_sync(0xffffffffu, sm_.base_reg, kk);
    const uint32_t m = __shfl_sync(0xffffffffu, sm_.mask_reg, kk);
    area = use_aa ? 1.0f : 0.0f;
    ok = true;
    is_part = false;
    if (k >= 0 && ((sm_.part >> k) & 1u)) {
        is_part = true;
        area = wc.parea[b + __popc(m & ((1u << lane) - 1u))];
        ok = !(area == -1.0f);
    }
    return k;
}

struct TileCtx {
    int b, tile_x, tile_y, warp, lane, px, py;
    bool inside;
    float fx0, fy0;
    size_t pix;
};

__device__ __forceinline__ TileCtx tile_ctx(const DM2RenderInputs& in, int gx, int gy) {
    TileCtx c;
    const int tiles = gx * gy;
    c.b = blockIdx.x / tiles;
    const int t_in = blockIdx.x % tiles;
    c.tile_x = t_in % gx; c.tile_y = t_in / gx;
    c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
    c.px = c.tile_x * kTile + (c.warp & 1) * 8 + (c.lane & 7);
    c.py = c.tile_y * kTile + (c.warp >> 1) * 4 + (c.lane >> 3);
    c.inside = (c.px < in.W) && (c.py < in.H);
    const uint32_t pmx = (uint32_t)in.patch_min[2 * c.b], pmy = (uint32_t)in.patch_min[2 * c.b + 1];
    c.fx0 = (float)((uint32_t)(c.tile_x * kTile + (c.warp & 1) * 8) + pmx);
    c.fy0 = (float)((uint32_t)(c.tile_y * kTile + (c.warp >> 1) * 4) + pmy);
    c.pix = c.inside ? ((size_t)c.b * in.H + c.py) * in.W + c.px : 0;
    return c;
}

// index of this pixel's ray: patch-local (B,H,W,3) buffers, or the renderer's full-resolution
// (V,ray_h,ray_w,3) buffers addressed through ray_view + patch_min (fused path)
__device__ __forceinline__ size_t ray_index(const DM2RenderInputs& in, const TileCtx& tc) {
    if (in.ray_view == nullptr) return tc.pix;
    const int pmx = in.patch_min[2 * tc.b], pmy = in.patch_min[2 * tc.b + 1];
    return ((size_t)in.ray_view[tc.b] * in.ray_h + (tc.py + pmy)) * in.ray_w + (tc.px + pmx);
}

}  // namespace dm2
