// dm2_raster_common.cuh — device code shared by the forward (dm2_raster_fwd.cu) and backward
// (dm2_raster_bwd.cu) tile kernels.
//
// Both kernels are WARP-AUTONOMOUS: a CTA of four warps renders half of a 16x16 tile (the tile size is part
// of the reference's binning contract), one warp per 8x4-pixel footprint, one pixel per lane - and after the
// prologue no warp ever waits for another one (no CTA barrier in the main loop).  Each warp
//   * scans the tile's sorted instance list 32 entries at a time through the 1-byte per-instance footprint
//     mask written by k_classify_instances (bit w: the face touches a pixel of footprint w) and pushes the
//     candidates into a private ring queue;
//   * stages the packed 256-byte records of the next sub-chunk of candidates into its private shared memory
//     with 16-byte cp.async (15 consecutive lanes fetch one contiguous 240-byte record), together with the
//     candidate's 16-byte classification word for this footprint;
//   * turns the classification words into per-lane hit / partial bit masks (lane k first reduces candidate
//     k's corner grid to a "partially covered pixels" mask, then every lane extracts its own bit from two
//     broadcast words per candidate), compacts the partially covered (pixel, face) pairs into a pair list and
//     runs the expensive polygon clip DENSE over that list: any lane clips any pixel's pair;
//   * finally each lane walks its own hits in list order through the order-dependent part (forward: hit
//     point, barycentrics, blend, termination; backward: transmittance recovery and dL/dalpha).
// The first versions of these kernels staged records per CTA behind two barriers per round: ncu showed
// 21-27 % of all stall samples on that barrier.
#pragma once
#include "dm2_common.cuh"
#include "dm2_aa.cuh"

namespace dm2 {

constexpr int kRecStride = 68;              // words between staged records (272 B, 16-byte aligned)
constexpr int kWarpsPerCta = 4;              // a CTA renders half a tile (4 of its 8 footprints): the warps never
constexpr int kRasterThreads = 32 * kWarpsPerCta;   // talk to each other, smaller CTAs free their SM slots sooner
constexpr int kCtasPerTile = 8 / kWarpsPerCta;
constexpr int kGradWords = 32;              // per-(view,face) gradient record (29 used)
constexpr int kQueue = 64;                  // candidate ring capacity (>= sub-chunk - 1 + 32)
constexpr uint32_t kSkipBits = 0xFFFFFFFFu; // table marker: the reference skips this (pixel, face) pair

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// per-warp candidate ring
struct WarpQueue {
    uint32_t fid[kQueue];                   // face id within the view
    uint32_t idx[kQueue];                   // position in the tile's sorted list
};

// Scans the tile list (forward: ascending from `cursor`; REVERSE: descending, entries [0, cursor)
// remain) until the ring holds at least SUB candidates or the list is exhausted.
template <bool REVERSE, int SUB>
__device__ __forceinline__ void fill_queue(WarpQueue& q, int head, int& count, int& cursor, int total,
                                           const uint8_t* __restrict__ imask, const uint32_t* __restrict__ list,
                                           int warp, int lane) {
    static_assert(SUB - 1 + 32 <= kQueue, "candidate ring too small");
    while (count < SUB && (REVERSE ? cursor > 0 : cursor < total)) {
        const int i = REVERSE ? cursor - 1 - lane : cursor + lane;
        const bool valid = REVERSE ? i >= 0 : i < total;
        bool hit = false;
        uint32_t fid = 0;
        if (valid) {
            hit = (imask[i] >> warp) & 1u;
            fid = list[i];
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int pos = (head + count + __popc(m & ((1u << lane) - 1u))) & (kQueue - 1);
            q.fid[pos] = fid;
            q.idx[pos] = (uint32_t)i;
        }
        count += __popc(m);
        cursor += REVERSE ? -32 : 32;
    }
    __syncwarp();
}

// Stages the records of the first `ncs` queued candidates into slots 0..ncs-1; the 16th 16-byte chunk
// of a slot receives the candidate's classification words (k_classify_instances) for this warp's
// footprint.  tmeta: meta + warp * R + range.x (nullptr when aa_temperature == 0).
constexpr int R_META = 60;                  // staged slot only: touched-pixel mask, corner grid lo / hi, cache block offset
constexpr int R_PMASK = 64;                 // staged slot only: partially-covered-pixel mask of the footprint
constexpr uint32_t kNoCache = 0xFFFFFFFFu;  // cache offset: this candidate has no block in the shading cache

// ------------------------------------------------------------------------------------------------
// Shading cache (the role of the reference's four oarea buffers, render.cu:143-146 - a zero-filled
// B*H*W*K*32-byte array there - as a compact pool, and more: HBM is ~90 % idle while these kernels are
// instruction bound, so the forward pass also hands the backward pass everything it would otherwise
// recompute per contribution).  For every staged (instance, footprint) the forward pass reserves ONE block
//     [ n_t hit records : float4 (uc, vc, ratio, code) ]   n_t = touched pixels of the footprint
//     [ n_p x 6 floats  : d(area)/d(aa verts)         ]   n_p = partially covered pixels
// ordered by pixel inside each part; `ratio == 0` marks a pair the reference skips (AA error code, zero
// overlap, denom == 0, zero coverage).  The block's offset (in 16-byte granules) lives in the 4th
// classification word of the instance.  The backward pass reads the records back instead of re-running
// Moeller-Trumbore, the barycentric clamp and the polygon clip, and falls back to recomputation for
// blocks that did not fit (kNoCache) or when the cache is off.
// Pool layout: [u32 granule counter | pad to 256 B][capacity granules of 16 bytes].
struct ShadeCache {
    uint32_t* counter;
    float4* data;
    uint32_t capacity;                      // granules
    uint32_t exact_clip;                    // 1: every partially covered pair takes the exact polygon clip (dm2_set_exact_clip)
};
constexpr size_t kCacheHeader = 256;
__host__ __device__ __forceinline__ ShadeCache shade_cache_view(void* pool, long long bytes) {
    ShadeCache c;
    c.counter = reinterpret_cast<uint32_t*>(pool);
    c.data = reinterpret_cast<float4*>(reinterpret_cast<char*>(pool) + kCacheHeader);
    long long cap = pool && bytes > (long long)kCacheHeader ? (bytes - (long long)kCacheHeader) / 16 : 0;
    if (cap > (1ll << 32) - (1ll << 28)) cap = (1ll << 32) - (1ll << 28);       // the 32-bit counter must not wrap
    c.capacity = (uint32_t)cap;
    c.exact_clip = 0u;
    return c;
}
__device__ __forceinline__ uint32_t cache_block_granules(uint32_t n_t, uint32_t n_p) {
    return n_t + ((6u * n_p + 3u) >> 2);
}

// partially covered pixels of a footprint from its classification words (touched & not all 4 corners inside)
__device__ __forceinline__ uint32_t partial_mask(const uint4 m) {
    const unsigned long long C = ((unsigned long long)m.z << 32) | m.y;
    uint32_t full = 0u;
#pragma unroll
    for (int y = 0; y < 4; y++) {
        const uint32_t r0 = (uint32_t)(C >> (9 * y)) & 0x1FFu, r1 = (uint32_t)(C >> (9 * (y + 1))) & 0x1FFu;
        full |= (r0 & (r0 >> 1) & r1 & (r1 >> 1) & 0xFFu) << (8 * y);
    }
    return m.x & ~full;
}
// `ahead`: candidates already queued behind these (the next sub-chunk): their records (2 lines each) and
// classification words are requested into L2 now, a whole sub-chunk of work before they are staged.
__device__ __forceinline__ void stage_candidates(float* __restrict__ srec, const float* __restrict__ records,
                                                 const uint4* __restrict__ tmeta, const WarpQueue& q, int head, int ncs,
                                                 size_t view_base, int lane, int ahead = 0) {
#ifndef DM2_NO_STAGE_PREFETCH
    if (lane < ahead) {
        const int pos = (head + ncs + lane) & (kQueue - 1);
        const float* r = records + (view_base + q.fid[pos]) * kRecWords;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 32));
        if (tmeta) asm volatile("prefetch.global.L2 [%0];" ::"l"(tmeta + q.idx[pos]));
    }
#endif
    for (int t = lane; t < ncs * 16; t += 32) {
        const int slot = t >> 4, chunk = t & 15;
        const int pos = (head + slot) & (kQueue - 1);
        if (chunk == 15) {
            if (tmeta) cp_async16(srec + slot * kRecStride + R_META, tmeta + q.idx[pos]);
            continue;
        }
        cp_async16(srec + slot * kRecStride + chunk * 4, records + (view_base + q.fid[pos]) * kRecWords + chunk * 4);
    }
    cp_async_wait_all();
    __syncwarp();
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

// Interpolated colour and NDC depth from the barycentric weights s.w0..w2 (forward.cu:383-395).
__device__ __forceinline__ void shade_interp(const float* __restrict__ rec, Shade& s) {
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
    s.iD = __fmaf_rn(s.w2, rec[R_P2 + 3], __fmaf_rn(s.w0, rec[R_P0 + 3], __fmul_rn(s.w1, rec[R_P1 + 3])));
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
    shade_interp(rec, s);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Per-lane view of the staged classification words: bit k of hit / part = candidate k touches / only
// partially covers my pixel; im holds the 4-bit inside-corner mask (aa.h:103-149 corner order) of every
// partial candidate.  lane_ok(k): extra per-lane admission test (forward: pixel not done; backward:
// inside n_contrib).  With aa_temperature == 0 every admitted pixel is a hit and nothing is partial.
// 4-bit inside-corner mask of pixel `l` of the footprint from the staged classification words
__device__ __forceinline__ unsigned corner_mask_of(const uint4 m, int l) {
    const int sh = 9 * (l >> 3) + (l & 7);
    const uint32_t c = (uint32_t)((((unsigned long long)m.z << 32) | m.y) >> sh);
    return (c & 3u) | ((c >> 8) & 4u) | ((c >> 6) & 8u);      // (x,y) (x+1,y) (x+1,y+1) (x,y+1)
}

// The staged slots must carry the partial masks (store_partial_masks) when use_aa.
__device__ __forceinline__ void store_partial_masks(float* __restrict__ srec, int ncs, int lane) {
    if (lane < ncs)
        srec[lane * kRecStride + R_PMASK] =
            __uint_as_float(partial_mask(*reinterpret_cast<const uint4*>(srec + lane * kRecStride + R_META)));
    __syncwarp();
}

template <typename LaneOk>
__device__ __forceinline__ void classify_subchunk(const float* __restrict__ srec, int ncs, bool use_aa, int lane,
                                                  uint32_t& hit, uint32_t& part, LaneOk lane_ok) {
    hit = 0u; part = 0u;
    // bit `lane` of every candidate's touched-pixel and partially-covered-pixel masks (two broadcast reads)
    for (int k = 0; k < ncs; k++) {
        uint32_t touched = 0xFFFFFFFFu, pm = 0u;
        if (use_aa) {
            touched = __float_as_uint(srec[k * kRecStride + R_META]);
            pm = __float_as_uint(srec[k * kRecStride + R_PMASK]);
        }
        const uint32_t on = lane_ok(k) ? ((touched >> lane) & 1u) : 0u;
        hit |= on << k;
        part |= (on & (pm >> lane)) << k;
    }
}

struct TileCtx {
    int tile, b, tile_x, tile_y, warp, lane, px, py;
    bool inside;
    float fx0, fy0;
    size_t pix;
};

// context of footprint `footprint` (0..7) of global tile `tile` (view * tiles + tile in view) for this lane
__device__ __forceinline__ TileCtx tile_ctx_of(const DM2RenderInputs& in, int gx, int gy, int tile, int footprint) {
    TileCtx c;
    const int tiles = gx * gy;
    c.tile = tile;
    c.b = c.tile / tiles;
    const int t_in = c.tile % tiles;
    c.tile_x = t_in % gx; c.tile_y = t_in / gx;
    c.warp = footprint;
    c.lane = threadIdx.x & 31;
    c.px = c.tile_x * kTile + (c.warp & 1) * 8 + (c.lane & 7);
    c.py = c.tile_y * kTile + (c.warp >> 1) * 4 + (c.lane >> 3);
    c.inside = (c.px < in.W) && (c.py < in.H);
    const uint32_t pmx = (uint32_t)in.patch_min[2 * c.b], pmy = (uint32_t)in.patch_min[2 * c.b + 1];
    c.fx0 = (float)((uint32_t)(c.tile_x * kTile + (c.warp & 1) * 8) + pmx);
    c.fy0 = (float)((uint32_t)(c.tile_y * kTile + (c.warp >> 1) * 4) + pmy);
    c.pix = c.inside ? ((size_t)c.b * in.H + c.py) * in.W + c.px : 0;
    return c;
}
// static assignment: CTA blockIdx.x renders kWarpsPerCta footprints of tile blockIdx.x / kCtasPerTile
__device__ __forceinline__ TileCtx tile_ctx(const DM2RenderInputs& in, int gx, int gy) {
    return tile_ctx_of(in, gx, gy, blockIdx.x / kCtasPerTile,
                       (blockIdx.x % kCtasPerTile) * kWarpsPerCta + (threadIdx.x >> 5));
}

// (out of line: once per thread, and the tile kernels' register allocation must not depend on this prologue)
static __device__ __noinline__ void make_ray_outlined(const float* prm, int view, int x, int y, int W, int H, int V,
                                                      int recipe, float* out6) {
    f3 o, d;
    make_ray(prm, view, x, y, W, H, V, recipe, o, d);
    out6[0] = o.x; out6[1] = o.y; out6[2] = o.z; out6[3] = d.x; out6[4] = d.y; out6[5] = d.z;
}

// This pixel's ray: from patch-local (B,H,W,3) buffers, from the renderer's full-resolution
// (V,ray_h,ray_w,3) buffers addressed through ray_view + patch_min (fused path), or - ray_recipe > 0 -
// computed from the view's 32 ray parameters (make_ray: bit-identical to the buffers by construction,
// verified by the host before it selects a recipe).
__device__ __forceinline__ void load_ray(const DM2RenderInputs& in, const TileCtx& tc, f3& o, f3& d) {
    size_t rix = tc.pix;
    if (in.ray_view != nullptr) {
        const int pmx = in.patch_min[2 * tc.b], pmy = in.patch_min[2 * tc.b + 1];
        const int view = in.ray_view[tc.b];
#ifndef DM2_EXP_NO_KERNEL_RAYS
        if (in.ray_recipe > 0) {
            float r6[6];
            make_ray_outlined(in.ray_params + 32 * (size_t)view, view, tc.px + pmx, tc.py + pmy, in.ray_w, in.ray_h, in.ray_v,
                              in.ray_recipe, r6);
            o = f3{r6[0], r6[1], r6[2]};
            d = f3{r6[3], r6[4], r6[5]};
            return;
        }
#endif
        rix = ((size_t)view * in.ray_h + (tc.py + pmy)) * in.ray_w + (tc.px + pmx);
    }
    const float* ro = in.image_ray_o + rix * 3;
    const float* rd = in.image_ray_d + rix * 3;
    o = f3{ro[0], ro[1], ro[2]};
    d = f3{rd[0], rd[1], rd[2]};
}

}  // namespace dm2
