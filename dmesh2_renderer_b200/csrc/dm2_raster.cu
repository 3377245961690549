// dm2_raster.cu — per-tile forward compositor and reverse-order backward pass.
// Behaviour of reference cuda_impl/forward.cu:139-432 (renderCUDA<3>) and
// cuda_impl/backward.cu:17-532, re-designed for sm_100a.
//
// Common structure (one CTA per 16x16 tile, one warp per 8x4 pixel block, one pixel per lane):
//   * the tile's sorted face list is consumed in rounds of kStage packed 256-byte records copied
//     global->shared with 16-byte cp.async (no register staging); records sit at a 272-byte stride
//     so lanes reading DIFFERENT faces spread over the banks;
//   * per round each warp ballots "face bbox overlaps my 8x4 footprint" into a compact candidate
//     list; then every LANE walks that list with its own cursor: a cheap scan (bbox + 12 half-plane
//     tests) until it finds the next face that touches ITS pixel, after which the warp reconverges
//     for the expensive part (polygon clip, ray/triangle, blend) with most lanes active on different
//     faces.  The reference walks all faces in lock-step with all 256 threads.
//   * no AA cache: the backward pass recomputes the overlap area (the reference zero-fills and
//     round-trips a B*H*W*K*32-byte cache, render.cu:143-146).
//
// Backward (the reference issues 29 global atomics per (pixel, face) contribution,
// backward.cu:494-529).  Here each round alternates two phases:
//   A  pixel-sequential: the per-lane cursor walk (back to front) recovers T and computes
//      dL/dalpha for each contribution - the only quantities that depend on the compositing order -
//      and pushes a 12-byte record (face slot, pixel, T, dL/dalpha) into a shared-memory queue
//      (at most kMaxPush per lane per phase, so the queue is bounded);
//   B  face-parallel: the queue is counting-sorted by face slot; one THREAD per contribution
//      recomputes the order-independent terms (coverage + its gradient, hit point, weights) and the
//      29 per-face gradient values; each warp transposes its 32x29 values through shared memory and
//      lane k sums column k over runs of equal face slot, carrying the sum across iterations, so a
//      (tile, face) pair costs ONE coalesced 29-lane atomic into a per-(view,face) 128-byte gradient
//      record.  A final kernel scatters those records to the six output tensors.
#include "dm2_common.cuh"
#include "dm2_aa.cuh"

namespace dm2 {

constexpr int kStage = 128;                 // faces staged per round
constexpr int kRecStride = 68;              // words between staged records (272 B, 16-byte aligned)
constexpr int kRasterThreads = 256;
constexpr int kMaxPush = 8;                 // backward: contributions a lane may queue per A phase
constexpr int kQueue = kRasterThreads * kMaxPush;
constexpr int kGradWords = 32;              // per-(view,face) gradient record (29 used)
constexpr int kPixWords = 11;               // staged per-pixel data in backward (10 used; odd stride)

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

// Per-lane bit masks over the warp's candidate list: which candidates touch this lane's pixel, and
// which of those only partially (need the polygon clip).  Popped lowest position first = list order.
struct HitMasks {
    uint32_t hit[kStage / 32], part[kStage / 32];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int ch = 0; ch < kStage / 32; ch++) hit[ch] = 0u;
    }
    __device__ __forceinline__ bool any() const {
        uint32_t a = 0;
#pragma unroll
        for (int ch = 0; ch < kStage / 32; ch++) a |= hit[ch];
        return a != 0;
    }
    __device__ __forceinline__ int pop(bool& is_part) {
        int cpos = -1;
        is_part = false;
#pragma unroll
        for (int ch = 0; ch < kStage / 32; ch++) {
            if (cpos < 0 && hit[ch]) {
                const int b = __ffs(hit[ch]) - 1;
                cpos = ch * 32 + b;
                is_part = (part[ch] >> b) & 1u;
                hit[ch] &= hit[ch] - 1u;
            }
        }
        return cpos;
    }
};

// `limit`: candidates whose slot is >= limit are ignored (backward: behind the pixel's last contributor
// after the slot -> list-position mapping has been applied by the caller through `slot_ok`).
template <typename SlotOk>
__device__ __forceinline__ void classify_candidates(const float* __restrict__ srec, const uint8_t* __restrict__ cand,
                                                    int n_cand, bool use_aa, float pxmin, float pymin, bool lane_on,
                                                    HitMasks& hm, SlotOk slot_ok) {
#pragma unroll
    for (int ch = 0; ch < kStage / 32; ch++) {
        uint32_t h = 0u, p = 0u;
        const int lim = min(32, n_cand - ch * 32);            // warp-uniform
        for (int k = 0; k < lim; k++) {
            const int j = cand[ch * 32 + k];                  // broadcast
            int cls = AA_FULL;
            unsigned inmask;
            if (use_aa) cls = aa_classify(srec + j * kRecStride, pxmin, pymin, inmask);
            const bool on = lane_on && (cls != AA_OUT) && slot_ok(j);
            h |= (on ? 1u : 0u) << k;
            p |= ((on && cls == AA_PART) ? 1u : 0u) << k;
        }
        hm.hit[ch] = h;
        hm.part[ch] = p;
    }
}
__device__ __forceinline__ void classify_candidates(const float* __restrict__ srec, const uint8_t* __restrict__ cand,
                                                    int n_cand, bool use_aa, float pxmin, float pymin, bool lane_on,
                                                    HitMasks& hm) {
    classify_candidates(srec, cand, n_cand, use_aa, pxmin, pymin, lane_on, hm, [](int) { return true; });
}

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

// =============================================================================== forward
__global__ void __launch_bounds__(kRasterThreads, 3)
k_raster_forward(DM2RenderInputs in, const float* __restrict__ records, const uint32_t* __restrict__ list,
                 const uint2* __restrict__ ranges, float* __restrict__ final_T, float* __restrict__ final_prev_T,
                 uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
                 int gx, int gy) {
    extern __shared__ __align__(16) float smem_f[];
    float* srec = smem_f;
    uint8_t* scand = reinterpret_cast<uint8_t*>(srec + kStage * kRecStride);
    const TileCtx tc = tile_ctx(in, gx, gy);
    const int warp = tc.warp, lane = tc.lane;
    const float tau = in.aa_temperature;
    const bool use_aa = tau > 0.0f;
    uint8_t* cand = scand + warp * kStage;
    WarpClip& wclip = reinterpret_cast<WarpClip*>(scand + 8 * kStage)[warp];

    PixelGeom pg;
    pg.pxmin = (float)((uint32_t)tc.px + (uint32_t)in.patch_min[2 * tc.b]);      // forward.cu:319-322
    pg.pymin = (float)((uint32_t)tc.py + (uint32_t)in.patch_min[2 * tc.b + 1]);
    pg.o = f3{0, 0, 0}; pg.d = f3{0, 0, 0};
    if (tc.inside) {
        const size_t rix = ray_index(in, tc);
        const float* ro = in.image_ray_o + rix * 3;
        const float* rd = in.image_ray_d + rix * 3;
        pg.o = f3{ro[0], ro[1], ro[2]};
        pg.d = f3{rd[0], rd[1], rd[2]};
    }

    const uint2 range = ranges[blockIdx.x];
    const int total = (int)(range.y - range.x);
    const size_t view_base = (size_t)tc.b * in.F;

    bool done = !tc.inside;
    float T = 1.0f, pT = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t last_contributor = 0;

    for (int base = 0; base < total; base += kStage) {
        // reference forward.cu:257-260: stop when the whole tile is done
        if (__syncthreads_count(done) == kRasterThreads) break;
        const int cnt = min(kStage, total - base);
        stage_records(srec, records, list, (long long)range.x + base, 1, cnt, view_base);
        __syncthreads();
        if (__all_sync(0xffffffffu, done)) continue;
        const int n_cand = build_candidates(srec, cnt, use_aa, tc.fx0, tc.fy0, cand, lane);
        for (int sub = 0; sub < n_cand; sub += kSub) {
            if (__all_sync(0xffffffffu, done)) break;
            const uint8_t* csub = cand + sub;
            SubMasks sbm;
            subchunk_prepare(srec, csub, min(kSub, n_cand - sub), use_aa, pg.pxmin, pg.pymin, !done, tc.fx0, tc.fy0,
                             wclip, lane, sbm, [](int) { return true; });
            for (;;) {
                // ---- every lane takes its own next hit (in list order); lanes work on different faces
                float area;
                bool ok, is_part;
                const int k = subchunk_pop(sbm, !done, wclip, lane, use_aa, area, ok, is_part);
                if (!__any_sync(0xffffffffu, k >= 0)) break;
                if (k >= 0 && ok) {
                    const int jj = csub[k];
                    const float* rec = srec + jj * kRecStride;
                    Shade s;
                    s.area = area;
                    if (shade_tail(rec, pg, tau, s)) {
                        // forward.cu:398-415
                        const float alpha = __fmul_rn(rec[R_MISC + 1], s.ratio);
                        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                        C0 = __fmaf_rn(__fmul_rn(s.iC[0], alpha), T, C0);   // C += iC*alpha*T
                        C1 = __fmaf_rn(__fmul_rn(s.iC[1], alpha), T, C1);
                        C2 = __fmaf_rn(__fmul_rn(s.iC[2], alpha), T, C2);
                        D  = __fmaf_rn(__fmul_rn(s.iD, alpha), T, D);
                        pT = T;
                        T = test_T;
                        last_contributor = (uint32_t)(base + jj + 1);
                        if (T < kTEps) done = true;
                    }
                }
            }
        }
    }

    if (tc.inside) {
        final_prev_T[tc.pix] = pT;
        final_T[tc.pix] = T;
        n_contrib[tc.pix] = last_contributor;
        out_color[tc.pix * 3 + 0] = __fmaf_rn(T, in.background[0], C0);
        out_color[tc.pix * 3 + 1] = __fmaf_rn(T, in.background[1], C1);
        out_color[tc.pix * 3 + 2] = __fmaf_rn(T, in.background[2], C2);
        out_depth[tc.pix] = __fadd_rn(D, T);
    }
}

constexpr size_t kFwdSmem = (size_t)kStage * kRecStride * 4 + 8 * kStage + 8 * sizeof(WarpClip);

int launch_raster_forward(const DM2RenderInputs& in, const FaceWs& fw, const BinWs& bw, const ImageWs& iw,
                          float* out_color, float* out_depth, cudaStream_t s) {
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    const long long nblk = (long long)in.B * gx * gy;
    if (nblk == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        attr_set = true;
    }
    ProfScope _ps("raster_forward", s);
    k_raster_forward<<<(unsigned)nblk, kRasterThreads, kFwdSmem, s>>>(in, fw.records, bw.vals, iw.ranges, iw.final_T,
                                                                      iw.final_prev_T, iw.n_contrib, out_color,
                                                                      out_depth, gx, gy);
    DM2_LAUNCH_CHECK();
    return 0;
}

// =============================================================================== backward
// reference cuda_impl/auxiliary.h:245-290 (note: `denom_inv` is taken BEFORE the max(), so the
// clamp is inert; and the "v" derivative is built from dot(cross(T,E1),E2) exactly as upstream).
__device__ __forceinline__ f3 scale3(const f3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 add3(const f3& a, const f3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 subn3(const f3& a, const f3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 neg3(const f3& a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 crossn(const f3& a, const f3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float dotn(const f3& a, const f3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// a*s - t*b per component with the contraction nvcc applies to `x*y - z*w`
__device__ __forceinline__ f3 fms3(const f3& a, float s, float t, const f3& b) {
    return {__fmaf_rn(a.x, s, -__fmul_rn(t, b.x)), __fmaf_rn(a.y, s, -__fmul_rn(t, b.y)),
            __fmaf_rn(a.z, s, -__fmul_rn(t, b.z))};
}
__device__ __forceinline__ f3 mul3(const f3& a, float s) { return {__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)}; }

// The big gradients this produces are ill-conditioned (1/denom^2 with cancelling terms), so the
// operation order is pinned to the reference's expressions with its FMA contraction pattern: results
// then agree with the reference (and the oracle) far better than free-form re-association would.
__device__ __forceinline__ void ray_tri_grad(const f3& o, const f3& d, const f3& p0, const f3& p1, const f3& p2,
                                             f3& du_dp0, f3& du_dp1, f3& du_dp2, f3& dv_dp0, f3& dv_dp1,
                                             f3& dv_dp2) {
    const f3 T = sub3(o, p0), E1 = sub3(p1, p0), E2 = sub3(p2, p0);
    const f3 PE = cross3(d, E2);
    const float denom_sqrt = dot3(PE, E1);
    const float denom = __fmul_rn(denom_sqrt, denom_sqrt);
    const float denom_inv = __fdiv_rn(1.0f, denom);
    const float v0 = dot3(PE, T);
    const float v1 = denom_sqrt;
    const float v2 = dot3(cross3(T, E1), E2);
    const f3 du_dE1 = mul3(mul3(neg3(PE), v0), denom_inv);
    const f3 du_dE2 = mul3(fms3(cross3(T, d), v1, v0, cross3(E1, d)), denom_inv);
    const f3 du_dT  = mul3(mul3(PE, v1), denom_inv);
    const f3 dv_dE1 = mul3(fms3(cross3(E2, T), v1, v2, PE), denom_inv);
    const f3 dv_dE2 = mul3(fms3(cross3(T, E1), v1, v2, cross3(E1, d)), denom_inv);
    const f3 dv_dT  = mul3(mul3(cross3(E1, E2), v1), denom_inv);
    du_dp0 = sub3(sub3(neg3(du_dE1), du_dE2), du_dT);
    dv_dp0 = sub3(sub3(neg3(dv_dE1), dv_dE2), dv_dT);
    du_dp1 = du_dE1; dv_dp1 = dv_dE1;
    du_dp2 = du_dE2; dv_dp2 = dv_dE2;
}
// dL_du * du + dL_dv * dv  ->  fma(dL_du, du, dL_dv * dv)
__device__ __forceinline__ f3 comb3(float a, const f3& x, float b, const f3& y) {
    return {__fmaf_rn(a, x.x, __fmul_rn(b, y.x)), __fmaf_rn(a, x.y, __fmul_rn(b, y.y)), __fmaf_rn(a, x.z, __fmul_rn(b, y.z))};
}

struct BwdSmem {
    float srec[kStage * kRecStride];
    float spix[kRasterThreads * kPixWords];      // o.xyz d.xyz dC.rgb dD per pixel of the tile
    float red[8][32 * 33];                       // per-warp 32x29 transpose tile (stride 33)
    uint32_t q_meta[kQueue];                     // slot | pixel << 8
    float q_T[kQueue];
    float q_g[kQueue];                           // dL/dalpha
    uint16_t perm[kQueue];
    int cnt[kStage], off[kStage], cur[kStage];
    int rslot[8][32];
    int qn;
    int max_last[8];
    uint8_t cand[8][kStage];
};

__global__ void __launch_bounds__(kRasterThreads, 2)
k_raster_backward(DM2RenderInputs in, const float* __restrict__ records, const uint32_t* __restrict__ list,
                  const uint2* __restrict__ ranges, const float* __restrict__ final_T,
                  const float* __restrict__ final_prev_T, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                  float* __restrict__ gface, int gx, int gy) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
    const TileCtx tc = tile_ctx(in, gx, gy);
    const int warp = tc.warp, lane = tc.lane, tid = threadIdx.x;
    const float tau = in.aa_temperature;
    const bool use_aa = tau > 0.0f;
    const uint32_t pmx = (uint32_t)in.patch_min[2 * tc.b], pmy = (uint32_t)in.patch_min[2 * tc.b + 1];
    const int tile_px0 = tc.tile_x * kTile, tile_py0 = tc.tile_y * kTile;

    const uint2 range = ranges[blockIdx.x];
    const size_t view_base = (size_t)tc.b * in.F;

    // ---- per-pixel state (backward.cu:133-160); the pixel's constants are also published to smem
    PixelGeom pg;
    pg.pxmin = (float)((uint32_t)tc.px + pmx);
    pg.pymin = (float)((uint32_t)tc.py + pmy);
    pg.o = f3{0, 0, 0}; pg.d = f3{0, 0, 0};
    float dC[3] = {0, 0, 0}, dD = 0.0f;
    float T_final = 0.0f, prev_T_final = 0.0f;
    int my_last = 0;
    if (tc.inside) {
        const size_t rix = ray_index(in, tc);
        const float* ro = in.image_ray_o + rix * 3;
        const float* rd = in.image_ray_d + rix * 3;
        pg.o = f3{ro[0], ro[1], ro[2]};
        pg.d = f3{rd[0], rd[1], rd[2]};
        dC[0] = dL_dcolor[tc.pix * 3 + 0]; dC[1] = dL_dcolor[tc.pix * 3 + 1]; dC[2] = dL_dcolor[tc.pix * 3 + 2];
        dD = dL_ddepth[tc.pix];
        T_final = final_T[tc.pix];
        prev_T_final = final_prev_T[tc.pix];
        my_last = (int)n_contrib[tc.pix];
    }
    {
        float* sp = sm.spix + tid * kPixWords;
        sp[0] = pg.o.x; sp[1] = pg.o.y; sp[2] = pg.o.z; sp[3] = pg.d.x; sp[4] = pg.d.y; sp[5] = pg.d.z;
        sp[6] = dC[0]; sp[7] = dC[1]; sp[8] = dC[2]; sp[9] = dD;
    }
    float T = prev_T_final;
    bool first_pass = true;
    float accum_rec[3] = {0, 0, 0}, accum_recd = 0.0f;
    float last_alpha = 0.0f, last_color[3] = {0, 0, 0}, last_depth = 0.0f;
    float bg_dot = 0.0f;
    bg_dot += in.background[0] * dC[0]; bg_dot += in.background[1] * dC[1]; bg_dot += in.background[2] * dC[2];
    const float bd_dot = dD;

    // faces behind every pixel's last contributor are never touched (backward.cu:219-221)
    const int warp_last = __reduce_max_sync(0xffffffffu, my_last);
    if (lane == 0) sm.max_last[warp] = warp_last;
    if (tid == 0) sm.qn = 0;
    for (int i = tid; i < kStage; i += kRasterThreads) { sm.cnt[i] = 0; sm.cur[i] = 0; }
    __syncthreads();
    int block_last = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) block_last = max(block_last, sm.max_last[i]);

    uint8_t* cand = sm.cand[warp];

    for (int hi = block_last; hi > 0; hi -= kStage) {
        const int cnt = min(kStage, hi);
        __syncthreads();
        // slot i holds list entry (hi-1-i): back to front
        stage_records(sm.srec, records, list, (long long)range.x + hi - 1, -1, cnt, view_base);
        __syncthreads();
        int n_cand = 0;
        if (hi - cnt < warp_last) n_cand = build_candidates(sm.srec, cnt, use_aa, tc.fx0, tc.fy0, cand, lane);
        // lock-step classification; faces behind this pixel's last contributor are masked out
        HitMasks hm;
        classify_candidates(sm.srec, cand, n_cand, use_aa, pg.pxmin, pg.pymin, tc.inside, hm,
                            [&](int j) { return hi - 1 - j < my_last; });

        for (;;) {        // alternate phase A (bounded) and phase B until the round is exhausted
            // ====================== phase A: pixel-sequential walk, queue the contributions
            int pushes = 0;
            for (;;) {
                bool is_part = false;
                int cpos = -1;
                if (pushes < kMaxPush) cpos = hm.pop(is_part);
                if (!__any_sync(0xffffffffu, cpos >= 0)) break;
                bool push = false;
                float rec_T = 0.f, rec_g = 0.f;
                int jj = 0;
                if (cpos >= 0) {
                    jj = cand[cpos];
                    const float* rec = sm.srec + jj * kRecStride;
                    Shade s;
                    if (shade<false>(rec, pg, tau, is_part, s, nullptr)) {
                        const float alpha = __fmul_rn(rec[R_MISC + 1], s.ratio);
                        // backward.cu:340-348
                        if (!first_pass) T = T / (1.0f - alpha);
                        first_pass = false;
                        float dL_dalpha = 0.0f;
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
                            last_color[ch] = s.iC[ch];
                            dL_dalpha += (s.iC[ch] - accum_rec[ch]) * dC[ch];
                        }
                        accum_recd = last_alpha * last_depth + (1.0f - last_alpha) * accum_recd;
                        last_depth = s.iD;
                        dL_dalpha += (s.iD - accum_recd) * dD;
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        if (alpha == 1.0f) {
                            dL_dalpha += (-prev_T_final) * bg_dot;
                            dL_dalpha += (-prev_T_final) * bd_dot;
                        } else {
                            dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                            dL_dalpha += (-T_final / (1.0f - alpha)) * bd_dot;
                        }
                        push = true; rec_T = T; rec_g = dL_dalpha;
                    }
                }
                // warp-aggregated queue allocation
                const unsigned pm = __ballot_sync(0xffffffffu, push);
                if (pm) {
                    int qbase = 0;
                    if (lane == __ffs(pm) - 1) qbase = atomicAdd(&sm.qn, __popc(pm));
                    qbase = __shfl_sync(0xffffffffu, qbase, __ffs(pm) - 1);
                    if (push) {
                        const int q = qbase + __popc(pm & ((1u << lane) - 1u));
                        sm.q_meta[q] = (uint32_t)jj | ((uint32_t)tid << 8) | ((is_part ? 1u : 0u) << 16);
                        sm.q_T[q] = rec_T;
                        sm.q_g[q] = rec_g;
                        atomicAdd(&sm.cnt[jj], 1);
                        pushes++;
                    }
                }
            }
            const bool lane_more = hm.any();
            __syncthreads();
            const int qn = sm.qn;
            // ====================== phase B: counting sort by slot, one thread per contribution
            if (qn > 0) {
                if (warp == 0) {        // exclusive scan of the per-slot counts (kStage = 128 = 4 per lane)
                    int v[kStage / 32], s = 0;
#pragma unroll
                    for (int k = 0; k < kStage / 32; k++) { v[k] = sm.cnt[lane * (kStage / 32) + k]; s += v[k]; }
                    int inc = s;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                    int run = inc - s;
#pragma unroll
                    for (int k = 0; k < kStage / 32; k++) { sm.off[lane * (kStage / 32) + k] = run; run += v[k]; }
                }
                __syncthreads();
                for (int i = tid; i < qn; i += kRasterThreads) {
                    const int j = sm.q_meta[i] & 0xffu;
                    const int p = sm.off[j] + atomicAdd(&sm.cur[j], 1);
                    sm.perm[p] = (uint16_t)i;
                }
                __syncthreads();
                // warp w owns the contiguous slice [lo, hi_p) of the sorted queue
                const int lo = (int)(((long long)qn * warp) / 8), hi_p = (int)(((long long)qn * (warp + 1)) / 8);
                int cur_slot = -1;
                float carry = 0.0f;
                float* red = sm.red[warp];
                for (int p0 = lo; p0 < hi_p; p0 += 32) {
                    const int p = p0 + lane;
                    float v[29];
#pragma unroll
                    for (int k = 0; k < 29; k++) v[k] = 0.0f;
                    int slot = -1;
                    if (p < hi_p) {
                        const int i = sm.perm[p];
                        const uint32_t meta = sm.q_meta[i];
                        slot = (int)(meta & 0xffu);
                        const int ptid = (int)((meta >> 8) & 0xffu);
                        const float Tq = sm.q_T[i], dL_dalpha = sm.q_g[i];
                        const float* rec = sm.srec + slot * kRecStride;
                        const float* sp = sm.spix + ptid * kPixWords;
                        // pixel of thread `ptid` (same mapping as tile_ctx)
                        const int pw = ptid >> 5, pl = ptid & 31;
                        PixelGeom qg;
                        qg.pxmin = (float)((uint32_t)(tile_px0 + (pw & 1) * 8 + (pl & 7)) + pmx);
                        qg.pymin = (float)((uint32_t)(tile_py0 + (pw >> 1) * 4 + (pl >> 3)) + pmy);
                        qg.o = f3{sp[0], sp[1], sp[2]};
                        qg.d = f3{sp[3], sp[4], sp[5]};
                        const float qdC[3] = {sp[6], sp[7], sp[8]}, qdD = sp[9];
                        const bool is_part = (meta >> 16) & 1u;
                        Shade s;
                        float g[6];
                        if (shade<true>(rec, qg, tau, is_part, s, g)) {
                            const float opacity = rec[R_MISC + 1], intense = rec[R_MISC + 2];
                            const float alpha = __fmul_rn(opacity, s.ratio);
                            float dL_diC[3];
#pragma unroll
                            for (int ch = 0; ch < 3; ch++) dL_diC[ch] = qdC[ch] * alpha * Tq;
                            const float dL_diD = qdD * alpha * Tq;
                            // backward.cu:408-415
                            v[21] = dL_dalpha * s.ratio;
                            const float dL_doarea = (dL_dalpha * opacity) * tau;
#pragma unroll
                            for (int k = 0; k < 6; k++) v[23 + k] = dL_doarea * g[k];
                            // backward.cu:420-455
                            const float4 a0 = *reinterpret_cast<const float4*>(rec + R_P0);
                            const float4 a1 = *reinterpret_cast<const float4*>(rec + R_P1);
                            const float4 a2 = *reinterpret_cast<const float4*>(rec + R_P2);
                            const float4 q0 = *reinterpret_cast<const float4*>(rec + R_COL);
                            const float4 q1 = *reinterpret_cast<const float4*>(rec + R_COL + 4);
                            const float q2b = rec[R_COL + 8];
                            const float col0[3] = {q0.x, q0.y, q0.z}, col1[3] = {q0.w, q1.x, q1.y},
                                        col2[3] = {q1.z, q1.w, q2b};
                            float di0 = 0.f, di1 = 0.f, di2 = 0.f, dint = 0.f;
#pragma unroll
                            for (int ch = 0; ch < 3; ch++) {
                                di0 += col0[ch] * dL_diC[ch] * intense;
                                di1 += col1[ch] * dL_diC[ch] * intense;
                                di2 += col2[ch] * dL_diC[ch] * intense;
                                v[9 + ch]  = s.w0 * dL_diC[ch] * intense;
                                v[12 + ch] = s.w1 * dL_diC[ch] * intense;
                                v[15 + ch] = s.w2 * dL_diC[ch] * intense;
                                dint += (s.w0 * col0[ch] + s.w1 * col1[ch] + s.w2 * col2[ch]) * dL_diC[ch];
                            }
                            v[22] = dint;
                            di0 += a0.w * dL_diD; di1 += a1.w * dL_diD; di2 += a2.w * dL_diD;
                            v[18] = s.w0 * dL_diD; v[19] = s.w1 * dL_diD; v[20] = s.w2 * dL_diD;
                            // backward.cu:460-488, auxiliary.h:331-357
                            float duc_du, duc_dv = 0.f, dvc_du = 0.f, dvc_dv;
                            if (s.code == 0) { duc_du = 1.f; dvc_dv = 1.f; }
                            else if (s.code <= 3) { duc_du = 0.f; dvc_dv = 0.f; }
                            else if (s.code == 4) { duc_du = 0.f; dvc_dv = 1.f; }
                            else if (s.code == 5) { duc_du = 1.f; dvc_dv = 0.f; }
                            else { duc_du = 0.5f; dvc_du = -0.5f; duc_dv = -0.5f; dvc_dv = 0.5f; }
                            const float di0_du = -duc_du - dvc_du, di0_dv = -duc_dv - dvc_dv;
                            const float dL_du = di0 * di0_du + di1 * duc_du + di2 * dvc_du;
                            const float dL_dv = di0 * di0_dv + di1 * duc_dv + di2 * dvc_dv;
                            f3 du0, du1, du2, dv0, dv1, dv2;
                            ray_tri_grad(qg.o, qg.d, f3{a0.x, a0.y, a0.z}, f3{a1.x, a1.y, a1.z},
                                         f3{a2.x, a2.y, a2.z}, du0, du1, du2, dv0, dv1, dv2);
                            const f3 g0 = comb3(dL_du, du0, dL_dv, dv0);
                            const f3 g1 = comb3(dL_du, du1, dL_dv, dv1);
                            const f3 g2 = comb3(dL_du, du2, dL_dv, dv2);
                            v[0] = g0.x; v[1] = g0.y; v[2] = g0.z;
                            v[3] = g1.x; v[4] = g1.y; v[5] = g1.z;
                            v[6] = g2.x; v[7] = g2.y; v[8] = g2.z;
                        }
                    }
                    // ---- transpose through smem; lane k sums column k over runs of equal slot
                    sm.rslot[warp][lane] = slot;
#pragma unroll
                    for (int k = 0; k < 29; k++) red[lane * 33 + k] = v[k];
                    __syncwarp();
                    if (lane < 29) {
#pragma unroll 4
                        for (int r = 0; r < 32; r++) {
                            const int sr = sm.rslot[warp][r];
                            if (sr != cur_slot) {
                                if (cur_slot >= 0 && carry != 0.0f) {
                                    const int fid = __float_as_int(sm.srec[cur_slot * kRecStride + R_FID]);
                                    atomicAdd(gface + (view_base + fid) * kGradWords + lane, carry);
                                }
                                cur_slot = sr;
                                carry = 0.0f;
                            }
                            if (sr >= 0) carry += red[r * 33 + lane];
                        }
                    }
                    __syncwarp();
                }
                if (lane < 29 && cur_slot >= 0 && carry != 0.0f) {
                    const int fid = __float_as_int(sm.srec[cur_slot * kRecStride + R_FID]);
                    atomicAdd(gface + (view_base + fid) * kGradWords + lane, carry);
                }
                __syncthreads();
                // reset the queue bookkeeping for the next phase
                if (tid == 0) sm.qn = 0;
                for (int i = tid; i < kStage; i += kRasterThreads) { sm.cnt[i] = 0; sm.cur[i] = 0; }
            }
            if (!__syncthreads_or(lane_more)) break;
        }
    }
}

// Scatter the per-(view,face) gradient records to the reference's six gradient tensors
// (render.cu:313-318).  Outputs are caller-zeroed; per-face outputs have a single writer.
__global__ void __launch_bounds__(256)
k_grad_scatter(int B, int P, int F, const int32_t* __restrict__ faces, const uint32_t* __restrict__ tiles_touched,
               const float* __restrict__ gface, DM2RenderGrads gr) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * F) return;
    if (tiles_touched[idx] == 0) return;
    const int b = (int)(idx / F), f = (int)(idx % F);
    const float4* g4 = reinterpret_cast<const float4*>(gface + (size_t)idx * kGradWords);
    float g[32];
#pragma unroll
    for (int k = 0; k < 8; k++) { const float4 t = g4[k]; g[4 * k] = t.x; g[4 * k + 1] = t.y; g[4 * k + 2] = t.z; g[4 * k + 3] = t.w; }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int v = faces[3 * (size_t)f + i];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (g[3 * i + c] != 0.0f) atomicAdd(gr.dL_dverts + 3 * (size_t)v + c, g[3 * i + c]);
            if (g[9 + 3 * i + c] != 0.0f) atomicAdd(gr.dL_dverts_color + 3 * (size_t)v + c, g[9 + 3 * i + c]);
        }
        if (g[18 + i] != 0.0f) atomicAdd(gr.dL_dverts_ndc + 3 * ((size_t)b * P + v) + 2, g[18 + i]);
    }
    if (g[21] != 0.0f) atomicAdd(gr.dL_dfaces_opacity + f, g[21]);
    gr.dL_dfaces_intense[idx] = g[22];
#pragma unroll
    for (int k = 0; k < 6; k++) gr.dL_daa_face_verts[(size_t)idx * 6 + k] = g[23 + k];
}

// Fused-path scatter: same per-face record, but the screen-space part (d/d aa verts, d/d ndc.z) is
// chained through the CCW swap, the image transform and the perspective divide down to
// d/d verts_proj (what torch autograd does for the reference's host prep, __init__.py:251-260 and
// pyrenderer.py:6-30), with one 16-byte vector reduction per (view, face, vertex).
__device__ __forceinline__ void red_add_v4(float* p, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__global__ void __launch_bounds__(256)
k_fused_grad_scatter(DM2FusedInputs in, const uint32_t* __restrict__ tiles_touched, const float* __restrict__ records,
                     const float* __restrict__ gface, DM2FusedGrads gr) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)in.B * in.F) return;
    if (tiles_touched[idx] == 0) return;
    const int b = (int)(idx / in.F), f = (int)(idx % in.F);
    const float4* g4 = reinterpret_cast<const float4*>(gface + (size_t)idx * kGradWords);
    float g[32];
#pragma unroll
    for (int k = 0; k < 8; k++) { const float4 t = g4[k]; g[4 * k] = t.x; g[4 * k + 1] = t.y; g[4 * k + 2] = t.z; g[4 * k + 3] = t.w; }
    const bool flip = (__float_as_uint(records[(size_t)idx * kRecWords + R_MISC + 3]) >> 8) & 1u;
    const float4* vp = reinterpret_cast<const float4*>(in.verts_proj) + (size_t)b * in.P;
    const float fw = (float)in.image_w, fh = (float)in.image_h;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int v = in.faces[3 * (size_t)f + i];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (g[3 * i + c] != 0.0f) atomicAdd(gr.dL_dverts + 3 * (size_t)v + c, g[3 * i + c]);
            if (g[9 + 3 * i + c] != 0.0f) atomicAdd(gr.dL_dverts_color + 3 * (size_t)v + c, g[9 + 3 * i + c]);
        }
        // aa verts are (v0, v2, v1) when the face was flipped to CCW
        const int k = (i == 0) ? 0 : (flip ? 3 - i : i);
        const float dix = (k == 0) ? g[23] : ((k == 1) ? g[25] : g[27]);
        const float diy = (k == 0) ? g[24] : ((k == 1) ? g[26] : g[28]);
        const float dz = g[18 + i];
        if (dix != 0.0f || diy != 0.0f || dz != 0.0f) {
            const float4 c = vp[v];
            float w = c.w;
            bool clamped = false;
            if (w >= 0.0f && w < 1e-4f) { w = 1e-4f; clamped = true; }
            if (w < 0.0f && w > -1e-4f) { w = -1e-4f; clamped = true; }
            const float nx = c.x / w, ny = c.y / w, nz = c.z / w;
            const float dnx = (dix * fw) * 0.5f, dny = (diy * fh) * 0.5f;
            const float dw = clamped ? 0.0f : -(dnx * nx + dny * ny + dz * nz) / w;
            red_add_v4(gr.dL_dverts_proj + 4 * ((size_t)b * in.P + v), dnx / w, dny / w, dz / w, dw);
        }
    }
    if (g[21] != 0.0f) atomicAdd(gr.dL_dfaces_opacity + f, g[21]);
    gr.dL_dfaces_intense[idx] = g[22];
}

size_t backward_ws_bytes(long long BF) { return align_up((size_t)BF * kGradWords * 4, 256) + 256; }

int launch_fused_grad_scatter(const DM2FusedInputs& in, const FaceWs& fw, const DM2FusedGrads& g, void* grad_ws,
                              cudaStream_t s) {
    const long long BF = (long long)in.B * in.F;
    if (BF == 0) return 0;
    ProfScope _ps("grad_scatter", s);
    k_fused_grad_scatter<<<cdiv(BF, 256), 256, 0, s>>>(in, fw.tiles_touched, fw.records,
                                                       reinterpret_cast<const float*>(grad_ws), g);
    DM2_LAUNCH_CHECK();
    return 0;
}

int launch_grad_scatter(const DM2RenderInputs& in, const FaceWs& fw, const DM2RenderGrads& g, void* grad_ws,
                        cudaStream_t s) {
    const long long BF = (long long)in.B * in.F;
    if (BF == 0) return 0;
    ProfScope _ps("grad_scatter", s);
    k_grad_scatter<<<cdiv(BF, 256), 256, 0, s>>>(in.B, in.P, in.F, in.faces, fw.tiles_touched,
                                                 reinterpret_cast<const float*>(grad_ws), g);
    DM2_LAUNCH_CHECK();
    return 0;
}

// zero the per-face gradient records and run the tile kernel; the caller scatters afterwards
int launch_raster_backward(const DM2RenderInputs& in, const FaceWs& fw, const BinWs& bw, const ImageWs& iw,
                           const float* dL_dcolor, const float* dL_ddepth, void* grad_ws, cudaStream_t s) {
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    const long long nblk = (long long)in.B * gx * gy;
    const long long BF = (long long)in.B * in.F;
    if (nblk == 0 || BF == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        DM2_CHECK(cudaFuncSetAttribute(k_raster_backward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(BwdSmem)));
        attr_set = true;
    }
    float* gface = reinterpret_cast<float*>(grad_ws);
    {
        ProfScope _ps("grad_zero", s);
        DM2_CHECK(cudaMemsetAsync(gface, 0, (size_t)BF * kGradWords * 4, s));
    }
    {
        ProfScope _ps("raster_backward", s);
        k_raster_backward<<<(unsigned)nblk, kRasterThreads, sizeof(BwdSmem), s>>>(
            in, fw.records, bw.vals, iw.ranges, iw.final_T, iw.final_prev_T, iw.n_contrib, dL_dcolor, dL_ddepth,
            gface, gx, gy);
        DM2_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace dm2
