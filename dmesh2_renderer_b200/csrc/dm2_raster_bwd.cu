// dm2_raster_bwd.cu — reverse-order backward pass of the tile compositor.
// Behaviour of reference cuda_impl/backward.cu:17-532 (29 global atomics per (pixel, face)
// contribution, 232 registers, a B*H*W*K*32-byte AA cache), re-designed for sm_100a.
//
// Warp-autonomous like the forward kernel (dm2_raster_common.cuh): one warp per 8x4-pixel footprint, one
// pixel per lane, the tile's sorted list scanned back to front.  Per sub-chunk of kBSub candidates a warp does:
//   1. per-lane hit / partial bit masks from the staged classification words;
//   2. a DENSE pass over the candidate-major list of all hit pairs: the (uc, vc, ratio, code) record the forward
//      pass left in the shading cache (recomputed - clip, hit point, clamp - only for blocks that are not
//      cached), interpolation, alpha and S = iC . dL/dC + iD . dL/dD into [candidate][pixel] tables;
//   3. the pixel-sequential walk (each lane its own hits, back to front): recovers T (T /= 1 - alpha)
//      and computes dL/dalpha - the only order-dependent quantities - into the same tables;
//   4. the face-parallel gradient phase over the same list: one THREAD per contribution evaluates the 29
//      per-face gradient values (pixel constants are fetched from the owning lane with shuffles, the record and
//      d(area)/d(aa verts) from the shading cache); the warp transposes them through shared
//      memory, lane q sums column q over the run of one candidate and issues ONE coalesced 29-lane atomic
//      per (warp, face) into the per-(view, face) 128-byte gradient record.
// A second kernel scatters the records to the output tensors (or, on the fused path, chains them
// through the perspective divide to d/d verts).
#include "dm2_raster_common.cuh"

namespace dm2 {

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


// Recomputation for pairs whose block is not in the shading cache (cache off, aa_temperature == 0, or pool
// overflow): kept out of line so that the rarely taken path does not sit in the instruction stream of the hot
// loops (the kernel showed "no instruction" fetch stalls with the clip inlined twice).  Returns the hit record
// the forward pass would have written: (uc, vc, ratio, code), ratio == 0 = skip.
__device__ __noinline__ float4 shade_recompute(const float* rec, float ox, float oy, float oz, float dx, float dy,
                                               float dz, float pxmin, float pymin, float tau, int l, bool partial,
                                               bool use_aa, bool exact_clip) {
    Shade s;
    s.area = use_aa ? 1.0f : 0.0f;
    if (partial) {
        float a;
        if (!aa_overlap<false>(rec, pxmin, pymin, exact_clip,
                               [&]() { return corner_mask_of(*reinterpret_cast<const uint4*>(rec + R_META), l); }, a, nullptr))
            return make_float4(0.f, 0.f, 0.f, 0.f);
        s.area = a;
    }
    PixelGeom qg;
    qg.pxmin = pxmin; qg.pymin = pymin;
    qg.o = f3{ox, oy, oz}; qg.d = f3{dx, dy, dz};
    if (!shade_tail(rec, qg, tau, s)) return make_float4(0.f, 0.f, 0.f, 0.f);
    return make_float4(s.w1, s.w2, s.ratio, __int_as_float(s.code));
}
__device__ __noinline__ void clip_gradient(const float* rec, float pxmin, float pymin, unsigned inmask, bool exact_clip, float* g) {
    float a, gg[6];
    aa_overlap<true>(rec, pxmin, pymin, exact_clip, [&]() { return inmask; }, a, gg);
#pragma unroll
    for (int q = 0; q < 6; q++) g[q] = gg[q];
}

// Transpose tile of 16 rows (contributions) x 32 columns (29 gradient values + padding).  Columns are
// stored in groups of four - element (q, row) at (q >> 2) * 68 + row * 4 + (q & 3) - so that a row is
// written with 8 STS.128 and a column is read at a fixed stride of 4 words; the group stride of 68 words
// makes both bank-conflict free.
constexpr int kRedGroup = 68;
constexpr int kRedWords = 8 * kRedGroup;

// Candidates per sub-chunk.  Measured at cfg2 (ms): 8 candidates / 6 CTAs of 4 warps per SM 1.86;
// 16 / 4 CTAs 1.69; 12 / 5 CTAs 1.66 - the dense passes want many pairs per sub-chunk (~110 here, 4
// batches), the latency-bound walk wants warps.
// (same-box A/B with the final kernels: 11 candidates 1.271, 12 1.251, 13 1.240 - 41.5 KB per CTA still fit 5 times)
#ifndef DM2_BSUB
#define DM2_BSUB 13
#endif
#ifndef DM2_BWD_MINBLOCKS
#define DM2_BWD_MINBLOCKS 5
#endif
constexpr int kBSub = DM2_BSUB;
constexpr int kBSlots = kBSub * 32;

struct BwdWarpSmem {
    float srec[kBSub * kRecStride];         // staged records of the sub-chunk
    float tA[kBSlots];                      // [candidate][pixel]: alpha (or skip)  -> T before the contribution
    float tS[kBSlots];                      //                     iC.dC + iD.dD     -> dL/dalpha
    float red[kRedWords];                   // half-warp transpose tile (layout: red_index)
    WarpQueue q;
    uint16_t plist[kBSlots];                // candidate-major list of all hit pairs: k | lane << 4 | partial << 9
};

__global__ void __launch_bounds__(kRasterThreads, DM2_BWD_MINBLOCKS)
k_raster_backward(DM2RenderInputs in, const float* __restrict__ records, const uint32_t* __restrict__ list,
                  const uint8_t* __restrict__ imask, const uint4* __restrict__ meta, uint32_t R, ShadeCache cache,
                  const uint2* __restrict__ ranges,
                  const float* __restrict__ final_T, const float* __restrict__ final_prev_T,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor,
                  const float* __restrict__ dL_ddepth, float* __restrict__ gface, int gx, int gy) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TileCtx tc = tile_ctx(in, gx, gy);
    const int warp = tc.warp, lane = tc.lane;
    BwdWarpSmem& ws = reinterpret_cast<BwdWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const float tau = in.aa_temperature;
    const bool use_aa = tau > 0.0f;
    const unsigned lt_mask = (1u << lane) - 1u;

    const uint2 range = ranges[tc.tile];
    const size_t view_base = (size_t)tc.b * in.F;
    const uint32_t* tlist = list + range.x;
    const uint8_t* tmask = imask + range.x;
    const uint4* tmeta = use_aa ? meta + (size_t)warp * R + range.x : nullptr;

    // ---- per-pixel state (backward.cu:133-160)
    PixelGeom pg;
    pg.pxmin = (float)((uint32_t)tc.px + (uint32_t)in.patch_min[2 * tc.b]);
    pg.pymin = (float)((uint32_t)tc.py + (uint32_t)in.patch_min[2 * tc.b + 1]);
    pg.o = f3{0, 0, 0}; pg.d = f3{0, 0, 0};
    float dC0 = 0.f, dC1 = 0.f, dC2 = 0.f, dD = 0.f;
    float T_final = 0.0f, prev_T_final = 0.0f;
    int my_last = 0;
    if (tc.inside) {
        load_ray(in, tc, pg.o, pg.d);
        dC0 = dL_dcolor[tc.pix * 3 + 0]; dC1 = dL_dcolor[tc.pix * 3 + 1]; dC2 = dL_dcolor[tc.pix * 3 + 2];
        dD = dL_ddepth[tc.pix];
        if (in.depth_unit) dD = -0.5f * dD;      // d(1 - (z + 1) / 2)/dz, exact
        T_final = final_T[tc.pix];
        prev_T_final = final_prev_T[tc.pix];
        my_last = (int)n_contrib[tc.pix];
    }
    float T = prev_T_final;
    bool first_pass = true;
    // backward.cu:357-405 keeps a running blend `accum_rec` per channel and sums (c - accum_rec) * dC
    // over the channels; the blend is linear, so the same quantity is carried here as ONE scalar
    // A = accum_rec . dC with S = c . dC per face (dense pass):  dL/dalpha = (S - A) * T + background term.
    float accA = 0.0f, last_alpha = 0.0f, last_S = 0.0f;
    float bg_dot = 0.0f;
    bg_dot += in.background[0] * dC0; bg_dot += in.background[1] * dC1; bg_dot += in.background[2] * dC2;
    const float bd_dot = dD;

    // faces behind every pixel's last contributor are never touched (backward.cu:219-221)
    const int warp_last = __reduce_max_sync(0xffffffffu, my_last);
    const bool cached = cache.capacity != 0u;

    int head = 0, count = 0, cursor = warp_last;
    for (;;) {
        fill_queue<true, kBSub>(ws.q, head, count, cursor, warp_last, tmask, tlist, warp, lane);
        if (count == 0) break;
        const int ncs = min(count, kBSub);
        stage_candidates(ws.srec, records, tmeta, ws.q, head, ncs, view_base, lane, count - ncs);

        if (use_aa) store_partial_masks(ws.srec, ncs, lane);
        // the candidates' cache blocks (written by the forward kernel: DRAM by now) are pulled into L2 while the
        // masks and the pair list are built: lane k asks for candidate k's block, one request per 128-byte line
        if (cached && lane < ncs) {
            const float* rec = ws.srec + lane * kRecStride;
            const uint32_t poff = __float_as_uint(rec[R_META + 3]);
            if (poff != kNoCache) {
                const uint32_t bytes = 16u * cache_block_granules(__popc(__float_as_uint(rec[R_META])),
                                                                  __popc(__float_as_uint(rec[R_PMASK])));
                const char* base = reinterpret_cast<const char*>(cache.data + (size_t)poff);
                for (uint32_t off = 0; off < bytes; off += 128u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
            }
        }

        // ====== 1. per-lane hit / partial masks from the classification words
        uint32_t hit, part;
        classify_subchunk(ws.srec, ncs, use_aa, lane, hit, part, [&](int k) {
            return tc.inside && (int)ws.q.idx[(head + k) & (kQueue - 1)] < my_last;
        });
        if (__any_sync(0xffffffffu, hit != 0u)) {
            // candidate-major list of all hit pairs (k | lane << 4 | partial << 9); the dense pass and the
            // gradient pass both run over it
            int npairs = 0;
            for (int k = 0; k < ncs; k++) {
                const bool h = (hit >> k) & 1u;
                const unsigned m = __ballot_sync(0xffffffffu, h);
                if (h) ws.plist[npairs + __popc(m & lt_mask)] = (uint16_t)(k | (lane << 4) | (((part >> k) & 1u) << 9));
                npairs += __popc(m);
            }
            __syncwarp();

            // ====== 2. dense pass over every hit pair: hit record (cached, else recomputed), alpha, S
            for (int p0 = 0; p0 < npairs; p0 += 32) {
                const int pid = p0 + lane;
                const bool act = pid < npairs;
                const unsigned e = act ? ws.plist[pid] : 0u;
                const int k = e & 15, l = (e >> 4) & 31;
                const float* rec = ws.srec + k * kRecStride;
                const int slot = k * 32 + l;
                // the record's load is issued first; the shuffles below cover part of its latency
                const uint32_t poff = act ? __float_as_uint(rec[R_META + 3]) : kNoCache;
                const bool from_cache = cached && poff != kNoCache;
                float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                if (from_cache)
                    h = __ldg(cache.data + (size_t)poff + __popc(__float_as_uint(rec[R_META]) & ((1u << l) - 1u)));
                const float qC0 = __shfl_sync(0xffffffffu, dC0, l), qC1 = __shfl_sync(0xffffffffu, dC1, l);
                const float qC2 = __shfl_sync(0xffffffffu, dC2, l), qD = __shfl_sync(0xffffffffu, dD, l);
                if (!__all_sync(0xffffffffu, from_cache || !act)) {
                    const float ox = __shfl_sync(0xffffffffu, pg.o.x, l), oy = __shfl_sync(0xffffffffu, pg.o.y, l);
                    const float oz = __shfl_sync(0xffffffffu, pg.o.z, l), dx = __shfl_sync(0xffffffffu, pg.d.x, l);
                    const float dy = __shfl_sync(0xffffffffu, pg.d.y, l), dz = __shfl_sync(0xffffffffu, pg.d.z, l);
                    if (act && !from_cache)
                        h = shade_recompute(rec, ox, oy, oz, dx, dy, dz, tc.fx0 + (float)(l & 7), tc.fy0 + (float)(l >> 3),
                                            tau, l, (e >> 9) & 1u, use_aa, cache.exact_clip != 0u);
                }
                if (!act) continue;
                if (h.z != 0.0f) {
                    Shade s;
                    s.w1 = h.x; s.w2 = h.y;
                    s.w0 = __fsub_rn(__fsub_rn(1.0f, h.x), h.y);
                    shade_interp(rec, s);
                    ws.tA[slot] = __fmul_rn(rec[R_MISC + 1], h.z);
                    ws.tS[slot] = s.iC[0] * qC0 + s.iC[1] * qC1 + s.iC[2] * qC2 + s.iD * qD;
                } else {
                    ws.tA[slot] = __uint_as_float(kSkipBits);
                }
            }
            __syncwarp();

            // ====== 3. pixel-sequential walk (back to front): T and dL/dalpha per contribution
            for (uint32_t hh = hit; hh != 0u; hh &= hh - 1u) {
                const int k = __ffs(hh) - 1;
                const int slot = k * 32 + lane;
                const float alpha = ws.tA[slot];
                if (__float_as_uint(alpha) == kSkipBits) continue;
                const float S = ws.tS[slot];
                // backward.cu:340-348: T = T / (1 - alpha); the same reciprocal serves the background term
                const float r1 = __frcp_rn(1.0f - alpha);
                if (!first_pass) T = T * r1;
                first_pass = false;
                // backward.cu:357-405
                accA = last_alpha * last_S + (1.0f - last_alpha) * accA;
                last_S = S;
                float dL_dalpha = (S - accA) * T;
                last_alpha = alpha;
                const float bgw = (alpha == 1.0f) ? -prev_T_final : -T_final * r1;
                dL_dalpha += bgw * bg_dot;
                dL_dalpha += bgw * bd_dot;
                ws.tA[slot] = T;
                ws.tS[slot] = dL_dalpha;
            }
            __syncwarp();

            // ====== 4. face-parallel gradients over the same list (skipped pairs contribute zero rows)
            const int total_c = npairs;
            float carry = 0.0f;
            for (int r0 = 0; r0 < total_c; r0 += 32) {
                const int r = r0 + lane;
                const bool in_list = r < total_c;
                const unsigned e = in_list ? ws.plist[r] : 0u;
                const int k = e & 15, src = (e >> 4) & 31;
                const bool act = in_list && (__float_as_uint(ws.tA[k * 32 + src]) != kSkipBits);
                // last row of its candidate's run?
                const bool run_end = in_list && ((r + 1 == total_c) || ((ws.plist[r + 1 < kBSlots ? r + 1 : r] & 15) != k));
                const unsigned flush = __ballot_sync(0xffffffffu, run_end);
                // pixel constants of the owning lane
                PixelGeom qg;
                qg.o.x = __shfl_sync(0xffffffffu, pg.o.x, src); qg.o.y = __shfl_sync(0xffffffffu, pg.o.y, src);
                qg.o.z = __shfl_sync(0xffffffffu, pg.o.z, src); qg.d.x = __shfl_sync(0xffffffffu, pg.d.x, src);
                qg.d.y = __shfl_sync(0xffffffffu, pg.d.y, src); qg.d.z = __shfl_sync(0xffffffffu, pg.d.z, src);
                const float qC0 = __shfl_sync(0xffffffffu, dC0, src), qC1 = __shfl_sync(0xffffffffu, dC1, src);
                const float qC2 = __shfl_sync(0xffffffffu, dC2, src), qD = __shfl_sync(0xffffffffu, dD, src);
                float v[29];
#pragma unroll
                for (int q = 0; q < 29; q++) v[q] = 0.0f;
                if (act) {
                    const float* rec = ws.srec + k * kRecStride;
                    const int slot = k * 32 + src;
                    qg.pxmin = tc.fx0 + (float)(src & 7);
                    qg.pymin = tc.fy0 + (float)(src >> 3);
                    const uint32_t poff = __float_as_uint(rec[R_META + 3]);
                    const bool from_cache = cached && poff != kNoCache;
                    const uint32_t touched = __float_as_uint(rec[R_META]);
                    // hit record and d(area)/d(aa verts) of a partially covered pixel: straight from the shading
                    // cache (global; the loads are issued first and consumed last), else recomputed
                    float4 h;
                    float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (from_cache) {
                        h = __ldg(cache.data + (size_t)poff + __popc(touched & ((1u << src) - 1u)));
                        if ((e >> 9) & 1u) {
                            const uint32_t j = __popc(__float_as_uint(rec[R_PMASK]) & ((1u << src) - 1u));
                            const float2* gj = reinterpret_cast<const float2*>(cache.data + (size_t)poff + __popc(touched)) + 3 * j;
                            const float2 g0 = __ldg(gj), g1 = __ldg(gj + 1), g2 = __ldg(gj + 2);
                            g[0] = g0.x; g[1] = g0.y; g[2] = g1.x; g[3] = g1.y; g[4] = g2.x; g[5] = g2.y;
                        }
                    } else {
                        h = shade_recompute(rec, qg.o.x, qg.o.y, qg.o.z, qg.d.x, qg.d.y, qg.d.z, qg.pxmin, qg.pymin, tau,
                                            src, (e >> 9) & 1u, use_aa, cache.exact_clip != 0u);
                        if ((e >> 9) & 1u) {
                            const unsigned inm = corner_mask_of(*reinterpret_cast<const uint4*>(rec + R_META), src);
                            float gt[6];
                            clip_gradient(rec, qg.pxmin, qg.pymin, inm, cache.exact_clip != 0u, gt);
#pragma unroll
                            for (int q = 0; q < 6; q++) g[q] = gt[q];
                        }
                    }
                    Shade s;
                    s.w1 = h.x; s.w2 = h.y; s.ratio = h.z; s.code = __float_as_int(h.w);
                    s.w0 = __fsub_rn(__fsub_rn(1.0f, h.x), h.y);
                    const float Tq = ws.tA[slot], dL_dalpha = ws.tS[slot];
                    const float opacity = rec[R_MISC + 1], intense = rec[R_MISC + 2];
                    const float alpha = __fmul_rn(opacity, s.ratio);
                    const float dL_diC[3] = {qC0 * alpha * Tq, qC1 * alpha * Tq, qC2 * alpha * Tq};
                    const float dL_diD = qD * alpha * Tq;
                    // backward.cu:408-415
                    v[21] = dL_dalpha * s.ratio;
                    // backward.cu:420-455
                    const float4 a0 = *reinterpret_cast<const float4*>(rec + R_P0);
                    const float4 a1 = *reinterpret_cast<const float4*>(rec + R_P1);
                    const float4 a2 = *reinterpret_cast<const float4*>(rec + R_P2);
                    const float4 q0 = *reinterpret_cast<const float4*>(rec + R_COL);
                    const float4 q1 = *reinterpret_cast<const float4*>(rec + R_COL + 4);
                    const float q2b = rec[R_COL + 8];
                    const float col0[3] = {q0.x, q0.y, q0.z}, col1[3] = {q0.w, q1.x, q1.y}, col2[3] = {q1.z, q1.w, q2b};
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
                    ray_tri_grad(qg.o, qg.d, f3{a0.x, a0.y, a0.z}, f3{a1.x, a1.y, a1.z}, f3{a2.x, a2.y, a2.z},
                                 du0, du1, du2, dv0, dv1, dv2);
                    const f3 g0 = comb3(dL_du, du0, dL_dv, dv0);
                    const f3 g1 = comb3(dL_du, du1, dL_dv, dv1);
                    const f3 g2 = comb3(dL_du, du2, dL_dv, dv2);
                    v[0] = g0.x; v[1] = g0.y; v[2] = g0.z;
                    v[3] = g1.x; v[4] = g1.y; v[5] = g1.z;
                    v[6] = g2.x; v[7] = g2.y; v[8] = g2.z;
                    const float dL_doarea = (dL_dalpha * opacity) * tau;
#pragma unroll
                    for (int q = 0; q < 6; q++) v[23 + q] = dL_doarea * g[q];
                }
                // ---- transpose through smem (two half-warps); lane c sums column c over the rows of one
                //      candidate's run (rows are candidate-major) and issues ONE coalesced 29-lane atomic
                //      per (warp, face) into the per-(view,face) gradient record
                const int nrows = min(32, total_c - r0);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (h * 16 >= nrows) break;
                    if ((lane >> 4) == h) {
                        const int rw = lane & 15;
#pragma unroll
                        for (int g4 = 0; g4 < 7; g4++)
                            *reinterpret_cast<float4*>(&ws.red[g4 * kRedGroup + rw * 4]) =
                                make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
                        *reinterpret_cast<float4*>(&ws.red[7 * kRedGroup + rw * 4]) = make_float4(v[28], 0.f, 0.f, 0.f);
                    }
                    __syncwarp();
                    // rows past the end of the list were stored as zeros (v = 0 for inactive lanes): all 16
                    // rows of the half are summed, run boundaries come from the warp-uniform flush mask
                    const float* colp = ws.red + (lane >> 2) * kRedGroup + (lane & 3);
                    const unsigned fl = (flush >> (h * 16)) & 0xFFFFu;
                    float x[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) x[j] = colp[j * 4];
#pragma unroll
                    for (int j0 = 0; j0 < 16; j0 += 4) {
                        if (((fl >> j0) & 0xFu) == 0u) {             // no run ends in these 4 rows: tree sum
                            carry += (x[j0] + x[j0 + 1]) + (x[j0 + 2] + x[j0 + 3]);
                            continue;
                        }
#pragma unroll
                        for (int j = j0; j < j0 + 4; j++) {
                            carry += x[j];
                            if ((fl >> j) & 1u) {                    // warp-uniform
                                const int kk = __shfl_sync(0xffffffffu, k, h * 16 + j);
                                const int fid = __float_as_int(ws.srec[kk * kRecStride + R_FID]);
                                if (lane < 29 && carry != 0.0f) atomicAdd(gface + (view_base + fid) * kGradWords + lane, carry);
                                carry = 0.0f;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
        }
        head = (head + ncs) & (kQueue - 1);
        count -= ncs;
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

#ifndef DM2_GS_MINB
#define DM2_GS_MINB 4
#endif
__global__ void __launch_bounds__(256, DM2_GS_MINB)
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
    const float* mvb = in.mv + 16 * (size_t)b;
    const float* prb = in.proj + 16 * (size_t)b;
    const float fw = (float)in.image_w, fh = (float)in.image_h;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int v = in.faces[3 * (size_t)f + i];
        float dv[3] = {g[3 * i], g[3 * i + 1], g[3 * i + 2]};        // hit-point path (d/d world-space vertex)
        // aa verts are (v0, v2, v1) when the face was flipped to CCW
        const int k = (i == 0) ? 0 : (flip ? 3 - i : i);
        const float dix = (k == 0) ? g[23] : ((k == 1) ? g[25] : g[27]);
        const float diy = (k == 0) ? g[24] : ((k == 1) ? g[26] : g[28]);
        const float dz = g[18 + i];
        if (dix != 0.0f || diy != 0.0f || dz != 0.0f) {
            // image -> ndc -> clip (w clamp blocks the gradient, as torch.where does) -> view -> world
            const float4 c = clip_vertex(in.verts + 3 * (size_t)v, mvb, prb);
            float w = c.w;
            bool clamped = false;
            if (w >= 0.0f && w < 1e-4f) { w = 1e-4f; clamped = true; }
            if (w < 0.0f && w > -1e-4f) { w = -1e-4f; clamped = true; }
            const float nx = c.x / w, ny = c.y / w, nz = c.z / w;
            const float dnx = (dix * fw) * 0.5f, dny = (diy * fh) * 0.5f;
            const float dc[4] = {dnx / w, dny / w, dz / w, clamped ? 0.0f : -(dnx * nx + dny * ny + dz * nz) / w};
            float dview[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                dview[q] = dc[0] * prb[q] + dc[1] * prb[4 + q] + dc[2] * prb[8 + q] + dc[3] * prb[12 + q];
#pragma unroll
            for (int q = 0; q < 3; q++)
                dv[q] += dview[0] * mvb[q] + dview[1] * mvb[4 + q] + dview[2] * mvb[8 + q] + dview[3] * mvb[12 + q];
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (dv[c] != 0.0f) atomicAdd(gr.dL_dverts + 3 * (size_t)v + c, dv[c]);
            if (g[9 + 3 * i + c] != 0.0f) atomicAdd(gr.dL_dverts_color + 3 * (size_t)v + c, g[9 + 3 * i + c]);
        }
    }
    if (g[21] != 0.0f) atomicAdd(gr.dL_dfaces_opacity + f, g[21]);
    gr.dL_dfaces_intense[idx] = g[22];
}

constexpr size_t kBwdSmem = sizeof(BwdWarpSmem) * (kRasterThreads / 32);

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
                           const float* dL_dcolor, const float* dL_ddepth, void* grad_ws, const void* shade_cache,
                           long long shade_cache_bytes, cudaStream_t s) {
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    const long long nblk = (long long)in.B * gx * gy;
    const long long BF = (long long)in.B * in.F;
    if (nblk == 0 || BF == 0) return 0;
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        DM2_CHECK(cudaFuncSetAttribute(k_raster_backward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kBwdSmem));
        attr_set.done();
    }
    float* gface = reinterpret_cast<float*>(grad_ws);
    {
        ProfScope _ps("grad_zero", s);
        DM2_CHECK(cudaMemsetAsync(gface, 0, (size_t)BF * kGradWords * 4, s));
    }
    {
        ProfScope _ps("raster_backward", s);
        ShadeCache cache = shade_cache_view(const_cast<void*>(shade_cache), shade_cache_bytes);
        cache.exact_clip = exact_clip_requested() ? 1u : 0u;
        k_raster_backward<<<(unsigned)(nblk * kCtasPerTile), kRasterThreads, kBwdSmem, s>>>(
            in, fw.records, bw.vals, bw.imask, reinterpret_cast<const uint4*>(bw.meta), (uint32_t)bw.R,
            cache, iw.ranges, iw.final_T, iw.final_prev_T,
            iw.n_contrib, dL_dcolor, dL_ddepth, gface, gx, gy);
        DM2_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace dm2
