// dm2_raster_fwd.cu — per-tile forward compositor.
// Behaviour of reference cuda_impl/forward.cu:139-432 (renderCUDA<3>), re-designed for sm_100a.
// Structure: see dm2_raster_common.cuh (warp-autonomous candidate queue, dense pair evaluation).
// The reference walks every listed face in lock-step with all 256 threads of the tile behind a
// barrier per 256-face round and re-reads ~40 AA floats per (pixel, face) from global memory.
#include "dm2_raster_common.cuh"

namespace dm2 {

// Candidates per sub-chunk.  Measured at cfg2 (ms, CTAs of 4 warps): 12 candidates / 7 CTAs per SM 1.23;
// 16 / 6 CTAs 1.155; 20 / 5 CTAs 1.17; 24 / 4 CTAs 1.21.
constexpr int kFSub = 16;
constexpr int kFSlots = kFSub * 32;

struct FwdWarpSmem {
    float srec[kFSub * kRecStride];         // staged records of the sub-chunk
    float parea[kFSlots];                   // [candidate][pixel]: overlap area of a partially covered pixel (-1: skip)
    WarpQueue q;
    uint16_t plist[kFSlots];                // partially covered (pixel, face) pairs: k | lane << 4
};

// CACHE: fill the shading cache for the backward pass (dm2_raster_common.cuh): d(area)/d(aa verts) of every
// clipped pair and one (uc, vc, ratio, code) record per visited (pixel, face) pair.
template <bool CACHE>
#ifndef DM2_FWD_MINBLOCKS
#define DM2_FWD_MINBLOCKS (3 * kCtasPerTile)
#endif
__global__ void __launch_bounds__(kRasterThreads, DM2_FWD_MINBLOCKS)
k_raster_forward(DM2RenderInputs in, const float* __restrict__ records, const uint32_t* __restrict__ list,
                 const uint8_t* __restrict__ imask, uint4* __restrict__ meta, uint32_t R, ShadeCache cache,
                 BinGuard guard, const uint2* __restrict__ ranges, float* __restrict__ final_T,
                 float* __restrict__ final_prev_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                 float* __restrict__ out_depth, int gx, int gy) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (bin_invalid(guard)) return;          // speculative launch on a workspace that turned out too small
    const TileCtx tc = tile_ctx(in, gx, gy);
    const int warp = tc.warp, lane = tc.lane;
    FwdWarpSmem& ws = reinterpret_cast<FwdWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const float tau = in.aa_temperature;
    const bool use_aa = tau > 0.0f;
    const unsigned lt_mask = (1u << lane) - 1u;

    PixelGeom pg;
    pg.pxmin = (float)((uint32_t)tc.px + (uint32_t)in.patch_min[2 * tc.b]);      // forward.cu:319-322
    pg.pymin = (float)((uint32_t)tc.py + (uint32_t)in.patch_min[2 * tc.b + 1]);
    pg.o = f3{0, 0, 0}; pg.d = f3{0, 0, 0};
    if (tc.inside) {
        load_ray(in, tc, pg.o, pg.d);
    }

    const uint2 range = ranges[tc.tile];
    const int total = (int)(range.y - range.x);
    const size_t view_base = (size_t)tc.b * in.F;
    const uint32_t* tlist = list + range.x;
    const uint8_t* tmask = imask + range.x;
    uint4* tmeta = use_aa ? meta + (size_t)warp * R + range.x : nullptr;

    bool done = !tc.inside;
    float T = 1.0f, pT = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t last_contributor = 0;

    int head = 0, count = 0, cursor = 0;
    bool pool_full = false;
    while (!__all_sync(0xffffffffu, done)) {
        fill_queue<false, kFSub>(ws.q, head, count, cursor, total, tmask, tlist, warp, lane);
        if (count == 0) break;
        const int ncs = min(count, kFSub);
        stage_candidates(ws.srec, records, tmeta, ws.q, head, ncs, view_base, lane, count - ncs);

        if (use_aa) store_partial_masks(ws.srec, ncs, lane);
        if (CACHE && use_aa) {
            // reserve one cache block per staged candidate: a hit record for every touched pixel of the
            // footprint + the area gradient of every partially covered one
            uint32_t cnt = 0u;
            if (lane < ncs)
                cnt = cache_block_granules(__popc(__float_as_uint(ws.srec[lane * kRecStride + R_META])),
                                           __popc(__float_as_uint(ws.srec[lane * kRecStride + R_PMASK])));
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < kFSub; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, ncs - 1);
            // (a warp that once found the pool full stops bumping it: every warp overshoots by at most one
            // sub-chunk, so the 32-bit counter cannot wrap - the capacity stays below 2^32 - 2^28 granules)
            uint32_t base = cache.capacity;
            if (lane == 0 && tot > 0u && !pool_full) base = atomicAdd(cache.counter, tot);
            base = __shfl_sync(0xffffffffu, base, 0);
            const bool fits = (unsigned long long)base + tot <= (unsigned long long)cache.capacity;
            pool_full = pool_full || !fits;
            if (lane < ncs) {
                const uint32_t poff = fits ? base + incl - cnt : kNoCache;
                ws.srec[lane * kRecStride + R_META + 3] = __uint_as_float(poff);
                reinterpret_cast<uint32_t*>(tmeta + ws.q.idx[(head + lane) & (kQueue - 1)])[3] = poff;
            }
            __syncwarp();
        }

        // ---- per-lane hit / partial masks from the classification words
        uint32_t hit, part;
        classify_subchunk(ws.srec, ncs, use_aa, lane, hit, part, [&](int) { return !done; });
        if (__any_sync(0xffffffffu, hit != 0u)) {
            // ---- the polygon clip - by far the most expensive step - runs DENSE over the list of all
            //      partially covered (pixel, face) pairs of the sub-chunk: any lane clips any pixel's pair
            int np = 0;
            if (__any_sync(0xffffffffu, part != 0u)) {
                // lane-major compaction: an exclusive scan of the per-lane pair counts, then every lane writes
                // its own pairs (the clip results are addressed by (candidate, pixel): the list order is free;
                // 16 ballot rounds, one per candidate, cost 0.03 ms more)
                {
                    const int mine = __popc(part);
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    np = __shfl_sync(0xffffffffu, incl, 31);
                    int w = incl - mine;
                    for (uint32_t pp = part; pp != 0u; pp &= pp - 1u)
                        ws.plist[w++] = (uint16_t)((__ffs(pp) - 1) | (lane << 4));
                }
                __syncwarp();
                for (int pid = lane; pid < np; pid += 32) {
                    const unsigned e = ws.plist[pid];
                    const int k = e & 15, l = (e >> 4) & 31;
                    const float* rec = ws.srec + k * kRecStride;
                    const unsigned inm = corner_mask_of(*reinterpret_cast<const uint4*>(rec + R_META), l);
                    float area, g[6];
                    bool chain = false;
                    const bool ok = aa_clip<CACHE>(rec, tc.fx0 + (float)(l & 7), tc.fy0 + (float)(l >> 3), inm, area, g, &chain);
                    if (CACHE && chain)        // kink / inconsistent polygon: the reference's Jacobian chain (rare, out of line)
                        aa_grad_reference(rec, tc.fx0 + (float)(l & 7), tc.fy0 + (float)(l >> 3), inm, g);
                    ws.parea[k * 32 + l] = ok ? area : -1.0f;
                    if (CACHE) {
                        const uint32_t poff = __float_as_uint(rec[R_META + 3]);
                        if (poff != kNoCache) {
                            const uint32_t n_t = __popc(__float_as_uint(rec[R_META]));
                            const uint32_t j = __popc(__float_as_uint(rec[R_PMASK]) & ((1u << l) - 1u));
                            float2* gj = reinterpret_cast<float2*>(cache.data + (size_t)poff + n_t) + 3 * j;
                            gj[0] = make_float2(g[0], g[1]); gj[1] = make_float2(g[2], g[3]); gj[2] = make_float2(g[4], g[5]);
                        }
                    }
                }
                __syncwarp();
            }

            // ---- every lane walks its own hits in list order: hit point, barycentrics, blend, termination
            for (uint32_t hh = hit; hh != 0u && !done; hh &= hh - 1u) {
                const int k = __ffs(hh) - 1;
                const float* rec = ws.srec + k * kRecStride;
                Shade s;
                s.area = use_aa ? 1.0f : 0.0f;
                // the backward pass reads the pair's record instead of re-running the hit point and the clamp;
                // pairs the reference skips leave a zero record (rare: out of the straight path)
                auto leave_record = [&](const float4 r4) {
#ifdef DM2_EXP_NO_HITREC                 /* timing experiment only: the backward pass then reads garbage */
                    return;
#endif
                    if (CACHE) {
                        const uint32_t poff = __float_as_uint(rec[R_META + 3]);
                        if (poff != kNoCache)
                            cache.data[(size_t)poff + __popc(__float_as_uint(rec[R_META]) & lt_mask)] = r4;
                    }
                };
                if ((part >> k) & 1u) {
                    s.area = ws.parea[k * 32 + lane];
                    if (s.area == -1.0f) { leave_record(make_float4(0.f, 0.f, 0.f, 0.f)); continue; }
                }
                if (!shade_tail(rec, pg, tau, s)) { leave_record(make_float4(0.f, 0.f, 0.f, 0.f)); continue; }
                leave_record(make_float4(s.w1, s.w2, s.ratio, __int_as_float(s.code)));
                // forward.cu:398-415
                const float alpha = __fmul_rn(rec[R_MISC + 1], s.ratio);
                C0 = __fmaf_rn(__fmul_rn(s.iC[0], alpha), T, C0);   // C += iC*alpha*T
                C1 = __fmaf_rn(__fmul_rn(s.iC[1], alpha), T, C1);
                C2 = __fmaf_rn(__fmul_rn(s.iC[2], alpha), T, C2);
                D  = __fmaf_rn(__fmul_rn(s.iD, alpha), T, D);
                pT = T;
                T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                last_contributor = ws.q.idx[(head + k) & (kQueue - 1)] + 1u;
                if (T < kTEps) done = true;                       // forward.cu:412-415
            }
            __syncwarp();
        }
        head = (head + ncs) & (kQueue - 1);
        count -= ncs;
    }

    if (tc.inside) {
        final_prev_T[tc.pix] = pT;
        final_T[tc.pix] = T;
        n_contrib[tc.pix] = last_contributor;
        out_color[tc.pix * 3 + 0] = __fmaf_rn(T, in.background[0], C0);
        out_color[tc.pix * 3 + 1] = __fmaf_rn(T, in.background[1], C1);
        out_color[tc.pix * 3 + 2] = __fmaf_rn(T, in.background[2], C2);
        const float dn = __fadd_rn(D, T);
        // depth_unit: the three torch ops of reference __init__.py:377-378, 1 - (depth + 1) / 2
        out_depth[tc.pix] = in.depth_unit ? __fsub_rn(1.0f, __fmul_rn(__fadd_rn(dn, 1.0f), 0.5f)) : dn;
    }
}

constexpr size_t kFwdSmem = sizeof(FwdWarpSmem) * (kRasterThreads / 32);

int launch_raster_forward(const DM2RenderInputs& in, const FaceWs& fw, const BinWs& bw, const ImageWs& iw,
                          float* out_color, float* out_depth, void* shade_cache, long long shade_cache_bytes,
                          cudaStream_t s) {
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    const long long nblk = (long long)in.B * gx * gy;
    if (nblk == 0) return 0;
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        attr_set.done();
    }
    const ShadeCache cache = shade_cache_view(shade_cache, shade_cache_bytes);
    const BinGuard guard{iw.tile_cursor + nblk, (uint32_t)bw.R, (uint32_t)bw.longest_cap};
    const bool use_cache = cache.capacity > 0 && in.aa_temperature > 0.0f && bw.R > 0;
    if (use_cache) DM2_CHECK(cudaMemsetAsync(cache.counter, 0, sizeof(uint32_t), s));
    ProfScope _ps("raster_forward", s);
    if (use_cache)
        k_raster_forward<true><<<(unsigned)(nblk * kCtasPerTile), kRasterThreads, kFwdSmem, s>>>(
            in, fw.records, bw.vals, bw.imask, reinterpret_cast<uint4*>(bw.meta), (uint32_t)bw.R, cache, guard, iw.ranges,
            iw.final_T, iw.final_prev_T, iw.n_contrib, out_color, out_depth, gx, gy);
    else
        k_raster_forward<false><<<(unsigned)(nblk * kCtasPerTile), kRasterThreads, kFwdSmem, s>>>(
            in, fw.records, bw.vals, bw.imask, reinterpret_cast<uint4*>(bw.meta), (uint32_t)bw.R, cache, guard, iw.ranges,
            iw.final_T, iw.final_prev_T, iw.n_contrib, out_color, out_depth, gx, gy);
    DM2_LAUNCH_CHECK();
    return 0;
}

}  // namespace dm2
