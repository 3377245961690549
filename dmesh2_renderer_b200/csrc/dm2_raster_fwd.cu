// dm2_raster_fwd.cu — per-tile forward compositor.
// Behaviour of reference cuda_impl/forward.cu:139-432 (renderCUDA<3>), re-designed for sm_100a.
// Structure: see dm2_raster_common.cuh (warp-autonomous candidate queue, dense pair evaluation).
// The reference walks every listed face in lock-step with all 256 threads of the tile behind a
// barrier per 256-face round and re-reads ~40 AA floats per (pixel, face) from global memory.
#include "dm2_raster_common.cuh"

namespace dm2 {

// Candidates per sub-chunk.  Measured at cfg2 (ms, CTAs of 4 warps): 12 candidates / 7 CTAs per SM 1.23;
// 16 / 6 CTAs 1.155; 20 / 5 CTAs 1.17; 24 / 4 CTAs 1.21.
#ifndef DM2_FSUB
#define DM2_FSUB 16
#endif
constexpr int kFSub = DM2_FSUB;
constexpr int kFSubBits = kFSub > 16 ? 5 : 4;      // candidate field of a pair-list entry
constexpr int kFSlots = kFSub * 32;

struct FwdWarpSmem {
    float srec[kFSub * kRecStride];         // staged records of the sub-chunk
    float parea[kFSlots];                   // [candidate][pixel]: overlap area of a partially covered pixel (-1: skip)
    WarpQueue q;
    uint16_t plist[kFSlots];                // partially covered (pixel, face) pairs: k | lane << kFSubBits
};

// CACHE: fill the shading cache for the backward pass (dm2_raster_common.cuh): d(area)/d(aa verts) of every
// clipped pair and one (uc, vc, ratio, code) record per visited (pixel, face) pair.
// ROUNDS: composite list entries [round_begin, round_end) only, resuming from / leaving the raw per-pixel state
// (T, previous T, last contributor in the image-state buffers; un-finalised C / D in out_color / out_depth); a
// footprint that has finished - every pixel terminated or its list exhausted - writes the final values and sets
// its warp_done byte, and returns at once in later rounds.  (A separate instantiation: the single-pass kernel's
// code generation is sensitive to every added live value, profiles/r02_ab_experiments.md.)
template <bool CACHE, bool ROUNDS>
__device__ __forceinline__ void
raster_forward_footprint(const TileCtx& tc, FwdWarpSmem& ws, const DM2RenderInputs& in, const float* __restrict__ records,
                         const uint32_t* __restrict__ list, const uint8_t* __restrict__ imask, uint4* __restrict__ meta,
                         uint32_t R, const ShadeCache& cache, const uint2* __restrict__ ranges, float* __restrict__ final_T,
                         float* __restrict__ final_prev_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                         float* __restrict__ out_depth, int round_begin, int round_end, uint8_t* __restrict__ warp_done) {
#include "dm2_raster_fwd_body.inc"
}

// One CTA = kWarpsPerCta footprints of tile blockIdx.x / kCtasPerTile (single pass, or the first round).
#ifndef DM2_FWD_MINBLOCKS
#define DM2_FWD_MINBLOCKS (3 * kCtasPerTile)
#endif
template <bool CACHE, bool ROUNDS>
__global__ void __launch_bounds__(kRasterThreads, DM2_FWD_MINBLOCKS)
k_raster_forward(DM2RenderInputs in, const float* __restrict__ records, const uint32_t* __restrict__ list,
                 const uint8_t* __restrict__ imask, uint4* __restrict__ meta, uint32_t R, ShadeCache cache,
                 BinGuard guard, const uint2* __restrict__ ranges, float* __restrict__ final_T,
                 float* __restrict__ final_prev_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                 float* __restrict__ out_depth, int gx, int gy, int round_begin, int round_end,
                 uint8_t* __restrict__ warp_done) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (bin_invalid(guard)) return;          // speculative launch on a workspace that turned out too small
    const TileCtx tc = tile_ctx(in, gx, gy);
    FwdWarpSmem& ws = reinterpret_cast<FwdWarpSmem*>(smem_raw)[threadIdx.x >> 5];
#include "dm2_raster_fwd_body.inc"
}

// Second round onwards: a persistent grid strides over the list of part-tiles (tile * kCtasPerTile + part) that are
// still running (k_round_worklist) - a round in which nearly every tile has finished must not cost a launch of
// tiles x kCtasPerTile CTAs (32 us at 2048^2).
template <bool CACHE>
__global__ void __launch_bounds__(kRasterThreads, DM2_FWD_MINBLOCKS)
k_raster_forward_listed(DM2RenderInputs in, const float* __restrict__ records, const uint32_t* __restrict__ list,
                        const uint8_t* __restrict__ imask, uint4* __restrict__ meta, uint32_t R, ShadeCache cache,
                        BinGuard guard, const uint2* __restrict__ ranges, float* __restrict__ final_T,
                        float* __restrict__ final_prev_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                        float* __restrict__ out_depth, int gx, int gy, int round_begin, int round_end,
                        uint8_t* __restrict__ warp_done, const uint32_t* __restrict__ work,
                        const uint32_t* __restrict__ work_n, uint32_t* __restrict__ next_item) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_item;
    if (bin_invalid(guard)) return;
    FwdWarpSmem& ws = reinterpret_cast<FwdWarpSmem*>(smem_raw)[threadIdx.x >> 5];
    const uint32_t n = *work_n;
    // items are handed out dynamically (they differ in length by orders of magnitude; a static stride costs 10 %
    // when most tiles are still running)
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(next_item, 1u);
        __syncthreads();
        const uint32_t w = s_item;
        if (w >= n) return;
        const uint32_t item = work[w];
        raster_forward_footprint<CACHE, true>(
            tile_ctx_of(in, gx, gy, (int)(item / kCtasPerTile), (int)(item % kCtasPerTile) * kWarpsPerCta + (int)(threadIdx.x >> 5)),
            ws, in, records, list, imask, meta, R, cache, ranges, final_T, final_prev_T, n_contrib, out_color, out_depth,
            round_begin, round_end, warp_done);
        __syncwarp();
    }
}

constexpr size_t kFwdSmem = sizeof(FwdWarpSmem) * (kRasterThreads / 32);

int launch_raster_forward(const DM2RenderInputs& in, const FaceWs& fw, const BinWs& bw, const ImageWs& iw,
                          float* out_color, float* out_depth, void* shade_cache, long long shade_cache_bytes,
                          int round_begin, int round_end, const RoundWork& work, cudaStream_t s) {
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    const long long nblk = (long long)in.B * gx * gy;
    if (nblk == 0) return 0;
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward_listed<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        DM2_CHECK(cudaFuncSetAttribute(k_raster_forward_listed<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
        attr_set.done();
    }
    ShadeCache cache = shade_cache_view(shade_cache, shade_cache_bytes);
    cache.exact_clip = exact_clip_requested() ? 1u : 0u;
    const BinGuard guard{iw.tile_cursor + nblk, (uint32_t)bw.R, (uint32_t)bw.longest_cap};
    const bool use_cache = cache.capacity > 0 && in.aa_temperature > 0.0f && bw.R > 0;
    const bool rounds = round_end > 0;
    if (use_cache && round_begin == 0) DM2_CHECK(cudaMemsetAsync(cache.counter, 0, sizeof(uint32_t), s));
    ProfScope _ps("raster_forward", s);
#define DM2_FWD_ARGS                                                                                                   \
    in, fw.records, bw.vals, bw.imask, reinterpret_cast<uint4*>(bw.meta), (uint32_t)bw.R, cache, guard, iw.ranges,     \
        iw.final_T, iw.final_prev_T, iw.n_contrib, out_color, out_depth, gx, gy, round_begin, round_end, iw.warp_done
    if (rounds && round_begin > 0) {
        // (launch_classify_round has listed the part-tiles still running)
        const unsigned grid = (unsigned)std::min<long long>(nblk * kCtasPerTile, (long long)sm_count() * DM2_FWD_MINBLOCKS);
        if (use_cache) k_raster_forward_listed<true><<<grid, kRasterThreads, kFwdSmem, s>>>(DM2_FWD_ARGS, work.fwd_list, work.n + 1, work.n + 2);
        else k_raster_forward_listed<false><<<grid, kRasterThreads, kFwdSmem, s>>>(DM2_FWD_ARGS, work.fwd_list, work.n + 1, work.n + 2);
    } else {
        const unsigned grid = (unsigned)(nblk * kCtasPerTile);
        if (use_cache) {
            if (rounds) k_raster_forward<true, true><<<grid, kRasterThreads, kFwdSmem, s>>>(DM2_FWD_ARGS);
            else k_raster_forward<true, false><<<grid, kRasterThreads, kFwdSmem, s>>>(DM2_FWD_ARGS);
        } else {
            if (rounds) k_raster_forward<false, true><<<grid, kRasterThreads, kFwdSmem, s>>>(DM2_FWD_ARGS);
            else k_raster_forward<false, false><<<grid, kRasterThreads, kFwdSmem, s>>>(DM2_FWD_ARGS);
        }
    }
#undef DM2_FWD_ARGS
    DM2_LAUNCH_CHECK();
    return 0;
}

}  // namespace dm2
