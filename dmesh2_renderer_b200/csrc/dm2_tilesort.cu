// dm2_tilesort.cu — tile binning as a two-level radix sort.
//
// The reference duplicates every face once per touched tile with the key (tile << 32 | depth bits),
// sorts all R keys with cub::DeviceRadixSort::SortPairs on the low 32 + log2(tiles) bits (six 8-bit
// passes at 1024^2) and then finds the tile ranges in the sorted keys (reference
// cuda_impl/renderer.cu:165-218, 415-492).  Here the tile bits are ONE most-significant "digit":
//   1. k_count_tiles        histogram of that digit: instances per tile (from the per-face tile rects);
//   2. k_tile_scan          exclusive scan of the histogram = the reference's tile ranges, for free, and
//                           num_rendered (the scan over per-face counts and identifyTileRanges disappear);
//   3. k_scatter_instances  every instance goes straight into its tile's segment (the reference's
//                           key / value pair, 12 bytes, written once);
//   4. k_tile_sort          each segment is sorted by (depth bits, face id) in shared memory.
// A stable sort of keys that are emitted in ascending face order (what CUB does for the reference)
// orders equal (tile, depth) pairs by face id, so sorting each segment on the 64-bit composite
// (depth bits << 32 | face id) yields exactly the reference's sorted key and value arrays whatever
// order the scatter arrived in.  Four launches and ~36 bytes of traffic per instance instead of
// thirty launches and ~170 bytes.
#include <algorithm>
#include "dm2_common.cuh"

namespace dm2 {

// Visits every tile of every lane's rect: small rects by their own lane, rects of more than 16 tiles
// cooperatively by the whole warp (one huge triangle must not serialise thousands of atomics behind
// a single lane).  Two-phase callbacks so that the atomics of up to four tiles are in flight together:
// a(global tile index) -> ticket, then b(global tile index, ticket, p0, p1) with the owning face's payload.
template <typename A, typename B>
__device__ __forceinline__ void for_each_tile(uint32_t n, uint32_t x0, uint32_t y0, uint32_t w, uint32_t tbase,
                                              uint32_t p0, uint32_t p1, int gx, int lane, A a, B b) {
    if (n && n <= 16) {
        for (uint32_t i0 = 0; i0 < n; i0 += 4) {
            uint32_t t[4], k[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t i = i0 + j;
                t[j] = tbase + (y0 + i / w) * gx + (x0 + i % w);
                if (i < n) k[j] = a(t[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (i0 + j < n) b(t[j], k[j], p0, p1);
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, n > 16);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t sn = __shfl_sync(0xffffffffu, n, src), st = __shfl_sync(0xffffffffu, tbase, src);
        const uint32_t sx = __shfl_sync(0xffffffffu, x0, src), sy = __shfl_sync(0xffffffffu, y0, src);
        const uint32_t sw = __shfl_sync(0xffffffffu, w, src);
        const uint32_t s0 = __shfl_sync(0xffffffffu, p0, src), s1 = __shfl_sync(0xffffffffu, p1, src);
        for (uint32_t i = lane; i < sn; i += 32) {
            const uint32_t t = st + (sy + i / sw) * gx + (sx + i % sw);
            b(t, a(t), s0, s1);
        }
    }
}

struct FaceRect { uint32_t n, x0, y0, w, tbase; };
// Count / scatter run with one grid row per view (blockIdx.y) so that a block only sees the gx*gy tiles
// of ONE view: when they fit, the per-tile counters are privatised in shared memory (shared-memory
// atomics per instance, one global atomic per touched tile and block - 2.3 M same-address-heavy global
// atomics per view otherwise).  kFacesPerThread faces per thread, strided.
constexpr int kBinThreads = 256;
constexpr int kMaxFacesPerThread = 16;          // fewer for small scenes, so that the grid still fills the GPU
constexpr int kMaxSmemTiles = 16384;            // 64 KB of counters (2048^2 frame)

__device__ __forceinline__ FaceRect load_rect_view(int f, int F, int view, int gx, int gy,
                                                   const uint32_t* __restrict__ tiles_touched,
                                                   const uint2* __restrict__ rects) {
    FaceRect r{0u, 0u, 0u, 1u, 0u};
    if (f < F) {
        const size_t idx = (size_t)view * F + f;
        r.n = tiles_touched[idx];
        if (r.n) {
            const uint2 q = rects[idx];
            r.x0 = q.x & 0xffffu; r.y0 = q.x >> 16;
            r.w = (q.y & 0xffffu) - r.x0;
        }
    }
    return r;
}

template <bool SMEM>
__global__ void __launch_bounds__(kBinThreads)
k_count_tiles(int F, int fpt, int gx, int gy, const uint32_t* __restrict__ tiles_touched,
              const uint2* __restrict__ rects, uint32_t* __restrict__ tile_count) {
    extern __shared__ uint32_t s_cnt[];
    const int view = blockIdx.y, tiles = gx * gy, lane = threadIdx.x & 31;
    uint32_t* gcnt = tile_count + (size_t)view * tiles;
    if (SMEM) {
        for (int i = threadIdx.x; i < tiles; i += kBinThreads) s_cnt[i] = 0u;
        __syncthreads();
    }
    uint32_t* cnt = SMEM ? s_cnt : gcnt;
    for (int it = 0; it < fpt; it++) {
        const int f = (blockIdx.x * fpt + it) * kBinThreads + threadIdx.x;
        const FaceRect r = load_rect_view(f, F, view, gx, gy, tiles_touched, rects);
        for_each_tile(r.n, r.x0, r.y0, r.w, 0u, 0u, 0u, gx, lane,
                      [&](uint32_t t) { atomicAdd(cnt + t, 1u); return 0u; },
                      [&](uint32_t, uint32_t, uint32_t, uint32_t) {});
    }
    if (SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < tiles; i += kBinThreads) {
            const uint32_t c = s_cnt[i];
            if (c) atomicAdd(gcnt + i, c);
        }
    }
}

// One block: exclusive scan of the per-tile counts.  ranges[t] = [start, end) (the reference leaves
// untouched tiles at (0,0), renderer.cu:211), cursor[t] = start (bumped by the scatter),
// cursor[num_tiles] = num_rendered, cursor[num_tiles + 2] = the longest segment (zeroed by the caller;
// the host skips the sort tiers no segment falls into).
__global__ void __launch_bounds__(1024)
k_tile_scan(const uint32_t* tile_count, uint32_t num_tiles, uint2* __restrict__ ranges, uint32_t* cursor) {
    __shared__ uint32_t wsum[32];
    const uint32_t per = (num_tiles + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, num_tiles), hi = min(lo + per, num_tiles);
    uint32_t s = 0, longest = 0;
    for (uint32_t i = lo; i < hi; i++) { const uint32_t c = tile_count[i]; s += c; longest = max(longest, c); }
    longest = __reduce_max_sync(0xffffffffu, longest);
    if ((threadIdx.x & 31) == 0 && longest) atomicMax(cursor + num_tiles + 2, longest);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const uint32_t v = wsum[lane];
        uint32_t iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        wsum[lane] = iv - v;
    }
    __syncthreads();
    uint32_t run = wsum[w] + inc - s;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t c = tile_count[i];
        ranges[i] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
        cursor[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) cursor[num_tiles] = run;
}

// reference renderer.cu:415-465: key = (tile + tiles*view) << 32 | bits(depth); value = face id -
// written into the tile's segment at a position drawn from the tile's cursor.  SMEM: the block first
// counts its own instances per tile in shared memory, reserves one range per touched tile with a single
// global atomic, then hands out positions inside its ranges with shared-memory atomics.
template <bool SMEM>
__global__ void __launch_bounds__(kBinThreads)
k_scatter_instances(int F, int fpt, int gx, int gy, const float* __restrict__ key_depths,
                    const uint32_t* __restrict__ tiles_touched, const uint2* __restrict__ rects,
                    uint32_t* __restrict__ cursor, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                    BinGuard guard) {
    extern __shared__ uint32_t s_cnt[];
    if (bin_invalid(guard)) return;
    const int view = blockIdx.y, tiles = gx * gy, lane = threadIdx.x & 31;
    const uint32_t tbase = (uint32_t)view * (uint32_t)tiles;
    uint32_t* gcur = cursor + (size_t)tbase;
    if (SMEM) {
        for (int i = threadIdx.x; i < tiles; i += kBinThreads) s_cnt[i] = 0u;
        __syncthreads();
        for (int it = 0; it < fpt; it++) {
            const int f = (blockIdx.x * fpt + it) * kBinThreads + threadIdx.x;
            const FaceRect r = load_rect_view(f, F, view, gx, gy, tiles_touched, rects);
            for_each_tile(r.n, r.x0, r.y0, r.w, 0u, 0u, 0u, gx, lane,
                          [&](uint32_t t) { atomicAdd(s_cnt + t, 1u); return 0u; },
                          [&](uint32_t, uint32_t, uint32_t, uint32_t) {});
        }
        __syncthreads();
        // s_cnt[t] := first position of this block's range in tile t's segment
        for (int i = threadIdx.x; i < tiles; i += kBinThreads) {
            const uint32_t c = s_cnt[i];
            if (c) s_cnt[i] = atomicAdd(gcur + i, c);
        }
        __syncthreads();
    }
    uint32_t* cur = SMEM ? s_cnt : gcur;
    for (int it = 0; it < fpt; it++) {
        const int f = (blockIdx.x * fpt + it) * kBinThreads + threadIdx.x;
        const FaceRect r = load_rect_view(f, F, view, gx, gy, tiles_touched, rects);
        uint32_t dbits = 0u;
        if (r.n) dbits = __float_as_uint(key_depths[(size_t)view * F + f]);
        for_each_tile(r.n, r.x0, r.y0, r.w, 0u, dbits, (uint32_t)f, gx, lane,
                      [&](uint32_t t) { return atomicAdd(cur + t, 1u); },
                      [&](uint32_t t, uint32_t pos, uint32_t d, uint32_t fid) {
                          keys[pos] = ((uint64_t)(tbase + t) << 32) | d;
                          vals[pos] = fid;
                      });
    }
}

// ------------------------------------------------------------------ per-tile segment sort
// Bitonic network in its all-ascending form (first step of every merge mirrors the partner index),
// so elements past the end of the segment behave as +infinity simply by being skipped.
// Generic over the cooperating group: (tid, nthr, sync) = a whole CTA with __syncthreads or one warp
// with __syncwarp.
template <typename Load, typename Store, typename Sync>
__device__ __forceinline__ void bitonic_ascending(uint32_t n, uint32_t n_pad, uint32_t tid, uint32_t nthr, Load ld,
                                                  Store st, Sync sync) {
    for (uint32_t k = 2; k <= n_pad; k <<= 1) {
        const uint32_t hk = k >> 1;
        for (uint32_t i = tid; i < (n_pad >> 1); i += nthr) {
            const uint32_t blk = i / hk, within = i % hk;
            const uint32_t a = blk * k + within, b = blk * k + k - 1u - within;
            if (b < n) {
                const uint64_t va = ld(a), vb = ld(b);
                if (va > vb) { st(a, vb); st(b, va); }
            }
        }
        sync();
        for (uint32_t j = hk >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < (n_pad >> 1); i += nthr) {
                const uint32_t a = 2u * j * (i / j) + (i % j), b = a + j;
                if (b < n) {
                    const uint64_t va = ld(a), vb = ld(b);
                    if (va > vb) { st(a, vb); st(b, va); }
                }
            }
            sync();
        }
    }
}

// Three tiers by segment length (one launch each, CTAs of the other tiers exit at once):
constexpr uint32_t kSortCap0 = 2048;          // tier 0:  256 threads,  16 KB of shared memory
constexpr uint32_t kSortCap1 = 8192;          // tier 1: 1024 threads,  64 KB
constexpr uint32_t kSortCap2 = 24576;         // tier 2: 1024 threads, 192 KB; longer segments: global memory
constexpr uint32_t kSortDirectMax = 256;      // up to here one block-wide network; above: depth bins first

__device__ __forceinline__ uint32_t next_pow2(uint32_t n) { return n <= 1u ? 1u : 1u << (32 - __clz(n - 1u)); }

// Sorts one tile segment (composite keys: depth bits << 32 | face id) into s[0..n) with the whole CTA.
// Segments of more than kSortDirectMax elements are first split into 2 * (number of warps) bins by a
// MONOTONE function of the depth bits (one more most-significant radix digit, linear between the
// segment's min and max depth): histogram and scatter straight from the global input, then every warp
// sorts its bins with the bitonic network on its own, without block barriers - O(n) + small sorts
// instead of log^2(n) block-wide stages.  A segment whose keys crowd into one bin (> n / 4) falls back
// to the block-wide network.
// LAZY sort (deep scenes, dm2_capi.cu forward_tail): `prefix` < n sorts only the bins that start before list
// position `prefix` - the rest of the segment stays in bin order (every key of a later bin is behind every key of an
// earlier one, the bins themselves unsorted).  The forward pass of such scenes terminates within the first few hundred
// entries of lists of ~10 000 (10 M triangles @2048^2: 94.5 % of the instances lie behind the deepest position any
// pixel of their tile visits), and the per-bin networks are ~85 % of this kernel's work; a tile that does run deeper
// gets the rest of its segment sorted before the round that needs it (k_tile_sort_finish).  Returns the length of the
// sorted prefix (n: everything).
template <int THREADS>
__device__ __forceinline__ uint32_t sort_segment_smem(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                      uint64_t* __restrict__ s, uint32_t n, uint32_t prefix) {
    constexpr int NW = THREADS / 32, NB = 2 * NW;
    __shared__ uint32_t s_bin[NB + 1];
    __shared__ uint32_t s_red[2 * NW];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    bool binned = n > kSortDirectMax;
    if (binned) {
        uint32_t lo = 0xFFFFFFFFu, hi = 0u;
        for (uint32_t i = tid; i < n; i += THREADS) {
            const uint32_t d = (uint32_t)kin[i];
            lo = min(lo, d); hi = max(hi, d);
        }
        lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) { s_red[warp] = lo; s_red[NW + warp] = hi; }
        if (tid <= NB) s_bin[tid] = 0u;
        __syncthreads();
        lo = 0xFFFFFFFFu; hi = 0u;
#pragma unroll
        for (int w = 0; w < NW; w++) { lo = min(lo, s_red[w]); hi = max(hi, s_red[NW + w]); }
        const uint32_t dmin = lo;
        const uint64_t span = (uint64_t)(hi - lo) + 1ull;
        auto bin_of = [&](uint32_t d) { return (uint32_t)(((uint64_t)(d - dmin) * NB) / span); };
        __syncthreads();
        for (uint32_t i = tid; i < n; i += THREADS) atomicAdd(&s_bin[bin_of((uint32_t)kin[i])], 1u);
        __syncthreads();
        // exclusive scan of the NB counters (warp 0), largest bin
        if (warp == 0) {
            const uint32_t c0 = lane < NB ? s_bin[lane] : 0u, c1 = (NB > 32 && lane + 32 < NB) ? s_bin[lane + 32] : 0u;
            uint32_t i0 = c0, i1 = c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
                if (lane >= o) { i0 += t0; i1 += t1; }
            }
            const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31);
            const uint32_t big = __reduce_max_sync(0xffffffffu, max(c0, c1));
            if (lane < NB) s_bin[lane] = i0 - c0;
            if (NB > 32 && lane + 32 < NB) s_bin[lane + 32] = tot0 + i1 - c1;
            if (lane == 0) s_red[0] = big;
        }
        __syncthreads();
        binned = s_red[0] <= n / 4u;
        if (binned) {
            for (uint32_t i = tid; i < n; i += THREADS) {
                const uint64_t k = kin[i];
                const uint32_t pos = atomicAdd(&s_bin[bin_of((uint32_t)k)], 1u);
                s[pos] = (k << 32) | vin[i];
            }
            __syncthreads();
            // s_bin[b] now holds the END of bin b; every warp sorts bins warp, warp + NW
            uint32_t upto = n;
            if (prefix < n) {                    // end of the last bin that starts before `prefix` (uniform)
                for (int b = NB - 1; b >= 0; b--)
                    if (s_bin[b] >= prefix) upto = s_bin[b];
            }
            for (uint32_t b = warp; b < (uint32_t)NB; b += NW) {
                const uint32_t e = s_bin[b], st0 = b == 0 ? 0u : s_bin[b - 1];
                const uint32_t m = e - st0;
                if (m > 1u && st0 < prefix) {
                    uint64_t* sb = s + st0;
                    bitonic_ascending(m, next_pow2(m), lane, 32u, [&](uint32_t i) { return sb[i]; },
                                      [&](uint32_t i, uint64_t x) { sb[i] = x; }, [] { __syncwarp(); });
                }
            }
            __syncthreads();
            return upto;
        }
    }
    for (uint32_t i = tid; i < n; i += THREADS) s[i] = (kin[i] << 32) | vin[i];          // depth bits << 32 | face id
    __syncthreads();
    bitonic_ascending(n, next_pow2(n), tid, (uint32_t)THREADS, [&](uint32_t i) { return s[i]; },
                      [&](uint32_t i, uint64_t x) { s[i] = x; }, [] { __syncthreads(); });
    return n;
}

// work == nullptr: one CTA per tile (blockIdx.x), first phase (`prefix`: see sort_segment_smem).  work != nullptr: the
// SECOND phase of the lazy sort - a persistent grid over a round's work list (k_round_worklist): the listed tiles of
// this tier whose sorted prefix ends before min(need_end, n) get their WHOLE segment sorted from the unsorted scatter
// output, which is still in the workspace.  Positions below the old prefix receive the values they already hold (the
// order is a total one), so the classification words and cache offsets recorded for them stay valid.
template <int TIER>
__global__ void __launch_bounds__(TIER == 0 ? 256 : 1024)
k_tile_sort(const uint2* __restrict__ ranges, const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, BinGuard guard, uint32_t prefix,
            uint32_t* __restrict__ sorted_upto, const uint32_t* __restrict__ work, const uint32_t* __restrict__ work_n,
            uint32_t need_end) {
    extern __shared__ __align__(16) unsigned char tsort_smem[];
    if (bin_invalid(guard)) return;
    constexpr int THREADS = TIER == 0 ? 256 : 1024;
    constexpr uint32_t LO = TIER == 0 ? 0u : (TIER == 1 ? kSortCap0 : kSortCap1);
    constexpr uint32_t HI = TIER == 0 ? kSortCap0 : (TIER == 1 ? kSortCap1 : 0xFFFFFFFFu);
    uint64_t* s = reinterpret_cast<uint64_t*>(tsort_smem);
    const bool listed = work != nullptr;
    const uint32_t n_items = listed ? *work_n : gridDim.x;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t tile = listed ? work[item] : item;
        const uint2 rg = ranges[tile];
        const uint32_t n = rg.y - rg.x;
        if (n == 0u || n <= LO || n > HI) continue;
        if (listed && sorted_upto[tile] >= min(need_end, n)) continue;             // (uniform per CTA)
        const uint64_t tile_hi = (uint64_t)tile << 32;
        if (TIER < 2 || n <= kSortCap2) {
            const uint32_t upto = sort_segment_smem<THREADS>(keys_in + rg.x, vals_in + rg.x, s, n, listed ? 0xFFFFFFFFu : prefix);
            for (uint32_t i = threadIdx.x; i < n; i += THREADS) {
                const uint64_t v = s[i];
                keys[rg.x + i] = tile_hi | (v >> 32);
                vals[rg.x + i] = (uint32_t)v;
            }
            if (sorted_upto && threadIdx.x == 0) sorted_upto[tile] = upto;
        } else {
            if (sorted_upto && threadIdx.x == 0) sorted_upto[tile] = n;
            // rare: a segment longer than kSortCap2 instances - block-wide network on the output key buffer
            uint64_t* g = keys + rg.x;
            for (uint32_t i = threadIdx.x; i < n; i += THREADS) g[i] = (keys_in[rg.x + i] << 32) | vals_in[rg.x + i];
            __syncthreads();
            bitonic_ascending(n, next_pow2(n), threadIdx.x, (uint32_t)THREADS, [&](uint32_t i) { return g[i]; },
                              [&](uint32_t i, uint64_t v) { g[i] = v; }, [] { __syncthreads(); });
            for (uint32_t i = threadIdx.x; i < n; i += THREADS) {
                const uint64_t v = g[i];
                g[i] = tile_hi | (v >> 32);
                vals[rg.x + i] = (uint32_t)v;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ launchers
// faces per thread of the count / scatter kernels: as many as keep ~4 blocks per SM busy
static int faces_per_thread(long long BF) {
    const long long want_blocks = 148 * 4;
    long long fpt = BF / (want_blocks * kBinThreads);
    return (int)std::max(1ll, std::min((long long)kMaxFacesPerThread, fpt));
}

// stage 1 of the binning (inside dm2_*_preprocess): per-tile counts -> ranges, cursors, num_rendered
int launch_tile_count_scan(long long BF, int F, int gx, int gy, int num_tiles, const uint32_t* tiles_touched,
                           const uint2* rects, uint2* ranges, uint32_t* cursor, cudaStream_t s) {
    // the counts live in the cursor array (zeroed by the caller) until the scan turns them into start offsets
    if (num_tiles == 0) return 0;
    if (BF > 0) {
        static PerDeviceOnce cattr;
        if (cattr.first()) {
            DM2_CHECK(cudaFuncSetAttribute(k_count_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTiles * 4));
            DM2_CHECK(cudaFuncSetAttribute(k_scatter_instances<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kMaxSmemTiles * 4));
            cattr.done();
        }
        ProfScope _ps("tile_count", s);
        const int tiles = gx * gy;
        const int fpt = faces_per_thread(BF);
        const dim3 grid(cdiv(F, kBinThreads * fpt), (unsigned)(BF / F));
        if (tiles <= kMaxSmemTiles)
            k_count_tiles<true><<<grid, kBinThreads, (size_t)tiles * 4, s>>>(F, fpt, gx, gy, tiles_touched, rects, cursor);
        else
            k_count_tiles<false><<<grid, kBinThreads, 0, s>>>(F, fpt, gx, gy, tiles_touched, rects, cursor);
        DM2_LAUNCH_CHECK();
    }
    {
        ProfScope _ps("tile_scan", s);
        k_tile_scan<<<1, 1024, 0, s>>>(cursor, (uint32_t)num_tiles, ranges, cursor);
        DM2_LAUNCH_CHECK();
    }
    return 0;
}

// stage 2: scatter into the tile segments and sort every segment
int launch_tile_sort(long long BF, int F, int gx, int gy, int num_tiles, const float* key_depths,
                     const uint32_t* tiles_touched, const uint2* rects, const uint2* ranges, uint32_t* cursor,
                     uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys, uint32_t* vals,
                     long long capacity, long long longest, const LazySort& lazy, cudaStream_t s) {
    if (BF == 0 || num_tiles == 0) return 0;
    const BinGuard guard{cursor + num_tiles, (uint32_t)capacity, 0xFFFFFFFFu};
    if (longest < 0) longest = 1ll << 40;            // unknown: launch every tier
    {
        ProfScope _ps("scatter_instances", s);
        const int tiles = gx * gy;
        const int fpt = faces_per_thread(BF);
        const dim3 grid(cdiv(F, kBinThreads * fpt), (unsigned)(BF / F));
        if (tiles <= kMaxSmemTiles)
            k_scatter_instances<true><<<grid, kBinThreads, (size_t)tiles * 4, s>>>(
                F, fpt, gx, gy, key_depths, tiles_touched, rects, cursor, keys_unsorted, vals_unsorted, guard);
        else
            k_scatter_instances<false><<<grid, kBinThreads, 0, s>>>(F, fpt, gx, gy, key_depths, tiles_touched, rects,
                                                                  cursor, keys_unsorted, vals_unsorted, guard);
        DM2_LAUNCH_CHECK();
    }
    const int tiers = 1 | (longest > kSortCap0 ? 2 : 0) | (longest > kSortCap1 ? 4 : 0);
    if (lazy.sorted_upto) DM2_CHECK(cudaMemsetAsync(lazy.sorted_upto, 0, (size_t)num_tiles * sizeof(uint32_t), s));
    return launch_tile_sort_tiers(num_tiles, ranges, cursor, keys_unsorted, vals_unsorted, keys, vals, capacity, tiers, lazy, s);
}

// longest segment covered by the tiers launch_tile_sort enqueues for a given hint (-1: all of them)
long long sort_tiers_longest_cap(long long longest_hint) {
    if (longest_hint < 0 || longest_hint > kSortCap1) return 0xFFFFFFFFll;
    return longest_hint > kSortCap0 ? (long long)kSortCap1 : (long long)kSortCap0;
}

// the per-segment sort alone; tiers: bit t = run tier t (segments of <= 2048 / <= 8192 / more instances)
int launch_tile_sort_tiers(int num_tiles, const uint2* ranges, const uint32_t* cursor, const uint64_t* keys_unsorted,
                           const uint32_t* vals_unsorted, uint64_t* keys, uint32_t* vals, long long capacity, int tiers,
                           const LazySort& lazy, cudaStream_t s) {
    if (num_tiles == 0 || tiers == 0) return 0;
    const BinGuard guard{cursor + num_tiles, (uint32_t)capacity, 0xFFFFFFFFu};
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        DM2_CHECK(cudaFuncSetAttribute(k_tile_sort<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kSortCap1 * sizeof(uint64_t))));
        DM2_CHECK(cudaFuncSetAttribute(k_tile_sort<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kSortCap2 * sizeof(uint64_t))));
        attr_set.done();
    }
    const uint32_t prefix = lazy.sorted_upto ? lazy.prefix : 0xFFFFFFFFu;
    ProfScope _ps("tile_sort", s);
    if (tiers & 1) {
        k_tile_sort<0><<<num_tiles, 256, kSortCap0 * sizeof(uint64_t), s>>>(ranges, keys_unsorted, vals_unsorted, keys, vals, guard, prefix,
                                                                            lazy.sorted_upto, nullptr, nullptr, 0u);
        DM2_LAUNCH_CHECK();
    }
    if (tiers & 2) {
        k_tile_sort<1><<<num_tiles, 1024, kSortCap1 * sizeof(uint64_t), s>>>(ranges, keys_unsorted, vals_unsorted, keys, vals, guard, prefix,
                                                                            lazy.sorted_upto, nullptr, nullptr, 0u);
        DM2_LAUNCH_CHECK();
    }
    if (tiers & 4) {
        k_tile_sort<2><<<num_tiles, 1024, kSortCap2 * sizeof(uint64_t), s>>>(ranges, keys_unsorted, vals_unsorted, keys, vals, guard, prefix,
                                                                            lazy.sorted_upto, nullptr, nullptr, 0u);
        DM2_LAUNCH_CHECK();
    }
    return 0;
}

// the lazy sort's second phase for the tiles of a round's work list (list + count on the device): the tier kernels as
// persistent grids over the list; tiers: as launch_tile_sort_tiers
int launch_tile_sort_finish(int num_tiles, const uint2* ranges, const uint32_t* cursor, const uint64_t* keys_unsorted,
                            const uint32_t* vals_unsorted, uint64_t* keys, uint32_t* vals, long long capacity, int tiers,
                            const uint32_t* work, const uint32_t* work_n, long long need_end, const LazySort& lazy,
                            cudaStream_t s) {
    // (every segment is sorted at least up to min(prefix, n): rounds that end there need nothing)
    if (num_tiles == 0 || !lazy.sorted_upto || lazy.prefix == 0xFFFFFFFFu || need_end <= (long long)lazy.prefix) return 0;
    const BinGuard guard{cursor + num_tiles, (uint32_t)capacity, 0xFFFFFFFFu};
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        DM2_CHECK(cudaFuncSetAttribute(k_tile_sort<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kSortCap1 * sizeof(uint64_t))));
        DM2_CHECK(cudaFuncSetAttribute(k_tile_sort<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kSortCap2 * sizeof(uint64_t))));
        attr_set.done();
    }
    ProfScope _ps("tile_sort", s);
    const uint32_t need = (uint32_t)std::min<long long>(need_end, 0xFFFFFFFFll);
    const int sms = sm_count();
    if (tiers & 1) {
        k_tile_sort<0><<<(unsigned)std::min(num_tiles, 8 * sms), 256, kSortCap0 * sizeof(uint64_t), s>>>(
            ranges, keys_unsorted, vals_unsorted, keys, vals, guard, 0xFFFFFFFFu, lazy.sorted_upto, work, work_n, need);
        DM2_LAUNCH_CHECK();
    }
    if (tiers & 2) {
        k_tile_sort<1><<<(unsigned)std::min(num_tiles, 2 * sms), 1024, kSortCap1 * sizeof(uint64_t), s>>>(
            ranges, keys_unsorted, vals_unsorted, keys, vals, guard, 0xFFFFFFFFu, lazy.sorted_upto, work, work_n, need);
        DM2_LAUNCH_CHECK();
    }
    if (tiers & 4) {
        k_tile_sort<2><<<(unsigned)std::min(num_tiles, sms), 1024, kSortCap2 * sizeof(uint64_t), s>>>(
            ranges, keys_unsorted, vals_unsorted, keys, vals, guard, 0xFFFFFFFFu, lazy.sorted_upto, work, work_n, need);
        DM2_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace dm2
