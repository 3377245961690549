// dm2_peer.cu — the one exchange step of the multi-view path: the per-step sum of the replicated-parameter
// gradients [dverts | dverts_color | dfaces_opacity] over the GPUs of one NVSwitch box, as a
// hand-written kernel over peer memory instead of a library allreduce.
//
// Every rank's bucket lives in a symmetric allocation whose peer mappings (and, when the fabric supports it,
// one multicast mapping) are visible in every process (torch.distributed._symmetric_memory is the plumbing
// that creates and exchanges the handles; the arithmetic and all data movement are the kernels below).
// Rank r owns slice r of the bucket:
//   * k_peer_allreduce_mc  (NVLS): multimem.ld_reduce pulls the slice from all ranks and adds it INSIDE the
//     switch, multimem.st broadcasts the sum to every rank's bucket - each GPU moves 1/W of the bucket in and
//     out of its SMs;
//   * k_peer_allreduce_p2p (fallback when there is no multicast mapping): the owner loads the slice from the W
//     peer buffers (rank order 0..W-1: the same, deterministic sum on every rank), adds, and stores the result
//     into all W buffers.
// Two cross-GPU barriers bracket the data movement (all buckets complete before the first load; all stores
// landed before anyone reads its bucket).  dm2_peer_allreduce_sum leaves them to the caller (two extra
// launches on every rank); dm2_peer_allreduce_sum_fused does them INSIDE the kernel with release / acquire
// flags at system scope in the symmetric allocation - one launch per step:
//   entry: block 0 stores `epoch` into slot [rank] of every peer's arrive[] array; every block spins until all
//          slots of its OWN arrive[] array carry the epoch (its local bucket is complete by stream order);
//   exit:  after its stores + a system fence every block bumps a local counter; the LAST block of the rank
//          stores `epoch` into slot [rank] of every peer's depart[] array and spins until its own depart[] array
//          is complete - the kernel cannot finish before every peer has finished writing this rank's bucket.
// Spins are bounded (~20 s of SM clocks), then the kernel traps: a dead peer fails the job instead of hanging it.
#include <stdlib.h>
#include <mutex>
#include "dm2_common.cuh"

namespace dm2 {

struct PeerPtrs { float* p[16]; };

__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_v4(float* p, const float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce_v4(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_v4(float* p, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- in-kernel cross-GPU barriers -----------------------------------------------------------------------
// flag area of one rank (u32 words): arrive[16] | depart[16] | done counter
struct PeerFlags { uint32_t* p[16]; int rank, world; uint32_t epoch; };      // p[0] == nullptr: caller-side barriers
constexpr int kFlagDepart = 16, kFlagDone = 32;
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t epoch) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
        if (clock64() - t0 > 40000000000ll) __trap();          // ~20 s: a peer never arrived
        __nanosleep(64);
    }
}
__device__ __forceinline__ void peer_barrier_entry(const PeerFlags& f) {
    if (f.p[0] == nullptr) return;
    if (blockIdx.x == 0 && threadIdx.x < f.world) st_release_sys(f.p[threadIdx.x] + f.rank, f.epoch);
    if (threadIdx.x < f.world) spin_until(f.p[f.rank] + threadIdx.x, f.epoch);
    __syncthreads();
}
__device__ __forceinline__ void peer_barrier_exit(const PeerFlags& f) {
    if (f.p[0] == nullptr) return;
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        uint32_t* done = f.p[f.rank] + kFlagDone;
        const uint32_t old = atomicAdd(done, 1u);
        s_last = old == gridDim.x - 1u;
        if (s_last) *done = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < f.world) {
        st_release_sys(f.p[threadIdx.x] + kFlagDepart + f.rank, f.epoch);
        spin_until(f.p[f.rank] + kFlagDepart + threadIdx.x, f.epoch);
    }
}

// [begin4, end4): this rank's slice in units of float4
template <int W>
__global__ void __launch_bounds__(256)
k_peer_allreduce_p2p(PeerPtrs ptrs, long long begin4, long long end4, PeerFlags flags) {
    peer_barrier_entry(flags);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = begin4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end4; i += stride) {
        float4 v[W];
#pragma unroll
        for (int q = 0; q < W; q++) v[q] = ld_sys_v4(ptrs.p[q] + 4 * i);
        float4 a = v[0];
#pragma unroll
        for (int q = 1; q < W; q++) { a.x += v[q].x; a.y += v[q].y; a.z += v[q].z; a.w += v[q].w; }
#pragma unroll
        for (int q = 0; q < W; q++) st_sys_v4(ptrs.p[q] + 4 * i, a);
    }
    peer_barrier_exit(flags);
}

// U independent 16-byte switch reductions in flight per thread (the round trip through the switch is long)
template <int U>
__global__ void __launch_bounds__(256)
k_peer_allreduce_mc(float* mc, long long begin4, long long end4, PeerFlags flags) {
    peer_barrier_entry(flags);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = begin4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < end4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = mc_ld_reduce_v4(mc + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < U; u++) mc_st_v4(mc + 4 * (i + u * stride), v[u]);
    }
    for (; i < end4; i += stride) mc_st_v4(mc + 4 * i, mc_ld_reduce_v4(mc + 4 * i));
    peer_barrier_exit(flags);
}

template <int W>
static void launch_p2p(const PeerPtrs& ptrs, long long b4, long long e4, int grid, const PeerFlags& fl, cudaStream_t s) {
    k_peer_allreduce_p2p<W><<<grid, 256, 0, s>>>(ptrs, b4, e4, fl);
}

int launch_peer_allreduce(const void* const* peer_ptrs, void* multicast_ptr, const void* const* peer_flag_ptrs,
                          uint32_t epoch, int world, int rank, long long n_floats, cudaStream_t s) {
    DM2_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "dm2_peer_allreduce_sum: bad world / rank");
    DM2_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, "dm2_peer_allreduce_sum: the bucket length must be a multiple of 4 floats");
    if (world == 1 || n_floats == 0) return 0;
    DM2_REQUIRE(peer_ptrs != nullptr, "dm2_peer_allreduce_sum: null peer pointer table");
    const long long n4 = n_floats / 4;
    const long long per = (n4 + world - 1) / world;
    const long long b4 = std::min(n4, per * rank), e4 = std::min(n4, per * (rank + 1));
    PeerFlags fl;
    for (int q = 0; q < 16; q++) fl.p[q] = nullptr;
    fl.rank = rank; fl.world = world; fl.epoch = epoch;
    if (peer_flag_ptrs) {
        for (int q = 0; q < world; q++) {
            DM2_REQUIRE(peer_flag_ptrs[q] != nullptr, "dm2_peer_allreduce_sum_fused: null flag pointer");
            fl.p[q] = reinterpret_cast<uint32_t*>(const_cast<void*>(peer_flag_ptrs[q]));
        }
    }
    // (with in-kernel barriers a rank with an empty slice still has to take part in both of them)
    if (e4 <= b4 && !peer_flag_ptrs) return 0;
    int dev = 0, num_sms = 0;
    DM2_CHECK(cudaGetDevice(&dev));
    DM2_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    static int mc_unroll = 4, grid_mult = 4;            // tuning knobs (defaults measured on 8 x B200)
    static std::once_flag knobs_once;
    std::call_once(knobs_once, [] {
        const char* u = getenv("DM2_PEER_MC_UNROLL");
        const char* g = getenv("DM2_PEER_GRID_MULT");
        if (u) mc_unroll = atoi(u);
        if (g) grid_mult = std::max(1, atoi(g));
    });
    const int grid = (int)std::max<long long>(1, std::min<long long>((e4 - b4 + 255) / 256, (long long)num_sms * grid_mult));
    ProfScope _ps("peer_allreduce", s);
    if (multicast_ptr) {
        float* mc = reinterpret_cast<float*>(multicast_ptr);
        if (mc_unroll >= 8) k_peer_allreduce_mc<8><<<grid, 256, 0, s>>>(mc, b4, e4, fl);
        else if (mc_unroll >= 4) k_peer_allreduce_mc<4><<<grid, 256, 0, s>>>(mc, b4, e4, fl);
        else if (mc_unroll >= 2) k_peer_allreduce_mc<2><<<grid, 256, 0, s>>>(mc, b4, e4, fl);
        else k_peer_allreduce_mc<1><<<grid, 256, 0, s>>>(mc, b4, e4, fl);
    } else {
        PeerPtrs ptrs;
        for (int q = 0; q < 16; q++) ptrs.p[q] = q < world ? reinterpret_cast<float*>(const_cast<void*>(peer_ptrs[q])) : nullptr;
        for (int q = 0; q < world; q++) DM2_REQUIRE(ptrs.p[q] != nullptr, "dm2_peer_allreduce_sum: null peer pointer");
        switch (world) {
            case 2: launch_p2p<2>(ptrs, b4, e4, grid, fl, s); break;
            case 3: launch_p2p<3>(ptrs, b4, e4, grid, fl, s); break;
            case 4: launch_p2p<4>(ptrs, b4, e4, grid, fl, s); break;
            case 5: launch_p2p<5>(ptrs, b4, e4, grid, fl, s); break;
            case 6: launch_p2p<6>(ptrs, b4, e4, grid, fl, s); break;
            case 7: launch_p2p<7>(ptrs, b4, e4, grid, fl, s); break;
            case 8: launch_p2p<8>(ptrs, b4, e4, grid, fl, s); break;
            case 16: launch_p2p<16>(ptrs, b4, e4, grid, fl, s); break;
            default: DM2_REQUIRE(false, "dm2_peer_allreduce_sum: unsupported world size (2..8, 16)");
        }
    }
    DM2_LAUNCH_CHECK();
    return 0;
}

}  // namespace dm2
