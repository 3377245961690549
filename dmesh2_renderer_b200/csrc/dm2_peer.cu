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
// The caller brackets the kernel with two cross-GPU barriers (all buckets complete before the first load; all
// stores landed before anyone reads its bucket).
#include <stdlib.h>
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

// [begin4, end4): this rank's slice in units of float4
template <int W>
__global__ void __launch_bounds__(256)
k_peer_allreduce_p2p(PeerPtrs ptrs, long long begin4, long long end4) {
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
}

// U independent 16-byte switch reductions in flight per thread (the round trip through the switch is long)
template <int U>
__global__ void __launch_bounds__(256)
k_peer_allreduce_mc(float* mc, long long begin4, long long end4) {
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
}

template <int W>
static void launch_p2p(const PeerPtrs& ptrs, long long b4, long long e4, int grid, cudaStream_t s) {
    k_peer_allreduce_p2p<W><<<grid, 256, 0, s>>>(ptrs, b4, e4);
}

int launch_peer_allreduce(const void* const* peer_ptrs, void* multicast_ptr, int world, int rank, long long n_floats,
                          cudaStream_t s) {
    DM2_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "dm2_peer_allreduce_sum: bad world / rank");
    DM2_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, "dm2_peer_allreduce_sum: the bucket length must be a multiple of 4 floats");
    if (world == 1 || n_floats == 0) return 0;
    DM2_REQUIRE(peer_ptrs != nullptr, "dm2_peer_allreduce_sum: null peer pointer table");
    const long long n4 = n_floats / 4;
    const long long per = (n4 + world - 1) / world;
    const long long b4 = std::min(n4, per * rank), e4 = std::min(n4, per * (rank + 1));
    if (e4 <= b4) return 0;
    int dev = 0, num_sms = 0;
    DM2_CHECK(cudaGetDevice(&dev));
    DM2_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    static int mc_unroll = 0, grid_mult = 0;
    if (mc_unroll == 0) {                 // tuning knobs (defaults measured on 8 x B200)
        const char* u = getenv("DM2_PEER_MC_UNROLL");
        const char* g = getenv("DM2_PEER_GRID_MULT");
        mc_unroll = u ? atoi(u) : 4;
        grid_mult = g ? atoi(g) : 4;
        if (grid_mult < 1) grid_mult = 1;
    }
    const int grid = (int)std::min<long long>((e4 - b4 + 255) / 256, (long long)num_sms * grid_mult);
    ProfScope _ps("peer_allreduce", s);
    if (multicast_ptr) {
        float* mc = reinterpret_cast<float*>(multicast_ptr);
        if (mc_unroll >= 8) k_peer_allreduce_mc<8><<<grid, 256, 0, s>>>(mc, b4, e4);
        else if (mc_unroll >= 4) k_peer_allreduce_mc<4><<<grid, 256, 0, s>>>(mc, b4, e4);
        else if (mc_unroll >= 2) k_peer_allreduce_mc<2><<<grid, 256, 0, s>>>(mc, b4, e4);
        else k_peer_allreduce_mc<1><<<grid, 256, 0, s>>>(mc, b4, e4);
    } else {
        PeerPtrs ptrs;
        for (int q = 0; q < 16; q++) ptrs.p[q] = q < world ? reinterpret_cast<float*>(const_cast<void*>(peer_ptrs[q])) : nullptr;
        for (int q = 0; q < world; q++) DM2_REQUIRE(ptrs.p[q] != nullptr, "dm2_peer_allreduce_sum: null peer pointer");
        switch (world) {
            case 2: launch_p2p<2>(ptrs, b4, e4, grid, s); break;
            case 3: launch_p2p<3>(ptrs, b4, e4, grid, s); break;
            case 4: launch_p2p<4>(ptrs, b4, e4, grid, s); break;
            case 5: launch_p2p<5>(ptrs, b4, e4, grid, s); break;
            case 6: launch_p2p<6>(ptrs, b4, e4, grid, s); break;
            case 7: launch_p2p<7>(ptrs, b4, e4, grid, s); break;
            case 8: launch_p2p<8>(ptrs, b4, e4, grid, s); break;
            case 16: launch_p2p<16>(ptrs, b4, e4, grid, s); break;
            default: DM2_REQUIRE(false, "dm2_peer_allreduce_sum: unsupported world size (2..8, 16)");
        }
    }
    DM2_LAUNCH_CHECK();
    return 0;
}

}  // namespace dm2
