// dm2_common.cuh — shared declarations for libdm2_b200 (sm_100a only).
//
// Numerics policy.  Everything that feeds a discrete decision the reference makes (tile bbox,
// depth keys, AA inside/clip tests, ray/triangle hits, tet-walk sign tests, early termination)
// is written with explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fmaf_rn, IEEE
// division and sqrt) so that the instruction sequence is fixed by the source, not by the
// optimiser.  The sequences follow the FMA-contraction pattern nvcc 12.9 produces for the
// reference's expressions on sm_100a:
//      a*b - c*d          ->  fma(a, b, -rn(c*d))
//      a*b + c*d          ->  fma(a, b,  rn(c*d))
//      a*b + c*d + e*f    ->  fma(e, f, fma(a, b, rn(c*d)))
//      x + a*b            ->  fma(a, b, x)
// The CPU oracle (oracle/dm2_oracle.cpp) restates the same sequences with fmaf().
#pragma once
#include <algorithm>
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dm2_renderer.h"

namespace dm2 {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
void count_launch();
// optional per-kernel timing with CUDA events on the launching stream (dm2_profile_* in the C-ABI)
int prof_begin(const char* name, cudaStream_t s);      // returns a handle (-1 when profiling is off)
void prof_end(int handle, cudaStream_t s);
struct ProfScope {
    cudaStream_t s;
    int h;
    ProfScope(const char* name, cudaStream_t st) : s(st), h(prof_begin(name, st)) {}
    ~ProfScope() { prof_end(h, s); }
};
#define DM2_CHECK(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            dm2::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,           \
                           cudaGetErrorString(_e));                                       \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)
#define DM2_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        dm2::count_launch();                                                              \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            dm2::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,       \
                           cudaGetErrorString(_e));                                       \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)
#define DM2_REQUIRE(cond, msg)                                                            \
    do {                                                                                  \
        if (!(cond)) { dm2::set_error("%s", msg); return 2; }                             \
    } while (0)

// cudaFuncSetAttribute applies to the CURRENT device: one-time kernel set-up is tracked per device, so that
// renderers on several GPUs of one process all get their shared-memory opt-in.  Thread-safe: first() stays
// true for a device until done() was called for it AFTER the attribute calls succeeded, so threads that race
// on the first launch each perform the (idempotent) set-up themselves before they launch; none can skip it.
struct PerDeviceOnce {
    std::atomic<unsigned long long> ready{0ull};
    static unsigned long long bit() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) return 0ull;
        return 1ull << (d & 63);
    }
    bool first() const { const unsigned long long b = bit(); return b == 0ull || !(ready.load(std::memory_order_acquire) & b); }
    void done() { ready.fetch_or(bit(), std::memory_order_release); }
};

// dm2_set_exact_clip(): every partially covered (pixel, face) pair takes the reference's polygon construction
// (dm2_aa.cuh) instead of the polygon-free area (dm2_aa_fast.cuh) - process-wide, defined in dm2_capi.cu
bool exact_clip_requested();

// SMs of the current device (persistent grids)
static inline int sm_count() {
    static std::atomic<int> cached[64];
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return 148;
    int v = cached[d & 63].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) return 148;
        cached[d & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kTile = DM2_TILE;
constexpr float kTEps = 1e-4f;           // T_EPS, reference cuda_impl/auxiliary.h:9

// ---------------------------------------------------------------- packed face record (render)
// One 256-byte record per (view, face), written once by the preprocess kernel and staged into
// shared memory by the tile kernels with 16-byte loads.  Word offsets (32-bit words):
constexpr int kRecWords   = 64;
constexpr int kRecBytes   = kRecWords * 4;
constexpr int R_BBOX      = 0;    // txmin txmax tymin tymax      (min/max of aa_face_verts)
constexpr int R_N01       = 4;    // n0.x n0.y n1.x n1.y          (aa_face_edges_normal)
constexpr int R_N2C       = 8;    // n2.x n2.y c0 c1              (…, aa_face_edges_normal_c)
constexpr int R_MISC      = 12;   // c2, opacity, intense, flags (bits 0..5: edges_iszero[e][k] at bit 2e+k; bit 8: the fused
                                  // preprocess flipped the face to CCW; bit 9: some screen coordinate is an integer;
                                  // bits 10, 11: kAafNearInt, kAafNever of dm2_aa_fast.cuh)
constexpr int R_P0        = 16;   // p0.xyz (world), z0 (ndc)
constexpr int R_P1        = 20;
constexpr int R_P2        = 24;
constexpr int R_COL       = 28;   // c0.rgb c1.rgb c2.rgb  (9 words: 28..36)
constexpr int R_VID       = 37;   // vertex ids (3 ints: 37..39)
constexpr int R_AAV       = 40;   // aa_face_verts   3 x (x,y)   40..45
constexpr int R_AAE       = 46;   // aa_face_edges   3 x (x,y)   46..51
constexpr int R_AAR       = 52;   // aa_face_edges_recip         52..57
constexpr int R_FID       = 58;   // face id (int)
constexpr int R_SAFE      = 59;   // thresholds of the polygon-free overlap area (dm2_aa_fast.cuh: err | tau_v as bf16)

// packed record of the layered path: 12 words = 48 bytes per (view, face)
constexpr int kLRecWords  = 12;   // p0.xyz p1.xyz p2.xyz min_depth max_depth face_id

// ---------------------------------------------------------------- device math
struct f3 { float x, y, z; };

__device__ __forceinline__ f3 sub3(const f3& a, const f3& b) {
    return { __fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z) };
}
// helper_math cross(): (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
__device__ __forceinline__ f3 cross3(const f3& a, const f3& b) {
    return { __fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)),
             __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
             __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)) };
}
// helper_math dot(): a.x*b.x + a.y*b.y + a.z*b.z
__device__ __forceinline__ float dot3(const f3& a, const f3& b) {
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}

// Moller-Trumbore without rejection (reference cuda_impl/auxiliary.h:212-243).
// Returns false iff denom == 0.  tuv = (t, u, v).
__device__ __forceinline__ bool ray_tri(const f3& o, const f3& d, const f3& p0, const f3& p1,
                                        const f3& p2, float& t, float& u, float& v) {
    f3 T  = sub3(o, p0);
    f3 E1 = sub3(p1, p0);
    f3 E2 = sub3(p2, p0);
    f3 Pv = cross3(d, E2);
    f3 Q  = cross3(T, E1);
    float denom = dot3(Pv, E1);
    if (denom == 0.0f) return false;
    float inv = __fdiv_rn(1.0f, denom);
    t = __fmul_rn(dot3(Q, E2), inv);
    u = __fmul_rn(dot3(Pv, T), inv);
    v = __fmul_rn(dot3(Q, d), inv);
    return true;
}

// 7-region barycentric clamp (reference cuda_impl/auxiliary.h:292-329).
__device__ __forceinline__ void clamp_bary(float u, float v, float& uc, float& vc, int& code) {
    if (u >= 0.0f && v >= 0.0f && __fadd_rn(u, v) <= 1.0f) { uc = u; vc = v; code = 0; }
    else if (u <= 0.0f && v <= 0.0f) { uc = 0.0f; vc = 0.0f; code = 1; }
    else if ((u >= 1.0f && v <= 0.0f) || (v >= 0.0f && v <= __fsub_rn(u, 1.0f))) { uc = 1.0f; vc = 0.0f; code = 2; }
    else if ((u <= 0.0f && v >= 1.0f) || (u >= 0.0f && v >= __fadd_rn(u, 1.0f))) { uc = 0.0f; vc = 1.0f; code = 3; }
    else if (u <= 0.0f && v <= 1.0f && v >= 0.0f) { uc = 0.0f; vc = v; code = 4; }
    else if (u <= 1.0f && u >= 0.0f && v <= 0.0f) { uc = u; vc = 0.0f; code = 5; }
    else {
        uc = __fmul_rn(__fsub_rn(__fadd_rn(1.0f, u), v), 0.5f);
        vc = __fmul_rn(__fadd_rn(__fsub_rn(1.0f, u), v), 0.5f);
        code = 6;
    }
}

// getPatchRectFromTri (reference cuda_impl/auxiliary.h:72-92).  No FMA in this chain.
__device__ __forceinline__ void patch_rect(uint32_t pmx, uint32_t pmy, float p0x, float p0y, float p1x,
                                           float p1y, float p2x, float p2y, int gx, int gy,
                                           int& x0, int& y0, int& x1, int& y1) {
    float fpx = (float)pmx, fpy = (float)pmy;          // float - uint  ->  uint converted to float
    float min_x = __fsub_rn(fminf(fminf(p0x, p1x), p2x), fpx);
    float min_y = __fsub_rn(fminf(fminf(p0y, p1y), p2y), fpy);
    float max_x = __fsub_rn(fmaxf(fmaxf(p0x, p1x), p2x), fpx);
    float max_y = __fsub_rn(fmaxf(fmaxf(p0y, p1y), p2y), fpy);
    int ix0 = (int)floorf(__fmul_rn(min_x, 0.0625f));
    int iy0 = (int)floorf(__fmul_rn(min_y, 0.0625f));
    int ix1 = (int)ceilf(__fmul_rn(max_x, 0.0625f));
    int iy1 = (int)ceilf(__fmul_rn(max_y, 0.0625f));
    x0 = min(gx, max(0, ix0));
    y0 = min(gy, max(0, iy0));
    x1 = min(gx, max(0, ix1));
    y1 = min(gy, max(0, iy1));
}

// clip = ([v,1] mv^T) proj^T with the accumulation order cuBLAS uses for these K=4 products on
// sm_100 (k = 0..3, first term a plain product, then fma) - bit-identical to torch.matmul
// (tools/probe_matmul_rounding.py).  m / p: row-major 4x4.
__device__ __forceinline__ float4 clip_vertex(const float* __restrict__ v, const float* __restrict__ m,
                                              const float* __restrict__ p) {
    float vw[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
        vw[j] = __fmaf_rn(1.0f, m[4 * j + 3], __fmaf_rn(v[2], m[4 * j + 2], __fmaf_rn(v[1], m[4 * j + 1], __fmul_rn(v[0], m[4 * j]))));
    float c[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
        c[j] = __fmaf_rn(vw[3], p[4 * j + 3], __fmaf_rn(vw[2], p[4 * j + 2], __fmaf_rn(vw[1], p[4 * j + 1], __fmul_rn(vw[0], p[4 * j]))));
    return make_float4(c[0], c[1], c[2], c[3]);
}

// Pixel ray of frame pixel (x, y) of view `view` from prm = [inv(proj) | inv(mv)] of that view (reference
// __init__.py:198-237), with the rounding choices of recipe r in [1, 36]: r - 1 = 18 * T + 9 * D + 3 * M + N
//   D: centre / size as an IEEE division (0) or as centre * rn(1 / size) (1: what torch's CUDA division by a
//      host scalar does);
//   M: accumulation of the two K = 4 products: fma chain k = 0..3 after a plain first product (0: what cuBLAS
//      runs for them), fma chain k = 3..0 (1), two fma pairs added (2);
//   N: squared norm: fma chain x, y, z (0), (rn(x^2) + rn(y^2)) + rn(z^2) (1), (rn(x^2) + rn(z^2)) + rn(y^2)
//      (2: torch's reduction kernel on sm_100);
//   T: 1 = the last (V H W) mod 65535 pixels of the whole ray tensor use (rn(p0) + rn(p2)) + (rn(p1) + rn(p3))
//      for both products: torch runs the (V H W)-batch matmul in chunks of 65535 (the grid limit), and cuBLAS
//      picks another kernel for a small last chunk (measured on B200, tools/probe_ray_stages.py).
__device__ __forceinline__ float ray_dot4(const float c0, const float c1, const float c2, const float c3,
                                          const float* __restrict__ m, int M) {
    if (M == 0) return __fmaf_rn(c3, m[3], __fmaf_rn(c2, m[2], __fmaf_rn(c1, m[1], __fmul_rn(c0, m[0]))));
    if (M == 1) return __fmaf_rn(c0, m[0], __fmaf_rn(c1, m[1], __fmaf_rn(c2, m[2], __fmul_rn(c3, m[3]))));
    if (M == 2) return __fadd_rn(__fmaf_rn(c1, m[1], __fmul_rn(c0, m[0])), __fmaf_rn(c3, m[3], __fmul_rn(c2, m[2])));
    return __fadd_rn(__fadd_rn(__fmul_rn(c0, m[0]), __fmul_rn(c2, m[2])), __fadd_rn(__fmul_rn(c1, m[1]), __fmul_rn(c3, m[3])));
}
__device__ __forceinline__ void make_ray(const float* __restrict__ prm, int view, int x, int y, int W, int H, int V,
                                         int recipe, f3& o, f3& d) {
    const int r = recipe - 1;
    const int T = r / 18, D = (r / 9) % 2, N = r % 3;
    int M = (r / 3) % 3;
    if (T) {
        const long long total = (long long)V * H * W, row = ((long long)view * H + y) * W + x;
        if (row >= total - total % 65535) M = 3;
    }
    const float cx = (float)x + 0.5f, cy = (float)y + 0.5f;
    const float fx = D ? __fmul_rn(cx, __fdiv_rn(1.0f, (float)W)) : __fdiv_rn(cx, (float)W);
    const float fy = D ? __fmul_rn(cy, __fdiv_rn(1.0f, (float)H)) : __fdiv_rn(cy, (float)H);
    const float nx = __fsub_rn(__fmul_rn(fx, 2.0f), 1.0f), ny = __fsub_rn(__fmul_rn(fy, 2.0f), 1.0f);
    const float* ip = prm;            // inv(proj), row-major: view[j] = sum_k clip[k] * ip[4 j + k]
    const float* im = prm + 16;       // inv(mv)
    float v[4], w[3];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = ray_dot4(nx, ny, -1.0f, 1.0f, ip + 4 * j, M);
#pragma unroll
    for (int j = 0; j < 3; j++) w[j] = ray_dot4(v[0], v[1], v[2], v[3], im + 4 * j, M);
    o = f3{im[3], im[7], im[11]};
    const float dx = __fsub_rn(w[0], o.x), dy = __fsub_rn(w[1], o.y), dz = __fsub_rn(w[2], o.z);
    float n2;
    if (N == 0) n2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    else if (N == 1) n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    else n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)), __fmul_rn(dy, dy));
    const float len = __fadd_rn(__fsqrt_rn(n2), 1e-6f);
    d = f3{__fdiv_rn(dx, len), __fdiv_rn(dy, len), __fdiv_rn(dz, len)};
}

__device__ __forceinline__ float clamp01_depth(float z) {
    float d = __fmul_rn(__fadd_rn(z, 1.0f), 0.5f);
    if (d < 0.0f) d = 0.0f;
    if (d > 1.0f) d = 1.0f;
    return d;
}

// ---------------------------------------------------------------- binning header + speculative launches
// k_tile_scan leaves three words behind the per-tile cursors: hdr[0] = num_rendered, hdr[1] = input flags
// raised by the preprocess kernel (kFlag*), hdr[2] = the longest tile segment.  Every kernel downstream of
// the scan can be enqueued BEFORE the host has read them back (dm2_*_forward_pipelined): it receives the
// capacity (in instances) of the binning workspace it was launched on and returns at once if the frame
// turned out not to fit or the inputs were rejected - the host then re-runs the tail on a larger workspace.
constexpr uint32_t kFlagPatchX = 1u;       // some patch_min.x + patch_width  > frame width  (reference __init__.py:283)
constexpr uint32_t kFlagPatchY = 2u;       // some patch_min.y + patch_height > frame height (reference __init__.py:284)
constexpr uint32_t kFlagPatchNeg = 4u;     // negative patch_min on the fused route (rays are read in place)
constexpr uint32_t kFlagVertexId = 8u;     // a face references a vertex id outside [0, P)
// longest_cap: the longest tile segment the sort tiers that were enqueued can handle (kernels that read the
// SORTED lists must not run on segments no launched tier has sorted; 0xFFFFFFFF for the binning kernels themselves)
struct BinGuard { const uint32_t* hdr; uint32_t capacity; uint32_t longest_cap; };
__device__ __forceinline__ bool bin_invalid(const BinGuard g) {
    return __ldg(g.hdr) > g.capacity || __ldg(g.hdr + 1) != 0u || __ldg(g.hdr + 2) > g.longest_cap;
}

// ---------------------------------------------------------------- internal launchers
struct FaceWs {
    float* depths; float* min_depths; float* max_depths;
    uint32_t* tiles_touched;
    uint2* rects;                 // packed tile bbox per (view,face): x = x0 | y0<<16, y = x1 | y1<<16
    float* records; size_t record_stride;
};
struct ImageWs {
    uint2* ranges; uint32_t* tile_cursor; float* final_T; float* final_prev_T; uint32_t* n_contrib;
    uint8_t* warp_done;           // u8[8 * tiles]: footprint finished (forward pass in rounds)
    uint32_t* round_work;         // u32[4 + 3 * tiles]: work lists of a round (RoundWork)
    uint32_t* sorted_upto;        // u32[tiles]: sorted prefix of every tile segment (lazy sort of deep scenes)
    int* first_face; int* first_tet;
};
// tiles (cls_list, n[0] entries) and half tiles (tile * 2 + half: fwd_list, n[1] entries) still running in a round;
// n[2]: next fwd_list item to hand out
struct RoundWork { uint32_t* n; uint32_t* cls_list; uint32_t* fwd_list; };
struct BinWs {
    uint64_t* keys_unsorted; uint32_t* vals_unsorted; uint64_t* keys; uint32_t* vals;
    uint8_t* imask;               // per sorted instance: bit w = footprint w has a touched pixel (render path)
    long long R;                  // capacity of this workspace in instances (>= num_rendered; = stride of meta)
    long long longest_cap = 0xFFFFFFFFll;   // longest tile segment the enqueued sort tiers cover (speculative launches)
    void* meta;                   // uint4[8][R]: per (footprint, instance) pixel classification (k_classify_instances)
};
struct TileMaskArgs {             // optional second output of the tile-range pass (render path)
    const uint32_t* vals; const float* records; const int32_t* patch_min; uint8_t* imask; void* meta;
    int F, gx, gy, use_bbox;
};

int launch_render_preprocess(const DM2RenderInputs& in, const FaceWs& fw, uint32_t* flags, cudaStream_t s);
int launch_layers_preprocess(const DM2LayerInputs& in, const FaceWs& fw, uint32_t* flags, cudaStream_t s);
int launch_scan(const uint32_t* in, uint32_t* out, long long n, uint32_t* partials, cudaStream_t s);
long long scan_partials_count(long long n);
int launch_tile_count_scan(long long BF, int F, int gx, int gy, int num_tiles, const uint32_t* tiles_touched,
                           const uint2* rects, uint2* ranges, uint32_t* cursor, cudaStream_t s);
// lazy segment sort of deep scenes (dm2_tilesort.cu): sorted_upto u32[num_tiles] receives the length of every tile's
// sorted prefix (nullptr: everything is sorted and nothing is recorded); prefix: list positions to sort at once
struct LazySort { uint32_t* sorted_upto; uint32_t prefix; };
int launch_tile_sort(long long BF, int F, int gx, int gy, int num_tiles, const float* key_depths,
                     const uint32_t* tiles_touched, const uint2* rects, const uint2* ranges, uint32_t* cursor,
                     uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys, uint32_t* vals,
                     long long capacity, long long longest, const LazySort& lazy, cudaStream_t s);
int launch_tile_sort_tiers(int num_tiles, const uint2* ranges, const uint32_t* cursor, const uint64_t* keys_unsorted,
                           const uint32_t* vals_unsorted, uint64_t* keys, uint32_t* vals, long long capacity, int tiers,
                           const LazySort& lazy, cudaStream_t s);
int launch_tile_sort_finish(int num_tiles, const uint2* ranges, const uint32_t* cursor, const uint64_t* keys_unsorted,
                            const uint32_t* vals_unsorted, uint64_t* keys, uint32_t* vals, long long capacity, int tiers,
                            const uint32_t* work, const uint32_t* work_n, long long need_end, const LazySort& lazy,
                            cudaStream_t s);
int launch_sort_pairs(const uint64_t* kin, const uint32_t* vin, uint64_t* kout, uint32_t* vout,
                      long long n, int end_bit, void* temp, size_t temp_bytes, cudaStream_t s);
size_t sort_temp_bytes(long long n);
int launch_classify(const uint64_t* keys, long long capacity, long long longest_cap, const uint32_t* hdr, const TileMaskArgs& tm,
                    cudaStream_t s);
long long sort_tiers_longest_cap(long long longest_hint);
// sort: (listed rounds only) the lazy sort's second phase runs between the work list and the classification
struct RoundSort { const uint64_t* keys_unsorted; const uint32_t* vals_unsorted; uint64_t* keys; uint32_t* vals;
                   const uint32_t* cursor; LazySort lazy; int tiers; };
int launch_classify_round(const uint64_t* keys, long long capacity, long long longest_cap, const uint32_t* hdr,
                          const TileMaskArgs& tm, const uint2* ranges, const uint8_t* warp_done, int num_tiles,
                          int round_begin, int round_len, const RoundWork& work, const RoundSort& sort, cudaStream_t s);
// round_end == 0: one pass over the whole lists; else list entries [round_begin, round_end) of every tile, resuming
// from / leaving the per-pixel state in out_color / out_depth / final_T / final_prev_T / n_contrib (iw.warp_done)
int launch_raster_forward(const DM2RenderInputs& in, const FaceWs& fw, const BinWs& bw,
                          const ImageWs& iw, float* out_color, float* out_depth, void* shade_cache,
                          long long shade_cache_bytes, int round_begin, int round_end, const RoundWork& work,
                          cudaStream_t s);
int launch_raster_backward(const DM2RenderInputs& in, const FaceWs& fw, const BinWs& bw,
                           const ImageWs& iw, const float* dL_dcolor, const float* dL_ddepth,
                           void* grad_ws, const void* shade_cache, long long shade_cache_bytes, cudaStream_t s);
int launch_generate_rays(const float* ray_params, int V, int H, int W, int recipe, float* out_o, float* out_d, cudaStream_t s);
int launch_peer_allreduce(const void* const* peer_ptrs, void* multicast_ptr, const void* const* peer_flag_ptrs,
                          uint32_t epoch, int world, int rank, long long n_floats, cudaStream_t s);
int launch_grad_scatter(const DM2RenderInputs& in, const FaceWs& fw, const DM2RenderGrads& g, void* grad_ws,
                        cudaStream_t s);
int launch_fused_preprocess(const DM2FusedInputs& in, const FaceWs& fw, uint32_t* patch_flags, cudaStream_t s);
int launch_fused_grad_scatter(const DM2FusedInputs& in, const FaceWs& fw, const DM2FusedGrads& g, void* grad_ws,
                              cudaStream_t s);
size_t backward_ws_bytes(long long BF);
int launch_layers(const DM2LayerInputs& in, const FaceWs& fw, const BinWs& bw, const ImageWs& iw,
                  void* geo_ws, int32_t* render_layers, int32_t* render_layers_cnt, bool walk_only, bool geometry_ready,
                  cudaStream_t s);
int launch_layers_geometry(const DM2LayerInputs& in, void* geo_ws, bool links_too, cudaStream_t s);
size_t layers_geo_bytes(int F, int T);
int launch_layers_composite_forward(int B, int H, int W, int L, int F, const int32_t* layers, const int32_t* cnt,
                                    const float* color, const float* opacity, const float* intense, const float* background,
                                    float* out_color, float* out_alpha, cudaStream_t s);
int launch_layers_composite_backward(int B, int H, int W, int L, int F, const int32_t* layers, const int32_t* cnt,
                                     const float* color, const float* opacity, const float* intense, const float* background,
                                     const float* g_color, const float* g_alpha, void* grad_ws, float* d_color,
                                     float* d_opacity, float* d_intense, cudaStream_t s);

}  // namespace dm2
