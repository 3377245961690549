// dm2_raster.cu — per-tile forward compositor and reverse-order backward pass.
// Behaviour of reference cuda_impl/forward.cu:139-432 (renderCUDA<3>) and
// cuda_impl/backward.cu:17-532, re-designed for sm_100a:
//   * one CTA per 16x16 tile, one warp per 8x4 pixel block (compact footprint);
//   * the tile's sorted face list is consumed in rounds of kStage packed 256-byte records,
//     copied global->shared with 16-byte cp.async (no register staging);
//   * per round each warp ballots "face bbox overlaps my 8x4 footprint" over the staged faces and
//     only walks the surviving faces (the reference lets all 256 threads test every face);
//   * AA operands are broadcast from shared memory; the clipper keeps no local-memory arrays;
//   * no AA cache: the backward pass recomputes the overlap area + gradient (the reference
//     zero-fills and round-trips a B*H*W*K*32-byte cache, render.cu:143-146);
//   * backward reduces the 29 per-face gradient values over the warp with a 31-shuffle
//     transposing butterfly, then issues ONE atomic per value per (warp, face) instead of 29
//     atomics per (pixel, face) (reference backward.cu:494-529).
#include "dm2_common.cuh"
#include "dm2_aa.cuh"

namespace dm2 {

constexpr int kStage = 128;                       // faces staged per round: 32 KB of records
constexpr int kRasterThreads = 256;

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
        cp_async16(smem + slot * kRecWords + chunk * 4, src);
    }
    cp_async_wait_all();
}

struct PixelGeom {
    bool inside;
    int px, py;             // pixel within the patch
    float pxmin, pymin;     // image-space lower-left corner (patch offset added)
    f3 o, d;
};

__device__ __forceinline__ bool footprint_hit(const float* rec, float fx0, float fy0) {
    const float4 bb = *reinterpret_cast<const float4*>(rec + R_BBOX);
    return !((fx0 + 8.0f < bb.x) || (fx0 > bb.y) || (fy0 + 4.0f < bb.z) || (fy0 > bb.w));
}

// Everything the forward and the backward pass must agree on bit-for-bit for one (pixel, face):
// coverage ratio, clamped barycentrics, interpolated colour/depth.  Returns false when the
// reference `continue`s (forward.cu:337-381).
struct Shade {
    float area, ratio, w0, w1, w2, iC[3], iD;
    int code;
};

template <bool GRAD>
__device__ __forceinline__ bool shade(const float* __restrict__ rec, const PixelGeom& pg, float tau,
                                      Shade& s, float* __restrict__ g) {
    s.area = 0.0f;
    if (tau > 0.0f) {
        if (!aa_overlap<GRAD>(rec, pg.pxmin, pg.pymin, s.area, g)) return false;
    } else if (GRAD) {
#pragma unroll
        for (int k = 0; k < 6; k++) g[k] = 0.0f;
    }
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

__device__ __forceinline__ PixelGeom pixel_geom(const DM2RenderInputs& in, int b, int tile_x, int tile_y,
                                                int warp, int lane) {
    PixelGeom pg;
    pg.px = tile_x * kTile + (warp & 1) * 8 + (lane & 7);
    pg.py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
    pg.inside = (pg.px < in.W) && (pg.py < in.H);
    const uint32_t pmx = (uint32_t)in.patch_min[2 * b], pmy = (uint32_t)in.patch_min[2 * b + 1];
    pg.pxmin = (float)((uint32_t)pg.px + pmx);        // forward.cu:319-322 (uint + uint -> float)
    pg.pymin = (float)((uint32_t)pg.py + pmy);
    pg.o = f3{0, 0, 0}; pg.d = f3{0, 0, 0};
    if (pg.inside) {
        const size_t pix = ((size_t)b * in.H + pg.py) * in.W + pg.px;
        const float* ro = in.image_ray_o + pix * 3;
        const float* rd = in.image_ray_d + pix * 3;
        pg.o = f3{ro[0], ro[1], ro[2]};
        pg.d = f3{rd[0], rd[1], rd[2]};
    }
    return pg;
}

// =============================================================================== forward
__global__ void __launch_bounds__(kRasterThreads, 2)
k_raster_forward(DM2RenderInputs in, const float* __restrict__ records, const uint32_t* __restrict__ list,
                 const uint2* __restrict__ ranges, float* __restrict__ final_T, float* __restrict__ final_prev_T,
                 uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
                 int gx, int gy) {
    __shared__ __align__(16) float srec[kStage * kRecWords];
    const int tiles = gx * gy;
    const int b = blockIdx.x / tiles, t_in = blockIdx.x % tiles;
    const int tile_x = t_in % gx, tile_y = t_in / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PixelGeom pg = pixel_geom(in, b, tile_x, tile_y, warp, lane);
    const float tau = in.aa_temperature;
    const float fx0 = (float)((uint32_t)(tile_x * kTile + (warp & 1) * 8) + (uint32_t)in.patch_min[2 * b]);
    const float fy0 = (float)((uint32_t)(tile_y * kTile + (warp >> 1) * 4) + (uint32_t)in.patch_min[2 * b + 1]);

    const uint2 range = ranges[blockIdx.x];
    const int total = (int)(range.y - range.x);
    const size_t view_base = (size_t)b * in.F;

    bool done = !pg.inside;
    float T = 1.0f, pT = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t last_contributor = 0;

    for (int base = 0; base < total; base += kStage) {
        // reference forward.cu:257-260: stop when the whole tile is done
        if (__syncthreads_count(done) == kRasterThreads) break;
        const int cnt = min(kStage, total - base);
        stage_records(srec, records, list, (long long)range.x + base, 1, cnt, view_base);
        __syncthreads();
        if (!__all_sync(0xffffffffu, done)) {
            for (int c = 0; c < cnt; c += 32) {
                const int j = c + lane;
                const bool cand = (j < cnt) && (tau > 0.0f ? footprint_hit(srec + j * kRecWords, fx0, fy0) : true);
                unsigned m = __ballot_sync(0xffffffffu, cand);
                while (m) {
                    const int jj = c + __ffs(m) - 1;
                    m &= m - 1;
                    if (!done) {
                        const float* rec = srec + jj * kRecWords;
                        Shade s;
                        if (shade<false>(rec, pg, tau, s, nullptr)) {
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
                    if (__all_sync(0xffffffffu, done)) { m = 0; c = cnt; }
                }
            }
        }
    }

    if (pg.inside) {
        const size_t pix = ((size_t)b * in.H + pg.py) * in.W + pg.px;
        final_prev_T[pix] = pT;
        final_T[pix] = T;
        n_contrib[pix] = last_contributor;
        out_color[pix * 3 + 0] = __fmaf_rn(T, in.background[0], C0);
        out_color[pix * 3 + 1] = __fmaf_rn(T, in.background[1], C1);
        out_color[pix * 3 + 2] = __fmaf_rn(T, in.background[2], C2);
        out_depth[pix] = __fadd_rn(D, T);
    }
}

int launch_raster_forward(const DM2RenderInputs& in, const FaceWs& fw, const BinWs& bw, const ImageWs& iw,
                          float* out_color, float* out_depth, cudaStream_t s) {
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    const long long nblk = (long long)in.B * gx * gy;
    if (nblk == 0) return 0;
    ProfScope _ps("raster_forward", s);
    k_raster_forward<<<(unsigned)nblk, kRasterThreads, 0, s>>>(in, fw.records, bw.vals, iw.ranges, iw.final_T,
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

__device__ __forceinline__ void ray_tri_grad(const f3& o, const f3& d, const f3& p0, const f3& p1, const f3& p2,
                                             f3& du_dp0, f3& du_dp1, f3& du_dp2, f3& dv_dp0, f3& dv_dp1,
                                             f3& dv_dp2) {
    const f3 T = subn3(o, p0), E1 = subn3(p1, p0), E2 = subn3(p2, p0);
    const f3 PE = crossn(d, E2);
    const float denom_sqrt = dotn(PE, E1);
    const float denom = denom_sqrt * denom_sqrt;
    const float denom_inv = 1.0f / denom;
    const float v0 = dotn(PE, T);
    const float v1 = denom_sqrt;
    const float v2 = dotn(crossn(T, E1), E2);
    const f3 du_dE1 = scale3(scale3(scale3(PE, -1.0f), v0), denom_inv);
    const f3 du_dE2 = scale3(subn3(scale3(crossn(T, d), v1), scale3(crossn(E1, d), v0)), denom_inv);
    const f3 du_dT  = scale3(scale3(PE, v1), denom_inv);
    const f3 dv_dE1 = scale3(subn3(scale3(crossn(E2, T), v1), scale3(PE, v2)), denom_inv);
    const f3 dv_dE2 = scale3(subn3(scale3(crossn(T, E1), v1), scale3(crossn(E1, d), v2)), denom_inv);
    const f3 dv_dT  = scale3(scale3(crossn(E1, E2), v1), denom_inv);
    du_dp0 = subn3(subn3(neg3(du_dE1), du_dE2), du_dT);
    dv_dp0 = subn3(subn3(neg3(dv_dE1), dv_dE2), dv_dT);
    du_dp1 = du_dE1; dv_dp1 = dv_dE1;
    du_dp2 = du_dE2; dv_dp2 = dv_dE2;
}

__global__ void __launch_bounds__(kRasterThreads, 1)
k_raster_backward(DM2RenderInputs in, const float* __restrict__ records, const uint32_t* __restrict__ list,
                  const uint2* __restrict__ ranges, const float* __restrict__ final_T,
                  const float* __restrict__ final_prev_T, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth, DM2RenderGrads gr,
                  int gx, int gy) {
    __shared__ __align__(16) float srec[kStage * kRecWords];
    __shared__ int s_max_last[8];
    const int tiles = gx * gy;
    const int b = blockIdx.x / tiles, t_in = blockIdx.x % tiles;
    const int tile_x = t_in % gx, tile_y = t_in / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PixelGeom pg = pixel_geom(in, b, tile_x, tile_y, warp, lane);
    const float tau = in.aa_temperature;
    const float fx0 = (float)((uint32_t)(tile_x * kTile + (warp & 1) * 8) + (uint32_t)in.patch_min[2 * b]);
    const float fy0 = (float)((uint32_t)(tile_y * kTile + (warp >> 1) * 4) + (uint32_t)in.patch_min[2 * b + 1]);

    const uint2 range = ranges[blockIdx.x];
    const size_t view_base = (size_t)b * in.F;
    const size_t pix = pg.inside ? ((size_t)b * in.H + pg.py) * in.W + pg.px : 0;

    // per-pixel state (backward.cu:133-160)
    const float T_final = pg.inside ? final_T[pix] : 0.0f;
    const float prev_T_final = pg.inside ? final_prev_T[pix] : 0.0f;
    const int my_last = pg.inside ? (int)n_contrib[pix] : 0;
    float T = prev_T_final;
    bool first_pass = true;
    float accum_rec[3] = {0, 0, 0}, accum_recd = 0.0f;
    float last_alpha = 0.0f, last_color[3] = {0, 0, 0}, last_depth = 0.0f;
    float dC[3] = {0, 0, 0}, dD = 0.0f;
    if (pg.inside) {
        dC[0] = dL_dcolor[pix * 3 + 0]; dC[1] = dL_dcolor[pix * 3 + 1]; dC[2] = dL_dcolor[pix * 3 + 2];
        dD = dL_ddepth[pix];
    }
    float bg_dot = 0.0f;
    bg_dot += in.background[0] * dC[0]; bg_dot += in.background[1] * dC[1]; bg_dot += in.background[2] * dC[2];
    const float bd_dot = dD;

    // faces behind every pixel's last contributor are never touched (backward.cu:219-221)
    const int warp_last = __reduce_max_sync(0xffffffffu, my_last);
    if (lane == 0) s_max_last[warp] = warp_last;
    __syncthreads();
    int block_last = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) block_last = max(block_last, s_max_last[i]);

    for (int hi = block_last; hi > 0; hi -= kStage) {
        const int cnt = min(kStage, hi);
        __syncthreads();
        // slot i holds list entry (hi-1-i): back-to-front
        stage_records(srec, records, list, (long long)range.x + hi - 1, -1, cnt, view_base);
        __syncthreads();
        if (hi - cnt >= warp_last) continue;                 // nothing for this warp in this round
        for (int c = 0; c < cnt; c += 32) {
            const int j = c + lane;
            const bool cand = (j < cnt) && (hi - 1 - j < warp_last) &&
                              (tau > 0.0f ? footprint_hit(srec + j * kRecWords, fx0, fy0) : true);
            unsigned m = __ballot_sync(0xffffffffu, cand);
            while (m) {
                const int jj = c + __ffs(m) - 1;
                m &= m - 1;
                const int idx = hi - 1 - jj;                  // position in the tile list
                const float* rec = srec + jj * kRecWords;
                float v[32];
#pragma unroll
                for (int k = 0; k < 32; k++) v[k] = 0.0f;
                bool contrib = false;
                if (idx < my_last) {
                    Shade s;
                    float g[6];
                    if (shade<true>(rec, pg, tau, s, g)) {
                        contrib = true;
                        const float opacity = rec[R_MISC + 1], intense = rec[R_MISC + 2];
                        const float alpha = __fmul_rn(opacity, s.ratio);
                        // backward.cu:340-348
                        if (!first_pass) T = T / (1.0f - alpha);
                        first_pass = false;
                        float dL_diC[3], dL_diD, dL_dalpha = 0.0f;
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
                            last_color[ch] = s.iC[ch];
                            dL_diC[ch] = dC[ch] * alpha * T;
                            dL_dalpha += (s.iC[ch] - accum_rec[ch]) * dC[ch];
                        }
                        accum_recd = last_alpha * last_depth + (1.0f - last_alpha) * accum_recd;
                        last_depth = s.iD;
                        dL_diD = dD * alpha * T;
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
                        ray_tri_grad(pg.o, pg.d, f3{a0.x, a0.y, a0.z}, f3{a1.x, a1.y, a1.z}, f3{a2.x, a2.y, a2.z},
                                     du0, du1, du2, dv0, dv1, dv2);
                        const f3 g0 = add3(scale3(du0, dL_du), scale3(dv0, dL_dv));
                        const f3 g1 = add3(scale3(du1, dL_du), scale3(dv1, dL_dv));
                        const f3 g2 = add3(scale3(du2, dL_du), scale3(dv2, dL_dv));
                        v[0] = g0.x; v[1] = g0.y; v[2] = g0.z;
                        v[3] = g1.x; v[4] = g1.y; v[5] = g1.z;
                        v[6] = g2.x; v[7] = g2.y; v[8] = g2.z;
                    }
                }
                if (!__any_sync(0xffffffffu, contrib)) continue;
                // transposing butterfly: afterwards lane l holds sum over the warp of v[l]
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < off; i++) {
                        const float send = upper ? v[i] : v[i + off];
                        const float keep = upper ? v[i + off] : v[i];
                        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                const float sum = v[0];
                if (lane < 29 && sum != 0.0f) {
                    const int fid = __float_as_int(rec[R_FID]);
                    float* dst;
                    if (lane < 9) dst = gr.dL_dverts + 3 * (size_t)__float_as_int(rec[R_VID + lane / 3]) + lane % 3;
                    else if (lane < 18) dst = gr.dL_dverts_color + 3 * (size_t)__float_as_int(rec[R_VID + (lane - 9) / 3]) + (lane - 9) % 3;
                    else if (lane < 21) dst = gr.dL_dverts_ndc + 3 * ((size_t)b * in.P + __float_as_int(rec[R_VID + lane - 18])) + 2;
                    else if (lane == 21) dst = gr.dL_dfaces_opacity + fid;
                    else if (lane == 22) dst = gr.dL_dfaces_intense + view_base + fid;
                    else dst = gr.dL_daa_face_verts + 6 * (view_base + fid) + (lane - 23);
                    atomicAdd(dst, sum);
                }
            }
        }
    }
}

int launch_raster_backward(const DM2RenderInputs& in, const FaceWs& fw, const BinWs& bw, const ImageWs& iw,
                           const float* dL_dcolor, const float* dL_ddepth, const DM2RenderGrads& g,
                           cudaStream_t s) {
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    const long long nblk = (long long)in.B * gx * gy;
    if (nblk == 0) return 0;
    ProfScope _ps("raster_backward", s);
    k_raster_backward<<<(unsigned)nblk, kRasterThreads, 0, s>>>(in, fw.records, bw.vals, iw.ranges, iw.final_T,
                                                                iw.final_prev_T, iw.n_contrib, dL_dcolor,
                                                                dL_ddepth, g, gx, gy);
    DM2_LAUNCH_CHECK();
    return 0;
}

}  // namespace dm2
