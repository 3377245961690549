// dm2_raster_fwd.cu — per-tile forward compositor.
// Behaviour of reference cuda_impl/forward.cu:139-432 (renderCUDA<3>), re-designed for sm_100a.
//
// One CTA per 16x16 tile (the tile size is part of the reference's binning contract), one warp per
// 8x4 pixel block, one pixel per lane.  Per round of kStage staged faces:
//   * 16-byte cp.async staging of the packed 256-byte records (272-byte stride in shared memory so
//     lanes reading DIFFERENT faces spread over the banks);
//   * each warp ballots "face bbox overlaps my 8x4 footprint" into a compact candidate list;
//   * per 16-candidate sub-chunk: lock-step classification (bbox + 12 half-plane tests), then the
//     polygon clip runs DENSE over the warp's list of partially covered (pixel, face) pairs, then
//     each lane walks its own hits in list order (ray/triangle, blend, early termination).
// The reference walks every listed face in lock-step with all 256 threads and re-reads ~40 AA
// floats per (pixel, face) from global memory.
#include "dm2_raster_common.cuh"

namespace dm2 {

constexpr int kStage = 128;                 // faces staged per round

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

}  // namespace dm2
