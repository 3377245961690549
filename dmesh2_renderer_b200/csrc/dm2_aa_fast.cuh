// dm2_aa_fast.cuh - pixel-square / triangle overlap area + gradient WITHOUT building the clip polygon.
//
// The reference (cuda_impl/aa.h:151-441) builds the clipped polygon vertex by vertex - 12 (triangle edge, pixel edge)
// intersection tests, runs of inside pixel corners, a triangle fan - and dm2_aa.cuh::aa_clip reproduces that construction
// bit for bit in ~640 predicated instructions per (pixel, face) pair: half of the forward kernel's instructions.
// This header evaluates the SAME area from the SAME rounded polygon vertices with Green's theorem over the vertical strip
// of the pixel instead:
//      area(T n S) = -sum over the triangle's edges of  INT clamp(y, 0, 1) dx   along the part of the edge inside the strip
// (pixel-local coordinates; T counter-clockwise in the reference's sense, aa.h:88-94).  The vertical strip lines contribute
// nothing (dx = 0), the pixel corners never appear, and per edge the integral needs only
//   * the strip end points - a triangle vertex inside the strip, or the crossing with x = pxmin / pxmax - and
//   * the crossings with y = pymin / pymax when the strip piece leaves the band,
// each computed with the reference's own expressions t = (c - p0) * recip, i = p0 + t * edge (aa.h:251-252) in frame
// coordinates, so every polygon vertex carries exactly the reference's rounding (~ulp(1000) = 6e-5 at coordinate 1000 -
// far above the 1e-7 at which two algebraically equal area formulas differ: the areas agree to ~1e-7, measured by
// tests/native/aa_fast_check.cpp against the oracle's transcription of aa.h on 1e8+ pairs).
// The gradient is the closed form dm2_aa.cuh documents (edge part [t_in, t_out] inside the pixel).
//
// The reference is NOT a function of the true geometry everywhere: where its independently rounded decisions disagree
// (inside-corner flags vs intersections, t within an ulp of 0 / 1, crossings that swap order on a pixel edge, "parallel"
// edges: |edge component| < 1e-3) it returns error codes (= the face is skipped for the pixel) or an inconsistent polygon.
// Those pairs must be reproduced, not corrected, so aa_area_fast() only answers for pairs that are robustly far from every
// such decision and returns false for the others - the caller then runs aa_clip() (+ aa_grad_reference()).  A pair is unsafe if
//   (S1) the face has a "parallel" edge component (edges_iszero) or non-finite / huge geometry (per-face flag);
//   (S2) a triangle vertex lies within tau_v of one of the pixel's boundary lines (and within the pixel's extent + tau_v
//        on the other axis).  tau_v is per face (bf16 in record word R_SAFE): 2 ulp of the face's largest coordinate,
//        divided by the smallest wedge factor |cot a - cot b| / |tan a - tan b| of the two edges at a vertex, capped at 1
//        (a thin wedge crossing a pixel edge just behind its apex produces two crossings closer than their rounding);
//        evaluated only for faces whose flag says that SOME coordinate is within tau_v of an integer;
//   (S3) a triangle edge's line passes within err (2.5 ulp of the face's largest coordinate) of a pixel corner;
//   (S4) the area is not in (err^2, 1 - 4e-7) (slivers, error 6 "area > pixel", NaN).
#pragma once

#ifdef __CUDACC__
#define AAF_FN __device__ __forceinline__
#define AAF_SUB(a, b) __fsub_rn(a, b)
#define AAF_MUL(a, b) __fmul_rn(a, b)
#define AAF_FMA(a, b, c) __fmaf_rn(a, b, c)
#define AAF_SAT(a) __saturatef(a)
#define AAF_BITS(f) __float_as_uint(f)
#define AAF_FLOAT(u) __uint_as_float(u)
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#define AAF_FN static inline
#define AAF_SUB(a, b) ((float)((a) - (b)))
#define AAF_MUL(a, b) ((float)((a) * (b)))
#define AAF_FMA(a, b, c) fmaf(a, b, c)
static inline float aaf_sat_host(float a) { return a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a); }
static inline uint32_t aaf_bits_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float aaf_float_host(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define AAF_SAT(a) aaf_sat_host(a)
#define AAF_BITS(f) aaf_bits_host(f)
#define AAF_FLOAT(u) aaf_float_host(u)
#endif

// thresholds in units of the face's largest |coordinate| (2^-23 = 1.19e-7 is one ulp at the top of a binade).
// Measured with tests/native/aa_fast_check.cpp (adversarial pairs: vertices on / next to pixel lines, edges through /
// next to corners, thin wedges behind a pixel edge, slivers, nearly axis-parallel edges, nearly covered pixels,
// coordinates up to 2040): contract violations appear at AAF_ERR_SCALE 1.2e-7 (5 in 1e7 pairs) and AAF_TAUV_SCALE
// 0.6e-7, none in 8e7 pairs at 2.4e-7 / 1.2e-7; the defaults keep a factor of 2.5 / 4 over the failing values.
#ifndef AAF_ERR_SCALE
#define AAF_ERR_SCALE 3.0e-7f
#endif
#ifndef AAF_TAUV_SCALE
#define AAF_TAUV_SCALE 2.4e-7f
#endif

namespace dm2 {
#ifdef AAF_DEBUG_REASON
static int aaf_reason = 0;       // host-side harness only: which test declined the last pair
#endif

// flag word of the record (R_MISC + 3): bits 0..5 edges_iszero, 8 flipped, 9 integer coordinate (see dm2_common.cuh)
constexpr unsigned kAafNearInt = 1u << 10;     // some screen coordinate within tau_v of an integer: run the (S2) test
constexpr unsigned kAafNever   = 1u << 11;     // never trust the fast path for this face
constexpr unsigned kAafSlowMask = 0x3Fu | kAafNever;

// per-face safety word: high half = err as bf16, low half = tau_v as bf16, both rounded UP
AAF_FN uint32_t aaf_bf16_up(float x) {
    uint32_t u = AAF_BITS(x);
    return (u & 0xFFFFu) ? ((u >> 16) + 1u) : (u >> 16);      // x >= 0 finite: next bf16 above
}
// verts: 3 x (x, y) frame coordinates (CCW), edges[i] = v[i+1] - v[i], flags: the record's flag word WITHOUT the two
// fast-path bits.  Returns the flag bits to OR in; *safe_word receives the packed thresholds.
AAF_FN unsigned aaf_face_safety(const float* verts, const float* edges, const float* recip, unsigned flags, uint32_t* safe_word) {
    float cmax = 1.0f;
    bool finite = true;
    for (int k = 0; k < 6; k++) {
        const float a = fabsf(verts[k]);
        finite = finite && (a < 1.0e6f) && (fabsf(edges[k]) < 1.0e6f);      // (NaN fails the comparison)
        cmax = fmaxf(cmax, a);
    }
    const float err = AAF_ERR_SCALE * cmax;
    unsigned out = 0u;
    if (!finite || (flags & 0x3Fu)) { *safe_word = 0u; return kAafNever; }
    // the fast path relies on edges = vertex differences and recip = 1 / edges (what Triangles computes,
    // pyrenderer.py:10-12); a caller of the 21-argument function who passes anything else gets the exact clip
    for (int i = 0; i < 3; i++) {
        const int j = (i + 1) % 3;
        const bool same = (edges[2 * i] == AAF_SUB(verts[2 * j], verts[2 * i])) & (edges[2 * i + 1] == AAF_SUB(verts[2 * j + 1], verts[2 * i + 1])) &
                          (fabsf(AAF_FMA(recip[2 * i], edges[2 * i], -1.0f)) <= 1.0e-6f) &
                          (fabsf(AAF_FMA(recip[2 * i + 1], edges[2 * i + 1], -1.0f)) <= 1.0e-6f);
        if (!same) { *safe_word = 0u; return kAafNever; }
    }
    // wedge factors: edges i-1 (arriving) and i (leaving) at vertex i
    float kmin = 1.0f;
    for (int i = 0; i < 3; i++) {
        const int p = (i + 2) % 3;
        const float cot_a = edges[2 * p] * recip[2 * p + 1], cot_b = edges[2 * i] * recip[2 * i + 1];
        const float tan_a = edges[2 * p + 1] * recip[2 * p], tan_b = edges[2 * i + 1] * recip[2 * i];
        kmin = fminf(kmin, fminf(fabsf(cot_a - cot_b), fabsf(tan_a - tan_b)));
    }
    if (!(kmin > 2.0e-3f)) { *safe_word = 0u; return kAafNever; }           // needle: tau_v would span pixels
    const float tauv = AAF_TAUV_SCALE * cmax / kmin;
    const uint32_t e16 = aaf_bf16_up(err), t16 = aaf_bf16_up(tauv);
    *safe_word = (e16 << 16) | t16;
    const float tv = AAF_FLOAT(t16 << 16);
    for (int k = 0; k < 6; k++)
        if (fabsf(verts[k] - rintf(verts[k])) <= tv) out |= kAafNearInt;
    return out;
}

// One (face, pixel) pair.  verts / edges / recip: the record's aa_face_verts / aa_face_edges / aa_face_edges_recip
// (3 x float2 each), flags + safe_word as written by aaf_face_safety.  Returns true when the pair is safe: area_out in
// (0, 1), g[6] = d area / d verts (GRAD).  false: evaluate the pair with aa_clip().
template <bool GRAD>
AAF_FN bool aa_area_fast(const float* __restrict__ verts, const float* __restrict__ edges, const float* __restrict__ recip,
                         unsigned flags, uint32_t safe_word, float pxmin, float pymin, float& area_out, float* __restrict__ g) {
#ifdef AAF_DEBUG_REASON
#define AAF_REASON(bit, cond) do { if (cond) aaf_reason |= (bit); } while (0)
    aaf_reason = 0;
#else
#define AAF_REASON(bit, cond) do { } while (0)
#endif
    AAF_REASON(1, flags & kAafSlowMask);
    if (flags & kAafSlowMask) return false;
    const float err = AAF_FLOAT(safe_word & 0xFFFF0000u);
    const float pxmax = pxmin + 1.0f, pymax = pymin + 1.0f;
    float vx[3], vy[3], lx[3], ly[3], sx[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        vx[k] = verts[2 * k]; vy[k] = verts[2 * k + 1];
        lx[k] = AAF_SUB(vx[k], pxmin); ly[k] = AAF_SUB(vy[k], pymin);
        sx[k] = AAF_SAT(lx[k]);
    }
    bool unsafe = false;
    if (flags & kAafNearInt) {                         // (S2), rare
        const float tv = AAF_FLOAT(safe_word << 16), hi = 1.0f + tv;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float dx = fminf(fabsf(lx[k]), fabsf(lx[k] - 1.0f)), dy = fminf(fabsf(ly[k]), fabsf(ly[k] - 1.0f));
            const bool inx = (lx[k] >= -tv) & (lx[k] <= hi), iny = (ly[k] >= -tv) & (ly[k] <= hi);
            unsafe |= ((dx <= tv) & iny) | ((dy <= tv) & inx);
            AAF_REASON(2, ((dx <= tv) & iny) | ((dy <= tv) & inx));
        }
    }
    float acc = 0.0f;
    float gr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ti = 0; ti < 3; ti++) {
        const int tj = (ti == 2) ? 0 : ti + 1;
        const float ax = vx[ti], ay = vy[ti];
        const float ex = edges[2 * ti], ey = edges[2 * ti + 1];
        const float rx = recip[2 * ti], ry = recip[2 * ti + 1];
        // the reference's line parameters (aa.h:251) for the four pixel boundary lines
        const float tx0 = AAF_MUL(AAF_SUB(pxmin, ax), rx), tx1 = AAF_MUL(AAF_SUB(pxmax, ax), rx);
        const float ty0 = AAF_MUL(AAF_SUB(pymin, ay), ry), ty1 = AAF_MUL(AAF_SUB(pymax, ay), ry);
        // strip end points: the vertex itself inside the strip, else the crossing with the strip line it lies behind
        const bool a_l = ax < pxmin, a_r = ax > pxmax, b_l = vx[tj] < pxmin, b_r = vx[tj] > pxmax;
        const float tA = a_l ? tx0 : (a_r ? tx1 : 0.0f);
        const float tB = b_l ? tx0 : (b_r ? tx1 : 1.0f);
        const float XA = sx[ti], XB = sx[tj];
        const float YA = AAF_SUB(AAF_FMA(tA, ey, ay), pymin);                 // fma(0, ey, ay) = ay exactly
        const float YB = (b_l | b_r) ? AAF_SUB(AAF_FMA(tB, ey, ay), pymin) : ly[tj];
        // crossings with the band lines (aa.h:252), pixel-local
        const float xc0 = AAF_SUB(AAF_FMA(ty0, ex, ax), pxmin), xc1 = AAF_SUB(AAF_FMA(ty1, ex, ax), pxmin);
        const bool a_lo = YA < 0.0f, a_hi = YA > 1.0f, b_lo = YB < 0.0f, b_hi = YB > 1.0f;
        const float XAp = a_lo ? xc0 : (a_hi ? xc1 : XA), XBp = b_lo ? xc0 : (b_hi ? xc1 : XB);
        const float YAp = AAF_SAT(YA), YBp = AAF_SAT(YB);
        const float inband = (XBp - XAp) * (0.5f * (YAp + YBp));
        float above = (a_hi ? XAp - XA : 0.0f) + (b_hi ? XB - XBp : 0.0f);
        above = (a_hi & b_hi) ? XB - XA : above;
        acc += inband + above;
        if (GRAD) {
            const float tAp = a_lo ? ty0 : (a_hi ? ty1 : tA), tBp = b_lo ? ty0 : (b_hi ? ty1 : tB);
            const float len = fmaxf(tBp - tAp, 0.0f);
            const float w1 = len * (0.5f * (tAp + tBp)), w0 = len - w1;
            gr[2 * ti] += ey * w0; gr[2 * ti + 1] -= ex * w0;
            gr[2 * tj] += ey * w1; gr[2 * tj + 1] -= ex * w1;
        }
        // (S3) distance of the edge's line from the four pixel corners, scaled by the edge length
        const float d00 = ey * lx[ti] - ex * ly[ti];
        const float d10 = d00 - ey, d01 = d00 + ex, d11 = d10 + ex;
        const float dmin = fminf(fminf(fabsf(d00), fabsf(d10)), fminf(fabsf(d01), fabsf(d11)));
        unsafe |= dmin <= err * (fabsf(ex) + fabsf(ey));
        AAF_REASON(4, dmin <= err * (fabsf(ex) + fabsf(ey)));
    }
    const float area = -acc;
    unsafe |= !((area > err * err) & (area < 1.0f - 4.0e-7f));                // (S4); NaN lands here
    AAF_REASON(8, !(area > err * err));
    AAF_REASON(16, !(area < 1.0f - 4.0e-7f));
    area_out = area;
    if (GRAD) {
#pragma unroll
        for (int k = 0; k < 6; k++) g[k] = gr[k];
    }
    return !unsafe;
}

}  // namespace dm2
