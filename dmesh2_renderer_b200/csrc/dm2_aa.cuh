// dm2_aa.cuh — exact pixel-square / triangle overlap area and its gradient w.r.t. the three
// screen-space triangle vertices (behaviour of reference cuda_impl/aa.h:96-504).
//
// Design differences from the reference (same results, different machine code):
//   * all per-face operands come from the packed record in shared memory instead of ~70 scalar
//     global loads through 64-bit accessor arithmetic;
//   * two stages: aa_classify() (bbox + 12 half-plane tests: outside / fully inside / partial) is the
//     cheap filter every (pixel, face) pair runs; aa_clip() (the polygon clip) only runs for
//     partially covered pixels;
//   * the clipped polygon is never materialised: vertices are streamed into a fan accumulator
//     (first / previous / current vertex), so there are no local-memory arrays (the reference keeps
//     polygon[10][2] + two [10][4] Jacobian arrays on the stack);
//   * the gradient is NOT chained through per-vertex 2x2 Jacobians (aa.h:280-294, 416-433).  For a
//     triangle edge whose part inside the pixel is the parameter interval [ta, tb], moving the edge
//     end points sweeps area  dA/dp_start = n (tb-ta - (tb^2-ta^2)/2),  dA/dp_end = n (tb^2-ta^2)/2
//     with n = (e.y, -e.x) the outward edge normal scaled by the edge length.  This is the same
//     derivative (it agrees with the reference's analytic gradient to 1e-6 on the reference's own
//     Python implementation, tests/golden/aa_overlap_golden.npz) at ~15 flops per edge.
// Error semantics: the reference returns codes 1..6 and the caller skips the face for that pixel
// on any non-zero code (forward.cu:337-338, backward.cu:286-289); here `false` = "skip".
#pragma once
#include "dm2_common.cuh"

namespace dm2 {

enum AAClass : int { AA_OUT = 0, AA_FULL = 1, AA_PART = 2 };

// Steps 1-3 of aa.h:446-496.  inmask bit k: pixel corner k lies inside all three edge half-planes.
__device__ __forceinline__ int aa_classify(const float* __restrict__ rec, float pxmin, float pymin,
                                           unsigned& inmask) {
    const float pxmax = __fadd_rn(pxmin, 1.0f);
    const float pymax = __fadd_rn(pymin, 1.0f);
    {   // triangle bounding box (aa.h:96-101)
        const float4 bb = *reinterpret_cast<const float4*>(rec + R_BBOX);
        if ((pxmax < bb.x) || (pxmin > bb.y) || (pymax < bb.z) || (pymin > bb.w)) return AA_OUT;
    }
    const float4 n01 = *reinterpret_cast<const float4*>(rec + R_N01);
    const float4 n2c = *reinterpret_cast<const float4*>(rec + R_N2C);
    const float c2 = rec[R_MISC];
    const float nx[3] = {n01.x, n01.z, n2c.x};
    const float ny[3] = {n01.y, n01.w, n2c.y};
    const float cc[3] = {n2c.z, n2c.w, c2};
    unsigned m = 0xF;
#pragma unroll
    for (int ti = 0; ti < 3; ti++) {
        // (vert[0]*n[0]) + (vert[1]*n[1]) - c >= 0, aa.h:20, for the 4 pixel corners (aa.h:103-149)
        const float ylo = __fmul_rn(pymin, ny[ti]);
        const float yhi = __fmul_rn(pymax, ny[ti]);
        const bool i0 = __fsub_rn(__fmaf_rn(pxmin, nx[ti], ylo), cc[ti]) >= 0.0f;
        const bool i1 = __fsub_rn(__fmaf_rn(pxmax, nx[ti], ylo), cc[ti]) >= 0.0f;
        const bool i2 = __fsub_rn(__fmaf_rn(pxmax, nx[ti], yhi), cc[ti]) >= 0.0f;
        const bool i3 = __fsub_rn(__fmaf_rn(pxmin, nx[ti], yhi), cc[ti]) >= 0.0f;
        if (!(i0 || i1 || i2 || i3)) return AA_OUT;          // all four corners outside one edge
        m &= (i0 ? 1u : 0u) | (i1 ? 2u : 0u) | (i2 ? 4u : 0u) | (i3 ? 8u : 0u);
    }
    inmask = m;
    return (m == 0xF) ? AA_FULL : AA_PART;
}

struct FanAcc {
    int np = 0;
    float area = 0.0f;
    float fx, fy, qx, qy;       // first and previous polygon vertex
    // returns false on polygon overflow (aa.h:45-48) or negative fan-triangle area (aa.h:410-412)
    __device__ __forceinline__ bool push(float x, float y) {
        if (np >= 10) return false;
        if (np == 0) { fx = x; fy = y; }
        else if (np == 1) { qx = x; qy = y; }
        else {
            // _tri_area(first, prev, v), aa.h:88-94
            const float ax = __fsub_rn(qx, fx), by = __fsub_rn(y, fy);
            const float cx = __fsub_rn(x, fx),  dy = __fsub_rn(qy, fy);
            const float s = __fmul_rn(0.5f, __fmaf_rn(ax, by, -__fmul_rn(cx, dy)));
            if (s < 0.0f) return false;
            area = __fadd_rn(area, s);
            qx = x; qy = y;
        }
        np++;
        return true;
    }
};

// Step 4 of aa.h:446-504 (= aa.h:151-441) for a partially covered pixel.  Returns false if the face
// must be skipped for this pixel (any reference error code, or zero overlap).  On true: area in
// (0, 1] (or NaN, which the reference also lets through); g[6] = d area / d tri verts when GRAD.
template <bool GRAD>
__device__ __forceinline__ bool aa_clip(const float* __restrict__ rec, float pxmin, float pymin, unsigned inmask,
                                        float& area_out, float* __restrict__ g) {
    const float pxmax = __fadd_rn(pxmin, 1.0f);
    const float pymax = __fadd_rn(pymin, 1.0f);
    FanAcc fan;
    const unsigned zflags = __float_as_uint(rec[R_MISC + 3]);
    float gr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

#pragma unroll 1
    for (int ti = 0; ti < 3; ti++) {
        const int tj = (ti == 2) ? 0 : ti + 1;
        const float2 p0 = *reinterpret_cast<const float2*>(rec + R_AAV + 2 * ti);
        const float2 p1 = *reinterpret_cast<const float2*>(rec + R_AAV + 2 * tj);
        const float2 ed = *reinterpret_cast<const float2*>(rec + R_AAE + 2 * ti);
        const float2 rc = *reinterpret_cast<const float2*>(rec + R_AAR + 2 * ti);
        const bool edge_vertical   = (zflags >> (2 * ti)) & 1u;      // edges_iszero[ti][0]
        const bool edge_horizontal = (zflags >> (2 * ti + 1)) & 1u;  // edges_iszero[ti][1]
        const bool in0 = (p0.x >= pxmin) && (p0.x <= pxmax) && (p0.y >= pymin) && (p0.y <= pymax);
        const bool in1 = (p1.x >= pxmin) && (p1.x <= pxmax) && (p1.y >= pymin) && (p1.y <= pymax);

        float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;     // up to two intersections, pixel-edge order
        float ta = 0.f, tb = 0.f;
        int pa = -1, pb = -1, n = 0;
#pragma unroll
        for (int pi = 0; pi < 4; pi++) {
            const bool horiz = (pi == 0) || (pi == 2);               // pixel edge y = const
            const bool parallel = horiz ? edge_horizontal : edge_vertical;
            const float cst  = (pi == 0) ? pymin : (pi == 1) ? pxmax : (pi == 2) ? pymax : pxmin;
            const float p0a0 = horiz ? p0.y : p0.x;                  // axis0 = the fixed axis
            const float p0a1 = horiz ? p0.x : p0.y;
            const float ra0  = horiz ? rc.y : rc.x;
            const float ea1  = horiz ? ed.x : ed.y;
            const float lo1  = horiz ? pxmin : pymin;
            const float hi1  = horiz ? pxmax : pymax;
            const float t  = __fmul_rn(__fsub_rn(cst, p0a0), ra0);   // aa.h:251
            const float i1 = __fmaf_rn(t, ea1, p0a1);                // aa.h:252
            const bool valid = (t >= 0.0f) && (t <= 1.0f) && (i1 >= lo1) && (i1 <= hi1) && !parallel;
            if (!valid) continue;
            if ((i1 == lo1) || (i1 == hi1)) return false;            // error 1: hits a pixel corner
            if (n >= 2) return false;                                // error 2: more than 2 intersections
            const float qx = horiz ? i1 : cst, qy = horiz ? cst : i1;
            if (n == 0) { ax = qx; ay = qy; ta = t; pa = pi; } else { bx = qx; by = qy; tb = t; pb = pi; }
            n++;
        }

        int final_pedge = -1;
        float t_in = 0.f, t_out = 0.f;        // part of the edge inside the pixel: [t_in, t_out]
        bool covered = true;
        if (n == 2) {
            // sort by t (aa.h:308-328); strict '>' keeps pixel-edge order on ties
            if (ta > tb) {
                if (!fan.push(bx, by) || !fan.push(ax, ay)) return false;
                final_pedge = pa; t_in = tb; t_out = ta;
            } else {
                if (!fan.push(ax, ay) || !fan.push(bx, by)) return false;
                final_pedge = pb; t_in = ta; t_out = tb;
            }
        } else if (n == 1) {
            if (!fan.push(ax, ay)) return false;
            if (!in0 && in1) { if (!fan.push(p1.x, p1.y)) return false; t_in = ta; t_out = 1.0f; }
            else if (in0 && !in1) { final_pedge = pa; t_in = 0.0f; t_out = ta; }
            else return false;                                       // error 3
        } else {
            if (in0 && in1) { if (!fan.push(p1.x, p1.y)) return false; t_in = 0.0f; t_out = 1.0f; }
            else if (in0 != in1) return false;                       // error 4
            else covered = false;                                    // edge misses the pixel
        }
        if (GRAD && covered) {
            const float w1 = 0.5f * (t_out * t_out - t_in * t_in);
            const float w0 = (t_out - t_in) - w1;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (k == ti) { gr[2 * k] += ed.y * w0; gr[2 * k + 1] -= ed.x * w0; }
                if (k == tj) { gr[2 * k] += ed.y * w1; gr[2 * k + 1] -= ed.x * w1; }
            }
        }
        if (final_pedge != -1) {
            // run of inside pixel corners after the last-hit pixel edge (aa.h:359-379)
            for (int pvi = 0; pvi < 4; pvi++) {
                const int cur = (final_pedge + 1 + pvi) & 3;
                if (!((inmask >> cur) & 1u)) break;
                if (!fan.push((cur == 1 || cur == 2) ? pxmax : pxmin, (cur >= 2) ? pymax : pymin)) return false;
            }
        }
    }
    if (fan.area > 1.0f) return false;                               // error 6
    if (fan.area == 0.0f) return false;                              // caller skips oarea == 0
    area_out = fan.area;
    if (GRAD) {
#pragma unroll
        for (int k = 0; k < 6; k++) g[k] = gr[k];
    }
    return true;
}

}  // namespace dm2
