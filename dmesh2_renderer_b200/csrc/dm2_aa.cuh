// dm2_aa.cuh — exact pixel-square / triangle overlap area and its gradient w.r.t. the three
// screen-space triangle vertices (behaviour of reference cuda_impl/aa.h:96-504).
//
// Design differences from the reference (same results, different machine code):
//   * all per-face operands come from the packed record in shared memory (one broadcast LDS per
//     operand) instead of ~70 scalar global loads through 64-bit accessor arithmetic;
//   * the clipped polygon is never materialised: vertices are streamed into a fan accumulator
//     (first / previous / current vertex), so there are no local-memory arrays
//     (the reference keeps polygon[10][2] + two [10][4] Jacobian arrays on the stack);
//   * the gradient is a template switch, so the forward kernel pays for the area only.
// Error semantics: the reference returns codes 1..6 and the caller skips the face for that pixel
// on any non-zero code (forward.cu:337-338, backward.cu:286-289); here `false` = "skip".
#pragma once
#include "dm2_common.cuh"

namespace dm2 {

template <bool GRAD>
struct PolyVert;
template <>
struct PolyVert<false> {
    float x, y;
};
template <>
struct PolyVert<true> {
    float x, y;
    // index of the triangle edge the vertex lies on (-1: pixel corner) and the two 2x2 Jacobians
    // d(vertex)/d(edge start), d(vertex)/d(edge end), laid out as in aa.h:286-294
    // (J[2*r + c] = d vertex[c] / d p[r]).
    int e;
    float j0[4], j1[4];
};

template <bool GRAD>
struct FanAcc {
    int np = 0;
    float area = 0.0f;
    PolyVert<GRAD> first, prev;
    float g[6] = {0, 0, 0, 0, 0, 0};     // GRAD: d area / d tri vert [3][2]

    __device__ __forceinline__ void scatter(const PolyVert<true>& pv, float gx, float gy) {
        if (pv.e < 0) return;
        // _update_grad_tri_verts, aa.h:67-86
        float a0 = __fmaf_rn(pv.j0[0], gx, __fmul_rn(pv.j0[1], gy));
        float a1 = __fmaf_rn(pv.j0[2], gx, __fmul_rn(pv.j0[3], gy));
        float b0 = __fmaf_rn(pv.j1[0], gx, __fmul_rn(pv.j1[1], gy));
        float b1 = __fmaf_rn(pv.j1[2], gx, __fmul_rn(pv.j1[3], gy));
        int i0 = pv.e, i1 = (pv.e == 2) ? 0 : pv.e + 1;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (i0 == k) { g[2 * k] += a0; g[2 * k + 1] += a1; }
            if (i1 == k) { g[2 * k] += b0; g[2 * k + 1] += b1; }
        }
    }

    // returns false on polygon overflow (aa.h:45-48) or negative fan-triangle area (aa.h:410-412)
    __device__ __forceinline__ bool push(const PolyVert<GRAD>& v) {
        if (np >= 10) return false;
        if (np == 0) first = v;
        else if (np == 1) prev = v;
        else {
            // _tri_area(first, prev, v), aa.h:88-94
            float ax = __fsub_rn(prev.x, first.x), by = __fsub_rn(v.y, first.y);
            float cx = __fsub_rn(v.x, first.x),    dy = __fsub_rn(prev.y, first.y);
            float s = __fmul_rn(0.5f, __fmaf_rn(ax, by, -__fmul_rn(cx, dy)));
            if (s < 0.0f) return false;
            area = __fadd_rn(area, s);
            if constexpr (GRAD) {
                // aa.h:416-433
                scatter(first, 0.5f * (prev.y - v.y),     0.5f * (-prev.x + v.x));
                scatter(prev,  0.5f * (v.y - first.y),    0.5f * (-v.x + first.x));
                scatter(v,     0.5f * (first.y - prev.y), 0.5f * (-first.x + prev.x));
            }
            prev = v;
        }
        np++;
        return true;
    }
};

// rec: packed face record in shared memory.  (pxmin, pymin): lower-left pixel corner in image
// coordinates (patch offset already added); the pixel is the unit square, pix_area = 1.
// Returns false if the face must be skipped for this pixel (error, or zero overlap).
// On true: area > 0 (or NaN, which the reference also lets through), g filled when GRAD.
template <bool GRAD>
__device__ __forceinline__ bool aa_overlap(const float* __restrict__ rec, float pxmin, float pymin,
                                           float& area_out, float* __restrict__ g_out) {
    const float pxmax = __fadd_rn(pxmin, 1.0f);
    const float pymax = __fadd_rn(pymin, 1.0f);

    // 1. triangle bounding box (aa.h:96-101, 476-478)
    {
        const float4 bb = *reinterpret_cast<const float4*>(rec + R_BBOX);
        if ((pxmax < bb.x) || (pxmin > bb.y) || (pymax < bb.z) || (pymin > bb.w)) return false;
    }

    // 2. pixel corners against the three edge half-planes (aa.h:103-149)
    unsigned inmask = 0xF;            // bit k: pixel corner k inside all edges so far
    {
        const float4 n01 = *reinterpret_cast<const float4*>(rec + R_N01);
        const float4 n2c = *reinterpret_cast<const float4*>(rec + R_N2C);
        const float c2 = rec[R_MISC];
        const float nx[3] = {n01.x, n01.z, n2c.x};
        const float ny[3] = {n01.y, n01.w, n2c.y};
        const float cc[3] = {n2c.z, n2c.w, c2};
#pragma unroll
        for (int ti = 0; ti < 3; ti++) {
            // (vert[0]*n[0]) + (vert[1]*n[1]) - c >= 0, aa.h:20
            const float ylo = __fmul_rn(pymin, ny[ti]);
            const float yhi = __fmul_rn(pymax, ny[ti]);
            const bool i0 = __fsub_rn(__fmaf_rn(pxmin, nx[ti], ylo), cc[ti]) >= 0.0f;
            const bool i1 = __fsub_rn(__fmaf_rn(pxmax, nx[ti], ylo), cc[ti]) >= 0.0f;
            const bool i2 = __fsub_rn(__fmaf_rn(pxmax, nx[ti], yhi), cc[ti]) >= 0.0f;
            const bool i3 = __fsub_rn(__fmaf_rn(pxmin, nx[ti], yhi), cc[ti]) >= 0.0f;
            if (!(i0 || i1 || i2 || i3)) return false;           // all four outside one edge
            inmask &= (i0 ? 1u : 0u) | (i1 ? 2u : 0u) | (i2 ? 4u : 0u) | (i3 ? 8u : 0u);
        }
    }

    // 3. pixel completely inside (aa.h:493-496): area = pix_area, zero gradient
    if (inmask == 0xF) {
        area_out = 1.0f;
        if (GRAD) {
#pragma unroll
            for (int k = 0; k < 6; k++) g_out[k] = 0.0f;
        }
        return true;
    }

    // 4. clip (aa.h:151-441), streaming the polygon into the fan accumulator
    FanAcc<GRAD> fan;
    const unsigned zflags = __float_as_uint(rec[R_MISC + 3]);

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

        PolyVert<GRAD> ia, ib;           // up to two intersections, in pixel-edge order
        float ta = 0.f, tb = 0.f;
        int pa = -1, pb = -1;
        int n = 0;

#pragma unroll
        for (int pi = 0; pi < 4; pi++) {
            const bool horiz = (pi == 0) || (pi == 2);             // pixel edge y = const
            const bool parallel = horiz ? edge_horizontal : edge_vertical;
            const float cst  = (pi == 0) ? pymin : (pi == 1) ? pxmax : (pi == 2) ? pymax : pxmin;
            const float p0a0 = horiz ? p0.y : p0.x;                // axis0 = fixed axis
            const float p0a1 = horiz ? p0.x : p0.y;
            const float p1a0 = horiz ? p1.y : p1.x;
            const float ra0  = horiz ? rc.y : rc.x;
            const float ea1  = horiz ? ed.x : ed.y;
            const float lo1  = horiz ? pxmin : pymin;
            const float hi1  = horiz ? pxmax : pymax;

            const float t  = __fmul_rn(__fsub_rn(cst, p0a0), ra0);     // aa.h:251
            const float i1 = __fmaf_rn(t, ea1, p0a1);                  // aa.h:252
            const bool valid = (t >= 0.0f) && (t <= 1.0f) && (i1 >= lo1) && (i1 <= hi1) && !parallel;
            if (!valid) continue;
            if ((i1 == lo1) || (i1 == hi1)) return false;             // error 1: hits a pixel corner
            if (n >= 2) return false;                                  // error 2: > 2 intersections
            PolyVert<GRAD> q;
            q.x = horiz ? i1 : cst;
            q.y = horiz ? cst : i1;
            if constexpr (GRAD) {
                q.e = ti;
                // aa.h:280-294; grad_t is non-zero only along axis0
                const float gt0 = __fmul_rn(__fmul_rn(__fsub_rn(cst, p1a0), ra0), ra0);
                const float gt1 = __fmul_rn(__fmul_rn(__fadd_rn(-cst, p0a0), ra0), ra0);
                const float g0x = horiz ? 0.0f : gt0, g0y = horiz ? gt0 : 0.0f;
                const float g1x = horiz ? 0.0f : gt1, g1y = horiz ? gt1 : 0.0f;
                const double omt = 1.0 - (double)t;                   // (1.0 - t): double literal
                q.j0[0] = (float)(omt + (double)__fmul_rn(g0x, ed.x));
                q.j0[1] = __fmul_rn(g0x, ed.y);
                q.j0[2] = __fmul_rn(g0y, ed.x);
                q.j0[3] = (float)(omt + (double)__fmul_rn(g0y, ed.y));
                q.j1[0] = __fmaf_rn(g1x, ed.x, t);
                q.j1[1] = __fmul_rn(g1x, ed.y);
                q.j1[2] = __fmul_rn(g1y, ed.x);
                q.j1[3] = __fmaf_rn(g1y, ed.y, t);
            }
            if (n == 0) { ia = q; ta = t; pa = pi; } else { ib = q; tb = t; pb = pi; }
            n++;
        }

        int final_pedge = -1;
        if (n == 2) {
            // sort by t (aa.h:308-328); strict '>' keeps pixel-edge order on ties
            if (ta > tb) {
                if (!fan.push(ib)) return false;
                if (!fan.push(ia)) return false;
                final_pedge = pa;
            } else {
                if (!fan.push(ia)) return false;
                if (!fan.push(ib)) return false;
                final_pedge = pb;
            }
        } else if (n == 1) {
            if (!fan.push(ia)) return false;
            if (!in0 && in1) {
                PolyVert<GRAD> q;
                q.x = p1.x; q.y = p1.y;
                if constexpr (GRAD) {
                    q.e = ti;
                    q.j0[0] = q.j0[1] = q.j0[2] = q.j0[3] = 0.0f;
                    q.j1[0] = 1.0f; q.j1[1] = 0.0f; q.j1[2] = 0.0f; q.j1[3] = 1.0f;
                }
                if (!fan.push(q)) return false;
            } else if (in0 && !in1) {
                final_pedge = pa;
            } else {
                return false;                                          // error 3
            }
        } else {
            if (in0 && in1) {
                PolyVert<GRAD> q;
                q.x = p1.x; q.y = p1.y;
                if constexpr (GRAD) {
                    q.e = ti;
                    q.j0[0] = q.j0[1] = q.j0[2] = q.j0[3] = 0.0f;
                    q.j1[0] = 1.0f; q.j1[1] = 0.0f; q.j1[2] = 0.0f; q.j1[3] = 1.0f;
                }
                if (!fan.push(q)) return false;
            } else if (in0 != in1) {
                return false;                                          // error 4
            }
        }

        if (final_pedge != -1) {
            // run of inside pixel corners after the last-hit pixel edge (aa.h:359-379)
            for (int pvi = 0; pvi < 4; pvi++) {
                const int cur = (final_pedge + 1 + pvi) & 3;
                if (!((inmask >> cur) & 1u)) break;
                PolyVert<GRAD> q;
                q.x = (cur == 1 || cur == 2) ? pxmax : pxmin;
                q.y = (cur >= 2) ? pymax : pymin;
                if constexpr (GRAD) {
                    q.e = -1;
                    q.j0[0] = q.j0[1] = q.j0[2] = q.j0[3] = 0.0f;
                    q.j1[0] = q.j1[1] = q.j1[2] = q.j1[3] = 0.0f;
                }
                if (!fan.push(q)) return false;
            }
        }
    }

    if (fan.area > 1.0f) return false;                                 // error 6
    if (fan.area == 0.0f) return false;                                // caller skips oarea == 0
    area_out = fan.area;
    if (GRAD) {
#pragma unroll
        for (int k = 0; k < 6; k++) g_out[k] = fan.g[k];
    }
    return true;
}

}  // namespace dm2
