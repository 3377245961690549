// dm2_aa.cuh — exact pixel-square / triangle overlap area and its gradient w.r.t. the three
// screen-space triangle vertices (behaviour of reference cuda_impl/aa.h:96-504).
//
// Design differences from the reference (same results, different machine code):
//   * all per-face operands come from the packed record in shared memory instead of ~70 scalar
//     global loads through 64-bit accessor arithmetic;
//   * two stages: the classification (bbox + 12 half-plane tests: outside / fully inside / partial,
//     aa.h:96-149) runs once per sorted instance and 8x4-pixel footprint, bit-parallel, in
//     k_classify_instances (dm2_binning.cu); aa_clip() (the polygon clip) only runs for partially
//     covered pixels and receives the 4-bit inside-corner mask from that pre-pass;
//   * the clipped polygon is never materialised: vertices are streamed into a fan accumulator
//     (first / previous / current vertex), so there are no local-memory arrays (the reference keeps
//     polygon[10][2] + two [10][4] Jacobian arrays on the stack);
//   * the gradient is normally NOT chained through per-vertex 2x2 Jacobians (aa.h:280-294, 416-433).  For a
//     triangle edge whose part inside the pixel is the parameter interval [ta, tb], moving the edge
//     end points sweeps area  dA/dp_start = n (tb-ta - (tb^2-ta^2)/2),  dA/dp_end = n (tb^2-ta^2)/2
//     with n = (e.y, -e.x) the outward edge normal scaled by the edge length.  This is the same
//     derivative (it agrees with the reference's analytic gradient to 1e-6 on the reference's own
//     Python implementation, tests/golden/aa_overlap_golden.npz) at ~15 flops per edge - WHENEVER the
//     polygon the reference builds is the true clip polygon.  Two situations make the reference's chained
//     gradient differ from the true derivative, and there the clip asks for aa_grad_reference(), an
//     out-of-line transcription of the reference's polygon + Jacobian chain:
//       (a) a triangle vertex exactly ON a pixel boundary line (a kink of the area: the reference
//           differentiates "the intersection slides along the pixel edge", a one-sided derivative) - flagged
//           per face by the preprocess kernels (bit 9 of the record's flag word: some screen coordinate of
//           the face is an integer), conservatively;
//       (b) an INCONSISTENT polygon: the inside-corner flags (half-plane tests, aa.h:103-149) and the
//           edge / pixel-edge intersections (aa.h:230-262) are rounded independently, so for an edge that
//           passes within an ulp of a pixel corner the run of inside corners behind an exit point can
//           disagree with the pixel edge on which the next edge enters - the reference then appends
//           spurious corners (or misses one), its area is off by ~1e-5 and its gradient by O(1)
//           (measured: 1 pair in ~16 M at 1 M triangles @1024^2; found by the full-size parity test
//           against the fp64-accumulated oracle).  Detected by counting corner pushes: in a true clip
//           polygon every inside corner is a vertex exactly once.
// Error semantics: the reference returns codes 1..6 and the caller skips the face for that pixel
// on any non-zero code (forward.cu:337-338, backward.cu:286-289); here `false` = "skip".
#pragma once
#include "dm2_common.cuh"
#include "dm2_aa_fast.cuh"

namespace dm2 {

// Fan accumulator with PREDICATED pushes: every lane executes the same instruction stream whatever
// its polygon looks like (the branchy formulation ran at 2-3 active lanes per instruction).
// Instruction-count reductions of the clip that were measured SLOWER on B200 in same-box A/Bs (ptxas' register
// allocation for the forward kernel is the binding constraint, not the instruction count) stay switchable:
//   DM2_CLIP_RELFAN  previous fan vertex kept relative to the first one (2 FADD less per push)
//   DM2_CLIP_CARRY   an edge's end-point inside test carried over to the next edge (4 compares less per edge)
//   DM2_CLIP_NCHECK  "more than two intersections" tested once per edge instead of once per pixel edge
#ifndef DM2_CLIP_RELFAN
#define DM2_CLIP_RELFAN 0
#endif
#ifndef DM2_CLIP_CARRY
#define DM2_CLIP_CARRY 0
#endif
#ifndef DM2_CLIP_NCHECK
#define DM2_CLIP_NCHECK 0
#endif
// DM2_CLIP_PFADD: conditional float moves written as predicated `x + 0.0f` (FMA pipe) instead of selects (ALU pipe,
// the pipe that limits the forward kernel: 72 % busy against 26 % for the FMA pipe)
#ifndef DM2_CLIP_PFADD
#define DM2_CLIP_PFADD 0
#endif
#if DM2_CLIP_PFADD
#define DM2_CMOV(p, dst, src) do { if (p) dst = __fadd_rn(src, 0.0f); } while (0)
#else
#define DM2_CMOV(p, dst, src) do { dst = (p) ? (src) : dst; } while (0)
#endif
struct FanAcc {
    int np = 0;
    float area = 0.0f;
    // first polygon vertex, and the previous one RELATIVE to it (qx - fx, qy - fy): exactly the two differences
    // the next fan triangle needs - they were computed as (x - fx, y - fy) when that vertex was pushed
    float fx = 0.f, fy = 0.f, qx = 0.f, qy = 0.f;
    bool bad = false;     // negative fan-triangle area (aa.h:410-412); polygon overflow is checked at the end
    __device__ __forceinline__ void push(bool cond, float x, float y) {
        const bool is_first = cond & (np == 0);
        const bool is_tri = cond & (np >= 2);
        // _tri_area(first, prev, v), aa.h:88-94
#if DM2_CLIP_RELFAN
        const float ax = qx, dy = qy;
        const float cx = __fsub_rn(x, fx), by = __fsub_rn(y, fy);
#else
        const float ax = __fsub_rn(qx, fx), by = __fsub_rn(y, fy);
        const float cx = __fsub_rn(x, fx),  dy = __fsub_rn(qy, fy);
#endif
        const float s = __fmul_rn(0.5f, __fmaf_rn(ax, by, -__fmul_rn(cx, dy)));
        bad |= is_tri & (s < 0.0f);
        area = is_tri ? __fadd_rn(area, s) : area;
        DM2_CMOV(is_first, fx, x);
        DM2_CMOV(is_first, fy, y);
        const bool set_prev = cond & (np >= 1);
#if DM2_CLIP_RELFAN
        DM2_CMOV(set_prev, qx, cx);
        DM2_CMOV(set_prev, qy, by);
#else
        DM2_CMOV(set_prev, qx, x);
        DM2_CMOV(set_prev, qy, y);
#endif
        np += cond ? 1 : 0;
    }
    // aa.h:45-48: adding a vertex to a polygon that already holds 10 is an error; np only grows on
    // accepted pushes, so "some push saw np >= 10" is the same as "np ended above 10"
    __device__ __forceinline__ bool overflow() const { return np > 10; }
};

// Step 4 of aa.h:446-504 (= aa.h:151-441) for a partially covered pixel.  Returns false if the face
// must be skipped for this pixel (any reference error code, or zero overlap).  On true: area in
// (0, 1] (or NaN, which the reference also lets through); g[6] = d area / d tri verts when GRAD.
// Straight-line predicated code: errors are accumulated in a flag instead of returning early (every
// reference error code has the same consequence - the face is skipped for this pixel).
// GRAD: *chain_needed (may be nullptr) is set when g must be replaced by aa_grad_reference() - see above.
template <bool GRAD>
__device__ __forceinline__ bool aa_clip(const float* __restrict__ rec, float pxmin, float pymin, unsigned inmask,
                                        float& area_out, float* __restrict__ g, bool* chain_needed = nullptr) {
    const float pxmax = __fadd_rn(pxmin, 1.0f);
    const float pymax = __fadd_rn(pymin, 1.0f);
    FanAcc fan;
    const unsigned zflags = __float_as_uint(rec[R_MISC + 3]);
    float gr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int corner_pushes = 0;        // GRAD only: see the consistency test at the end

#ifndef DM2_CLIP_UNROLL
#define DM2_CLIP_UNROLL 1
#endif
    constexpr int kClipUnroll = DM2_CLIP_UNROLL;
#if DM2_CLIP_CARRY
    // (an edge's end point is the next edge's start point: its inside test is carried over)
    bool in_next;
    {
        const float2 v0 = *reinterpret_cast<const float2*>(rec + R_AAV);
        in_next = (v0.x >= pxmin) & (v0.x <= pxmax) & (v0.y >= pymin) & (v0.y <= pymax);
    }
#endif
#pragma unroll kClipUnroll
    for (int ti = 0; ti < 3; ti++) {
        const int tj = (ti == 2) ? 0 : ti + 1;
        const float2 p0 = *reinterpret_cast<const float2*>(rec + R_AAV + 2 * ti);
        const float2 p1 = *reinterpret_cast<const float2*>(rec + R_AAV + 2 * tj);
        const float2 ed = *reinterpret_cast<const float2*>(rec + R_AAE + 2 * ti);
        const float2 rc = *reinterpret_cast<const float2*>(rec + R_AAR + 2 * ti);
        const bool edge_vertical   = (zflags >> (2 * ti)) & 1u;      // edges_iszero[ti][0]
        const bool edge_horizontal = (zflags >> (2 * ti + 1)) & 1u;  // edges_iszero[ti][1]
#if DM2_CLIP_CARRY
        const bool in0 = in_next;
        const bool in1 = (p1.x >= pxmin) & (p1.x <= pxmax) & (p1.y >= pymin) & (p1.y <= pymax);
        in_next = in1;
#else
        const bool in0 = (p0.x >= pxmin) & (p0.x <= pxmax) & (p0.y >= pymin) & (p0.y <= pymax);
        const bool in1 = (p1.x >= pxmin) & (p1.x <= pxmax) & (p1.y >= pymin) & (p1.y <= pymax);
#endif

        float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;     // up to two intersections, pixel-edge order
        float ta = 0.f, tb = 0.f;
        int pa = 0, pb = 0, n = 0;
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
            const bool valid = (t >= 0.0f) & (t <= 1.0f) & (i1 >= lo1) & (i1 <= hi1) & !parallel;
            // error 1: hits a pixel corner; error 2: more than two intersections
#if DM2_CLIP_NCHECK
            fan.bad |= valid & ((i1 == lo1) | (i1 == hi1));           // (error 2: n > 2 after the loop)
#else
            fan.bad |= valid & ((i1 == lo1) | (i1 == hi1) | (n >= 2));
#endif
            const float qx = horiz ? i1 : cst, qy = horiz ? cst : i1;
            const bool is_a = valid & (n == 0), is_b = valid & (n == 1);
            DM2_CMOV(is_a, ax, qx); DM2_CMOV(is_a, ay, qy); DM2_CMOV(is_a, ta, t); pa = is_a ? pi : pa;
            DM2_CMOV(is_b, bx, qx); DM2_CMOV(is_b, by, qy); DM2_CMOV(is_b, tb, t); pb = is_b ? pi : pb;
            n += valid ? 1 : 0;
        }
#if DM2_CLIP_NCHECK
        fan.bad |= n > 2;
#endif
        // aa.h:299-400 as predicates
        const bool two = (n >= 2), one = (n == 1), none = (n == 0);
        const bool swap = two & (ta > tb);                           // sort by t; strict '>' keeps order on ties
        const bool enter = one & !in0 & in1;                         // intersection, then the end point
        const bool leave = one & in0 & !in1;                         // start point inside, intersection, corners
        const bool whole = none & in0 & in1;                         // edge entirely inside the pixel
        fan.bad |= (one & (in0 == in1)) | (none & (in0 != in1));     // errors 3 and 4
        fan.push(!none, swap ? bx : ax, swap ? by : ay);
        fan.push(two | enter | whole, two ? (swap ? ax : bx) : p1.x, two ? (swap ? ay : by) : p1.y);
        // run of inside pixel corners after the last-hit pixel edge (aa.h:359-379); a partially
        // covered pixel has at most 3 inside corners
        bool run = two | leave;
        const int final_pedge = two ? (swap ? pa : pb) : pa;
#pragma unroll
        for (int pvi = 0; pvi < 3; pvi++) {
            const int cur = (final_pedge + 1 + pvi) & 3;
            run = run & (((inmask >> cur) & 1u) != 0u);
            fan.push(run, (cur == 1 || cur == 2) ? pxmax : pxmin, (cur >= 2) ? pymax : pymin);
            if (GRAD) corner_pushes += run ? 1 : 0;
        }
        if (GRAD) {

            // part of the edge inside the pixel: [t_in, t_out]
            const bool covered = !none || whole;
            const float t_in  = two ? (swap ? tb : ta) : (enter ? ta : 0.0f);
            const float t_out = two ? (swap ? ta : tb) : (leave ? ta : 1.0f);
            const float w1 = covered ? 0.5f * (t_out * t_out - t_in * t_in) : 0.0f;
            const float w0 = covered ? (t_out - t_in) - w1 : 0.0f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float wk = (k == ti) ? w0 : ((k == tj) ? w1 : 0.0f);
                gr[2 * k] += ed.y * wk;
                gr[2 * k + 1] -= ed.x * wk;
            }
        }
    }
    const bool ok = !fan.bad & !fan.overflow() & !(fan.area > 1.0f) & !(fan.area == 0.0f);   // error 6; caller skips oarea == 0
    area_out = fan.area;
    if (GRAD) {
#pragma unroll
        for (int k = 0; k < 6; k++) g[k] = gr[k];
        // In the true clip polygon every inside pixel corner is a vertex exactly once.  The reference's corner
        // runs (aa.h:359-379) can push a corner twice / not at all when flags and intersections disagree, and a
        // face with a vertex on an integer pixel line (flag bit 9, set by the preprocess kernels) has kinks.
#ifdef DM2_EXP_NO_CHAINDET                /* timing experiment only */
        if (chain_needed) *chain_needed = false;
        else
#endif
        if (chain_needed)
            *chain_needed = ok & ((corner_pushes != (int)__popc(inmask & 0xFu)) | (((zflags >> 9) & 1u) != 0u));
    }
    return ok;
}

// Out-of-line transcription of the reference's polygon construction and Jacobian-chain gradient
// (aa.h:151-441) for the rare pairs aa_clip flags (kinks, inconsistent polygons): d(area)/d(tri verts) exactly
// as the reference chains it - polygon array, per-vertex 2x2 Jacobians d ip/d p0 and d ip/d p1 (aa.h:280-294;
// the `1.0 - t` terms are evaluated in double there), fan triangles from polygon[0] (aa.h:403-433).  Only
// called for pairs whose clip succeeded (no error code), so the error exits are not repeated.
static __device__ __noinline__ void aa_grad_reference(const float* __restrict__ rec, float pxmin, float pymin, unsigned inmask,
                                               float* __restrict__ g) {
    const float pxmax = __fadd_rn(pxmin, 1.0f), pymax = __fadd_rn(pymin, 1.0f);
    const unsigned zflags = __float_as_uint(rec[R_MISC + 3]);
    float px[10], py[10], j0[10][4], j1[10][4];
    int pe[10], n = 0;
    auto add = [&](float x, float y, int e, const float* a, const float* b) {
        if (n >= 10) return;
        px[n] = x; py[n] = y; pe[n] = e;
#pragma unroll
        for (int k = 0; k < 4; k++) { j0[n][k] = a ? a[k] : 0.0f; j1[n][k] = b ? b[k] : 0.0f; }
        n++;
    };
    const float eye[4] = {1.f, 0.f, 0.f, 1.f};
    for (int ti = 0; ti < 3; ti++) {
        const int tj = (ti == 2) ? 0 : ti + 1;
        const float p0[2] = {rec[R_AAV + 2 * ti], rec[R_AAV + 2 * ti + 1]};
        const float p1[2] = {rec[R_AAV + 2 * tj], rec[R_AAV + 2 * tj + 1]};
        const float e[2] = {rec[R_AAE + 2 * ti], rec[R_AAE + 2 * ti + 1]};
        const float r[2] = {rec[R_AAR + 2 * ti], rec[R_AAR + 2 * ti + 1]};
        const bool vertical = (zflags >> (2 * ti)) & 1u, horizontal = (zflags >> (2 * ti + 1)) & 1u;
        const bool in0 = p0[0] >= pxmin && p0[0] <= pxmax && p0[1] >= pymin && p0[1] <= pymax;
        const bool in1 = p1[0] >= pxmin && p1[0] <= pxmax && p1[1] >= pymin && p1[1] <= pymax;
        float ipx[2], ipy[2], it[2], g0[2][4], g1[2][4];
        int ipe[2], ni = 0;
        for (int pi = 0; pi < 4 && ni < 2; pi++) {
            const bool ph = (pi == 0) || (pi == 2);
            if ((horizontal && ph) || (vertical && !ph)) continue;
            const int a0 = ph ? 1 : 0, a1 = 1 - a0;
            const float lo = ph ? pxmin : pymin, hi = ph ? pxmax : pymax;
            const float c = (pi == 0) ? pymin : (pi == 1) ? pxmax : (pi == 2) ? pymax : pxmin;
            const float tt = __fmul_rn(__fsub_rn(c, p0[a0]), r[a0]);
            const float i1 = __fmaf_rn(tt, e[a1], p0[a1]);
            if (!((tt >= 0.0f) && (tt <= 1.0f) && (i1 >= lo) && (i1 <= hi))) continue;
            ipx[ni] = ph ? i1 : c; ipy[ni] = ph ? c : i1; it[ni] = tt; ipe[ni] = pi;
            float gt0[2] = {0.f, 0.f}, gt1[2] = {0.f, 0.f};
            gt0[a0] = __fmul_rn(__fmul_rn(__fsub_rn(c, p1[a0]), r[a0]), r[a0]);
            gt1[a0] = __fmul_rn(__fmul_rn(__fadd_rn(-c, p0[a0]), r[a0]), r[a0]);
            g0[ni][0] = (float)((1.0 - (double)tt) + (double)__fmul_rn(gt0[0], e[0]));
            g0[ni][1] = __fmul_rn(gt0[0], e[1]);
            g0[ni][2] = __fmul_rn(gt0[1], e[0]);
            g0[ni][3] = (float)((1.0 - (double)tt) + (double)__fmul_rn(gt0[1], e[1]));
            g1[ni][0] = __fmaf_rn(gt1[0], e[0], tt);
            g1[ni][1] = __fmul_rn(gt1[0], e[1]);
            g1[ni][2] = __fmul_rn(gt1[1], e[0]);
            g1[ni][3] = __fmaf_rn(gt1[1], e[1], tt);
            ni++;
        }
        int final_pedge = -1;
        if (ni == 2) {
            const int s0 = (it[0] > it[1]) ? 1 : 0, s1 = 1 - s0;
            add(ipx[s0], ipy[s0], ti, g0[s0], g1[s0]);
            add(ipx[s1], ipy[s1], ti, g0[s1], g1[s1]);
            final_pedge = ipe[s1];
        } else if (ni == 1) {
            add(ipx[0], ipy[0], ti, g0[0], g1[0]);
            if (!in0 && in1) add(p1[0], p1[1], ti, nullptr, eye);
            else if (in0 && !in1) final_pedge = ipe[0];
        } else if (in0 && in1) {
            add(p1[0], p1[1], ti, nullptr, eye);
        }
        if (final_pedge >= 0) {
            for (int pvi = 0; pvi < 4; pvi++) {
                const int cur = (final_pedge + 1 + pvi) & 3;
                if (!((inmask >> cur) & 1u)) break;
                add((cur == 1 || cur == 2) ? pxmax : pxmin, (cur >= 2) ? pymax : pymin, -1, nullptr, nullptr);
            }
        }
    }
    float gr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto upd = [&](int v, float gax, float gay) {            // aa.h:67-86
        const int e = pe[v];
        if (e < 0) return;
        const int i0 = e, i1 = (e == 2) ? 0 : e + 1;
        gr[2 * i0]     += __fmaf_rn(j0[v][0], gax, __fmul_rn(j0[v][1], gay));
        gr[2 * i0 + 1] += __fmaf_rn(j0[v][2], gax, __fmul_rn(j0[v][3], gay));
        gr[2 * i1]     += __fmaf_rn(j1[v][0], gax, __fmul_rn(j1[v][1], gay));
        gr[2 * i1 + 1] += __fmaf_rn(j1[v][2], gax, __fmul_rn(j1[v][3], gay));
    };
    for (int si = 0; si + 2 < n; si++) {
        const int a = 0, b = si + 1, c = si + 2;
        upd(a, 0.5f * (py[b] - py[c]), 0.5f * (-px[b] + px[c]));
        upd(b, 0.5f * (py[c] - py[a]), 0.5f * (-px[c] + px[a]));
        upd(c, 0.5f * (py[a] - py[b]), 0.5f * (-px[a] + px[b]));
    }
#pragma unroll
    for (int k = 0; k < 6; k++) g[k] = gr[k];
}

// The exact construction for one pair, OUT OF LINE: aa_clip (+ the reference's Jacobian chain where it asks for it).
// The tile kernels call it for the pairs aa_area_fast() declines (~0.5 % at 1 M triangles @1024^2) and for every pair
// when the exact clip was requested (dm2_set_exact_clip).  gs: the caller's scratch array - its address escapes into
// this call, so it lives in local memory; callers copy it into registers instead of handing over their own g[].
template <bool GRAD>
static __device__ __noinline__ bool aa_clip_exact(const float* __restrict__ rec, float pxmin, float pymin, unsigned inmask,
                                                  float* __restrict__ area_out, float* __restrict__ gs) {
    float a, g[6];
    bool chain = false;
    const bool ok = aa_clip<GRAD>(rec, pxmin, pymin, inmask, a, g, GRAD ? &chain : nullptr);
    if (GRAD) {
        if (chain) aa_grad_reference(rec, pxmin, pymin, inmask, g);
#pragma unroll
        for (int k = 0; k < 6; k++) gs[k] = g[k];
    }
    *area_out = a;
    return ok;
}

// Overlap area (+ gradient) of one partially covered (pixel, face) pair: the polygon-free evaluation where it is
// the reference's result to ~1e-7 (dm2_aa_fast.cuh: pairs far from each of its rounding-dependent decisions), the exact construction otherwise.  inmask_of(): the 4-bit
// inside-corner mask, only evaluated on the exact path.  Returns false = the reference skips the face for this pixel.
template <bool GRAD, typename InMask>
__device__ __forceinline__ bool aa_overlap(const float* __restrict__ rec, float pxmin, float pymin, bool exact_only,
                                           InMask inmask_of, float& area, float* __restrict__ g) {
    if (!exact_only &&
        aa_area_fast<GRAD>(rec + R_AAV, rec + R_AAE, rec + R_AAR, __float_as_uint(rec[R_MISC + 3]), __float_as_uint(rec[R_SAFE]),
                           pxmin, pymin, area, g))
        return true;
    float as, gs[6];
    const bool ok = aa_clip_exact<GRAD>(rec, pxmin, pymin, inmask_of(), &as, gs);
    area = as;
    if (GRAD) {
#pragma unroll
        for (int k = 0; k < 6; k++) g[k] = gs[k];
    }
    return ok;
}

}  // namespace dm2
