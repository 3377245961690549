// dm2_layers.cu — LayeredRenderer hot path: per-pixel first hit over the min-depth-sorted tile
// list, choice of the entry tetrahedron, then the per-ray walk through the tet mesh recording the
// first L existing faces.  Behaviour of reference cuda_impl/forward.cu:538-709
// (firstIntersectCUDA), :744-1000 (generateRenderLayersCUDA) and auxiliary.h:382-431
// (tet_face_outward_normal); re-designed:
//   * unit face normals, the two adjacent tets and the existence flag of every face go into ONE
//     64-byte record per face, tet centroid sums + the four face ids into a 32-byte record per tet,
//     once per call: a walk step is two dependent loads (the tet, then its three other faces in
//     parallel) instead of ~30 scattered vertex / index loads in four dependent rounds;
//   * first hit and walk run in ONE kernel (first_face/first_tet stay in registers; they are still
//     written out as debug taps);
//   * out-of-image threads are bounds-checked (the reference writes first_face[-1..] for them when
//     W or H is not a multiple of 16, forward.cu:565-585).
// All sign/hit decisions use the fixed FMA sequences of dm2_common.cuh, so layer ids are
// reproducible bit-for-bit against the reference and the CPU oracle.
#include "dm2_common.cuh"

namespace dm2 {

constexpr int kLStage = 256;
constexpr int kFGeoWords = 16;     // p0.xyz E1.xyz E2.xyz n.xyz (unit normal, not yet oriented) face_tets[2] existence pad
// per tet, 256 bytes (float4 units):
//   [0] centroid sum.xyz, flip bits (slot i: the stored face normal points INTO the tet)   } entry-tet choice and the
//   [1] tet_faces[4]                                                                        } reference-shaped step
//   [2] id of the vertex opposite the face in slot i (-1: slot i is not a proper face of this tet)
//   [4 + i]  position of that opposite vertex, L1 length of the longest tet edge (+inf: no fast steps in this tet)
//   [8 + i]  stored (not yet oriented) unit normal of the face in slot i
//   [12 + i] int4: face id of slot i, the tet behind it (the reference's rule, forward.cu:985-992), link word,
//            existence flag of the face (refreshed by k_tet_exist on every call)
//            link word: bits 0-7 slot j of this tet -> slot of the same vertex in the tet behind (2 bits each),
//            bits 8-9 slot of the shared face in the tet behind, bit 10 flip bit of that slot, bit 12 flip bit of slot i
// A fast walk step reads [4 + entry slot], then [8 + exit slot] and [12 + exit slot] - all it needs besides the
// existence flag of the face it leaves through.
constexpr int kTGeoWords = 64;

size_t layers_geo_bytes(int F, int T) {
    return align_up((size_t)F * kFGeoWords * 4, 256) + align_up((size_t)T * kTGeoWords * 4, 256) + 256;
}

__global__ void __launch_bounds__(256)
k_face_geo(int F, const float* __restrict__ verts, const int32_t* __restrict__ faces,
           const int32_t* __restrict__ face_tets, const int32_t* __restrict__ face_existence, float* __restrict__ fgeo) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int v0 = faces[3 * (size_t)f], v1 = faces[3 * (size_t)f + 1], v2 = faces[3 * (size_t)f + 2];
    const f3 p0{verts[3 * (size_t)v0], verts[3 * (size_t)v0 + 1], verts[3 * (size_t)v0 + 2]};
    const f3 p1{verts[3 * (size_t)v1], verts[3 * (size_t)v1 + 1], verts[3 * (size_t)v1 + 2]};
    const f3 p2{verts[3 * (size_t)v2], verts[3 * (size_t)v2 + 1], verts[3 * (size_t)v2 + 2]};
    // auxiliary.h:403-408
    f3 n = cross3(sub3(p1, p0), sub3(p2, p0));
    float nn = __fsqrt_rn(dot3(n, n));
    nn = fmaxf(nn, 0.0001f);
    n = f3{__fdiv_rn(n.x, nn), __fdiv_rn(n.y, nn), __fdiv_rn(n.z, nn)};
    float4* o = reinterpret_cast<float4*>(fgeo + (size_t)f * kFGeoWords);
    // the two edge vectors of Moeller-Trumbore (auxiliary.h:218-219) are ray independent: stored, not p1 / p2
    const f3 e1 = sub3(p1, p0), e2 = sub3(p2, p0);
    o[0] = make_float4(p0.x, p0.y, p0.z, e1.x);
    o[1] = make_float4(e1.y, e1.z, e2.x, e2.y);
    o[2] = make_float4(e2.z, n.x, n.y, n.z);
    o[3] = make_float4(__int_as_float(face_tets[2 * (size_t)f]), __int_as_float(face_tets[2 * (size_t)f + 1]),
                       __int_as_float(face_existence[f]), 0.0f);
}

// (runs after k_face_geo: the orientation test of auxiliary.h:421-427 - is the face normal on the
// centroid's side? - does not depend on the ray either and is evaluated here once per (tet, face slot))
__global__ void __launch_bounds__(256)
k_tet_geo(int T, const float* __restrict__ verts, const int32_t* __restrict__ tets, const int32_t* __restrict__ faces,
          const int32_t* __restrict__ tet_faces, const int32_t* __restrict__ face_tets, const float* __restrict__ fgeo, int F,
          int P, float* __restrict__ tgeo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int4 q = *reinterpret_cast<const int4*>(tets + 4 * (size_t)t);
    const float* a = verts + 3 * (size_t)q.x;
    const float* b = verts + 3 * (size_t)q.y;
    const float* c = verts + 3 * (size_t)q.z;
    const float* d = verts + 3 * (size_t)q.w;
    // auxiliary.h:421: (q0 + q1 + q2 + q3) * 0.25f, left to right
    const float sx = __fadd_rn(__fadd_rn(__fadd_rn(a[0], b[0]), c[0]), d[0]);
    const float sy = __fadd_rn(__fadd_rn(__fadd_rn(a[1], b[1]), c[1]), d[1]);
    const float sz = __fadd_rn(__fadd_rn(__fadd_rn(a[2], b[2]), c[2]), d[2]);
    const int4 tf = *reinterpret_cast<const int4*>(tet_faces + 4 * (size_t)t);
    const int tfi[4] = {tf.x, tf.y, tf.z, tf.w};
    unsigned flip = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (tfi[i] < 0 || tfi[i] >= F) continue;
        const float4* fp = reinterpret_cast<const float4*>(fgeo + (size_t)tfi[i] * kFGeoWords);
        const float4 fa = fp[0], fc = fp[2];
        // d = centroid - p0 with centroid = sum * 0.25  ->  fma(sum, 0.25, -p0);  side = dot(n, d)
        const f3 dd{__fmaf_rn(sx, 0.25f, -fa.x), __fmaf_rn(sy, 0.25f, -fa.y), __fmaf_rn(sz, 0.25f, -fa.z)};
        if (dot3(f3{fc.y, fc.z, fc.w}, dd) > 0.0f) flip |= 1u << i;
    }
    // walk fast path: the tet vertex opposite each face slot (exactly three of the four tet vertices must be the
    // face's three distinct vertices), and the length of the longest edge (scale of the robustness margin)
    const int qv[4] = {q.x, q.y, q.z, q.w};
    int opp[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        opp[i] = -1;
        if (tfi[i] < 0 || tfi[i] >= F) continue;
        const int f0 = faces[3 * (size_t)tfi[i]], f1 = faces[3 * (size_t)tfi[i] + 1], f2 = faces[3 * (size_t)tfi[i] + 2];
        if (f0 == f1 || f1 == f2 || f0 == f2) continue;
        int n_in = 0, out = -1;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool on = qv[k] == f0 || qv[k] == f1 || qv[k] == f2;
            n_in += on ? 1 : 0;
            if (!on) out = qv[k];
        }
        if (n_in == 3 && out >= 0 && out < P) opp[i] = out;
    }
    if (q.x == q.y || q.x == q.z || q.x == q.w || q.y == q.z || q.y == q.w || q.z == q.w) opp[0] = opp[1] = opp[2] = opp[3] = -1;
    const float* pv[4] = {a, b, c, d};
    float lmax = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int k = i + 1; k < 4; k++)
            lmax = fmaxf(lmax, (pv[i][0] - pv[k][0]) * (pv[i][0] - pv[k][0]) + (pv[i][1] - pv[k][1]) * (pv[i][1] - pv[k][1]) +
                                   (pv[i][2] - pv[k][2]) * (pv[i][2] - pv[k][2]));
    lmax = sqrtf(lmax) * 1.000001f;          // Euclidean length of the longest edge, rounded up
    float4* o = reinterpret_cast<float4*>(tgeo + (size_t)t * kTGeoWords);
    o[0] = make_float4(sx, sy, sz, __uint_as_float(flip));
    o[1] = make_float4(__int_as_float(tf.x), __int_as_float(tf.y), __int_as_float(tf.z), __int_as_float(tf.w));
    o[2] = make_float4(__int_as_float(opp[0]), __int_as_float(opp[1]), __int_as_float(opp[2]), __int_as_float(opp[3]));
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float4 ov = make_float4(0.f, 0.f, 0.f, lmax), nv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tfi[i] >= 0 && tfi[i] < F) {
            const float4 fc = reinterpret_cast<const float4*>(fgeo + (size_t)tfi[i] * kFGeoWords)[2];
            nv = make_float4(fc.y, fc.z, fc.w, 0.0f);
        }
        if (opp[i] >= 0) ov = make_float4(verts[3 * (size_t)opp[i]], verts[3 * (size_t)opp[i] + 1], verts[3 * (size_t)opp[i] + 2], lmax);
        o[4 + i] = ov;
        o[8 + i] = nv;
    }
    o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Second pass over the tets (needs [1], [2] of the neighbours): the per-slot link records [12 + i], and +inf as the
// longest edge of every tet that is not REGULAR - four distinct valid faces, four distinct opposite vertices, and
// behind every face either nothing or a tet that lists the face exactly once and each of its three vertices
// exactly once among its own (valid, distinct) opposite vertices.  In a regular tet the reference's "curr_face
// appears exactly once" (forward.cu:880-893) holds by construction and slots can be carried instead of searched.
__device__ __forceinline__ bool tet_self_valid(const int tf[4], const int ov[4], int F) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ok &= tf[i] >= 0 && tf[i] < F && ov[i] >= 0;
#pragma unroll
        for (int k = i + 1; k < 4; k++) ok &= tf[i] != tf[k] && ov[i] != ov[k];
    }
    return ok;
}
__global__ void __launch_bounds__(256)
k_tet_link(int T, int F, const int32_t* __restrict__ face_tets, float* __restrict__ tgeo) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    float4* o = reinterpret_cast<float4*>(tgeo + (size_t)t * kTGeoWords);
    const float4 q0 = o[0], q1 = o[1], q2 = o[2];
    const int tf[4] = {__float_as_int(q1.x), __float_as_int(q1.y), __float_as_int(q1.z), __float_as_int(q1.w)};
    const int ov[4] = {__float_as_int(q2.x), __float_as_int(q2.y), __float_as_int(q2.z), __float_as_int(q2.w)};
    const unsigned flip = __float_as_uint(q0.w);
    bool regular = tet_self_valid(tf, ov, F);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        int nb = -1;
        unsigned link = ((flip >> e) & 1u) << 12;
        if (tf[e] >= 0 && tf[e] < F) {
            const int ft0 = face_tets[2 * (size_t)tf[e]], ft1 = face_tets[2 * (size_t)tf[e] + 1];
            nb = ft0 != t ? ft0 : ft1 != t ? ft1 : -1;
        }
        if (nb < -1 || nb >= T) { regular = false; }
        else if (nb >= 0) {
            const float4* p = reinterpret_cast<const float4*>(tgeo + (size_t)nb * kTGeoWords);
            const float4 n0 = p[0], n1 = p[1], n2 = p[2];
            const int tf2[4] = {__float_as_int(n1.x), __float_as_int(n1.y), __float_as_int(n1.z), __float_as_int(n1.w)};
            const int ov2[4] = {__float_as_int(n2.x), __float_as_int(n2.y), __float_as_int(n2.z), __float_as_int(n2.w)};
            bool ok = tet_self_valid(tf2, ov2, F);
            int cnt = 0, bs = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) if (tf2[k] == tf[e]) { cnt++; bs = k; }
            ok &= cnt == 1;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j == e) continue;
                int c2 = 0, jj = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) if (ov2[k] == ov[j]) { c2++; jj = k; }
                ok &= c2 == 1 && jj != bs;
                link |= (unsigned)jj << (2 * j);
            }
            link |= (unsigned)bs << 8 | ((__float_as_uint(n0.w) >> bs) & 1u) << 10;
            regular &= ok;
        }
        reinterpret_cast<int4*>(o)[12 + e] = make_int4(tf[e], nb, (int)link, 0);
    }
    if (!regular) {
#pragma unroll
        for (int e = 0; e < 4; e++) o[4 + e].w = __int_as_float(0x7f800000);
    }
}

struct FGeo { f3 p0, e1, e2, n; int ft0, ft1, exist; };
__device__ __forceinline__ FGeo load_fgeo(const float* __restrict__ fgeo, int f) {
    const float4* p = reinterpret_cast<const float4*>(fgeo + (size_t)f * kFGeoWords);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    return FGeo{f3{a.x, a.y, a.z}, f3{a.w, b.x, b.y}, f3{b.z, b.w, c.x}, f3{c.y, c.z, c.w},
                __float_as_int(d.x), __float_as_int(d.y), __float_as_int(d.z)};
}

// dot(outward normal of the face in slot `slot` of a tet, ray_d)   (auxiliary.h:382-431 + the caller's dot)
__device__ __forceinline__ float outward_dot(const FGeo& g, unsigned flip, int slot, const f3& rd) {
    const float nd = dot3(g.n, rd);
    return ((flip >> slot) & 1u) ? -nd : nd;     // dot(-n, rd) == -dot(n, rd) exactly
}

// Moeller-Trumbore (auxiliary.h:212-243) with the stored edge vectors: same operations as ray_tri()
__device__ __forceinline__ bool hit_test(const f3& o, const f3& d, const FGeo& g) {
    const f3 T = sub3(o, g.p0);
    const f3 Pv = cross3(d, g.e2);
    const f3 Q = cross3(T, g.e1);
    const float denom = dot3(Pv, g.e1);
    if (denom == 0.0f) return false;
    const float inv = __fdiv_rn(1.0f, denom);
    const float t = __fmul_rn(dot3(Q, g.e2), inv);
    const float u = __fmul_rn(dot3(Pv, T), inv);
    const float v = __fmul_rn(dot3(Q, d), inv);
    return t >= 0.0f && u >= 0.0f && v >= 0.0f && __fadd_rn(u, v) <= 1.0f;
}

__device__ __forceinline__ int slot_of(const float4 tfw, int face) {
    return __float_as_int(tfw.x) == face ? 0 : __float_as_int(tfw.y) == face ? 1 : __float_as_int(tfw.z) == face ? 2
         : __float_as_int(tfw.w) == face ? 3 : -1;
}
// the same quantity from the geometry (face not listed by the tet: inconsistent adjacency input)
__device__ __forceinline__ float outward_dot_geo(const FGeo& g, const float4 tsum, const f3& rd) {
    const f3 d{__fmaf_rn(tsum.x, 0.25f, -g.p0.x), __fmaf_rn(tsum.y, 0.25f, -g.p0.y), __fmaf_rn(tsum.z, 0.25f, -g.p0.z)};
    const float nd = dot3(g.n, rd);
    return (dot3(g.n, d) > 0.0f) ? -nd : nd;
}

#ifdef DM2_WALK_STATS          /* debug build (tools/build_variant.py walkstats -DDM2_WALK_STATS): step counters */
__device__ unsigned long long g_walk_stats[8];
#define DM2_WSTAT(i) atomicAdd(&g_walk_stats[i], 1ull)
extern "C" void dm2_debug_walk_stats(unsigned long long* out8, int reset) {
    cudaMemcpyFromSymbol(out8, g_walk_stats, sizeof(g_walk_stats));
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_walk_stats, z, sizeof(z)); }
}
#else
#define DM2_WSTAT(i) do { } while (0)
#endif

// State of the walk's fast path: the entry face's vertices, ordered so that the ray sees them counter-clockwise
// ([d, a-o, b-o], [d, b-o, c-o], [d, c-o, a-o] > 0), their ids, and the magnitudes of those three edge functions.
#ifndef DM2_WALK_K
#define DM2_WALK_K 16.0f
#endif
constexpr float kWalkMargin = DM2_WALK_K * 5.9604645e-8f;
struct WalkFace { f3 a, b, c; int ia, ib, ic; float mab, mbc, mca, zmin; bool ok; int cslot, pa, pb, pc; unsigned eflip; };
__device__ __forceinline__ void walk_face_init(WalkFace& wf, const float* __restrict__ verts, const int32_t* __restrict__ faces,
                                               int face, int P, const f3& ro, const f3& rd) {
    wf.ok = false;
    int i0 = __ldg(faces + 3 * (size_t)face), i1 = __ldg(faces + 3 * (size_t)face + 1), i2 = __ldg(faces + 3 * (size_t)face + 2);
    if ((unsigned)i0 >= (unsigned)P || (unsigned)i1 >= (unsigned)P || (unsigned)i2 >= (unsigned)P) return;
    f3 a{__ldg(verts + 3 * (size_t)i0), __ldg(verts + 3 * (size_t)i0 + 1), __ldg(verts + 3 * (size_t)i0 + 2)};
    f3 b{__ldg(verts + 3 * (size_t)i1), __ldg(verts + 3 * (size_t)i1 + 1), __ldg(verts + 3 * (size_t)i1 + 2)};
    f3 c{__ldg(verts + 3 * (size_t)i2), __ldg(verts + 3 * (size_t)i2 + 1), __ldg(verts + 3 * (size_t)i2 + 2)};
    // [d, x - o, y - o] = (d x (x - o)) . (y - x)
    const f3 ao = sub3(a, ro), bo = sub3(b, ro), co = sub3(c, ro);
    float eab = dot3(cross3(rd, ao), sub3(b, a));
    float ebc = dot3(cross3(rd, bo), sub3(c, b));
    float eca = dot3(cross3(rd, co), sub3(a, c));
    wf.zmin = fminf(dot3(rd, ao), fminf(dot3(rd, bo), dot3(rd, co)));       // smallest depth (x - o) . d of a vertex
    if (eab < 0.0f && ebc < 0.0f && eca < 0.0f) {            // seen clockwise: swap b and c
        const f3 t = b; b = c; c = t;
        const int ti = i1; i1 = i2; i2 = ti;
        const float e = eab; eab = -eca; eca = -e; ebc = -ebc;
    }
    if (!(eab > 0.0f && ebc > 0.0f && eca > 0.0f)) return;
    wf.a = a; wf.b = b; wf.c = c; wf.ia = i0; wf.ib = i1; wf.ic = i2;
    wf.mab = eab; wf.mbc = ebc; wf.mca = eca;
    wf.ok = true;
}

__device__ __forceinline__ void load_layer_ray(const DM2LayerInputs& in, int b, int px, int py, size_t pix, f3& ro, f3& rd) {
    if (in.ray_view && in.ray_recipe > 0) {
        make_ray(in.ray_params + 32 * (size_t)in.ray_view[b], in.ray_view[b], px, py, in.W, in.H, in.ray_v, in.ray_recipe, ro, rd);
    } else {
        const size_t rix = in.ray_view ? ((size_t)in.ray_view[b] * in.H + py) * in.W + px : pix;
        const float* o = in.image_ray_o + rix * 3;
        const float* d = in.image_ray_d + rix * 3;
        ro = f3{o[0], o[1], o[2]};
        rd = f3{d[0], d[1], d[2]};
    }
}

// Every call (face_existence is the one input an incremental re-generate changes): the existence flag of the face in
// each slot, next to the link record the fast step reads anyway.
__global__ void __launch_bounds__(256)
k_tet_exist(int T, int F, const int32_t* __restrict__ face_existence, float* __restrict__ tgeo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;        // (tet, slot)
    if (i >= 4 * T) return;
    int4* r = reinterpret_cast<int4*>(tgeo + (size_t)(i >> 2) * kTGeoWords) + 12 + (i & 3);
    const int f = r->x;
    r->w = (f >= 0 && f < F) ? face_existence[f] : 0;
}

// Completes the fast-path state for the step that enters tet `tet` through `face`: the slot of the entry face and
// the slots (position among the tet's opposite vertices) of the entry face's three vertices.
__device__ __forceinline__ void walk_enter_tet(WalkFace& wf, const float* __restrict__ tgeo, int tet, int face) {
    if (!wf.ok) return;
    wf.ok = false;
    if (tet < 0) return;
    const float4* tp = reinterpret_cast<const float4*>(tgeo + (size_t)tet * kTGeoWords);
    const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
    if (!(__ldg(reinterpret_cast<const float*>(tp + 4) + 3) < 3.0e38f)) return;          // not a regular tet
    const int cs = slot_of(q1, face);
    const int pa = slot_of(q2, wf.ia), pb = slot_of(q2, wf.ib), pc = slot_of(q2, wf.ic);
    if (cs < 0 || pa < 0 || pb < 0 || pc < 0) return;
    wf.cslot = cs; wf.pa = pa; wf.pb = pb; wf.pc = pc;
    wf.eflip = (__float_as_uint(q0.w) >> cs) & 1u;
    wf.ok = true;
}

// First hit over the min-depth-sorted tile list + choice of the entry tet: one CTA per 16x16 tile (the list is staged
// through shared memory by all 256 threads); results go to first_face / first_tet (also the state an incremental
// re-generate starts from).
__global__ void __launch_bounds__(256)
k_layers_first_hit(DM2LayerInputs in, const float* __restrict__ lrec, const uint32_t* __restrict__ list,
                   const uint2* __restrict__ ranges, const float* __restrict__ fgeo, const float* __restrict__ tgeo,
                   int* __restrict__ first_face_out, int* __restrict__ first_tet_out, int gx, int gy) {
    __shared__ __align__(16) float srec[kLStage * kLRecWords];
    const int tiles = gx * gy;
    const int b = blockIdx.x / tiles, t_in = blockIdx.x % tiles;
    const int tile_x = t_in % gx, tile_y = t_in / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = tile_x * kTile + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < in.W && py < in.H;
    const size_t pix = inside ? ((size_t)b * in.H + py) * in.W + px : 0;
    f3 ro{0, 0, 0}, rd{0, 0, 0};
    if (inside) load_layer_ray(in, b, px, py, pix, ro, rd);
    const uint2 range = ranges[blockIdx.x];
    const int total = (int)(range.y - range.x);
    const size_t view_base = (size_t)b * in.F;

    // ---------------- first hit (forward.cu:603-681)
    bool done = !inside;
    float min_T = -1.0f, min_T_max_depth = -1.0f;
    int first_face = -1;
    for (int base = 0; base < total; base += kLStage) {
        if (__syncthreads_count(done) == 256) break;
        const int cnt = min(kLStage, total - base);
        for (int q = threadIdx.x; q < cnt * 3; q += 256) {
            const int slot = q / 3, chunk = q - slot * 3;
            const uint32_t fid = list[(size_t)range.x + base + slot];
            reinterpret_cast<float4*>(srec)[slot * 3 + chunk] =
                __ldg(reinterpret_cast<const float4*>(lrec + (view_base + fid) * kLRecWords) + chunk);
        }
        __syncthreads();
        for (int j = 0; j < cnt; j++) {
            if (__all_sync(0xffffffffu, done)) break;
            if (done) continue;
            const float* r = srec + j * kLRecWords;
            if (min_T >= 0.0f && r[9] > min_T_max_depth) { done = true; continue; }
            float t, u, v;
            if (!ray_tri(ro, rd, f3{r[0], r[1], r[2]}, f3{r[3], r[4], r[5]}, f3{r[6], r[7], r[8]}, t, u, v)) continue;
            const bool hit = t >= 0.0f && u >= 0.0f && v >= 0.0f && __fadd_rn(u, v) <= 1.0f;
            if (!hit) continue;
            if (min_T < 0.0f || t < min_T) {
                min_T = t;
                min_T_max_depth = r[10];
                first_face = __float_as_int(r[11]);
            }
        }
    }
    if (!inside) return;

    // ---------------- entry tet (forward.cu:683-708)
    int first_tet = -1;
    if (first_face >= 0) {
        const FGeo cur = load_fgeo(fgeo, first_face);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int tet = i == 0 ? cur.ft0 : cur.ft1;
            if (tet < 0) continue;
            const float4* tp = reinterpret_cast<const float4*>(tgeo + (size_t)tet * kTGeoWords);
            const float4 ts = __ldg(tp), tfw = __ldg(tp + 1);
            const int sl = slot_of(tfw, first_face);
            const float od = sl >= 0 ? outward_dot(cur, __float_as_uint(ts.w), sl, rd) : outward_dot_geo(cur, ts, rd);
            if (od < 0.0f) first_tet = tet;
        }
    }
    first_face_out[pix] = first_face;
    first_tet_out[pix] = first_tet;
}

// The walk: one WARP per 8x4-pixel footprint, CTAs of kWalkThreads (the rays of a tile take 0 .. 100+ steps: with one
// 256-thread CTA per tile a third of the SM's warp slots waited for the slowest warp of their CTA).
#ifndef DM2_WALK_THREADS
#define DM2_WALK_THREADS 64
#endif
constexpr int kWalkThreads = DM2_WALK_THREADS;
// register budget: measured at cfg3 (ms per walk-only generate, through the public API) 16 CTAs / SM (64 registers)
// 0.595, 24 CTAs 0.623, 8 CTAs (128 registers) 0.635 in a same-box A/B (an earlier A/B had shown 0.571 for 8: not
// reproducible); CTAs of 32 / 128 threads 0.597 / 0.597
#ifndef DM2_WALK_MINB
#define DM2_WALK_MINB 16
#endif
__global__ void __launch_bounds__(kWalkThreads, DM2_WALK_MINB)
k_layers_walk(DM2LayerInputs in, const float* __restrict__ fgeo, const float* __restrict__ tgeo,
              const int* __restrict__ first_face_in, const int* __restrict__ first_tet_in, int32_t* __restrict__ layers,
              int32_t* __restrict__ layers_cnt, int gx, int gy) {
    const int tiles = gx * gy;
    const int fp_id = blockIdx.x * (kWalkThreads / 32) + (threadIdx.x >> 5);      // global footprint: tile * 8 + footprint
    const int tile = fp_id >> 3, warp = fp_id & 7, lane = threadIdx.x & 31;
    if (tile >= in.B * tiles) return;
    const int b = tile / tiles, t_in = tile % tiles;
    const int tile_x = t_in % gx, tile_y = t_in / gx;
    const int px = tile_x * kTile + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * kTile + (warp >> 1) * 4 + (lane >> 3);
    if (!(px < in.W && py < in.H)) return;
    const size_t pix = ((size_t)b * in.H + py) * in.W + px;
    f3 ro, rd;
    load_layer_ray(in, b, px, py, pix, ro, rd);
    const int first_face = first_face_in[pix], first_tet = first_tet_in[pix];
    float4 c2 = make_float4(0.f, 0.f, 0.f, 0.f), c3 = c2;
    if (first_face >= 0) {
        const float4* fp = reinterpret_cast<const float4*>(fgeo + (size_t)first_face * kFGeoWords);
        c2 = __ldg(fp + 2); c3 = __ldg(fp + 3);
    }
    struct { f3 n; int exist; } cur{f3{c2.y, c2.z, c2.w}, __float_as_int(c3.z)};

    // ---------------- walk (forward.cu:842-999)
    // The reference decides every step with three Moeller-Trumbore tests (the three other faces of the tet) and
    // stops unless exactly one of them is hit with the ray leaving the tet through it.  FAST PATH: with the entry
    // face (a, b, c) ordered so that the ray sees it counter-clockwise and w the tet's fourth vertex, the signs of
    //     s_x = [d, w - o, x - o] = (d x (w - o)) . (x - w),   x = a, b, c
    // name the face the ray leaves through - 24 flops instead of 3 x 45 + an IEEE division each.  The three
    // numerators of a Moeller-Trumbore test ARE these edge functions (u, v, 1-u-v = [d, x-o, y-o] / denom for the
    // face's three edges), evaluated with an absolute error below ~6 eps |d| |T| |e|; the fast path is taken only
    // when every edge function it relies on (three new ones, and the three of the entry face, which it carries
    // along) exceeds kWalkMargin = 16 u |d|_1 (|w - o|_1 + l) l (u = 2^-24, l = length of the longest tet edge; the
    // worst-case bound of the two evaluations together is 14 u in Euclidean norms, and |.|_1 >= |.|_2) and every vertex of the exit face lies
    // clearly in front of the ray origin.  Then the reference's test of the exit face passes and its tests of the two other faces fail
    // whatever their rounding (each of those has one robustly positive and one robustly negative edge function, so
    // one of u, v, 1-u-v is robustly negative for either sign of the denominator).  The orientation test of the
    // exit face (outward normal . d > 0) is evaluated exactly as the reference does.  Anything else - a margin
    // not met, an irregular tet, an unexpected sign pattern - takes the reference's three tests (SLOW PATH) for
    // that step, and the fast-path state is rebuilt from the face it arrives at.
    const int L = in.num_layers;
    int curr_face = first_face, curr_tet = first_tet, n_done = 0;
    const bool wdone = (curr_face == -1 || curr_tet == -1);
    unsigned guard = (unsigned)min((long long)in.T + in.F + 16, 0x7fffffffll);   // safety net against cyclic walks
    float nd_cur = 0.0f;                                     // dot(stored normal of curr_face, ray_d)
    int exist_cur = 0;
    WalkFace wf;
    wf.ok = false;
    wf.cslot = 0;
    const float rd1 = fabsf(rd.x) + fabsf(rd.y) + fabsf(rd.z);
    int32_t* const lp = layers + pix * L;
    if (!wdone) {
        nd_cur = dot3(cur.n, rd);
        exist_cur = cur.exist;
        walk_face_init(wf, in.verts, in.faces, curr_face, in.P, ro, rd);
        walk_enter_tet(wf, tgeo, curr_tet, curr_face);
    }
    // (the apex record of the tet a step enters is requested at the end of the step before, together with the
    // existence flag of the face it arrives at: one memory round trip per step)
    const float4* tp = reinterpret_cast<const float4*>(tgeo + (size_t)max(curr_tet, 0) * kTGeoWords);
    float4 wq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!wdone) wq = __ldg(tp + 4 + wf.cslot);
    while (!wdone && guard-- > 0u) {
        if (exist_cur) {
            if (n_done < L) lp[n_done] = curr_face;
            n_done++;
            if (n_done >= L) break;
        }
        if (curr_tet == -1) break;
        int next_face = -1, next_tet = -1;
        bool stepped = false;
        DM2_WSTAT(0);                                         // steps
        if (!wf.ok) DM2_WSTAT(1);                             // no fast state
        if (wf.ok) {
            // (a regular tet lists curr_face exactly once, in slot wf.cslot)
            if ((wf.eflip ? -nd_cur : nd_cur) >= 0.0f) break;                  // outward_dot(curr_face) >= 0
            const f3 w{wq.x, wq.y, wq.z};
            const float lmax = wq.w;
            const f3 wo = sub3(w, ro);
            const f3 m = cross3(rd, wo);
            const float sa = dot3(m, sub3(wf.a, w)), sb = dot3(m, sub3(wf.b, w)), sc = dot3(m, sub3(wf.c, w));
            const float thr = kWalkMargin * rd1 * (fabsf(wo.x) + fabsf(wo.y) + fabsf(wo.z) + lmax) * lmax;
            const float lo = fminf(fminf(fabsf(sa), fabsf(sb)), fminf(fabsf(sc), fminf(wf.mab, fminf(wf.mbc, wf.mca))));
            // leaves through (a, b, w): drop c; through (b, c, w): drop a; through (c, a, w): drop b
            const bool dc = sa > 0.0f && sb < 0.0f, da = sb > 0.0f && sc < 0.0f, db = sc > 0.0f && sa < 0.0f;
            // every vertex of the exit face lies in front of the ray origin by more than the error of the
            // reference's hit distance (~0.1 lmax |d| at this margin): its `t >= 0` holds
            const float zmin = fminf(wf.zmin, dot3(rd, wo));
            const bool front = zmin > 0.25f * rd1 * rd1 * lmax;
            if (!(lo > thr)) DM2_WSTAT(2);                    // margin (or not a regular tet: lmax = inf)
            else if (!front) DM2_WSTAT(3);
            else if (!(dc | da | db)) DM2_WSTAT(4);
            if (lo > thr && front && (dc | da | db)) {        // (the three patterns exclude each other)
                const int eslot = dc ? wf.pc : da ? wf.pa : wf.pb;
                const int4 sr = __ldg(reinterpret_cast<const int4*>(tp) + 12 + eslot);
                const float4 nq = __ldg(tp + 8 + eslot);
                const unsigned link = (unsigned)sr.z;
                const float nd = dot3(f3{nq.x, nq.y, nq.z}, rd);
                if ((((link >> 12) & 1u) ? -nd : nd) > 0.0f) {
                    next_face = sr.x;
                    next_tet = sr.y;
                    nd_cur = nd;
                    exist_cur = sr.w;
                    stepped = true;
                    DM2_WSTAT(5);                             // fast steps
                    wf.zmin = zmin;
                    // the dropped vertex is replaced in place by w (slot cslot of this tet): (a, b, w), (w, b, c),
                    // (a, w, c) keep the orientation; then the three slots are renamed for the tet behind
                    if (dc) { wf.c = w; wf.pc = wf.cslot; wf.mbc = fabsf(sb); wf.mca = fabsf(sa); }
                    if (da) { wf.a = w; wf.pa = wf.cslot; wf.mab = fabsf(sb); wf.mca = fabsf(sc); }
                    if (db) { wf.b = w; wf.pb = wf.cslot; wf.mab = fabsf(sa); wf.mbc = fabsf(sc); }
                    wf.pa = (link >> (2 * wf.pa)) & 3u;
                    wf.pb = (link >> (2 * wf.pb)) & 3u;
                    wf.pc = (link >> (2 * wf.pc)) & 3u;
                    wf.cslot = (link >> 8) & 3u;
                    wf.eflip = (link >> 10) & 1u;
                }
            }
        }
        if (!stepped) {
            // the reference's step: the 3 other faces of the tet, in slot order (forward.cu:880-893: stop unless
            // curr_face appears exactly once), three hit tests, exactly one of them leaving the tet
            const float4* tq = reinterpret_cast<const float4*>(tgeo + (size_t)curr_tet * kTGeoWords);
            const float4 ts = __ldg(tq), tfw = __ldg(tq + 1);
            const int tf[4] = {__float_as_int(tfw.x), __float_as_int(tfw.y), __float_as_int(tfw.z), __float_as_int(tfw.w)};
            const unsigned flip = __float_as_uint(ts.w);
            const int m0 = tf[0] == curr_face, m1 = tf[1] == curr_face, m2 = tf[2] == curr_face, m3 = tf[3] == curr_face;
            if (m0 + m1 + m2 + m3 != 1) break;
            const int cslot = m0 ? 0 : m1 ? 1 : m2 ? 2 : 3;
            if ((((flip >> cslot) & 1u) ? -nd_cur : nd_cur) >= 0.0f) break;    // outward_dot(curr_face) >= 0
            const int oslot[3] = {cslot == 0 ? 1 : 0, cslot <= 1 ? 2 : 1, cslot <= 2 ? 3 : 2};
            const int others[3] = {cslot == 0 ? tf[1] : tf[0], cslot <= 1 ? tf[2] : tf[1], cslot <= 2 ? tf[3] : tf[2]};
            int next_cnt = 0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const FGeo g = load_fgeo(fgeo, others[i]);
                if (hit_test(ro, rd, g) && outward_dot(g, flip, oslot[i], rd) > 0.0f) { next_face = others[i]; next_cnt++; }
            }
            if (next_cnt != 1) { DM2_WSTAT(6); break; }
            const float4* fp = reinterpret_cast<const float4*>(fgeo + (size_t)next_face * kFGeoWords);
            const float4 nrec2 = __ldg(fp + 2), nrec3 = __ldg(fp + 3);       // just read: L1 hits
            nd_cur = dot3(f3{nrec2.y, nrec2.z, nrec2.w}, rd);
            exist_cur = __float_as_int(nrec3.z);
            const int ft0 = __float_as_int(nrec3.x), ft1 = __float_as_int(nrec3.y);
            if (ft0 != curr_tet) next_tet = ft0;
            else if (ft1 != curr_tet) next_tet = ft1;
            walk_face_init(wf, in.verts, in.faces, next_face, in.P, ro, rd);
            walk_enter_tet(wf, tgeo, next_tet, next_face);
        }
        curr_face = next_face;
        curr_tet = next_tet;
        tp = reinterpret_cast<const float4*>(tgeo + (size_t)max(curr_tet, 0) * kTGeoWords);
        wq = __ldg(tp + 4 + wf.cslot);
    }
    layers_cnt[pix] = n_done;
}

__global__ void __launch_bounds__(256)
k_generate_rays(const float* __restrict__ prm, int V, int H, int W, int recipe, float* __restrict__ out_o,
                float* __restrict__ out_d) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)V * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), v = (int)(i / ((long long)W * H));
    f3 o, d;
    make_ray(prm + 32 * (size_t)v, v, x, y, W, H, V, recipe, o, d);
    out_o[3 * i] = o.x; out_o[3 * i + 1] = o.y; out_o[3 * i + 2] = o.z;
    out_d[3 * i] = d.x; out_d[3 * i + 1] = d.y; out_d[3 * i + 2] = d.z;
}

int launch_generate_rays(const float* ray_params, int V, int H, int W, int recipe, float* out_o, float* out_d, cudaStream_t s) {
    const long long n = (long long)V * H * W;
    if (n == 0) return 0;
    k_generate_rays<<<cdiv(n, 256), 256, 0, s>>>(ray_params, V, H, W, recipe, out_o, out_d);
    DM2_LAUNCH_CHECK();
    return 0;
}

// The per-geometry records (face / tet geometry, link records, existence flags): they depend on the geometry tensors
// only, not on the binning - dm2_layers_preprocess_with_geometry enqueues them behind the header read-back, so that
// they run while the host waits for num_rendered instead of after it.  links_too = false: existence flags only
// (incremental re-generation).
int launch_layers_geometry(const DM2LayerInputs& in, void* geo_ws, bool links_too, cudaStream_t s) {
    float* fgeo = reinterpret_cast<float*>(geo_ws);
    float* tgeo = reinterpret_cast<float*>(reinterpret_cast<char*>(geo_ws) + align_up((size_t)in.F * kFGeoWords * 4, 256));
    ProfScope _ps("layers_geometry", s);
    if (in.F > 0) {
        k_face_geo<<<cdiv(in.F, 256), 256, 0, s>>>(in.F, in.verts, in.faces, in.face_tets, in.face_existence, fgeo);
        DM2_LAUNCH_CHECK();
    }
    if (in.T > 0 && links_too) {
        k_tet_geo<<<cdiv(in.T, 256), 256, 0, s>>>(in.T, in.verts, in.tets, in.faces, in.tet_faces, in.face_tets, fgeo, in.F, in.P, tgeo);
        DM2_LAUNCH_CHECK();
        k_tet_link<<<cdiv(in.T, 256), 256, 0, s>>>(in.T, in.F, in.face_tets, tgeo);
        DM2_LAUNCH_CHECK();
    }
    if (in.T > 0) {
        k_tet_exist<<<cdiv(4ll * in.T, 256), 256, 0, s>>>(in.T, in.F, in.face_existence, tgeo);
        DM2_LAUNCH_CHECK();
    }
    return 0;
}

int launch_layers(const DM2LayerInputs& in, const FaceWs& fw, const BinWs& bw, const ImageWs& iw, void* geo_ws,
                  int32_t* render_layers, int32_t* render_layers_cnt, bool walk_only, bool geometry_ready, cudaStream_t s) {
    const int gx = cdiv(in.W, kTile), gy = cdiv(in.H, kTile);
    const long long nblk = (long long)in.B * gx * gy;
    if (nblk == 0) return 0;
    float* fgeo = reinterpret_cast<float*>(geo_ws);
    float* tgeo = reinterpret_cast<float*>(reinterpret_cast<char*>(geo_ws) + align_up((size_t)in.F * kFGeoWords * 4, 256));
    if (!geometry_ready && launch_layers_geometry(in, geo_ws, !walk_only, s)) return 1;
    ProfScope _ps(walk_only ? "layers_walk" : "layers_first_hit+walk", s);
    if (!walk_only) {
        k_layers_first_hit<<<(unsigned)nblk, 256, 0, s>>>(in, fw.records, bw.vals, iw.ranges, fgeo, tgeo, iw.first_face,
                                                          iw.first_tet, gx, gy);
        DM2_LAUNCH_CHECK();
    }
    k_layers_walk<<<(unsigned)cdiv(nblk * 8, kWalkThreads / 32), kWalkThreads, 0, s>>>(
        in, fgeo, tgeo, iw.first_face, iw.first_tet, render_layers, render_layers_cnt, gx, gy);
    DM2_LAUNCH_CHECK();
    return 0;
}

// =====================================================================================================
// Layer consumer: the step DMesh++ runs after LayeredRenderer.generate() - gather the
// attributes of the faces recorded per ray and composite them front to back - as ONE kernel per direction
// instead of L gathers + a cumulative product in torch (the reference returns ids only, __init__.py:404-451).
//   T_0 = 1;  a_l = opacity[f_l];  c_l = color[f_l] * intense[b, f_l] (intense optional)
//   C = sum_l T_l a_l c_l + T_L background;  T_{l+1} = T_l (1 - a_l);  alpha = 1 - T_L;   l < min(cnt, L)
// Backward: dL/dc_l = T_l a_l g;  dL/da_l = T_l (c_l . g) - (R_l - T_L g_alpha) / (1 - a_l) with
// R_l = sum_{m>l} T_m a_m (c_m . g) + T_L (background . g); per (ray, layer) ONE 16-byte vector reduction into a
// packed per-face record (d color, d opacity), unpacked by a second kernel.
constexpr int kCompMaxL = 32;

__global__ void __launch_bounds__(256)
k_layers_composite_fwd(long long npix, int HW, int L, int F, const int32_t* __restrict__ layers,
                       const int32_t* __restrict__ cnt, const float* __restrict__ color, const float* __restrict__ opacity,
                       const float* __restrict__ intense, const float* __restrict__ background,
                       float* __restrict__ out_color, float* __restrict__ out_alpha) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int n = min(cnt[i], L);
    const size_t vb = (size_t)(i / HW) * F;
    float T = 1.0f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int l = 0; l < n; l++) {
        const int f = layers[i * L + l];
        if ((unsigned)f >= (unsigned)F) continue;
        const float a = opacity[f];
        const float w = T * a * (intense ? intense[vb + f] : 1.0f);
        c0 = fmaf(w, color[3 * (size_t)f], c0);
        c1 = fmaf(w, color[3 * (size_t)f + 1], c1);
        c2 = fmaf(w, color[3 * (size_t)f + 2], c2);
        T *= 1.0f - a;
    }
    out_color[3 * i] = fmaf(T, background[0], c0);
    out_color[3 * i + 1] = fmaf(T, background[1], c1);
    out_color[3 * i + 2] = fmaf(T, background[2], c2);
    out_alpha[i] = 1.0f - T;
}

__global__ void __launch_bounds__(256)
k_layers_composite_bwd(long long npix, int HW, int L, int F, const int32_t* __restrict__ layers,
                       const int32_t* __restrict__ cnt, const float* __restrict__ color, const float* __restrict__ opacity,
                       const float* __restrict__ intense, const float* __restrict__ background,
                       const float* __restrict__ g_color, const float* __restrict__ g_alpha, float* __restrict__ gface,
                       float* __restrict__ g_intense) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int n = min(min(cnt[i], L), kCompMaxL);
    const size_t vb = (size_t)(i / HW) * F;
    const float g0 = g_color[3 * i], g1 = g_color[3 * i + 1], g2 = g_color[3 * i + 2], ga = g_alpha ? g_alpha[i] : 0.0f;
    float Ts[kCompMaxL];
    float T = 1.0f;
    for (int l = 0; l < n; l++) {
        Ts[l] = T;
        const int f = layers[i * L + l];
        if ((unsigned)f < (unsigned)F) T *= 1.0f - opacity[f];
    }
    // R = everything behind layer l as seen through dL/dC, minus the alpha output's share
    float R = T * (background[0] * g0 + background[1] * g1 + background[2] * g2) - T * ga;
    for (int l = n - 1; l >= 0; l--) {
        const int f = layers[i * L + l];
        if ((unsigned)f >= (unsigned)F) continue;
        const float a = opacity[f], it = intense ? intense[vb + f] : 1.0f;
        const float q0 = color[3 * (size_t)f], q1 = color[3 * (size_t)f + 1], q2 = color[3 * (size_t)f + 2];
        const float cg = q0 * g0 + q1 * g1 + q2 * g2;           // (colour . g), without the intensity
        const float Tl = Ts[l];
        const float w = Tl * a * it;
        const float da = Tl * it * cg - R / (1.0f - a);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(gface + 4 * (size_t)f), "f"(w * g0), "f"(w * g1),
                     "f"(w * g2), "f"(da) : "memory");
        if (g_intense) atomicAdd(g_intense + vb + f, Tl * a * cg);
        R += w * cg;
    }
}

__global__ void __launch_bounds__(256)
k_layers_composite_unpack(int F, const float* __restrict__ gface, float* __restrict__ d_color, float* __restrict__ d_opacity) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const float4 g = reinterpret_cast<const float4*>(gface)[f];
    d_color[3 * (size_t)f] = g.x; d_color[3 * (size_t)f + 1] = g.y; d_color[3 * (size_t)f + 2] = g.z;
    d_opacity[f] = g.w;
}

int launch_layers_composite_forward(int B, int H, int W, int L, int F, const int32_t* layers, const int32_t* cnt,
                                    const float* color, const float* opacity, const float* intense, const float* background,
                                    float* out_color, float* out_alpha, cudaStream_t s) {
    const long long npix = (long long)B * H * W;
    if (npix == 0) return 0;
    ProfScope _ps("layers_composite_forward", s);
    k_layers_composite_fwd<<<cdiv(npix, 256), 256, 0, s>>>(npix, H * W, L, F, layers, cnt, color, opacity, intense, background,
                                                          out_color, out_alpha);
    DM2_LAUNCH_CHECK();
    return 0;
}

int launch_layers_composite_backward(int B, int H, int W, int L, int F, const int32_t* layers, const int32_t* cnt,
                                     const float* color, const float* opacity, const float* intense, const float* background,
                                     const float* g_color, const float* g_alpha, void* grad_ws, float* d_color,
                                     float* d_opacity, float* d_intense, cudaStream_t s) {
    const long long npix = (long long)B * H * W;
    if (F == 0) return 0;
    DM2_REQUIRE(L <= kCompMaxL, "dm2_layers_composite_backward: at most 32 layers");
    float* gface = reinterpret_cast<float*>(grad_ws);
    ProfScope _ps("layers_composite_backward", s);
    DM2_CHECK(cudaMemsetAsync(gface, 0, (size_t)F * 16, s));
    if (npix > 0) {
        k_layers_composite_bwd<<<cdiv(npix, 256), 256, 0, s>>>(npix, H * W, L, F, layers, cnt, color, opacity, intense,
                                                              background, g_color, g_alpha, gface, d_intense);
        DM2_LAUNCH_CHECK();
    }
    k_layers_composite_unpack<<<cdiv(F, 256), 256, 0, s>>>(F, gface, d_color, d_opacity);
    DM2_LAUNCH_CHECK();
    return 0;
}

}  // namespace dm2
