// aa_fast_check.cpp - CPU validation of dm2_aa_fast.cuh (the polygon-free overlap area) against the oracle's transcription
// of the reference's polygon clip (oracle/dm2_oracle.cpp: orc_aa_overlap, reference cuda_impl/aa.h:96-504).
// TEST INFRASTRUCTURE: built and run by tests/test_aa_fast.py only.
//
// For random and adversarial (triangle, pixel) pairs it checks the contract of aa_area_fast():
//   whenever it answers (returns true), the reference answers too (error code 0, area > 0), the areas agree to 2e-6 and
//   the gradients to 2e-5 * (1 + |g|_inf);
// and reports how often it declines (the pairs that take the exact clip).
// usage: aa_fast_check <pairs per mode> <seed>      prints one JSON line; exit code 1 on a contract violation
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#define AAF_DEBUG_REASON 1
#include "../../dmesh2_renderer_b200/csrc/dm2_aa_fast.cuh"

extern "C" int orc_aa_overlap(const float* verts, const float* edges, const uint8_t* iszero, const float* recip,
                              const float* normal, const float* normal_c, float pxmin, float pymin, float* area, float* grad6);

// exact (double precision) overlap area by Sutherland-Hodgman clipping of the triangle against the pixel square: the
// ground truth for the gradient (central differences), which the reference's Jacobian chain only approximates - its
// polygon vertices carry ulp(coordinate) errors that the chain multiplies by the edge slopes
static double area_exact(const double* v, double px, double py) {
    double ax[16], ay[16], bx[16], by[16];
    int n = 3;
    for (int i = 0; i < 3; i++) { ax[i] = v[2 * i] - px; ay[i] = v[2 * i + 1] - py; }
    for (int side = 0; side < 4; side++) {
        int m = 0;
        auto inside = [&](double x, double y) { return side == 0 ? x >= 0 : side == 1 ? x <= 1 : side == 2 ? y >= 0 : y <= 1; };
        for (int i = 0; i < n; i++) {
            const int j = (i + 1) % n;
            const bool ii = inside(ax[i], ay[i]), ij = inside(ax[j], ay[j]);
            if (ii) { bx[m] = ax[i]; by[m] = ay[i]; m++; }
            if (ii != ij) {
                double t;
                if (side < 2) { const double c = side == 0 ? 0 : 1; t = (c - ax[i]) / (ax[j] - ax[i]); }
                else { const double c = side == 2 ? 0 : 1; t = (c - ay[i]) / (ay[j] - ay[i]); }
                bx[m] = ax[i] + t * (ax[j] - ax[i]); by[m] = ay[i] + t * (ay[j] - ay[i]); m++;
            }
        }
        n = m;
        for (int i = 0; i < n; i++) { ax[i] = bx[i]; ay[i] = by[i]; }
        if (n == 0) return 0;
    }
    double a = 0;
    for (int i = 0; i < n; i++) { const int j = (i + 1) % n; a += ax[i] * ay[j] - ax[j] * ay[i]; }
    return 0.5 * fabs(a);
}

struct Tri { float v[6], e[6], r[6], n[6], c[3]; uint8_t z[6]; unsigned flags; uint32_t safe; };

static void setup(Tri& t, float p0x, float p0y, float p1x, float p1y, float p2x, float p2y) {
    // Triangles.__init__ (pyrenderer.py:6-30): CCW re-ordering by the sign of tri_area, then fp32 set-up
    const float ax = p1x - p0x, by = p2y - p0y, cx = p2x - p0x, dy = p1y - p0y;
    const float sa = 0.5f * (ax * by - cx * dy);
    if (sa < 0) { std::swap(p1x, p2x); std::swap(p1y, p2y); }
    const float v[6] = {p0x, p0y, p1x, p1y, p2x, p2y};
    memcpy(t.v, v, sizeof v);
    for (int i = 0; i < 3; i++) {
        const int j = (i + 1) % 3;
        t.e[2 * i] = v[2 * j] - v[2 * i]; t.e[2 * i + 1] = v[2 * j + 1] - v[2 * i + 1];
        t.n[2 * i] = -t.e[2 * i + 1]; t.n[2 * i + 1] = t.e[2 * i];
        const float a = t.n[2 * i] * v[2 * i], b = t.n[2 * i + 1] * v[2 * i + 1];
        t.c[i] = a + b;
    }
    t.flags = 0;
    for (int k = 0; k < 6; k++) {
        t.z[k] = fabsf(t.e[k]) < 1e-3f;
        t.flags |= (unsigned)t.z[k] << k;
        t.r[k] = 1.0f / t.e[k];
    }
    t.flags |= dm2::aaf_face_safety(t.v, t.e, t.r, t.flags, &t.safe);
}

// unit checks of the per-face safety word (aaf_face_safety): thresholds packed as bf16 never round DOWN, faces the fast
// path must never answer for are flagged, a near-integer coordinate raises the flag that switches the vertex test on
static int unit_checks() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<float> U(0.0f, 1.0f);
    int bad = 0;
    for (int i = 0; i < 2000000; i++) {
        const float x = ldexpf(U(rng) + 1e-9f, (int)(U(rng) * 40) - 30);
        const float up = aaf_float_host(dm2::aaf_bf16_up(x) << 16);
        if (!(up >= x) || !(up <= x * 1.0079f)) bad |= 1;                       // one bf16 ulp = 2^-7
    }
    auto flags_of = [&](float p0x, float p0y, float p1x, float p1y, float p2x, float p2y, uint32_t* sw) {
        Tri t; setup(t, p0x, p0y, p1x, p1y, p2x, p2y); *sw = t.safe; return t.flags;
    };
    uint32_t sw;
    if (!(flags_of(10.3f, 10.2f, 14.7f, 10.2004f, 12.1f, 15.9f, &sw) & dm2::kAafNever)) bad |= 2;          // |e.y| < 1e-3
    if (!(flags_of(10.3f, 10.2f, 14.7f, NAN, 12.1f, 15.9f, &sw) & dm2::kAafNever)) bad |= 4;               // NaN
    if (!(flags_of(10.3f, 10.2f, 2.0e6f, 11.0f, 12.1f, 15.9f, &sw) & dm2::kAafNever)) bad |= 8;            // huge
    if (!(flags_of(10.3f, 10.2f, 1010.3f, 1010.2001f, 510.3f, 510.7f, &sw) & dm2::kAafNever)) bad |= 1 << 10;   // needle (wedge factor < 2e-3)
    unsigned f = flags_of(10.3f, 10.2f, 14.7f, 11.4f, 12.1f, 15.9f, &sw);
    if (f & (dm2::kAafNever | dm2::kAafNearInt)) bad |= 16;                                                  // ordinary face
    if (!(sw >> 16) || !(sw & 0xFFFFu)) bad |= 32;
    f = flags_of(10.3f, 10.2f, 14.0000001f, 11.4f, 12.1f, 15.9f, &sw);
    if (!(f & dm2::kAafNearInt) || (f & dm2::kAafNever)) bad |= 64;                                          // x = 14 (+ 1e-7)
    // edge data that is not v[i+1] - v[i] (a caller of the 21-argument function could pass anything): never
    {
        Tri t; setup(t, 10.3f, 10.2f, 14.7f, 11.4f, 12.1f, 15.9f);
        t.e[0] += 0.25f;
        unsigned fl = 0; for (int k = 0; k < 6; k++) fl |= (unsigned)t.z[k] << k;
        uint32_t w;
        if (!(dm2::aaf_face_safety(t.v, t.e, t.r, fl, &w) & dm2::kAafNever)) bad |= 128;
    }
    printf("{\"unit_failures\": %d}\n", bad);
    return bad ? 1 : 0;
}

// --pairs <file>: one record of 26 floats per pair (verts 6, edges 6, recip 6, iszero 6 as 0 / 1, pxmin, pymin), as
// tests/test_aa_fast.py writes them from tests/golden/aa_overlap_golden.npz (the reference's own pyrenderer.py);
// prints "answered area g0 .. g5" per pair
static int eval_pairs(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return 2;
    float r[26];
    while (fread(r, sizeof(float), 26, f) == 26) {
        unsigned flags = 0;
        for (int k = 0; k < 6; k++) flags |= (r[18 + k] != 0.0f ? 1u : 0u) << k;
        uint32_t sw = 0;
        flags |= dm2::aaf_face_safety(r, r + 6, r + 12, flags, &sw);
        float area = 0, g[6] = {0, 0, 0, 0, 0, 0};
        const bool ok = dm2::aa_area_fast<true>(r, r + 6, r + 12, flags, sw, r[24], r[25], area, g);
        printf("%d %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", ok ? 1 : 0, area, g[0], g[1], g[2], g[3], g[4], g[5]);
    }
    fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "--unit")) return unit_checks();
    if (argc > 2 && !strcmp(argv[1], "--pairs")) return eval_pairs(argv[2]);
    const long long n_per_mode = argc > 1 ? atoll(argv[1]) : 1000000;
    const uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<float> U(0.0f, 1.0f);
    std::normal_distribution<float> N(0.0f, 1.0f);
    long long total = 0, partial = 0, answered = 0, viol = 0, declined_ok = 0, declined_err = 0;
    double max_da = 0, max_dg = 0, max_dfd = 0;
    long long grad_vs_truth = 0;
    const int kModes = 12;
    long long mode_partial[kModes] = {0}, mode_answered[kModes] = {0};
    static long long reasons[kModes][5];
    for (int mode = 0; mode < kModes; mode++) {
        for (long long it = 0; it < n_per_mode; it++) {
            const int span = (it & 1) ? 2040 : 1000;
            float px = (float)(int)(U(rng) * span), py = (float)(int)(U(rng) * span);
            float p[6];
            const float scales[5] = {0.6f, 1.5f, 4.0f, 20.0f, 300.0f};
            const float s = scales[(int)(U(rng) * 5) % 5];
            for (int k = 0; k < 6; k++) p[k] = ((k & 1) ? py : px) + 0.5f + s * N(rng);
            const float tiny[8] = {0.0f, 6e-5f, 1.3e-4f, 3e-4f, 1e-3f, 3e-3f, 1e-2f, -1.0f};
            auto small = [&]() { float d = tiny[(int)(U(rng) * 8) % 8]; if (d < 0) d = U(rng) * 0.1f; return (U(rng) < 0.5f ? -d : d) * U(rng) * 2.0f; };
            if (mode == 10) {           // huge triangles, vertices far off screen (negative coordinates included)
                for (int k = 0; k < 6; k++) p[k] = ((k & 1) ? py : px) + 0.5f + 3000.0f * N(rng);
            } else if (mode == 11) {    // a patch far from the origin of a large frame (coordinates 4096 .. 16384)
                const float ox = 4096.0f + (float)(int)(U(rng) * 12288.0f), oy = 4096.0f + (float)(int)(U(rng) * 12288.0f);
                for (int k = 0; k < 6; k++) p[k] = ((k & 1) ? oy : ox) + 0.5f + s * N(rng);
                px = ox; py = oy;
            } else if (mode >= 8) {            // the benchmark scenes: cfg2-like (1024^2, sigma 4.7 px), cfg5-like (2048^2, sigma 6.6 px)
                const int sp = mode == 8 ? 1024 : 2048;
                const float sg = mode == 8 ? 4.7f : 6.6f;
                const float qx = (float)(int)(U(rng) * sp), qy = (float)(int)(U(rng) * sp);
                const float ccx = qx + 0.5f + 2.0f * N(rng), ccy = qy + 0.5f + 2.0f * N(rng);
                for (int k = 0; k < 6; k++) p[k] = ((k & 1) ? ccy : ccx) + sg * N(rng);
                px = qx; py = qy;
            } else if (mode == 1) {            // a vertex (nearly) on a pixel boundary line
                const int k = (int)(U(rng) * 6) % 6;
                p[k] = ((k & 1) ? py : px) + (float)(int)(U(rng) * 3 - 1) + small();
                p[k ^ 1] = ((k & 1) ? px : py) + U(rng) * 1.2f - 0.1f;
            } else if (mode == 2) {     // an edge (nearly) through a pixel corner
                const float cx = px + (float)((int)(U(rng) * 2) % 2), cy = py + (float)((int)(U(rng) * 2) % 2);
                const float ang = U(rng) * 6.2831853f, l0 = s * (0.05f + U(rng)), l1 = s * (0.05f + U(rng));
                const float d = small();
                p[0] = cx + cosf(ang) * l0 - sinf(ang) * d; p[1] = cy + sinf(ang) * l0 + cosf(ang) * d;
                p[2] = cx - cosf(ang) * l1 - sinf(ang) * d; p[3] = cy - sinf(ang) * l1 + cosf(ang) * d;
            } else if (mode == 3) {     // thin wedge, apex near a boundary line
                const bool horiz = U(rng) < 0.5f;
                const float h = small() * 5.0f, spread = 0.002f + U(rng) * U(rng) * 0.5f, len = 0.5f + U(rng) * 6.0f;
                const float ang = (U(rng) - 0.5f) * 1.0f;
                const float along = U(rng), sgn = U(rng) < 0.5f ? 1.0f : -1.0f;
                const float side = (float)((int)(U(rng) * 2) % 2);
                float ax_, ay_;
                if (horiz) { ax_ = px + along; ay_ = py + side + h; } else { ax_ = px + side + h; ay_ = py + along; }
                const float a1 = ang - spread * 0.5f, a2 = ang + spread * 0.5f;
                float d1x = sinf(a1), d1y = cosf(a1), d2x = sinf(a2), d2y = cosf(a2);
                if (!horiz) { std::swap(d1x, d1y); std::swap(d2x, d2y); }
                p[0] = ax_; p[1] = ay_;
                p[2] = ax_ + sgn * len * d1x; p[3] = ay_ + sgn * len * d1y;
                p[4] = ax_ + sgn * len * d2x; p[5] = ay_ + sgn * len * d2y;
            } else if (mode == 4) {     // nearly axis-parallel edge
                const int k = (int)(U(rng) * 2) % 2;
                p[2 + k] = p[k] + small() * 3.0f;
            } else if (mode == 5) {     // lattice-snapped vertices (quarter pixels)
                for (int k = 0; k < 6; k++) p[k] = rintf(p[k] * 4.0f) * 0.25f + (U(rng) < 0.3f ? small() : 0.0f);
            } else if (mode == 6) {     // sliver triangles
                const float tt = U(rng);
                p[4] = p[0] + tt * (p[2] - p[0]) + small(); p[5] = p[1] + tt * (p[3] - p[1]) + small();
            } else if (mode == 7) {     // pixel almost fully covered: a big triangle with one edge cutting a corner off
                const float cx = px + (float)((int)(U(rng) * 2) % 2), cy = py + (float)((int)(U(rng) * 2) % 2);
                const float ang = U(rng) * 6.2831853f, d = fabsf(small()) * 3.0f;
                const float nx = -sinf(ang), ny = cosf(ang);
                p[0] = cx + cosf(ang) * 30 + nx * d; p[1] = cy + sinf(ang) * 30 + ny * d;
                p[2] = cx - cosf(ang) * 30 + nx * d; p[3] = cy - sinf(ang) * 30 + ny * d;
                const float sg = ((px + 0.5f - cx) * nx + (py + 0.5f - cy) * ny) > 0 ? 1.0f : -1.0f;
                p[4] = cx + sg * nx * 60; p[5] = cy + sg * ny * 60;
            }
            Tri t;
            setup(t, p[0], p[1], p[2], p[3], p[4], p[5]);
            total++;
            float ra = 0, rg[6];
            const int err = orc_aa_overlap(t.v, t.e, t.z, t.r, t.n, t.c, px, py, &ra, rg);
            // classify like the pre-pass: only partially covered pixels reach the clip.  The oracle entry runs steps 1-3
            // itself: outside -> (0, area 0); fully inside -> (0, area 1, zero gradient).
            bool all_in = true, any_edge_all_out = false;
            for (int e = 0; e < 3; e++) {
                bool all_out = true;
                for (int cidx = 0; cidx < 4; cidx++) {
                    const float cx = (cidx == 1 || cidx == 2) ? px + 1 : px, cy = (cidx >= 2) ? py + 1 : py;
                    const float a = cx * t.n[2 * e], b = cy * t.n[2 * e + 1];
                    const bool in = (a + b) - t.c[e] >= 0;
                    all_out = all_out && !in; all_in = all_in && in;
                }
                any_edge_all_out = any_edge_all_out || all_out;
            }
            const float txmin = fminf(fminf(t.v[0], t.v[2]), t.v[4]), txmax = fmaxf(fmaxf(t.v[0], t.v[2]), t.v[4]);
            const float tymin = fminf(fminf(t.v[1], t.v[3]), t.v[5]), tymax = fmaxf(fmaxf(t.v[1], t.v[3]), t.v[5]);
            const bool bb_out = (px + 1 < txmin) || (px > txmax) || (py + 1 < tymin) || (py > tymax);
            if (bb_out || any_edge_all_out || all_in) continue;       // not a clip case (the flags here may differ from the
                                                                      // oracle's fma pattern by an ulp: harmless for the test)
            partial++; mode_partial[mode]++;
            float fa = 0, fg[6];
            const bool ok = dm2::aa_area_fast<true>(t.v, t.e, t.r, t.flags, t.safe, px, py, fa, fg);
            if (!ok) for (int b = 0; b < 5; b++) if (dm2::aaf_reason & (1 << b)) reasons[mode][b]++;
            if (!ok) { (err == 0 && ra != 0.0f) ? declined_ok++ : declined_err++; continue; }
            answered++; mode_answered[mode]++;
            bool bad = (err != 0) || (ra == 0.0f);
            float emax = 0; for (int k = 0; k < 6; k++) emax = fmaxf(emax, fabsf(t.e[k]));
            double da = fabs((double)fa - ra), dg = 0, gm = 0;
            for (int k = 0; k < 6; k++) { dg = fmax(dg, fabs((double)fg[k] - rg[k])); gm = fmax(gm, fabs((double)rg[k])); }
            if (!bad) { max_da = fmax(max_da, da); max_dg = fmax(max_dg, dg / (1 + gm)); }
            if (!bad && dg > 2e-5 * (1 + gm)) {
                // not the reference's chain to 2e-5: then it must be the TRUE derivative (central differences in double)
                double dv[6], dfd = 0;
                for (int k = 0; k < 6; k++) dv[k] = t.v[k];
                for (int k = 0; k < 6; k++) {
                    const double h = 2e-6, keep = dv[k];
                    dv[k] = keep + h; const double ap = area_exact(dv, px, py);
                    dv[k] = keep - h; const double am = area_exact(dv, px, py);
                    dv[k] = keep;
                    dfd = fmax(dfd, fabs((ap - am) / (2 * h) - fg[k]));
                }
                grad_vs_truth++;
                max_dfd = fmax(max_dfd, dfd);
                if (dfd <= 2e-5 * (1 + gm) + 5e-7 * emax) dg = 0;
                else dg = dfd;
            }
            if (bad || da > 2e-6 || dg > 2e-5 * (1 + gm) + 5e-7 * emax) {
                if (viol < 12)
                    fprintf(stderr, "VIOLATION mode %d err %d ref %.9g fast %.9g dg %.3g gm %.3g px %.1f py %.1f v %.9g %.9g %.9g %.9g %.9g %.9g\n",
                            mode, err, ra, fa, dg, gm, px, py, t.v[0], t.v[1], t.v[2], t.v[3], t.v[4], t.v[5]);
                viol++;
            }
        }
    }
    printf("{\"pairs\": %lld, \"partial\": %lld, \"answered\": %lld, \"declined_ref_ok\": %lld, \"declined_ref_error\": %lld, "
           "\"violations\": %lld, \"max_area_diff\": %.3g, \"max_grad_diff_vs_reference_chain\": %.3g, \"grads_checked_against_fp64_differences\": %lld, "
           "\"max_grad_diff_vs_fp64\": %.3g, \"answered_frac_by_mode\": [",
           total, partial, answered, declined_ok, declined_err, viol, max_da, max_dg, grad_vs_truth, max_dfd);
    for (int m = 0; m < kModes; m++) printf("%s%.4f", m ? ", " : "", mode_partial[m] ? (double)mode_answered[m] / mode_partial[m] : 0.0);
    printf("]");
    const char* names[5] = {"face_flag", "vertex_near_line", "edge_near_corner", "tiny_area", "full_area"};
    const int rm[3] = {0, 8, 9};
    const char* rn[3] = {"mixed", "cfg2_like", "cfg5_like"};
    for (int q = 0; q < 3; q++) {
        printf(", \"declined_by_reason_%s\": {", rn[q]);
        for (int b = 0; b < 5; b++) printf("%s\"%s\": %.5f", b ? ", " : "", names[b], mode_partial[rm[q]] ? (double)reasons[rm[q]][b] / mode_partial[rm[q]] : 0.0);
        printf("}");
    }
    printf("}\n");
    return viol ? 1 : 0;
}
