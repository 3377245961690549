// dm2_capi.cu — the extern "C" surface declared in include/dm2_renderer.h: argument validation
// (mirrors the AT_ERROR checks of reference render.cu:62-118, 247-303, 397-429 that can be done
// without torch shapes), workspace carving (reference state.h / renderer.cu:40-76, 499-507) and
// kernel sequencing (reference renderer.cu:78-269, 271-392, 509-674) — on the caller's stream,
// with a single stream wait for the num_rendered read-back and no cudaDeviceSynchronize.
#include <stdarg.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "dm2_common.cuh"

namespace dm2 {

static thread_local char g_err[1024] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch() { g_launches++; }

// ---- optional per-stage timing (CUDA events on the launching stream) -------------------------
// Shared by all threads of the process (one report); every access goes through g_prof_mu.
struct ProfEntry { const char* name; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfEntry> g_prof_pending;
static std::vector<cudaEvent_t> g_prof_pool;
static std::map<std::string, std::pair<double, long long>> g_prof_acc;
static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
int prof_begin(const char* name, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return -1;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfEntry p{name, prof_event(), prof_event()};
    cudaEventRecord(p.a, s);
    g_prof_pending.push_back(p);
    return (int)g_prof_pending.size() - 1;
}
void prof_end(int handle, cudaStream_t s) {
    if (handle < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (handle >= (int)g_prof_pending.size()) return;     // collected in between (scopes may nest)
    cudaEventRecord(g_prof_pending[handle].b, s);
}
static void prof_collect() {          // caller holds g_prof_mu
    for (auto& p : g_prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            auto& acc = g_prof_acc[p.name];
            acc.first += ms; acc.second += 1;
        }
        g_prof_pool.push_back(p.a); g_prof_pool.push_back(p.b);
    }
    g_prof_pending.clear();
}

static uint32_t higher_msb(uint32_t n) {          // reference renderer.cu:396-411
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

static void face_layout(long long BF, bool layered, DM2FaceLayout* L) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L->depths        = take((size_t)BF * 4);
    L->min_depths    = take((size_t)BF * 4);
    L->max_depths    = take((size_t)BF * 4);
    L->tiles_touched = take((size_t)BF * 4);
    L->rects         = take((size_t)BF * 8);
    L->record_stride = layered ? (size_t)kLRecWords * 4 : (size_t)kRecBytes;
    L->records       = take((size_t)BF * L->record_stride);
    L->total         = off + 256;
}

// tile_cursor area: u32[nt + 3] (cursors + the 12-byte header), padded to 16 bytes, then - used when the forward
// pass runs in rounds - u8[8 nt]: one "this footprint's compositing is finished" flag per (tile, 8x4 footprint), and
// u32[4 + 3 nt]: the work lists of a round (RoundWork)
static size_t cursor_words(size_t nt) { return (nt + 3 + 3) / 4 * 4; }
static size_t cursor_zero_bytes(size_t nt) { return cursor_words(nt) * 4 + nt * 8; }     // zeroed every frame
static size_t cursor_area_bytes(size_t nt) { return cursor_zero_bytes(nt) + (4 + 4 * nt) * 4; }   // + work lists, sorted_upto

static void image_layout(long long B, int H, int W, bool layered, DM2ImageLayout* L) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const long long tiles = (long long)cdiv(W, kTile) * cdiv(H, kTile);
    const size_t npix = (size_t)B * H * W;
    L->ranges = take((size_t)B * tiles * 8);
    L->tile_cursor = take(cursor_area_bytes((size_t)B * tiles));
    if (layered) {
        L->final_T = L->final_prev_T = L->n_contrib = 0;
        L->first_face = take(npix * 4);
        L->first_tet  = take(npix * 4);
    } else {
        L->final_T      = take(npix * 4);
        L->final_prev_T = take(npix * 4);
        L->n_contrib    = take(npix * 4);
        L->first_face = L->first_tet = 0;
    }
    L->total = off + 256;
}

static void binning_layout(long long R, DM2BinningLayout* L) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L->keys_unsorted = take((size_t)R * 8);
    L->vals_unsorted = take((size_t)R * 4);
    L->keys          = take((size_t)R * 8);
    L->vals          = take((size_t)R * 4);
    L->sort_temp     = take(0);                         // the pipeline's tile-segmented sort needs no scratch
    L->imask         = take((size_t)R);
    L->meta          = take((size_t)R * 8 * 16);
    L->total         = off + 256;
}

static FaceWs carve_face(void* ws, long long BF, bool layered) {
    DM2FaceLayout L; face_layout(BF, layered, &L);
    char* p = reinterpret_cast<char*>(ws);
    FaceWs f;
    f.depths = (float*)(p + L.depths); f.min_depths = (float*)(p + L.min_depths);
    f.max_depths = (float*)(p + L.max_depths); f.tiles_touched = (uint32_t*)(p + L.tiles_touched);
    f.rects = (uint2*)(p + L.rects); f.records = (float*)(p + L.records); f.record_stride = L.record_stride;
    return f;
}
static ImageWs carve_image(void* ws, long long B, int H, int W, bool layered) {
    DM2ImageLayout L; image_layout(B, H, W, layered, &L);
    char* p = reinterpret_cast<char*>(ws);
    ImageWs i;
    i.ranges = (uint2*)(p + L.ranges); i.tile_cursor = (uint32_t*)(p + L.tile_cursor); i.final_T = (float*)(p + L.final_T);
    const size_t nt = (size_t)B * cdiv(W, kTile) * cdiv(H, kTile);
    i.warp_done = (uint8_t*)(i.tile_cursor + cursor_words(nt));
    i.round_work = (uint32_t*)(i.warp_done + nt * 8);
    i.sorted_upto = i.round_work + 4 + 3 * nt;
    i.final_prev_T = (float*)(p + L.final_prev_T); i.n_contrib = (uint32_t*)(p + L.n_contrib);
    i.first_face = (int*)(p + L.first_face); i.first_tet = (int*)(p + L.first_tet);
    return i;
}
static BinWs carve_bin(void* ws, long long R) {
    DM2BinningLayout L; binning_layout(R, &L);
    char* p = reinterpret_cast<char*>(ws);
    BinWs b;
    b.keys_unsorted = (uint64_t*)(p + L.keys_unsorted); b.vals_unsorted = (uint32_t*)(p + L.vals_unsorted);
    b.keys = (uint64_t*)(p + L.keys); b.vals = (uint32_t*)(p + L.vals);
    b.imask = (uint8_t*)(p + L.imask);
    b.meta = p + L.meta;
    b.R = R;
    return b;
}

static int check_render_inputs(const DM2RenderInputs* in) {
    DM2_REQUIRE(in != nullptr, "inputs must not be null");
    DM2_REQUIRE(in->B >= 0 && in->P >= 0 && in->F >= 0 && in->H >= 0 && in->W >= 0, "negative dimension");
    DM2_REQUIRE(in->aa_temperature >= 0.0f && in->aa_temperature <= 1.0f, "aa_temperature must be in the range [0, 1]");
    DM2_REQUIRE(in->len_oarea_buffer >= 0, "len_oarea_buffer must be non-negative");
    DM2_REQUIRE((long long)in->B * in->F < (1ll << 31), "B*F must be < 2^31");
    DM2_REQUIRE((long long)in->B * in->H * in->W < (1ll << 31), "B*H*W must be < 2^31");
    DM2_REQUIRE(cdiv(in->W, kTile) < 65536 && cdiv(in->H, kTile) < 65536, "patch too large");
    return 0;
}

// Pinned read-back slot + event per (thread, device) for the 12-byte binning header.
struct Readback { uint32_t* host = nullptr; std::map<int, cudaEvent_t> ev; };
static thread_local Readback g_rb;
static int readback_slot(uint32_t** host, cudaEvent_t* ev) {
    if (!g_rb.host) DM2_CHECK(cudaHostAlloc((void**)&g_rb.host, 4 * sizeof(uint32_t), cudaHostAllocPortable));
    int d = 0;
    DM2_CHECK(cudaGetDevice(&d));
    auto it = g_rb.ev.find(d);
    if (it == g_rb.ev.end()) {
        cudaEvent_t e;
        DM2_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        it = g_rb.ev.emplace(d, e).first;
    }
    *host = g_rb.host; *ev = it->second;
    return 0;
}

// shared tail of the preprocess entry points: per-tile counts -> tile ranges, then the read-back of the
// binning header (num_rendered, input flags, longest segment) is ENQUEUED; wait_header() completes it.
// The caller zeroes iw.tile_cursor[0 .. num_tiles + 2] BEFORE its preprocess kernel (which may raise
// flags in hdr[1]).
static int count_and_scan(const FaceWs& fw, const ImageWs& iw, long long BF, int B, int F, int gx, int gy, cudaStream_t s) {
    const int num_tiles = B * gx * gy;
    if (launch_tile_count_scan(BF, F, gx, gy, num_tiles, fw.tiles_touched, fw.rects, iw.ranges, iw.tile_cursor, s))
        return 1;
    uint32_t* host; cudaEvent_t ev;
    if (readback_slot(&host, &ev)) return 1;
    DM2_CHECK(cudaMemcpyAsync(host, iw.tile_cursor + num_tiles, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    DM2_CHECK(cudaEventRecord(ev, s));
    return 0;
}
static int wait_header(int64_t* num_rendered, uint32_t* flags, int64_t* longest) {
    uint32_t* host; cudaEvent_t ev;
    if (readback_slot(&host, &ev)) return 1;
    DM2_CHECK(cudaEventSynchronize(ev));
    *num_rendered = (int64_t)host[0];
    *flags = host[1];
    if (longest) *longest = (int64_t)host[2];
    return 0;
}
static int check_flags(uint32_t flags) {
    // same messages as the reference's host-side asserts (__init__.py:283-284)
    DM2_REQUIRE(!(flags & kFlagPatchX), "Some b_patch_max_x exceed self.width");
    DM2_REQUIRE(!(flags & kFlagPatchY), "Some b_patch_max_y exceed self.height");
    DM2_REQUIRE(!(flags & kFlagPatchNeg), "negative patch_min is not supported by the fused route (set renderer.fused = False)");
    DM2_REQUIRE(!(flags & kFlagVertexId), "faces reference a vertex id outside [0, P)");
    return 0;
}
static int zero_cursor(const ImageWs& iw, int B, int gx, int gy, cudaStream_t s) {
    DM2_CHECK(cudaMemsetAsync(iw.tile_cursor, 0, cursor_zero_bytes((size_t)B * gx * gy), s));
    return 0;
}

// nothing to bin: every tile range is (0, 0)
static int zero_ranges(void* image_ws, int B, int H, int W, bool layered, cudaStream_t s) {
    ImageWs iw = carve_image(image_ws, B, H, W, layered);
    const size_t nt = (size_t)B * cdiv(W, kTile) * cdiv(H, kTile);
    DM2_CHECK(cudaMemsetAsync(iw.ranges, 0, nt * sizeof(uint2), s));
    DM2_CHECK(cudaMemsetAsync(iw.tile_cursor, 0, cursor_zero_bytes(nt), s));
    return 0;
}

// shared middle of both pipelines: scatter into the tile segments + per-segment sort.  capacity: instances
// the binning workspace was laid out for; longest: longest tile segment (-1 = unknown: every sort tier runs).
static int bin_faces(long long BF, int B, int F, int gx, int gy, const float* key_depths, const FaceWs& fw,
                     const BinWs& bw, const ImageWs& iw, long long longest, const LazySort& lazy, cudaStream_t s) {
    if (bw.R <= 0) return 0;
    return launch_tile_sort(BF, F, gx, gy, B * gx * gy, key_depths, fw.tiles_touched, fw.rects, iw.ranges,
                            iw.tile_cursor, bw.keys_unsorted, bw.vals_unsorted, bw.keys, bw.vals, bw.R, longest, lazy, s);
}

}  // namespace dm2

namespace dm2 {
static std::atomic<int> g_exact_clip{-1};          // -1: not set yet (DM2_EXACT_CLIP in the environment decides)
bool exact_clip_requested() {
    int v = g_exact_clip.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("DM2_EXACT_CLIP");
        v = (e && atoi(e) != 0) ? 1 : 0;
        int expected = -1;
        g_exact_clip.compare_exchange_strong(expected, v);
        v = g_exact_clip.load(std::memory_order_relaxed);
    }
    return v != 0;
}
}  // namespace dm2

using namespace dm2;

extern "C" {

void dm2_set_exact_clip(int32_t on) { g_exact_clip.store(on != 0 ? 1 : 0, std::memory_order_relaxed); }
int32_t dm2_get_exact_clip(void) { return exact_clip_requested() ? 1 : 0; }
const char* dm2_last_error(void) { return g_err; }
// Measured (tools/perf_configs.py --rounds-ab, profiles/r02_forward_rounds.md): doubling rounds beat equal ones (the
// launches of rounds in which every tile has finished are not free); 512 / 6144 keeps the 8-view 512^2 micro-batches
// (longest lists 3 800 .. 4 400 depending on the cameras, almost no early termination) on the single pass, where
// rounds cost 3-8 % - and, since the lazy segment sort, a second sort of every tile that runs to the end of its list.
// Known points: lists of <= 4 400 without early termination lose, lists of 11 400 with it (10 M triangles) gain 30 %.
static long long kRoundLen = 512, kRoundThreshold = 6144;
static void round_knobs() {             // DM2_ROUND_LEN (multiple of 256) / DM2_ROUND_THRESHOLD: measurement knobs
    static std::once_flag once;
    std::call_once(once, [] {
        if (const char* e = getenv("DM2_ROUND_LEN")) kRoundLen = std::max(256ll, atoll(e) / 256 * 256);
        if (const char* e = getenv("DM2_ROUND_THRESHOLD")) kRoundThreshold = atoll(e);
    });
}
int dm2_abi_version(void) { return DM2_ABI_VERSION; }
void dm2_set_forward_rounds(int64_t round_len, int64_t threshold) {
    round_knobs();                      // (the environment is read first so that it cannot override this call later)
    kRoundLen = std::max<long long>(256, round_len / 256 * 256);
    kRoundThreshold = threshold;
}
uint32_t dm2_higher_msb(uint32_t n) { return higher_msb(n); }
void dm2_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_collect();
    g_prof_on = on != 0;
}
void dm2_profile_reset(void) { std::lock_guard<std::mutex> lk(g_prof_mu); prof_collect(); g_prof_acc.clear(); }
// Writes "name total_ms count\n" lines into buf; returns the number of bytes needed.
int64_t dm2_profile_report(char* buf, int64_t cap) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_collect();
    std::string out;
    char line[256];
    for (auto& kv : g_prof_acc) {
        snprintf(line, sizeof(line), "%s %.6f %lld\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && cap > 0) { size_t n = std::min((size_t)cap - 1, out.size()); memcpy(buf, out.data(), n); buf[n] = 0; }
    return (int64_t)out.size() + 1;
}
int64_t dm2_launch_count(int32_t reset) { long long v = g_launches; if (reset) g_launches = 0; return v; }

int dm2_face_layout(int32_t B, int32_t F, int32_t layered, DM2FaceLayout* out) {
    DM2_REQUIRE(out && B >= 0 && F >= 0, "dm2_face_layout: bad arguments");
    face_layout((long long)B * F, layered != 0, out);
    return 0;
}
int dm2_image_layout(int32_t B, int32_t H, int32_t W, int32_t layered, DM2ImageLayout* out) {
    DM2_REQUIRE(out && B >= 0 && H >= 0 && W >= 0, "dm2_image_layout: bad arguments");
    image_layout(B, H, W, layered != 0, out);
    return 0;
}
int dm2_binning_layout(int64_t num_rendered, DM2BinningLayout* out) {
    DM2_REQUIRE(out && num_rendered >= 0 && num_rendered < (1ll << 31), "dm2_binning_layout: bad arguments");
    binning_layout(num_rendered, out);
    return 0;
}

int dm2_render_preprocess(const DM2RenderInputs* in, void* face_ws, void* image_ws, int64_t* num_rendered_host,
                          int64_t* longest_host, void* stream) {
    if (check_render_inputs(in)) return 2;
    DM2_REQUIRE(face_ws && image_ws && num_rendered_host, "dm2_render_preprocess: null workspace");
    cudaStream_t s = (cudaStream_t)stream;
    const long long BF = (long long)in->B * in->F;
    *num_rendered_host = 0;
    if (longest_host) *longest_host = 0;
    if (in->P == 0 || BF == 0) return zero_ranges(image_ws, in->B, in->H, in->W, false, s);    // reference render.cu:149
    FaceWs fw = carve_face(face_ws, BF, false);
    ImageWs iw = carve_image(image_ws, in->B, in->H, in->W, false);
    const int gx = cdiv(in->W, kTile), gy = cdiv(in->H, kTile);
    if (zero_cursor(iw, in->B, gx, gy, s)) return 1;
    if (launch_render_preprocess(*in, fw, iw.tile_cursor + (size_t)in->B * gx * gy + 1, s)) return 1;
    if (count_and_scan(fw, iw, BF, in->B, in->F, gx, gy, s)) return 1;
    uint32_t flags = 0;
    if (wait_header(num_rendered_host, &flags, longest_host)) return 1;
    return check_flags(flags);
}

// everything behind the scan: scatter + segment sort, classification, forward compositor.
// Deep scenes (a tile segment longer than kRoundThreshold instances) run classification + compositing in ROUNDS of
// kRoundLen, then doubling, list entries per tile: the forward pass terminates long before the end of such lists (T < 1e-4), and
// classifying entries nobody composites is the most expensive stage there (10 M triangles @2048^2: 2.9 of 10.8 ms).
// A round classifies entries [k C, (k + 1) C) of every tile that still has an unfinished footprint, then the
// compositor resumes from the per-pixel state it left in the output / image-state buffers.
static long long round_end(long long k) {        // list position behind round k
    return kRoundLen << std::min(k, 40ll);
}
// longest: the segment length the launch has to cover (a hint with slack in the pipelined path); basis: the length
// the rounds-or-not decision is taken on (without slack, so that it does not depend on the calling path)
static long long rounds_for(long long longest, long long basis) {
    round_knobs();
    if (longest < 0 || basis <= kRoundThreshold) return 1;             // unknown / shallow: one pass over everything
    long long n = 1;
    while (round_end(n - 1) < longest) n++;
    return n;
}
// longest tile segment a forward_tail launched for (longest, basis) handles correctly
static long long tail_longest_cap(long long longest, long long basis, bool sort_too) {
    long long cap = sort_too ? sort_tiers_longest_cap(longest) : 0xFFFFFFFFll;
    const long long rounds = rounds_for(longest, basis);
    if (rounds > 1) cap = std::min(cap, round_end(rounds - 1));
    return cap;
}
// Lazy segment sort (dm2_tilesort.cu): a frame that runs in rounds sorts only the first two rounds' worth of every
// tile list at once; tiles that run deeper are sorted to the end before the round that needs it.  DM2_LAZY_SORT=0: off.
static bool lazy_sort_enabled() {
    static std::once_flag once;
    static bool on = true;
    std::call_once(once, [] { if (const char* e = getenv("DM2_LAZY_SORT")) on = atoi(e) != 0; });
    return on;
}
static LazySort lazy_for(const ImageWs& iw, long long rounds) {
    return LazySort{iw.sorted_upto, rounds > 1 && lazy_sort_enabled() ? (uint32_t)round_end(1) : 0xFFFFFFFFu};
}
static int forward_tail(const DM2RenderInputs* in, const FaceWs& fw, BinWs bw, const ImageWs& iw, long long longest,
                        long long basis, float* out_color, float* out_depth, void* shade_cache,
                        long long shade_cache_bytes, bool sort_too, cudaStream_t s) {
    // kernels that read the SORTED lists return at once if a segment is longer than what was enqueued covers
    bw.longest_cap = tail_longest_cap(longest, basis, sort_too);
    const long long BF = (long long)in->B * in->F;
    const int gx = cdiv(in->W, kTile), gy = cdiv(in->H, kTile);
    const TileMaskArgs tm{bw.vals, fw.records, in->patch_min, bw.imask, bw.meta, in->F, gx, gy, in->aa_temperature > 0.0f ? 1 : 0};
    const long long rounds = rounds_for(longest, basis);
    const LazySort lazy = lazy_for(iw, rounds);
    if (sort_too && bin_faces(BF, in->B, in->F, gx, gy, fw.depths, fw, bw, iw, longest, lazy, s)) return 1;
    const uint32_t* hdr = iw.tile_cursor + (size_t)in->B * gx * gy;
    if (rounds == 1) {
        if (launch_classify(bw.keys, bw.R, bw.longest_cap, hdr, tm, s)) return 1;
        return launch_raster_forward(*in, fw, bw, iw, out_color, out_depth, shade_cache, shade_cache ? shade_cache_bytes : 0,
                                     0, 0, RoundWork{nullptr, nullptr, nullptr}, s);
    }
    const long long nt = (long long)in->B * gx * gy;
    const RoundWork work{iw.round_work, iw.round_work + 4, iw.round_work + 4 + nt};
    const RoundSort rsort{bw.keys_unsorted, bw.vals_unsorted, bw.keys, bw.vals, iw.tile_cursor, lazy,
                          1 | (longest < 0 ? 6 : (longest > 2048 ? 2 : 0) | (longest > 8192 ? 4 : 0))};
    for (long long k = 0; k < rounds; k++) {
        const int rb = k ? (int)round_end(k - 1) : 0, re = (int)round_end(k);
        if (launch_classify_round(bw.keys, bw.R, bw.longest_cap, hdr, tm, iw.ranges, iw.warp_done, (int)nt,
                                  rb, re - rb, work, rsort, s)) return 1;
        if (launch_raster_forward(*in, fw, bw, iw, out_color, out_depth, shade_cache, shade_cache ? shade_cache_bytes : 0,
                                  rb, re, work, s)) return 1;
    }
    return 0;
}

int dm2_render_forward(const DM2RenderInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                       int64_t capacity, int64_t longest, float* out_color, float* out_depth, void* shade_cache,
                       int64_t shade_cache_bytes, void* stream) {
    if (check_render_inputs(in)) return 2;
    DM2_REQUIRE(face_ws && binning_ws && image_ws && out_color && out_depth, "dm2_render_forward: null buffer");
    DM2_REQUIRE(capacity >= 0 && capacity < (1ll << 31), "dm2_render_forward: bad num_rendered");
    DM2_REQUIRE(shade_cache_bytes >= 0, "dm2_render_forward: bad shade_cache_bytes");
    cudaStream_t s = (cudaStream_t)stream;
    const long long BF = (long long)in->B * in->F;
    FaceWs fw = carve_face(face_ws, BF, false);
    BinWs bw = carve_bin(binning_ws, capacity);
    ImageWs iw = carve_image(image_ws, in->B, in->H, in->W, false);
    if (in->P == 0) {                                          // reference render.cu:128-129, 149: zeros
        const size_t npix = (size_t)in->B * in->H * in->W;
        DM2_CHECK(cudaMemsetAsync(out_color, 0, npix * 3 * sizeof(float), s));
        DM2_CHECK(cudaMemsetAsync(out_depth, 0, npix * sizeof(float), s));
        return 0;
    }
    return forward_tail(in, fw, bw, iw, longest, longest, out_color, out_depth, shade_cache, shade_cache_bytes, true, s);
}

int dm2_render_backward(const DM2RenderInputs* in, const void* face_ws, const void* binning_ws,
                        const void* image_ws, int64_t capacity, const float* dL_dout_color,
                        const float* dL_dout_depth, const DM2RenderGrads* grads, void* grad_ws,
                        const void* shade_cache, int64_t shade_cache_bytes, void* stream) {
    if (check_render_inputs(in)) return 2;
    DM2_REQUIRE(face_ws && binning_ws && image_ws && dL_dout_color && dL_dout_depth && grads && grad_ws,
                "dm2_render_backward: null buffer");
    DM2_REQUIRE(grads->dL_dverts && grads->dL_dverts_color && grads->dL_dfaces_opacity && grads->dL_dverts_ndc &&
                grads->dL_dfaces_intense && grads->dL_daa_face_verts, "dm2_render_backward: null gradient buffer");
    if (in->F == 0 || capacity == 0) return 0;                 // reference render.cu:320
    cudaStream_t s = (cudaStream_t)stream;
    const long long BF = (long long)in->B * in->F;
    FaceWs fw = carve_face(const_cast<void*>(face_ws), BF, false);
    BinWs bw = carve_bin(const_cast<void*>(binning_ws), capacity);
    ImageWs iw = carve_image(const_cast<void*>(image_ws), in->B, in->H, in->W, false);
    if (launch_raster_backward(*in, fw, bw, iw, dL_dout_color, dL_dout_depth, grad_ws, shade_cache,
                               shade_cache ? shade_cache_bytes : 0, s)) return 1;
    return launch_grad_scatter(*in, fw, *grads, grad_ws, s);
}

// ---- fused host prep path -------------------------------------------------------------------
static int check_fused_inputs(const DM2FusedInputs* in) {
    DM2_REQUIRE(in != nullptr, "inputs must not be null");
    DM2_REQUIRE(in->B >= 0 && in->P >= 0 && in->F >= 0 && in->H >= 0 && in->W >= 0, "negative dimension");
    DM2_REQUIRE(in->image_h >= in->H && in->image_w >= in->W, "patch larger than the frame");
    DM2_REQUIRE(in->B == 0 || (in->mv && in->proj), "mv / proj must not be null");
    DM2_REQUIRE(in->aa_temperature >= 0.0f && in->aa_temperature <= 1.0f, "aa_temperature must be in the range [0, 1]");
    DM2_REQUIRE(in->ray_recipe >= 0 && in->ray_recipe <= DM2_RAY_RECIPES, "ray_recipe out of range");
    DM2_REQUIRE((long long)in->B * in->F < (1ll << 31), "B*F must be < 2^31");
    DM2_REQUIRE((long long)in->B * in->H * in->W < (1ll << 31), "B*H*W must be < 2^31");
    DM2_REQUIRE(cdiv(in->W, kTile) < 65536 && cdiv(in->H, kTile) < 65536, "patch too large");
    return 0;
}
static DM2RenderInputs as_render_inputs(const DM2FusedInputs& f) {
    DM2RenderInputs r;
    memset(&r, 0, sizeof(r));
    r.B = f.B; r.P = f.P; r.F = f.F; r.H = f.H; r.W = f.W;
    r.aa_temperature = f.aa_temperature;
    r.background = f.background; r.patch_min = f.patch_min; r.verts = f.verts; r.faces = f.faces;
    r.verts_color = f.verts_color; r.faces_opacity = f.faces_opacity; r.faces_intense = f.faces_intense;
    r.image_ray_o = f.ray_o; r.image_ray_d = f.ray_d; r.ray_view = f.ray_view; r.ray_h = f.image_h; r.ray_w = f.image_w;
    r.depth_unit = f.depth_unit;
    r.ray_recipe = f.ray_params ? f.ray_recipe : 0;
    r.ray_params = f.ray_params;
    r.ray_v = f.ray_v;
    return r;
}

static int fused_front(const DM2FusedInputs* in, void* face_ws, void* image_ws, cudaStream_t s) {
    const long long BF = (long long)in->B * in->F;
    FaceWs fw = carve_face(face_ws, BF, false);
    ImageWs iw = carve_image(image_ws, in->B, in->H, in->W, false);
    const int gx = cdiv(in->W, kTile), gy = cdiv(in->H, kTile);
    if (zero_cursor(iw, in->B, gx, gy, s)) return 1;
    if (launch_fused_preprocess(*in, fw, iw.tile_cursor + (size_t)in->B * gx * gy + 1, s)) return 1;
    return count_and_scan(fw, iw, BF, in->B, in->F, gx, gy, s);
}

int dm2_fused_preprocess(const DM2FusedInputs* in, void* face_ws, void* image_ws, int64_t* num_rendered_host,
                         int64_t* longest_host, void* stream) {
    if (check_fused_inputs(in)) return 2;
    DM2_REQUIRE(face_ws && image_ws && num_rendered_host, "dm2_fused_preprocess: null workspace");
    cudaStream_t s = (cudaStream_t)stream;
    const long long BF = (long long)in->B * in->F;
    *num_rendered_host = 0;
    if (longest_host) *longest_host = 0;
    if (in->P == 0 || BF == 0) return zero_ranges(image_ws, in->B, in->H, in->W, false, s);
    if (fused_front(in, face_ws, image_ws, s)) return 1;
    uint32_t flags = 0;
    if (wait_header(num_rendered_host, &flags, longest_host)) return 1;
    return check_flags(flags);
}

int dm2_fused_forward(const DM2FusedInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                      int64_t capacity, int64_t longest, float* out_color, float* out_depth, void* shade_cache,
                      int64_t shade_cache_bytes, void* stream) {
    if (check_fused_inputs(in)) return 2;
    DM2_REQUIRE(in->ray_view && ((in->ray_o && in->ray_d) || (in->ray_params && in->ray_recipe > 0)),
                "dm2_fused_forward: null ray buffers");
    const DM2RenderInputs r = as_render_inputs(*in);
    return dm2_render_forward(&r, face_ws, binning_ws, image_ws, capacity, longest, out_color, out_depth, shade_cache,
                              shade_cache_bytes, stream);
}

// which sort tiers launch_tile_sort runs for a given `longest` (hint or exact): bit 1 = tier 1, bit 2 = tier 2
static int tiers_for(long long longest) {
    if (longest < 0) return 6;
    return (longest > 2048 ? 2 : 0) | (longest > 8192 ? 4 : 0);
}

int dm2_fused_forward_pipelined(const DM2FusedInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                                int64_t capacity, int64_t longest_hint, float* out_color, float* out_depth,
                                void* shade_cache, int64_t shade_cache_bytes, int64_t* num_rendered_host,
                                int64_t* longest_host, int32_t* fits_host, void* stream) {
    if (check_fused_inputs(in)) return 2;
    DM2_REQUIRE(in->ray_view && ((in->ray_o && in->ray_d) || (in->ray_params && in->ray_recipe > 0)),
                "dm2_fused_forward_pipelined: null ray buffers");
    DM2_REQUIRE(face_ws && binning_ws && image_ws && out_color && out_depth && num_rendered_host && longest_host && fits_host,
                "dm2_fused_forward_pipelined: null buffer");
    DM2_REQUIRE(capacity > 0 && capacity < (1ll << 31), "dm2_fused_forward_pipelined: bad capacity");
    cudaStream_t s = (cudaStream_t)stream;
    const long long BF = (long long)in->B * in->F;
    *num_rendered_host = 0; *longest_host = 0; *fits_host = 1;
    const DM2RenderInputs r = as_render_inputs(*in);
    if (in->P == 0 || BF == 0) {
        if (zero_ranges(image_ws, in->B, in->H, in->W, false, s)) return 1;
        return dm2_render_forward(&r, face_ws, binning_ws, image_ws, 0, 0, out_color, out_depth, nullptr, 0, stream);
    }
    // 1. preprocess + per-tile counts + scan; the 12-byte header read-back is enqueued behind them
    if (fused_front(in, face_ws, image_ws, s)) return 1;
    // 2. the whole tail is enqueued on the workspace the caller sized from the previous frame: every kernel
    //    checks the header on the device and returns at once if the frame does not fit
    FaceWs fw = carve_face(face_ws, BF, false);
    BinWs bw = carve_bin(binning_ws, capacity);
    ImageWs iw = carve_image(image_ws, in->B, in->H, in->W, false);
    // (the previous frame's longest segment + 25 %: a little growth must not cost a repair)
    const long long covered = longest_hint < 0 ? -1 : longest_hint + longest_hint / 4;
    if (forward_tail(&r, fw, bw, iw, covered, longest_hint, out_color, out_depth, shade_cache, shade_cache_bytes, true, s)) return 1;
    // 3. only now wait for the header (its event completed long before the raster kernel ends)
    uint32_t flags = 0;
    if (wait_header(num_rendered_host, &flags, longest_host)) return 1;
    if (check_flags(flags)) return 2;
    if (*num_rendered_host > capacity) { *fits_host = 0; return 0; }      // caller re-runs dm2_fused_forward
    const int missing = tiers_for(*longest_host) & ~tiers_for(covered);
    if (missing || *longest_host > tail_longest_cap(covered, longest_hint, true)) {
        // a segment longer than the hinted sort tier / the enqueued rounds: the scatter is valid (and must not run
        // twice) - sort the long segments now and redo classification + compositing on the same workspace
        const int gx = cdiv(in->W, kTile), gy = cdiv(in->H, kTile);
        if (missing && launch_tile_sort_tiers(in->B * gx * gy, iw.ranges, iw.tile_cursor, bw.keys_unsorted, bw.vals_unsorted,
                                              bw.keys, bw.vals, bw.R, missing,
                                              lazy_for(iw, rounds_for(*longest_host, *longest_host)), s)) return 1;
        DM2_CHECK(cudaMemsetAsync(iw.warp_done, 0, (size_t)in->B * gx * gy * 8, s));
        if (forward_tail(&r, fw, bw, iw, *longest_host, *longest_host, out_color, out_depth, shade_cache, shade_cache_bytes, false, s))
            return 1;
    }
    return 0;
}

int dm2_fused_backward(const DM2FusedInputs* in, const void* face_ws, const void* binning_ws,
                       const void* image_ws, int64_t capacity, const float* dL_dout_color,
                       const float* dL_dout_depth, const DM2FusedGrads* grads, void* grad_ws,
                       const void* shade_cache, int64_t shade_cache_bytes, void* stream) {
    if (check_fused_inputs(in)) return 2;
    DM2_REQUIRE(face_ws && binning_ws && image_ws && dL_dout_color && dL_dout_depth && grads && grad_ws,
                "dm2_fused_backward: null buffer");
    DM2_REQUIRE(grads->dL_dverts && grads->dL_dverts_color && grads->dL_dfaces_opacity && grads->dL_dfaces_intense,
                "dm2_fused_backward: null gradient buffer");
    if (in->F == 0 || capacity == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const long long BF = (long long)in->B * in->F;
    const DM2RenderInputs r = as_render_inputs(*in);
    FaceWs fw = carve_face(const_cast<void*>(face_ws), BF, false);
    BinWs bw = carve_bin(const_cast<void*>(binning_ws), capacity);
    ImageWs iw = carve_image(const_cast<void*>(image_ws), in->B, in->H, in->W, false);
    if (launch_raster_backward(r, fw, bw, iw, dL_dout_color, dL_dout_depth, grad_ws, shade_cache,
                               shade_cache ? shade_cache_bytes : 0, s)) return 1;
    return launch_fused_grad_scatter(*in, fw, *grads, grad_ws, s);
}

int dm2_generate_rays(const float* ray_params, int32_t V, int32_t H, int32_t W, int32_t recipe, float* out_o,
                      float* out_d, void* stream) {
    DM2_REQUIRE(ray_params && out_o && out_d && V >= 0 && H >= 0 && W >= 0, "dm2_generate_rays: bad arguments");
    DM2_REQUIRE(recipe >= 1 && recipe <= DM2_RAY_RECIPES, "dm2_generate_rays: recipe out of range");
    DM2_REQUIRE((long long)V * H * W < (1ll << 31), "dm2_generate_rays: V*H*W must be < 2^31");
    return launch_generate_rays(ray_params, V, H, W, recipe, out_o, out_d, (cudaStream_t)stream);
}

int dm2_peer_allreduce_sum(const void* const* peer_ptrs, void* multicast_ptr, int32_t world, int32_t rank,
                           int64_t n_floats, void* stream) {
    return launch_peer_allreduce(peer_ptrs, multicast_ptr, nullptr, 0u, world, rank, n_floats, (cudaStream_t)stream);
}
int dm2_peer_allreduce_sum_fused(const void* const* peer_ptrs, void* multicast_ptr, const void* const* peer_flag_ptrs,
                                 int32_t world, int32_t rank, int64_t n_floats, uint32_t epoch, void* stream) {
    DM2_REQUIRE(peer_flag_ptrs != nullptr, "dm2_peer_allreduce_sum_fused: null flag pointer table");
    return launch_peer_allreduce(peer_ptrs, multicast_ptr, peer_flag_ptrs, epoch, world, rank, n_floats, (cudaStream_t)stream);
}

int dm2_backward_workspace_bytes(int32_t B, int32_t F, size_t* bytes) {
    DM2_REQUIRE(bytes && B >= 0 && F >= 0, "dm2_backward_workspace_bytes: bad arguments");
    *bytes = backward_ws_bytes((long long)B * F);
    return 0;
}

static int check_layer_inputs(const DM2LayerInputs* in) {
    DM2_REQUIRE(in != nullptr, "inputs must not be null");
    DM2_REQUIRE(in->B >= 0 && in->P >= 0 && in->F >= 0 && in->T >= 0 && in->H >= 0 && in->W >= 0, "negative dimension");
    DM2_REQUIRE(in->num_layers >= 0, "num_layers must be non-negative");
    DM2_REQUIRE(in->ray_recipe >= 0 && in->ray_recipe <= DM2_RAY_RECIPES, "ray_recipe out of range");
    DM2_REQUIRE(in->ray_recipe == 0 || (in->ray_params && in->ray_view), "ray_recipe needs ray_params and ray_view");
    DM2_REQUIRE((long long)in->B * in->F < (1ll << 31), "B*F must be < 2^31");
    DM2_REQUIRE((long long)in->B * in->H * in->W < (1ll << 31), "B*H*W must be < 2^31");
    return 0;
}

static int layers_preprocess(const DM2LayerInputs* in, void* face_ws, void* image_ws, void* geo_ws,
                             int64_t* num_rendered_host, int64_t* longest_host, void* stream);
int dm2_layers_preprocess(const DM2LayerInputs* in, void* face_ws, void* image_ws, int64_t* num_rendered_host,
                          int64_t* longest_host, void* stream) {
    return layers_preprocess(in, face_ws, image_ws, nullptr, num_rendered_host, longest_host, stream);
}
int dm2_layers_preprocess_with_geometry(const DM2LayerInputs* in, void* face_ws, void* image_ws, void* geo_ws,
                                        int64_t* num_rendered_host, int64_t* longest_host, void* stream) {
    DM2_REQUIRE(geo_ws, "dm2_layers_preprocess_with_geometry: null geometry workspace");
    return layers_preprocess(in, face_ws, image_ws, geo_ws, num_rendered_host, longest_host, stream);
}
static int layers_preprocess(const DM2LayerInputs* in, void* face_ws, void* image_ws, void* geo_ws,
                             int64_t* num_rendered_host, int64_t* longest_host, void* stream) {
    if (check_layer_inputs(in)) return 2;
    DM2_REQUIRE(face_ws && image_ws && num_rendered_host, "dm2_layers_preprocess: null workspace");
    cudaStream_t s = (cudaStream_t)stream;
    const long long BF = (long long)in->B * in->F;
    *num_rendered_host = 0;
    if (longest_host) *longest_host = 0;
    if (BF == 0) {
        // (no faces: nothing is binned, but a caller of the *_prepared entry point still finds its tet records built)
        if (geo_ws && launch_layers_geometry(*in, geo_ws, true, s)) return 1;
        return zero_ranges(image_ws, in->B, in->H, in->W, true, s);
    }
    FaceWs fw = carve_face(face_ws, BF, true);
    ImageWs iw = carve_image(image_ws, in->B, in->H, in->W, true);
    const int gx = cdiv(in->W, kTile), gy = cdiv(in->H, kTile);
    if (zero_cursor(iw, in->B, gx, gy, s)) return 1;
    if (launch_layers_preprocess(*in, fw, iw.tile_cursor + (size_t)in->B * gx * gy + 1, s)) return 1;
    if (count_and_scan(fw, iw, BF, in->B, in->F, gx, gy, s)) return 1;
    // the geometry / link records do not depend on the binning: they run while the host waits for the header
    if (geo_ws && launch_layers_geometry(*in, geo_ws, true, s)) return 1;
    uint32_t flags = 0;
    if (wait_header(num_rendered_host, &flags, longest_host)) return 1;
    return check_flags(flags);
}

int dm2_layers_geo_bytes(int32_t F, int32_t T, size_t* bytes) {
    DM2_REQUIRE(bytes && F >= 0 && T >= 0, "dm2_layers_geo_bytes: bad arguments");
    *bytes = layers_geo_bytes(F, T);
    return 0;
}

static int generate_render_layers(const DM2LayerInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                                  void* geo_ws, int64_t num_rendered, int64_t longest, int32_t* render_layers,
                                  int32_t* render_layers_cnt, bool geometry_ready, void* stream);
int dm2_generate_render_layers(const DM2LayerInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                               void* geo_ws, int64_t num_rendered, int64_t longest, int32_t* render_layers,
                               int32_t* render_layers_cnt, void* stream) {
    return generate_render_layers(in, face_ws, binning_ws, image_ws, geo_ws, num_rendered, longest, render_layers,
                                  render_layers_cnt, false, stream);
}
int dm2_generate_render_layers_prepared(const DM2LayerInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                                        void* geo_ws, int64_t num_rendered, int64_t longest, int32_t* render_layers,
                                        int32_t* render_layers_cnt, void* stream) {
    return generate_render_layers(in, face_ws, binning_ws, image_ws, geo_ws, num_rendered, longest, render_layers,
                                  render_layers_cnt, true, stream);
}
static int generate_render_layers(const DM2LayerInputs* in, void* face_ws, void* binning_ws, void* image_ws,
                                  void* geo_ws, int64_t num_rendered, int64_t longest, int32_t* render_layers,
                                  int32_t* render_layers_cnt, bool geometry_ready, void* stream) {
    if (check_layer_inputs(in)) return 2;
    DM2_REQUIRE(face_ws && binning_ws && image_ws && geo_ws && render_layers_cnt, "dm2_generate_render_layers: null buffer");
    DM2_REQUIRE(render_layers || in->num_layers == 0, "dm2_generate_render_layers: null render_layers");
    DM2_REQUIRE(num_rendered >= 0 && num_rendered < (1ll << 31), "dm2_generate_render_layers: bad num_rendered");
    cudaStream_t s = (cudaStream_t)stream;
    const long long BF = (long long)in->B * in->F;
    const int gx = cdiv(in->W, kTile), gy = cdiv(in->H, kTile);
    FaceWs fw = carve_face(face_ws, BF, true);
    BinWs bw = carve_bin(binning_ws, num_rendered);
    ImageWs iw = carve_image(image_ws, in->B, in->H, in->W, true);
    // key = min depth (reference renderer.cu:603)
    if (bin_faces(BF, in->B, in->F, gx, gy, fw.min_depths, fw, bw, iw, longest, LazySort{nullptr, 0xFFFFFFFFu}, s)) return 1;
    return launch_layers(*in, fw, bw, iw, geo_ws, render_layers, render_layers_cnt, false, geometry_ready, s);
}

int dm2_regenerate_render_layers(const DM2LayerInputs* in, void* image_ws, void* geo_ws, int32_t* render_layers,
                                 int32_t* render_layers_cnt, void* stream) {
    if (check_layer_inputs(in)) return 2;
    DM2_REQUIRE(image_ws && geo_ws && render_layers_cnt, "dm2_regenerate_render_layers: null buffer");
    DM2_REQUIRE(render_layers || in->num_layers == 0, "dm2_regenerate_render_layers: null render_layers");
    ImageWs iw = carve_image(image_ws, in->B, in->H, in->W, true);
    FaceWs fw{};
    BinWs bw{};
    return launch_layers(*in, fw, bw, iw, geo_ws, render_layers, render_layers_cnt, true, false, (cudaStream_t)stream);
}

int dm2_layers_composite_forward(int32_t B, int32_t H, int32_t W, int32_t L, int32_t F, const int32_t* render_layers,
                                 const int32_t* render_layers_cnt, const float* faces_color, const float* faces_opacity,
                                 const float* faces_intense, const float* background, float* out_color, float* out_alpha,
                                 void* stream) {
    DM2_REQUIRE(B >= 0 && H >= 0 && W >= 0 && L >= 0 && F >= 0, "dm2_layers_composite_forward: negative dimension");
    DM2_REQUIRE((long long)B * H * W < (1ll << 31), "B*H*W must be < 2^31");
    if ((long long)B * H * W == 0) return 0;
    DM2_REQUIRE(render_layers_cnt && (render_layers || L == 0) && background && out_color && out_alpha &&
                (F == 0 || (faces_color && faces_opacity)), "dm2_layers_composite_forward: null buffer");
    return launch_layers_composite_forward(B, H, W, L, F, render_layers, render_layers_cnt, faces_color, faces_opacity,
                                           faces_intense, background, out_color, out_alpha, (cudaStream_t)stream);
}
int dm2_layers_composite_backward(int32_t B, int32_t H, int32_t W, int32_t L, int32_t F, const int32_t* render_layers,
                                  const int32_t* render_layers_cnt, const float* faces_color, const float* faces_opacity,
                                  const float* faces_intense, const float* background, const float* dL_dout_color,
                                  const float* dL_dout_alpha, void* grad_ws, float* dL_dfaces_color,
                                  float* dL_dfaces_opacity, float* dL_dfaces_intense, void* stream) {
    DM2_REQUIRE(B >= 0 && H >= 0 && W >= 0 && L >= 0 && F >= 0, "dm2_layers_composite_backward: negative dimension");
    DM2_REQUIRE((long long)B * H * W < (1ll << 31), "B*H*W must be < 2^31");
    if (F == 0) return 0;
    DM2_REQUIRE(render_layers_cnt && (render_layers || L == 0) && background && dL_dout_color && grad_ws && faces_color &&
                faces_opacity && dL_dfaces_color && dL_dfaces_opacity, "dm2_layers_composite_backward: null buffer");
    DM2_REQUIRE((faces_intense == nullptr) == (dL_dfaces_intense == nullptr) || dL_dfaces_intense == nullptr,
                "dm2_layers_composite_backward: dL_dfaces_intense without faces_intense");
    return launch_layers_composite_backward(B, H, W, L, F, render_layers, render_layers_cnt, faces_color, faces_opacity,
                                            faces_intense, background, dL_dout_color, dL_dout_alpha, grad_ws, dL_dfaces_color,
                                            dL_dfaces_opacity, dL_dfaces_intense, (cudaStream_t)stream);
}

int dm2_sort_temp_bytes(int64_t n, size_t* bytes) {
    DM2_REQUIRE(bytes && n >= 0, "dm2_sort_temp_bytes: bad arguments");
    *bytes = sort_temp_bytes(n);
    return 0;
}
int dm2_sort_pairs(const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out,
                   int64_t n, int32_t end_bit, void* temp, size_t temp_bytes, void* stream) {
    DM2_REQUIRE(n == 0 || (keys_in && vals_in && keys_out && vals_out && temp), "dm2_sort_pairs: null buffer");
    return launch_sort_pairs(keys_in, vals_in, keys_out, vals_out, n, end_bit, temp, temp_bytes, (cudaStream_t)stream);
}
int dm2_scan_partials_count(int64_t n, int64_t* count) {
    DM2_REQUIRE(count && n >= 0, "dm2_scan_partials_count: bad arguments");
    *count = scan_partials_count(n);
    return 0;
}
int dm2_inclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* partials, void* stream) {
    DM2_REQUIRE(n == 0 || (in && out && partials), "dm2_inclusive_scan_u32: null buffer");
    return launch_scan(in, out, n, partials, (cudaStream_t)stream);
}

}  // extern "C"
