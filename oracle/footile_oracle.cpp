// footile_oracle.cpp — CPU restatement of footile's fill/stroke hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing in footile_b200/ (the product)
// may include, link, import or call this file.  It is used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
// legs as the checker and the CPU baseline, never as the thing shipped.
//
// The Rust reference cannot be built in this environment (no rustc/cargo and
// its `pix 0.14` / `pointy 0.7` dependencies are un-vendored), so this is a
// single-threaded C++ restatement of the reference's algorithm, pinned to the
// reference by its own in-tree known-answer tests:
//   src/fixed.rs:166-313   (15 Fixed tests)        -> tests/test_oracle_kat.py
//   src/fig.rs:691-794     (fixed_pt + 6 rasters)  -> tests/test_oracle_kat.py
//   src/imgbuf.rs:205-232  (accumulate vectors)    -> tests/test_oracle_kat.py
// Parity status:
//   * Matte8 fill of already-flattened points: PINNED by the fig.rs rasters.
//   * Fixed arithmetic and row accumulate: PINNED.
//   * Curve flattening / transform (pointy f32 op order), stroker outlines and
//     partial-alpha SrcOver (pix Ch8 arithmetic): PARITY UNPINNED — the
//     reference holds no test asserting any of them and the crates' sources
//     are absent; the recalled semantics are isolated in the `pointy_compat`
//     and `pix_compat` namespaces below.
//
// Two fill implementations live here on purpose:
//   fill_sequential() follows src/fig.rs:480-666 (vertex sort, active-edge
//     list, per-row resolve) and is THE oracle;
//   fill_orderfree()  is the per-(edge,row) closed form the GPU design relies
//     on; tests prove it equals fill_sequential() and it is used only where the
//     sequential form's linear find (fig.rs:610-617) would take hours (the
//     10M-edge raster).
//
// Build: see oracle/Makefile (g++ -O3 -mssse3 -ffp-contract=off).

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <thread>
#include <numeric>
#include <vector>
#if defined(__SSSE3__)
#include <tmmintrin.h>
#endif

namespace {

// ---------------------------------------------------------------------------
// Fixed 16.16 — src/fixed.rs:10-159.  All i32 ops wrap (Rust release mode).
// ---------------------------------------------------------------------------
typedef int32_t fx_t;
const fx_t FX_ONE = 1 << 16, FX_HALF = 1 << 15, FX_MASK = (1 << 16) - 1;

inline fx_t fx_add(fx_t a, fx_t b) { return (fx_t)((uint32_t)a + (uint32_t)b); }   // fixed.rs:24-30
inline fx_t fx_sub(fx_t a, fx_t b) { return (fx_t)((uint32_t)a - (uint32_t)b); }   // fixed.rs:32-38
inline fx_t fx_mul(fx_t a, fx_t b) { return (fx_t)(((int64_t)a * (int64_t)b) >> 16); }  // fixed.rs:40-47
inline fx_t fx_div(fx_t a, fx_t b) { return (fx_t)((int64_t)((uint64_t)(int64_t)a << 16) / (int64_t)b); }  // fixed.rs:49-56
inline fx_t fx_shl(fx_t a, uint32_t s) { return (fx_t)((uint32_t)a << s); }        // fixed.rs:58-64
inline fx_t fx_shr(fx_t a, uint32_t s) { return a >> s; }                          // fixed.rs:66-72
inline fx_t fx_from_i32(int32_t i) { return (fx_t)((uint32_t)i << 16); }           // fixed.rs:74-79
inline int32_t fx_to_i32(fx_t a) { return a >> 16; }                               // fixed.rs:81-86
inline fx_t fx_from_f32(float f) {                                                 // fixed.rs:88-93
    // Rust `as i32`: truncate toward zero, saturate, NaN -> 0.
    float v = f * 65536.0f;
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (fx_t)v;
}
inline float fx_to_f32(fx_t a) { return (float)a / 65536.0f; }                     // fixed.rs:95-100
inline fx_t fx_abs(fx_t a) { return a < 0 ? (fx_t)(0u - (uint32_t)a) : a; }        // fixed.rs:122-124
inline fx_t fx_floor(fx_t a) { return a & ~FX_MASK; }                              // fixed.rs:127-129
inline fx_t fx_ceil(fx_t a) { return fx_floor(fx_sub(fx_add(a, FX_ONE), 1)); }     // fixed.rs:132-134
inline fx_t fx_round(fx_t a) { return fx_floor(fx_add(a, FX_HALF)); }              // fixed.rs:137-139
inline fx_t fx_trunc(fx_t a) { return a >= 0 ? fx_floor(a) : fx_ceil(a); }         // fixed.rs:142-148
inline fx_t fx_fract(fx_t a) { return a & FX_MASK; }                               // fixed.rs:151-153
inline fx_t fx_avg(fx_t a, fx_t b) { return fx_add(a, b) >> 1; }                   // fixed.rs:156-158

// ---------------------------------------------------------------------------
// pointy 0.7 semantics (source absent: RECALLED, parity unpinned).  f32,
// round-to-nearest, never fused (-ffp-contract=off).
// ---------------------------------------------------------------------------
namespace pointy_compat {
struct Pt { float x, y; };
inline Pt add(Pt a, Pt b) { return {a.x + b.x, a.y + b.y}; }
inline Pt sub(Pt a, Pt b) { return {a.x - b.x, a.y - b.y}; }
inline Pt scale(Pt a, float s) { return {a.x * s, a.y * s}; }
inline Pt midpoint(Pt a, Pt b) { return {(a.x + b.x) / 2.0f, (a.y + b.y) / 2.0f}; }
inline float distance_sq(Pt a, Pt b) { float dx = a.x - b.x, dy = a.y - b.y; return dx * dx + dy * dy; }
inline Pt right(Pt v) { return {v.y, -v.x}; }
inline Pt normalize(Pt v) {
    float m = hypotf(v.x, v.y);
    if (m > 0.0f) return {v.x / m, v.y / m};
    return {0.0f, 0.0f};
}
inline float angle_rel(Pt a, Pt b) {
    const float pi = 3.14159265358979323846f;
    float th = atan2f(a.y, a.x) - atan2f(b.y, b.x);
    if (th < -pi) return th + 2.0f * pi;
    if (th > pi) return th - 2.0f * pi;
    return th;
}
inline float cross(Pt a, Pt b) { return a.x * b.y - a.y * b.x; }
inline bool line_intersection(Pt a0, Pt a1, Pt b0, Pt b1, Pt *out) {
    Pt av = sub(a0, a1), bv = sub(b0, b1);
    float den = cross(av, bv);
    if (den != 0.0f) {
        float ca = cross(a0, a1), cb = cross(b0, b1);
        float xn = bv.x * ca - av.x * cb;
        float yn = bv.y * ca - av.y * cb;
        *out = {xn / den, yn / den};
        return true;
    }
    return false;
}
inline Pt transform(const float e[6], Pt p) {
    return {e[0] * p.x + e[1] * p.y + e[2], e[3] * p.x + e[4] * p.y + e[5]};
}
}  // namespace pointy_compat
using pointy_compat::Pt;

// ---------------------------------------------------------------------------
// pix 0.14 semantics (source absent: RECALLED, parity unpinned except the
// alpha=255-over-clear case of fig.rs:702-721).
// ---------------------------------------------------------------------------
namespace pix_compat {
inline uint8_t ch8_mul(uint8_t a, uint8_t b) {
    uint32_t l = a; l = (l << 4) | (l >> 4);
    uint32_t r = b; r = (r << 4) | (r >> 4);
    return (uint8_t)((l * r) >> 16);
}
inline uint8_t ch8_add(uint8_t a, uint8_t b) { unsigned s = (unsigned)a + b; return s > 255 ? 255 : (uint8_t)s; }
// dst.composite_channels_alpha(&src, SrcOver, &alpha) for an n-channel
// premultiplied pixel whose LAST channel is alpha (Graya8p n=2, Rgba8p n=4).
inline void src_over_alpha(uint8_t *dst, const uint8_t *src, int n, uint8_t alpha) {
    uint8_t sa1 = (uint8_t)(255 - ch8_mul(alpha, src[n - 1]));
    for (int c = 0; c < n; c++)
        dst[c] = ch8_add(ch8_mul(src[c], alpha), ch8_mul(dst[c], sa1));
}
}  // namespace pix_compat

// ---------------------------------------------------------------------------
// Row accumulate — src/imgbuf.rs:38-199
// ---------------------------------------------------------------------------
inline uint8_t sat_u8(int16_t v) { return v < 0 ? 0 : (v > 255 ? 255 : (uint8_t)v); }  // imgbuf.rs:64-66

void accumulate_non_zero_scalar(uint8_t *dst, int16_t *src, size_t n) {   // imgbuf.rs:54-61
    int16_t sum = 0;
    for (size_t i = 0; i < n; i++) {
        sum = (int16_t)(sum + src[i]);
        src[i] = 0;
        dst[i] = sat_u8(sum);
    }
}
void accumulate_even_odd_scalar(uint8_t *dst, int16_t *src, size_t n) {   // imgbuf.rs:157-167
    int16_t sum = 0;
    for (size_t i = 0; i < n; i++) {
        sum = (int16_t)(sum + src[i]);
        src[i] = 0;
        int16_t v = sum & 0xFF, odd = sum & 0x100;
        int16_t c = (int16_t)(v - odd);
        if (c < 0) c = (int16_t)-c;
        dst[i] = sat_u8(c);
    }
}
#if defined(__SSSE3__)
inline __m128i scan8(__m128i a) {                                         // imgbuf.rs:103-117
    a = _mm_add_epi16(a, _mm_slli_si128(a, 8));
    a = _mm_add_epi16(a, _mm_slli_si128(a, 4));
    return _mm_add_epi16(a, _mm_slli_si128(a, 2));
}
// The reference's SSSE3 body walks 8 lanes at a time past `len` when
// len % 8 != 0 (imgbuf.rs:78-93); here the SIMD loop covers the multiple-of-8
// prefix and a scalar tail finishes the row, which yields the same visible
// bytes without the out-of-bounds access.
void accumulate_simd(uint8_t *dst, int16_t *src, size_t n, bool even_odd) {  // imgbuf.rs:71-98,172-199
    const __m128i zero = _mm_setzero_si128();
    const __m128i bcast = _mm_set1_epi16(0x0F0E);
    __m128i sum = zero;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        __m128i a = _mm_loadu_si128((const __m128i *)(src + i));
        _mm_storeu_si128((__m128i *)(src + i), zero);
        a = _mm_add_epi16(scan8(a), sum);
        __m128i val = a;
        if (even_odd) {
            __m128i v = _mm_and_si128(a, _mm_set1_epi16(0xFF));
            __m128i odd = _mm_and_si128(a, _mm_set1_epi16(0x100));
            val = _mm_abs_epi16(_mm_sub_epi16(v, odd));
        }
        _mm_storel_epi64((__m128i *)(dst + i), _mm_packus_epi16(val, val));
        sum = _mm_shuffle_epi8(a, bcast);
    }
    if (i < n) {
        int16_t s = (int16_t)_mm_extract_epi16(sum, 0);
        for (; i < n; i++) {
            s = (int16_t)(s + src[i]);
            src[i] = 0;
            int16_t c = s;
            if (even_odd) {
                c = (int16_t)((s & 0xFF) - (s & 0x100));
                if (c < 0) c = (int16_t)-c;
            }
            dst[i] = sat_u8(c);
        }
    }
}
#endif
void accumulate(uint8_t *dst, int16_t *src, size_t n, bool even_odd, bool simd) {
#if defined(__SSSE3__)
    if (simd) { accumulate_simd(dst, src, n, even_odd); return; }
#endif
    (void)simd;
    if (even_odd) accumulate_even_odd_scalar(dst, src, n);
    else accumulate_non_zero_scalar(dst, src, n);
}

// ---------------------------------------------------------------------------
// Fig — src/fig.rs
// ---------------------------------------------------------------------------
struct FxPt { fx_t x, y; };
inline bool operator==(FxPt a, FxPt b) { return a.x == b.x && a.y == b.y; }

inline int16_t pixel_cov(fx_t fcov) {                                     // fig.rs:677-682
    return (int16_t)fx_to_i32(fx_round(fx_shl(fcov, 8)));
}
inline bool widdershins(FxPt a, FxPt b) {                                  // fig.rs:116-119
    return fx_mul(a.x, b.y) > fx_mul(b.x, a.y);
}

enum { FWD = 0, REV = 1 };
enum { FMT_MATTE8 = 0, FMT_GRAYA8P = 1, FMT_RGBA8P = 2 };
enum { RULE_NONZERO = 0, RULE_EVENODD = 1 };
inline int fmt_bpp(int fmt) { return fmt == FMT_MATTE8 ? 1 : (fmt == FMT_GRAYA8P ? 2 : 4); }

struct SubFig { uint32_t start, n; bool done; };

struct Fig {
    std::vector<FxPt> points;
    std::vector<SubFig> subs;
    uint32_t vid_cap;  // 65535 = strict Vid(u16) (vid.rs:10-24, fig.rs:430); larger = documented u32 extension
    explicit Fig(uint32_t cap) : vid_cap(cap) {
        points.reserve(1024); subs.reserve(16);
        subs.push_back({0, 0, false});                                    // fig.rs:339-344
    }
    bool coincident(FxPt p) const { return !points.empty() && points.back() == p; }  // fig.rs:445-451
    void add_point(Pt p) {                                                 // fig.rs:428-442
        if (points.size() < vid_cap) {
            bool done = subs.back().done;
            if (done) subs.push_back({(uint32_t)points.size(), 0, false});
            FxPt fp = {fx_from_f32(p.x), fx_from_f32(p.y)};
            if (done || !coincident(fp)) { points.push_back(fp); subs.back().n++; }
        }
    }
    void close() {                                                         // fig.rs:457-461,373-383
        if (points.empty()) return;
        SubFig &s = subs.back();
        if (s.n > 0) {
            if (coincident(points[s.start])) { points.pop_back(); s.n--; }
            s.done = true;
        }
    }
};

// Ring neighbour inside a sub-figure — fig.rs:143-163 (vertex -> sub lookup is
// a table here instead of the linear scan of fig.rs:386-394; same result).
struct Ring {
    const Fig &fig;
    std::vector<uint32_t> sub_of;
    explicit Ring(const Fig &f) : fig(f), sub_of(f.points.size()) {
        for (uint32_t s = 0; s < f.subs.size(); s++)
            for (uint32_t k = 0; k < f.subs[s].n; k++) sub_of[f.subs[s].start + k] = s;
    }
    uint32_t next(uint32_t v, int dir) const {
        const SubFig &s = fig.subs[sub_of[v]];
        if (dir == FWD) { uint32_t w = v + 1; return w < s.start + s.n ? w : s.start; }
        if (v > s.start) return v - 1;
        return s.n > 0 ? s.start + s.n - 1 : s.start;
    }
};

struct Edge {                                                              // fig.rs:47-66
    uint32_t v1; fx_t y_upper, y_lower; int dir;
    fx_t step_pix, inv_slope, x_bot, min_x, max_x;
};

inline Edge edge_new(uint32_t v1, FxPt p0, FxPt p1, int dir) {             // fig.rs:179-210
    Edge e;
    fx_t dx = fx_sub(p1.x, p0.x), dy = fx_sub(p1.y, p0.y);
    e.step_pix = dx != 0 ? std::min(fx_abs(fx_div(dy, dx)), FX_ONE) : 0;
    e.inv_slope = fx_div(dx, dy);
    e.y_upper = p0.y; e.y_lower = p1.y;
    fx_t y_bot = fx_sub(fx_floor(fx_add(p0.y, FX_ONE)), p0.y);
    e.x_bot = fx_add(p0.x, fx_mul(e.inv_slope, y_bot));
    e.v1 = v1; e.dir = dir; e.min_x = e.max_x = 0;
    return e;
}
inline int16_t continuing_cov(const Edge &e, int32_t y_row) {              // fig.rs:252-259
    return fx_to_i32(e.y_lower) == y_row ? pixel_cov(fx_fract(e.y_lower)) : (int16_t)256;
}
inline void set_x_limits(Edge &e, fx_t x0, int32_t y_row) {                // fig.rs:269-278
    fx_t x1 = e.x_bot;
    if (fx_to_i32(e.y_lower) == y_row) {
        fx_t y1 = fx_sub(fx_ceil(e.y_lower), e.y_lower);
        x1 = fx_sub(e.x_bot, fx_mul(e.inv_slope, y1));
    }
    e.min_x = std::min(x0, x1); e.max_x = std::max(x0, x1);
}
inline fx_t step_cov(const Edge &e, fx_t r) { return e.step_pix > 0 ? fx_mul(r, e.step_pix) : r; }  // fig.rs:315-321
inline void scan_area(const Edge &e, int fig_dir, int16_t cov, int16_t *area, int32_t width) {  // fig.rs:285-312
    int16_t ed = e.dir == fig_dir ? 1 : -1;
    fx_t full_cov = fx_from_f32((float)cov / 256.0f);
    int32_t min_pix = fx_to_i32(e.min_x), max_pix = fx_to_i32(e.max_x);
    fx_t r = min_pix == max_pix
        ? fx_mul(fx_sub(FX_ONE, fx_fract(fx_avg(e.max_x, e.min_x))), full_cov)
        : fx_mul(fx_sub(FX_ONE, fx_fract(e.min_x)), FX_HALF);
    fx_t x_cov = step_cov(e, r);
    fx_t step = step_cov(e, FX_ONE);
    int16_t sum_pix = 0;
    for (int32_t x = min_pix; x < width; x++) {
        int16_t x_pix = std::min(pixel_cov(x_cov), cov);
        int16_t p = (int16_t)(x_pix - sum_pix);
        int16_t &cell = area[x > 0 ? x : 0];
        cell = (int16_t)(cell + (int16_t)(p * ed));
        sum_pix = (int16_t)(sum_pix + p);
        if (sum_pix >= cov) break;
        x_cov = std::min(fx_add(x_cov, step), FX_ONE);
    }
}

struct RasterRef { uint8_t *px; uint32_t w, h; int fmt; };

// Resolve one row — fig.rs:621-665 (+ imgbuf for Matte8).  Zeroes `area`.
void rasterize_row(uint8_t *row, int16_t *area, uint32_t w, int fmt, int rule, const uint8_t *clr, bool simd) {
    if (fmt == FMT_MATTE8) {                                               // fig.rs:632-636,650-654: colour ignored
        accumulate(row, area, w, rule == RULE_EVENODD, simd);
        return;
    }
    int n = fmt_bpp(fmt);
    int16_t sum = 0;
    for (uint32_t i = 0; i < w; i++) {
        sum = (int16_t)(sum + area[i]);
        area[i] = 0;
        int16_t c = sum;
        if (rule == RULE_EVENODD) {
            c = (int16_t)((sum & 0xFF) - (sum & 0x100));
            if (c < 0) c = (int16_t)-c;
        }
        pix_compat::src_over_alpha(row + (size_t)i * n, clr, n, sat_u8(c));
    }
}

struct FillInfo { int dir; int32_t top_row; uint32_t n_points; };

// The oracle proper: fig.rs:480-626.
FillInfo fill_sequential(const Fig &fig, int rule, RasterRef ras, const uint8_t *clr, bool simd,
                         std::vector<int16_t> *area_dump /*nullable: H*W signed area before resolve*/) {
    FillInfo info = {FWD, 0, (uint32_t)fig.points.size()};
    uint32_t n = (uint32_t)fig.points.size();
    if (n == 0) return info;
    const std::vector<FxPt> &P = fig.points;
    std::vector<uint32_t> vids(n);
    std::iota(vids.begin(), vids.end(), 0u);
    std::stable_sort(vids.begin(), vids.end(), [&](uint32_t a, uint32_t b) {    // fig.rs:464-472,494
        if (P[a].y != P[b].y) return P[a].y < P[b].y;
        return P[a].x < P[b].x;
    });
    Ring ring(fig);
    uint32_t v0 = vids[0];
    FxPt p = P[v0], pf = P[ring.next(v0, FWD)], pr = P[ring.next(v0, REV)];     // fig.rs:402-411
    FxPt a = {fx_sub(pr.x, p.x), fx_sub(pr.y, p.y)}, b = {fx_sub(pf.x, p.x), fx_sub(pf.y, p.y)};
    int dir = widdershins(a, b) ? FWD : REV;
    int32_t top_row = fx_to_i32(p.y);                                           // fig.rs:496
    info.dir = dir; info.top_row = top_row;
    std::vector<int16_t> area(((size_t)ras.w + 7) & ~(size_t)7, 0);
    std::vector<Edge> edges; edges.reserve(16);
    size_t vi = 0;
    int32_t y_row = top_row;
    size_t bpr = (size_t)ras.w * fmt_bpp(ras.fmt);
    for (int64_t ry = std::max(top_row, 0); ry < (int64_t)ras.h; ry++) {        // fig.rs:497-498,539
        for (Edge &e : edges) {                                                 // fig.rs:557-566
            int16_t cov = continuing_cov(e, y_row);
            if (cov > 0) {
                set_x_limits(e, fx_sub(e.x_bot, e.inv_slope), y_row);           // fig.rs:262-266
                scan_area(e, dir, cov, area.data(), (int32_t)ras.w);
            }
        }
        while (vi < n && fx_to_i32(P[vids[vi]].y) <= y_row) {                   // fig.rs:541-549
            uint32_t v = vids[vi++];
            for (int dd = FWD; dd <= REV; dd++) {                               // fig.rs:576-586
                uint32_t w = ring.next(v, dd);
                if (w == v) continue;
                if (P[w].y > P[v].y) {                                          // fig.rs:589-600
                    Edge e = edge_new(w, P[v], P[w], dd);
                    int32_t r0 = fx_to_i32(e.y_upper);
                    int16_t cov = (int16_t)(continuing_cov(e, r0) - pixel_cov(fx_fract(e.y_upper)));
                    if (cov > 0) {
                        fx_t y0 = fx_sub(FX_ONE, fx_fract(e.y_upper));          // fig.rs:244-249
                        set_x_limits(e, fx_sub(e.x_bot, fx_mul(e.inv_slope, y0)), r0);
                        scan_area(e, dir, cov, area.data(), (int32_t)ras.w);
                    }
                    edges.push_back(e);
                } else if (P[w].y < P[v].y) {                                   // fig.rs:603-617
                    int odir = dd == FWD ? REV : FWD;
                    for (size_t i = 0; i < edges.size(); i++)
                        if (edges[i].v1 == v && edges[i].dir == odir) {
                            edges[i] = edges.back(); edges.pop_back(); break;
                        }
                }
            }
        }
        if (area_dump) std::copy(area.begin(), area.begin() + ras.w, area_dump->begin() + (size_t)ry * ras.w);
        rasterize_row(ras.px + (size_t)ry * bpr, area.data(), ras.w, ras.fmt, rule, clr, simd);
        for (Edge &e : edges) e.x_bot = fx_add(e.x_bot, e.inv_slope);           // fig.rs:569-573
        y_row++;
    }
    return info;
}

// Order-free closed form.  Checked against
// fill_sequential() by tests/test_oracle_orderfree.py.
// rows_lo/rows_hi: only raster rows in [rows_lo, rows_hi) are drawn (stripe checks of huge rasters).
FillInfo fill_orderfree(const Fig &fig, int rule, RasterRef ras, const uint8_t *clr, bool simd, int64_t rows_lo = 0,
                        int64_t rows_hi = INT64_MAX) {
    FillInfo info = {FWD, 0, (uint32_t)fig.points.size()};
    uint32_t n = (uint32_t)fig.points.size();
    if (n == 0) return info;
    const std::vector<FxPt> &P = fig.points;
    uint32_t v0 = 0;
    for (uint32_t v = 1; v < n; v++)
        if (P[v].y < P[v0].y || (P[v].y == P[v0].y && P[v].x < P[v0].x)) v0 = v;
    Ring ring(fig);
    FxPt p = P[v0], pf = P[ring.next(v0, FWD)], pr = P[ring.next(v0, REV)];
    FxPt a = {fx_sub(pr.x, p.x), fx_sub(pr.y, p.y)}, b = {fx_sub(pf.x, p.x), fx_sub(pf.y, p.y)};
    int dir = widdershins(a, b) ? FWD : REV;
    int32_t top = fx_to_i32(p.y);
    info.dir = dir; info.top_row = top;
    int64_t first_row = std::max<int64_t>(std::max(top, 0), rows_lo);
    int64_t last_row = std::min<int64_t>(ras.h, rows_hi);
    if (first_row >= last_row) return info;
    int32_t W = (int32_t)ras.w;
    size_t rows = (size_t)(last_row - first_row);
    // i32 accumulators, truncated to i16 at resolve: the i16 wrapping sums of
    // the reference are a ring homomorphism image of these.
    std::vector<int32_t> acc(rows * (size_t)W, 0);
    int32_t shift = std::min(top, 0);
    for (uint32_t v = 0; v < n; v++) {
        for (int dd = FWD; dd <= REV; dd++) {
            uint32_t w = ring.next(v, dd);
            if (w == v || !(P[w].y > P[v].y)) continue;
            Edge e = edge_new(w, P[v], P[w], dd);
            int32_t ed = dd == dir ? 1 : -1;
            int32_t r0 = fx_to_i32(e.y_upper), r1 = fx_to_i32(e.y_lower);
            for (int64_t r = std::max<int64_t>(r0, first_row + shift); r <= r1; r++) {
                int64_t ry = r - shift;
                if (ry < first_row) continue;
                if (ry >= last_row) break;
                fx_t x_bot = (fx_t)((uint32_t)e.x_bot + (uint32_t)(r - r0) * (uint32_t)e.inv_slope);
                int32_t cov = (r == r1 ? pixel_cov(fx_fract(e.y_lower)) : 256) - (r == r0 ? pixel_cov(fx_fract(e.y_upper)) : 0);
                if (cov <= 0) continue;
                fx_t x0 = r == r0 ? fx_sub(x_bot, fx_mul(e.inv_slope, fx_sub(FX_ONE, fx_fract(e.y_upper))))
                                  : fx_sub(x_bot, e.inv_slope);
                fx_t x1 = r == r1 ? fx_sub(x_bot, fx_mul(e.inv_slope, fx_sub(fx_ceil(e.y_lower), e.y_lower))) : x_bot;
                fx_t min_x = std::min(x0, x1), max_x = std::max(x0, x1);
                int32_t min_pix = fx_to_i32(min_x), max_pix = fx_to_i32(max_x);
                fx_t full = (fx_t)(cov << 8);
                fx_t rr = min_pix == max_pix ? fx_mul(fx_sub(FX_ONE, fx_fract(fx_avg(max_x, min_x))), full)
                                             : fx_mul(fx_sub(FX_ONE, fx_fract(min_x)), FX_HALF);
                fx_t first = e.step_pix > 0 ? fx_mul(rr, e.step_pix) : rr;
                fx_t step = e.step_pix > 0 ? e.step_pix : FX_ONE;
                auto X = [&](int64_t k) -> int32_t {
                    if (k < 0) return 0;
                    int64_t xc = std::min<int64_t>((int64_t)first + k * (int64_t)step, FX_ONE);
                    return std::min<int32_t>(pixel_cov((fx_t)xc), cov);
                };
                int32_t *row = acc.data() + (size_t)(ry - first_row) * W;
                int64_t k = 0;
                int32_t prev = 0;
                if (min_pix < 0) {            // columns < 0 fold into column 0 (fig.rs:295)
                    k = -(int64_t)min_pix;
                    int32_t xk = X(k);
                    if (W > 0) row[0] += ed * xk;
                    prev = xk; k++;
                    if (prev >= cov) continue;
                }
                for (;; k++) {
                    int64_t c = (int64_t)min_pix + k;
                    if (c >= W) break;
                    int32_t xk = X(k);
                    row[c] += ed * (xk - prev);
                    prev = xk;
                    if (xk >= cov) break;
                }
            }
        }
    }
    size_t bpr = (size_t)ras.w * fmt_bpp(ras.fmt);
    std::vector<int16_t> area(((size_t)ras.w + 7) & ~(size_t)7, 0);
    for (size_t i = 0; i < rows; i++) {
        for (int32_t x = 0; x < W; x++) area[x] = (int16_t)acc[i * W + x];
        rasterize_row(ras.px + (first_row + i) * bpr, area.data(), ras.w, ras.fmt, rule, clr, simd);
    }
    return info;
}

// ---------------------------------------------------------------------------
// Path ops, Plotter sinks — src/path.rs:18-31, src/plotter.rs:59-332
// ---------------------------------------------------------------------------
struct PathOp { uint32_t tag; float v[6]; };
enum { OP_CLOSE = 0, OP_MOVE = 1, OP_LINE = 2, OP_QUAD = 3, OP_CUBIC = 4, OP_PENWIDTH = 5 };
enum { JOIN_MITER = 0, JOIN_BEVEL = 1, JOIN_ROUND = 2 };
const int MAX_DEPTH = 16;   // the reference recurses without bound (README.md:32-33); SURVEY A.6-12.  4^16 covers any in-range curve at the minimum tolerance 0.01

struct WidePt { Pt p; float w; };
inline WidePt wmid(WidePt a, WidePt b) { return {pointy_compat::midpoint(a.p, b.p), (a.w + b.w) / 2.0f}; }  // geom.rs:31-35
inline float float_lerp(float a, float b, float t) { return b + (a - b) * t; }                           // geom.rs:14-16

struct Sink {
    virtual void add_point(WidePt p) = 0;
    virtual void close(bool joined) = 0;
    virtual ~Sink() {}
};

struct PlotState {
    float e[6] = {1, 0, 0, 0, 1, 0};
    float tol_sq = 0.3f * 0.3f;       // plotter.rs:97,111
    float s_width = 1.0f;             // plotter.rs:112
    int join = JOIN_MITER; float miter_limit = 4.0f;  // plotter.rs:113
    WidePt pen = {{0, 0}, 1.0f};
};

struct Flattener {
    PlotState &st; Sink &dst;
    Flattener(PlotState &s, Sink &d) : st(s), dst(d) {}
    WidePt tp(WidePt p) const { return {pointy_compat::transform(st.e, p.p), p.w}; }       // plotter.rs:169-172
    bool within(WidePt a, WidePt b) const { return pointy_compat::distance_sq(a.p, b.p) <= st.tol_sq; }  // plotter.rs:273-276
    void quad(WidePt a, WidePt b, WidePt c, int depth) {                                     // plotter.rs:248-265
        WidePt ab = wmid(a, b), bc = wmid(b, c), ab_bc = wmid(ab, bc), ac = wmid(a, c);
        if (within(ab_bc, ac) || depth >= MAX_DEPTH) dst.add_point(c);
        else { quad(a, ab, ab_bc, depth + 1); quad(ab_bc, bc, c, depth + 1); }
    }
    void cubic(WidePt pa, WidePt pb, WidePt pc, WidePt pd, int depth) {                      // plotter.rs:311-332
        WidePt ab = wmid(pa, pb), bc = wmid(pb, pc), cd = wmid(pc, pd);
        WidePt ab_bc = wmid(ab, bc), bc_cd = wmid(bc, cd), pe = wmid(ab_bc, bc_cd), ad = wmid(pa, pd);
        if (within(pe, ad) || depth >= MAX_DEPTH) dst.add_point(pd);
        else { cubic(pa, ab, ab_bc, pe, depth + 1); cubic(pe, bc_cd, cd, pd, depth + 1); }
    }
    void run(const PathOp *ops, size_t n) {                                                  // plotter.rs:175-197
        st.pen = {{0, 0}, st.s_width};                                                       // plotter.rs:128-130
        for (size_t i = 0; i < n; i++) {
            const PathOp &op = ops[i];
            switch (op.tag) {
            case OP_CLOSE: dst.close(true); st.pen = {{0, 0}, st.s_width}; break;            // plotter.rs:200-203
            case OP_MOVE: {                                                                  // plotter.rs:208-214
                WidePt p = {{op.v[0], op.v[1]}, st.s_width};
                dst.close(false); dst.add_point(tp(p)); st.pen = p; break; }
            case OP_LINE: {                                                                  // plotter.rs:219-224
                WidePt p = {{op.v[0], op.v[1]}, st.s_width};
                dst.add_point(tp(p)); st.pen = p; break; }
            case OP_QUAD: {                                                                  // plotter.rs:233-242
                WidePt pen = st.pen;
                WidePt bb = {{op.v[0], op.v[1]}, (pen.w + st.s_width) / 2.0f};
                WidePt cc = {{op.v[2], op.v[3]}, st.s_width};
                quad(tp(pen), tp(bb), tp(cc), 0); st.pen = cc; break; }
            case OP_CUBIC: {                                                                 // plotter.rs:286-305
                WidePt pen = st.pen;
                float w0 = float_lerp(pen.w, st.s_width, 1.0f / 3.0f);
                float w1 = float_lerp(pen.w, st.s_width, 2.0f / 3.0f);
                WidePt bb = {{op.v[0], op.v[1]}, w0}, cc = {{op.v[2], op.v[3]}, w1};
                WidePt dd = {{op.v[4], op.v[5]}, st.s_width};
                cubic(tp(pen), tp(bb), tp(cc), tp(dd), 0); st.pen = dd; break; }
            case OP_PENWIDTH: st.s_width = op.v[0]; break;                                   // plotter.rs:151-153
            default: break;
            }
        }
    }
};

struct FigSink : Sink {                                                                      // plotter.rs:71-78
    Fig fig;
    explicit FigSink(uint32_t cap) : fig(cap) {}
    void add_point(WidePt p) override { fig.add_point(p.p); }
    void close(bool) override { fig.close(); }
};

// ---------------------------------------------------------------------------
// Stroker — src/stroker.rs
// ---------------------------------------------------------------------------
struct SubStroke { uint32_t start, n; bool joined, done; };
struct Stroke : Sink {
    int join; float miter_limit, tol_sq; uint32_t vid_cap;
    std::vector<WidePt> points; std::vector<SubStroke> subs;
    Stroke(int j, float ml, float tol, uint32_t cap) : join(j), miter_limit(ml), tol_sq(tol), vid_cap(cap) {
        subs.push_back({0, 0, false, false});                                                // stroker.rs:114-124
    }
    void add_point(WidePt pt) override {                                                     // stroker.rs:204-216
        if (points.size() < vid_cap) {
            bool done = subs.back().done;
            if (done) subs.push_back({(uint32_t)points.size(), 0, false, false});
            bool coin = !points.empty() && pt.p.x == points.back().p.x && pt.p.y == points.back().p.y;
            if (done || !coin) { points.push_back(pt); subs.back().n++; }
        }
    }
    void close(bool joined) override {                                                       // stroker.rs:230-236
        if (!points.empty()) { subs.back().joined = joined; subs.back().done = true; }
    }
    uint32_t next(const SubStroke &s, uint32_t v, bool fwd) const {                          // stroker.rs:67-85
        if (fwd) { uint32_t w = v + 1; return w < s.start + s.n ? w : s.start; }
        return v > s.start ? v - 1 : s.start + s.n - 1;
    }
    static uint32_t sub_len(const SubStroke &s) { return s.joined ? s.n + 1 : (s.n > 0 ? s.n - 1 : 0); }  // stroker.rs:88-96
    void pt(std::vector<PathOp> &ops, Pt p) const { PathOp o = {OP_LINE, {p.x, p.y, 0, 0, 0, 0}}; ops.push_back(o); }
    void offset(WidePt p0, WidePt p1, Pt *r0, Pt *r1) const {                                // stroker.rs:301-309
        using namespace pointy_compat;
        Pt vr = normalize(right(sub(p1.p, p0.p)));
        *r0 = add(p0.p, scale(vr, p0.w / 2.0f));
        *r1 = add(p1.p, scale(vr, p1.w / 2.0f));
    }
    void arc(std::vector<PathOp> &ops, WidePt p, Pt a, Pt b, int depth) const {              // stroker.rs:399-416
        using namespace pointy_compat;
        Pt vr = normalize(right(sub(b, a)));
        Pt c = add(p.p, scale(vr, p.w / 2.0f));
        Pt ab = midpoint(a, b);
        if (distance_sq(c, ab) <= tol_sq || depth >= MAX_DEPTH) pt(ops, b);
        else { arc(ops, p, a, c, depth + 1); arc(ops, p, c, b, depth + 1); }
    }
    void do_join(std::vector<PathOp> &ops, WidePt p, Pt a0, Pt a1, Pt b0, Pt b1) const {     // stroker.rs:323-396
        using namespace pointy_compat;
        if (join == JOIN_MITER) {
            float ml = miter_limit;
            if (ml > 0.0f) {
                float sm_min = 1.0f / ml;
                float th = angle_rel(sub(a1, a0), sub(b0, b1));
                float sm = fabsf(sinf(th / 2.0f));
                Pt xp;
                if (sm >= sm_min && sm < 1.0f && line_intersection(a0, a1, b0, b1, &xp)) { pt(ops, xp); return; }
            }
            pt(ops, a1); pt(ops, b0);
        } else if (join == JOIN_BEVEL) {
            pt(ops, a1); pt(ops, b0);
        } else {
            float th = angle_rel(sub(a1, a0), sub(b0, b1));
            if (th <= 0.0f) { pt(ops, a1); pt(ops, b0); }
            else { pt(ops, a1); arc(ops, p, a1, b0, 0); }
        }
    }
    void side(std::vector<PathOp> &ops, const SubStroke &s, uint32_t start, bool fwd) const {  // stroker.rs:265-295
        bool have = false; Pt xr0 = {0, 0}, xr1 = {0, 0};
        uint32_t v0 = start, v1 = next(s, v0, fwd);
        uint32_t len = sub_len(s);
        for (uint32_t i = 0; i < len; i++) {
            WidePt p0 = points[v0], p1 = points[v1];
            Pt pr0, pr1; offset(p0, p1, &pr0, &pr1);
            if (have) do_join(ops, p0, xr0, xr1, pr0, pr1);
            else if (!s.joined) pt(ops, pr0);
            have = true; xr0 = pr0; xr1 = pr1;
            v0 = v1; v1 = next(s, v1, fwd);
        }
        if (!s.joined && have) pt(ops, xr1);
    }
    std::vector<PathOp> path_ops() const {                                                   // stroker.rs:239-262
        std::vector<PathOp> ops;
        for (const SubStroke &s : subs) {
            if (sub_len(s) == 0) continue;
            uint32_t end = next(s, s.start, false);
            side(ops, s, s.start, true);
            if (s.joined) ops.push_back({OP_CLOSE, {0, 0, 0, 0, 0, 0}});
            side(ops, s, end, false);
            ops.push_back({OP_CLOSE, {0, 0, 0, 0, 0, 0}});
        }
        return ops;
    }
};

// ---------------------------------------------------------------------------
// Plotter — src/plotter.rs:38-380
// ---------------------------------------------------------------------------
struct Plotter {
    uint32_t w, h; int fmt;
    std::vector<uint8_t> px;
    PlotState st;
    uint32_t vid_cap = 65535;
    bool simd = true;
    bool orderfree = false;
    int64_t rows_lo = 0, rows_hi = INT64_MAX;
    FillInfo last = {FWD, 0, 0};
    RasterRef ras() { return {px.data(), w, h, fmt}; }
    void fill(int rule, const PathOp *ops, size_t n, const uint8_t *clr) {                   // plotter.rs:339-350
        FigSink sink(vid_cap);
        Flattener(st, sink).run(ops, n);
        sink.fig.close();
        last = orderfree ? fill_orderfree(sink.fig, rule, ras(), clr, simd, rows_lo, rows_hi)
                         : fill_sequential(sink.fig, rule, ras(), clr, simd, nullptr);
    }
    std::vector<PathOp> stroke_ops(const PathOp *ops, size_t n) {                            // plotter.rs:361-363
        Stroke s(st.join, st.miter_limit, st.tol_sq, vid_cap);
        Flattener(st, s).run(ops, n);
        return s.path_ops();
    }
    void stroke(const PathOp *ops, size_t n, const uint8_t *clr) {                           // plotter.rs:356-365
        std::vector<PathOp> o = stroke_ops(ops, n);
        fill(RULE_NONZERO, o.data(), o.size(), clr);
    }
};

}  // namespace

// ===========================================================================
// C interface for ctypes (tests / bench cpu_baseline only)
// ===========================================================================
extern "C" {

typedef struct { uint32_t tag; float v[6]; } orc_path_op;

// ---- Fixed KAT hooks (fixed.rs tests) ----
int32_t orc_fx_from_f32(float f) { return fx_from_f32(f); }
int32_t orc_fx_from_i32(int32_t i) { return fx_from_i32(i); }
float orc_fx_to_f32(int32_t a) { return fx_to_f32(a); }
int32_t orc_fx_to_i32(int32_t a) { return fx_to_i32(a); }
// op: 0 add 1 sub 2 mul 3 div 4 shl 5 shr 6 abs 7 floor 8 ceil 9 round 10 trunc 11 fract 12 avg
int32_t orc_fx_op(int op, int32_t a, int32_t b) {
    switch (op) {
    case 0: return fx_add(a, b); case 1: return fx_sub(a, b); case 2: return fx_mul(a, b);
    case 3: return fx_div(a, b); case 4: return fx_shl(a, (uint32_t)b); case 5: return fx_shr(a, (uint32_t)b);
    case 6: return fx_abs(a); case 7: return fx_floor(a); case 8: return fx_ceil(a);
    case 9: return fx_round(a); case 10: return fx_trunc(a); case 11: return fx_fract(a);
    case 12: return fx_avg(a, b);
    }
    return 0;
}
int orc_widdershins(int32_t ax, int32_t ay, int32_t bx, int32_t by) { return widdershins({ax, ay}, {bx, by}) ? 1 : 0; }
int orc_pixel_cov(int32_t f) { return pixel_cov(f); }

// ---- accumulate KAT hook (imgbuf.rs tests) ----
void orc_accumulate(int rule, uint8_t *dst, int16_t *src, size_t n, int simd) {
    accumulate(dst, src, n, rule == RULE_EVENODD, simd != 0);
}
int orc_has_ssse3(void) {
#if defined(__SSSE3__)
    return 1;
#else
    return 0;
#endif
}
// Output conversion of examples/fishy.rs:33 (Raster::<SRgba8>::with_raster) and its Graya8p analogue: Premultiplied ->
// Straight (Ch8 division, RECALLED: min((c << 8) / a, 255), 0 for a = 0), then the sRGB transfer function through the
// 256-entry table round(255 * srgb(i / 255)); alpha copied.  fmt: 1 Graya8p -> SGraya8, 2 Rgba8p -> SRgba8, 0 Matte8 -> SGray8 (bytes kept).
void orc_convert_srgb(int fmt, const uint8_t *src, uint8_t *dst, size_t n_pixels) {
    uint8_t enc[256];
    for (int i = 0; i < 256; i++) {
        const double u = i / 255.0;
        const double e = u <= 0.0031308 ? 12.92 * u : 1.055 * pow(u, 1.0 / 2.4) - 0.055;
        const double r = floor(e * 255.0 + 0.5);
        enc[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
    auto div8 = [](uint32_t c, uint32_t a) -> uint32_t {
        if (a == 0) return 0;
        uint32_t q = (c << 8) / a;
        return q > 255 ? 255 : q;
    };
    const int bpp = fmt == 0 ? 1 : (fmt == 1 ? 2 : 4);
    for (size_t i = 0; i < n_pixels; i++) {
        const uint8_t *s = src + i * bpp;
        uint8_t *d = dst + i * bpp;
        if (fmt == 0) d[0] = s[0];
        else {
            const uint32_t a = s[bpp - 1];
            for (int ch = 0; ch < bpp - 1; ch++) d[ch] = enc[div8(s[ch], a)];
            d[bpp - 1] = (uint8_t)a;
        }
    }
}

// pix_compat hook
void orc_src_over(uint8_t *dst, const uint8_t *src, int n, uint8_t alpha) { pix_compat::src_over_alpha(dst, src, n, alpha); }

// ---- Fig-level hook (fig.rs tests call Fig::add_point / close / fill directly) ----
// pts: n points (x,y) f32; sub_len: lengths of consecutive sub-figures (each closed).
// mode: 0 sequential, 1 order-free.  area_out (nullable): H*W i16 signed area before resolve (sequential only).
int orc_fig_fill(uint32_t w, uint32_t h, int fmt, int rule, const float *pts, const uint32_t *sub_len, uint32_t n_subs,
                 const uint8_t *clr, uint8_t *raster_io, int mode, int simd, uint32_t vid_cap,
                 int16_t *area_out, int32_t *info_out /*dir, top_row, n_points*/) {
    Fig fig(vid_cap);
    size_t k = 0;
    for (uint32_t s = 0; s < n_subs; s++) {
        for (uint32_t i = 0; i < sub_len[s]; i++, k++) fig.add_point({pts[2 * k], pts[2 * k + 1]});
        fig.close();
    }
    RasterRef ras = {raster_io, w, h, fmt};
    FillInfo info;
    if (mode == 1) info = fill_orderfree(fig, rule, ras, clr, simd != 0);
    else {
        std::vector<int16_t> dump;
        if (area_out) dump.assign((size_t)w * h, 0);
        info = fill_sequential(fig, rule, ras, clr, simd != 0, area_out ? &dump : nullptr);
        if (area_out) memcpy(area_out, dump.data(), dump.size() * sizeof(int16_t));
    }
    if (info_out) { info_out[0] = info.dir; info_out[1] = info.top_row; info_out[2] = (int32_t)info.n_points; }
    return 0;
}

// ---- Plotter ----
void *orc_plotter_new(uint32_t w, uint32_t h, int fmt, const uint8_t *init) {
    Plotter *p = new Plotter();
    p->w = w; p->h = h; p->fmt = fmt;
    p->px.assign((size_t)w * h * fmt_bpp(fmt), 0);
    if (init) memcpy(p->px.data(), init, p->px.size());
    return p;
}
void orc_plotter_free(void *h) { delete (Plotter *)h; }
void orc_set_tolerance(void *h, float t) { float tol = t > 0.01f ? t : 0.01f; ((Plotter *)h)->st.tol_sq = tol * tol; }  // plotter.rs:133-137
void orc_set_transform(void *h, const float *e) { memcpy(((Plotter *)h)->st.e, e, 6 * sizeof(float)); }
void orc_set_join(void *h, int kind, float ml) { ((Plotter *)h)->st.join = kind; ((Plotter *)h)->st.miter_limit = ml; }
// vid_cap 65535 = strict reference; simd 1 = SSSE3 accumulate; orderfree 1 = closed-form fill
void orc_set_options(void *h, uint32_t vid_cap, int simd, int orderfree) {
    Plotter *p = (Plotter *)h; p->vid_cap = vid_cap; p->simd = simd != 0; p->orderfree = orderfree != 0;
}
int orc_fill(void *h, int rule, const orc_path_op *ops, size_t n, const uint8_t *clr) {
    ((Plotter *)h)->fill(rule, (const PathOp *)ops, n, clr); return 0;
}
int orc_stroke(void *h, const orc_path_op *ops, size_t n, const uint8_t *clr) {
    ((Plotter *)h)->stroke((const PathOp *)ops, n, clr); return 0;
}
void orc_read_raster(void *h, uint8_t *dst) { Plotter *p = (Plotter *)h; memcpy(dst, p->px.data(), p->px.size()); }
void orc_write_raster(void *h, const uint8_t *src) { Plotter *p = (Plotter *)h; memcpy(p->px.data(), src, p->px.size()); }
void orc_last_info(void *h, int32_t *out) { Plotter *p = (Plotter *)h; out[0] = p->last.dir; out[1] = p->last.top_row; out[2] = (int32_t)p->last.n_points; }
// order-free mode only: restrict drawing to raster rows [lo, hi)
void orc_set_rows(void *h, int64_t lo, int64_t hi) { Plotter *p = (Plotter *)h; p->rows_lo = lo; p->rows_hi = hi; }
float orc_get_pen_width(void *h) { return ((Plotter *)h)->st.s_width; }

// Probe: flattened Fixed points of a fill (after Fig intake + final close).
// Returns the number of points; writes min(n,cap) (x,y) pairs and sub (start,n) pairs.
size_t orc_debug_flatten(void *h, const orc_path_op *ops, size_t n, int32_t *xy, size_t cap, uint32_t *subs, size_t sub_cap, size_t *n_subs) {
    Plotter *p = (Plotter *)h;
    FigSink sink(p->vid_cap);
    Flattener(p->st, sink).run((const PathOp *)ops, n);
    sink.fig.close();
    size_t np = sink.fig.points.size();
    for (size_t i = 0; i < np && i < cap; i++) { xy[2 * i] = sink.fig.points[i].x; xy[2 * i + 1] = sink.fig.points[i].y; }
    size_t ns = 0;
    for (const SubFig &s : sink.fig.subs) {
        if (s.n == 0) continue;
        if (ns < sub_cap) { subs[2 * ns] = s.start; subs[2 * ns + 1] = s.n; }
        ns++;
    }
    if (n_subs) *n_subs = ns;
    return np;
}
// Probe: raw f32 flattened points with widths as the Stroke sink sees them (before de-dup).
size_t orc_debug_flatten_wide(void *h, const orc_path_op *ops, size_t n, float *xyw, size_t cap) {
    struct Rec : Sink { std::vector<WidePt> v; void add_point(WidePt p) override { v.push_back(p); } void close(bool) override {} } rec;
    Plotter *p = (Plotter *)h;
    Flattener(p->st, rec).run((const PathOp *)ops, n);
    for (size_t i = 0; i < rec.v.size() && i < cap; i++) { xyw[3 * i] = rec.v[i].p.x; xyw[3 * i + 1] = rec.v[i].p.y; xyw[3 * i + 2] = rec.v[i].w; }
    return rec.v.size();
}
// Probe: the outline ops Plotter::stroke would hand to fill.
size_t orc_debug_stroke_ops(void *h, const orc_path_op *ops, size_t n, orc_path_op *out, size_t cap) {
    std::vector<PathOp> o = ((Plotter *)h)->stroke_ops((const PathOp *)ops, n);
    for (size_t i = 0; i < o.size() && i < cap; i++) memcpy(&out[i], &o[i], sizeof(PathOp));
    return o.size();
}

// Probe: the edges Fig::fill builds for these ops (Edge::new, fig.rs:179-210), each with the sign
// Edge::scan_area gives it (fig.rs:286: +1 when the edge runs with the figure's direction), in ring order.
// 6 int32 per edge: x_bot, inv_slope, step_pix, y_upper, y_lower, sign.  Returns the number of edges.
size_t orc_debug_edges(void *h, const orc_path_op *ops, size_t n, int32_t *rec, size_t cap) {
    Plotter *p = (Plotter *)h;
    FigSink sink(p->vid_cap);
    Flattener(p->st, sink).run((const PathOp *)ops, n);
    sink.fig.close();
    const Fig &fig = sink.fig;
    const std::vector<FxPt> &P = fig.points;
    const uint32_t np = (uint32_t)P.size();
    if (np == 0) return 0;
    uint32_t v0 = 0;
    for (uint32_t v = 1; v < np; v++)
        if (P[v].y < P[v0].y || (P[v].y == P[v0].y && P[v].x < P[v0].x)) v0 = v;
    Ring ring(fig);
    FxPt q = P[v0], pf = P[ring.next(v0, FWD)], pr = P[ring.next(v0, REV)];
    FxPt a = {fx_sub(pr.x, q.x), fx_sub(pr.y, q.y)}, b = {fx_sub(pf.x, q.x), fx_sub(pf.y, q.y)};
    const int dir = widdershins(a, b) ? FWD : REV;
    size_t ne = 0;
    for (uint32_t v = 0; v < np; v++)
        for (int dd = FWD; dd <= REV; dd++) {
            uint32_t w = ring.next(v, dd);
            if (w == v || !(P[w].y > P[v].y)) continue;
            Edge e = edge_new(w, P[v], P[w], dd);
            if (ne < cap) {
                int32_t *r = rec + 6 * ne;
                r[0] = e.x_bot; r[1] = e.inv_slope; r[2] = e.step_pix; r[3] = e.y_upper; r[4] = e.y_lower; r[5] = dd == dir ? 1 : -1;
            }
            ne++;
        }
    return ne;
}

// Timed multi-threaded batch (bench.py cpu_baseline / --impl reference): job j = ops[offs[j], offs[j+1])
// filled into its own pre-allocated w x h raster with rules[j], transforms[6j..] (or identity), colour
// clr; jobs are dealt round-robin to n_threads std::threads, one Plotter per job (the reference is
// single-threaded per Plotter).  Returns the wall time of the fills in seconds (allocation excluded).
double orc_batch_fill_timed(uint32_t w, uint32_t h, int fmt, uint32_t n_jobs, const orc_path_op *ops, const uint64_t *offs,
                            const uint8_t *rules, const float *transforms, const uint8_t *clr, uint32_t n_threads, uint32_t repeats) {
    std::vector<Plotter *> ps(n_jobs);
    for (uint32_t j = 0; j < n_jobs; j++) {
        Plotter *p = new Plotter();
        p->w = w; p->h = h; p->fmt = fmt;
        p->px.assign((size_t)w * h * fmt_bpp(fmt), 0);
        if (transforms) memcpy(p->st.e, transforms + 6 * (size_t)j, 6 * sizeof(float));
        ps[j] = p;
    }
    if (n_threads < 1) n_threads = 1;
    auto work = [&](uint32_t t) {
        for (uint32_t r = 0; r < repeats; r++)
            for (uint32_t j = t; j < n_jobs; j += n_threads)
                ps[j]->fill(rules ? rules[j] : 0, (const PathOp *)ops + offs[j], (size_t)(offs[j + 1] - offs[j]), clr);
    };
    auto t0 = std::chrono::steady_clock::now();
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < n_threads; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (Plotter *p : ps) delete p;
    return dt;
}

// Checksum of every job's raster (the 256-lane FNV-1a fold of ftl_batch_checksums) without keeping the
// rasters: one Plotter per thread, cleared before each job.  Test infrastructure for full-size batches.
void orc_batch_fill_checksums(uint32_t w, uint32_t h, int fmt, uint32_t n_jobs, const orc_path_op *ops, const uint64_t *offs,
                              const uint8_t *rules, const float *transforms, const uint8_t *clr, uint32_t n_threads, uint64_t *out) {
    if (n_threads < 1) n_threads = 1;
    auto work = [&](uint32_t t) {
        Plotter p;
        p.w = w; p.h = h; p.fmt = fmt;
        const size_t bytes = (size_t)w * h * fmt_bpp(fmt);
        for (uint32_t j = t; j < n_jobs; j += n_threads) {
            p.px.assign(bytes, 0);
            if (transforms) memcpy(p.st.e, transforms + 6 * (size_t)j, 6 * sizeof(float));
            p.fill(rules ? rules[j] : 0, (const PathOp *)ops + offs[j], (size_t)(offs[j + 1] - offs[j]), clr);
            uint64_t lane[256];
            for (int i = 0; i < 256; i++) lane[i] = 0xcbf29ce484222325ull;
            const uint8_t *b = p.px.data();
            for (size_t i = 0; i < bytes; i++) lane[i & 255] = (lane[i & 255] ^ b[i]) * 0x100000001b3ull;
            uint64_t g = 0xcbf29ce484222325ull;
            for (int i = 0; i < 256; i++)
                for (int k = 0; k < 8; k++) g = (g ^ ((lane[i] >> (8 * k)) & 0xFF)) * 0x100000001b3ull;
            out[j] = g;
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; t++) th.emplace_back(work, t);
    for (auto &t : th) t.join();
}

}  // extern "C"
