// stroker.cpp — host-side stroke outline generation.
//
// footile strokes a path by flattening it into a polyline of (point, width),
// offsetting both sides by half the width, joining consecutive offsets and
// handing the outline back to fill() as Line/Close ops
// (reference: src/stroker.rs:204-416, src/plotter.rs:356-365).  The outline
// pass is sequential f32 arithmetic over libm's hypotf/atan2f/sinf, whose bits
// must be the ones Rust's f32::{hypot,atan2,sin} return on this platform
// (glibc), so it stays on the host; the flattening before it
// and the fill after it run on the device.
#include <math.h>

#include "engine.h"
#include "pointy_compat.cuh"

namespace ftl {

using pointy::Pt;

namespace {

struct Wide {
    Pt p;
    float w;
};
struct Sub {  // SubStroke (stroker.rs:31-40)
    uint32_t start, n;
    bool joined, done;
};

inline Pt normalize(Pt v) {  // pointy Pt::normalize (RECALLED): v / hypot, zero stays zero
    float m = hypotf(v.x, v.y);
    if (m > 0.0f) return {v.x / m, v.y / m};
    return {0.0f, 0.0f};
}
inline float angle_rel(Pt a, Pt b) {  // pointy Pt::angle_rel (RECALLED): wrapped to (-pi, pi]
    const float pi = 3.14159265358979323846f;
    float th = atan2f(a.y, a.x) - atan2f(b.y, b.x);
    if (th < -pi) return th + 2.0f * pi;
    if (th > pi) return th - 2.0f * pi;
    return th;
}
inline bool intersection(Pt a0, Pt a1, Pt b0, Pt b1, Pt *out) {  // pointy Line::intersection (RECALLED)
    Pt av = pointy::sub(a0, a1), bv = pointy::sub(b0, b1);
    float den = pointy::cross(av, bv);
    if (den == 0.0f) return false;
    float ca = pointy::cross(a0, a1), cb = pointy::cross(b0, b1);
    float xn = bv.x * ca - av.x * cb;
    float yn = bv.y * ca - av.y * cb;
    *out = {xn / den, yn / den};
    return true;
}

struct Outline {
    const StrokeParams &sp;
    std::vector<Wide> pts;
    std::vector<Sub> subs;
    std::vector<ftl_path_op> *out;

    explicit Outline(const StrokeParams &s, std::vector<ftl_path_op> *o) : sp(s), out(o) { subs.push_back({0, 0, false, false}); }

    void add_point(Wide q) {  // Stroke::add_point (stroker.rs:204-216): 65535-point cap, f32 de-dup
        if (pts.size() >= 65535) return;
        bool done = subs.back().done;
        if (done) subs.push_back({(uint32_t)pts.size(), 0, false, false});
        bool same = !pts.empty() && q.p.x == pts.back().p.x && q.p.y == pts.back().p.y;
        if (done || !same) {
            pts.push_back(q);
            subs.back().n++;
        }
    }
    void close(bool joined) {  // Stroke::close (stroker.rs:230-236)
        if (pts.empty()) return;
        subs.back().joined = joined;
        subs.back().done = true;
    }

    void line(Pt p) { out->push_back({FTL_OP_LINE, {p.x, p.y, 0, 0, 0, 0}}); }
    void close_op() { out->push_back({FTL_OP_CLOSE, {0, 0, 0, 0, 0, 0}}); }

    static uint32_t step(const Sub &s, uint32_t v, bool fwd) {  // SubStroke::next (stroker.rs:67-85)
        if (fwd) return v + 1 < s.start + s.n ? v + 1 : s.start;
        return v > s.start ? v - 1 : s.start + s.n - 1;
    }
    static uint32_t seg_count(const Sub &s) {  // SubStroke::len (stroker.rs:88-96)
        if (s.joined) return s.n + 1;
        return s.n > 0 ? s.n - 1 : 0;
    }

    // stroke_arc (stroker.rs:399-416), depth-first with an explicit stack
    void arc(Wide p, Pt a, Pt b) {
        struct Item { Pt a, b; int depth; };
        std::vector<Item> todo;
        todo.push_back({a, b, 0});
        while (!todo.empty()) {
            Item it = todo.back();
            todo.pop_back();
            Pt vr = normalize(pointy::right(pointy::sub(it.b, it.a)));
            Pt c = pointy::add(p.p, pointy::scale(vr, p.w / 2.0f));
            Pt ab = pointy::midpoint(it.a, it.b);
            if (pointy::distance_sq(c, ab) <= sp.tol_sq || it.depth >= 16) line(it.b);  // same depth cap as the curve flattening
            else {
                todo.push_back({c, it.b, it.depth + 1});
                todo.push_back({it.a, c, it.depth + 1});
            }
        }
    }
    void bevel(Pt a1, Pt b0) { line(a1); line(b0); }  // stroker.rs:370-373
    void join(Wide p, Pt a0, Pt a1, Pt b0, Pt b1) {   // stroke_join (stroker.rs:323-396)
        if (sp.join == FTL_JOIN_MITER) {
            float ml = sp.miter_limit;
            if (ml > 0.0f) {
                float sm_min = 1.0f / ml;
                float th = angle_rel(pointy::sub(a1, a0), pointy::sub(b0, b1));
                float sm = fabsf(sinf(th / 2.0f));
                Pt xp;
                if (sm >= sm_min && sm < 1.0f && intersection(a0, a1, b0, b1, &xp)) {
                    line(xp);
                    return;
                }
            }
            bevel(a1, b0);
        } else if (sp.join == FTL_JOIN_BEVEL) {
            bevel(a1, b0);
        } else {
            float th = angle_rel(pointy::sub(a1, a0), pointy::sub(b0, b1));
            if (th <= 0.0f) bevel(a1, b0);
            else {
                line(a1);
                arc(p, a1, b0);
            }
        }
    }
    // stroke_side (stroker.rs:265-295)
    void side(const Sub &s, uint32_t start, bool fwd) {
        bool have = false;
        Pt x0 = {0, 0}, x1 = {0, 0};
        uint32_t v0 = start, v1 = step(s, v0, fwd);
        for (uint32_t i = 0, n = seg_count(s); i < n; i++) {
            Wide p0 = pts[v0], p1 = pts[v1];
            Pt vr = normalize(pointy::right(pointy::sub(p1.p, p0.p)));  // stroke_offset (stroker.rs:301-309)
            Pt r0 = pointy::add(p0.p, pointy::scale(vr, p0.w / 2.0f));
            Pt r1 = pointy::add(p1.p, pointy::scale(vr, p1.w / 2.0f));
            if (have) join(p0, x0, x1, r0, r1);
            else if (!s.joined) line(r0);
            have = true;
            x0 = r0;
            x1 = r1;
            v0 = v1;
            v1 = step(s, v1, fwd);
        }
        if (!s.joined && have) line(x1);
    }
    void emit() {  // path_ops / stroke_sub (stroker.rs:239-262)
        for (const Sub &s : subs) {
            if (seg_count(s) == 0) continue;
            side(s, s.start, true);
            if (s.joined) close_op();
            side(s, step(s, s.start, false), false);
            close_op();
        }
    }
};

}  // namespace

float stroke_widths(float s_width, const ftl_path_op *ops, size_t n_ops, std::vector<float> *opw) {
    opw->assign(2 * n_ops, 0.0f);
    float pen_w = s_width;  // Plotter::reset (plotter.rs:128-130)
    for (size_t i = 0; i < n_ops; i++) {
        (*opw)[2 * i] = pen_w;
        switch (ops[i].tag) {
        case FTL_OP_PENWIDTH: s_width = ops[i].v[0]; break;  // plotter.rs:151-153
        case FTL_OP_CLOSE: pen_w = s_width; break;           // plotter.rs:200-203
        case FTL_OP_MOVE: case FTL_OP_LINE: case FTL_OP_QUAD: case FTL_OP_CUBIC: pen_w = s_width; break;  // move_pen
        default: break;
        }
        (*opw)[2 * i + 1] = s_width;
    }
    return s_width;
}

void stroke_outline(const StrokeParams &sp, const ftl_path_op *ops, size_t n_ops, const WideFlat &flat, std::vector<ftl_path_op> *out) {
    out->clear();
    Outline o(sp, out);
    size_t at = 0;
    for (size_t i = 0; i < n_ops; i++) {
        if (ops[i].tag == FTL_OP_CLOSE) o.close(true);     // plotter.rs:200-203
        else if (ops[i].tag == FTL_OP_MOVE) o.close(false);  // plotter.rs:210
        for (uint32_t k = 0; k < flat.counts[i]; k++, at++) o.add_point({{flat.xyw[3 * at], flat.xyw[3 * at + 1]}, flat.xyw[3 * at + 2]});
    }
    o.emit();
}

}  // namespace ftl
