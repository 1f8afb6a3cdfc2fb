// stroker.cpp — host-side stroke outline generation.
//
// footile strokes a path by flattening it into a polyline of (point, width),
// offsetting both sides by half the width, joining consecutive offsets and
// handing the outline back to fill() as Line/Close ops
// (reference: src/stroker.rs:204-416, src/plotter.rs:356-365).  The outline
// pass is f32 arithmetic over libm's hypotf/atan2f/sinf, whose bits must be the
// ones Rust's f32::{hypot,atan2,sin} return on this platform (glibc).  This file
// is the sequential form: it outlines small strokes in
// microseconds and is the ordered fallback of the device stroker
// (stroke_kernels.cuh), which produces the same ops bit for bit for large
// strokes and batches; the fill after it always runs on the device.
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

// Stroke-side flattening on the host: the same op sequence as the device kernel flatten_ops<WIDE = true>
// (front_kernels.cuh; plotter.rs:175-332 with WidePt, geom.rs:14-35) - IEEE f32 adds, multiplies and divisions by two,
// never contracted (-ffp-contract=off here, -fmad=false there), so both produce the same bits.  A stroke of a few
// thousand ops is flattened faster here than one device round trip takes.
namespace {
struct WFlat {
    const float *e;
    float tol_sq;
    WideFlat *out;
    uint32_t n;
    void put(Wide q) {
        out->xyw.push_back(q.p.x);
        out->xyw.push_back(q.p.y);
        out->xyw.push_back(q.w);
        n++;
    }
    static Wide mid(Wide a, Wide b) { return {pointy::midpoint(a.p, b.p), (a.w + b.w) / 2.0f}; }  // WidePt::midpoint (geom.rs:31-35)
    void quad(Wide a, Wide b, Wide c, int depth) {  // plotter.rs:248-265
        Wide ab = mid(a, b), bc = mid(b, c), ab_bc = mid(ab, bc), ac = mid(a, c);
        if (pointy::distance_sq(ab_bc.p, ac.p) <= tol_sq || depth >= 16) put(c);
        else {
            quad(a, ab, ab_bc, depth + 1);
            quad(ab_bc, bc, c, depth + 1);
        }
    }
    void cubic(Wide a, Wide b, Wide c, Wide d, int depth) {  // plotter.rs:311-332
        Wide ab = mid(a, b), bc = mid(b, c), cd = mid(c, d), ab_bc = mid(ab, bc), bc_cd = mid(bc, cd), pe = mid(ab_bc, bc_cd), ad = mid(a, d);
        if (pointy::distance_sq(pe.p, ad.p) <= tol_sq || depth >= 16) put(d);
        else {
            cubic(a, ab, ab_bc, pe, depth + 1);
            cubic(pe, bc_cd, cd, d, depth + 1);
        }
    }
};
}  // namespace

void flatten_wide_host(const float e[6], float tol_sq, const ftl_path_op *ops, size_t n_ops, const float *opw, WideFlat *out) {
    out->counts.assign(n_ops, 0);
    out->xyw.clear();
    WFlat f{e, tol_sq, out, 0};
    Pt pen = {0.0f, 0.0f};  // Plotter::reset / close move the pen to the origin (plotter.rs:128-130,200-203)
    for (size_t i = 0; i < n_ops; i++) {
        const ftl_path_op &op = ops[i];
        f.n = 0;
        if (op.tag == FTL_OP_CLOSE) pen = {0.0f, 0.0f};
        else if (op.tag >= FTL_OP_MOVE && op.tag <= FTL_OP_CUBIC) {
            const float w_pen = opw[2 * i], w_now = opw[2 * i + 1];
            const Wide a = {pointy::transform(e, pen), w_pen};
            if (op.tag == FTL_OP_MOVE || op.tag == FTL_OP_LINE) {
                f.put({pointy::transform(e, {op.v[0], op.v[1]}), w_now});
                pen = {op.v[0], op.v[1]};
            } else if (op.tag == FTL_OP_QUAD) {
                const Wide b = {pointy::transform(e, {op.v[0], op.v[1]}), (w_pen + w_now) / 2.0f};
                const Wide c = {pointy::transform(e, {op.v[2], op.v[3]}), w_now};
                f.quad(a, b, c, 0);
                pen = {op.v[2], op.v[3]};
            } else {  // float_lerp(a, b, t) = b + (a - b) * t (geom.rs:14-16)
                const Wide b = {pointy::transform(e, {op.v[0], op.v[1]}), w_now + (w_pen - w_now) * (1.0f / 3.0f)};
                const Wide c = {pointy::transform(e, {op.v[2], op.v[3]}), w_now + (w_pen - w_now) * (2.0f / 3.0f)};
                const Wide d = {pointy::transform(e, {op.v[4], op.v[5]}), w_now};
                f.cubic(a, b, c, d, 0);
                pen = {op.v[4], op.v[5]};
            }
        }
        out->counts[i] = f.n;
    }
}

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

void stroke_sub_table(const ftl_path_op *ops, uint32_t op_begin, uint32_t op_end, uint32_t job, std::vector<uint32_t> *subs, uint32_t *op_sub) {
    bool have_points = false, pending = true;  // pending: the next drawing op starts a sub-stroke
    size_t cur = 0;
    for (uint32_t i = op_begin; i < op_end; i++) {
        const uint32_t tag = ops[i].tag;
        op_sub[i] = 0xFFFFFFFFu;
        if (tag == FTL_OP_PENWIDTH) continue;
        if (tag == FTL_OP_CLOSE || tag == FTL_OP_MOVE) {  // Stroke::close(joined) acts on the current sub-stroke, also when it is already done
            if (have_points) {
                (*subs)[cur + 2] = tag == FTL_OP_CLOSE ? 1u : 0u;
                pending = true;
            }
            if (tag == FTL_OP_CLOSE) continue;
        }
        if (pending) {
            cur = subs->size();
            const uint32_t rec[4] = {i, i + 1, 0u, job};
            subs->insert(subs->end(), rec, rec + 4);
            pending = false;
        }
        op_sub[i] = (uint32_t)(cur / 4);
        (*subs)[cur + 1] = i + 1;
        have_points = true;
    }
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
