// capi.cpp — the C ABI of include/footile_b200.h over ftl::Engine.
// Plain pointers and sizes only; never throws across the boundary.
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <algorithm>
#include <new>
#include <thread>
#include <utility>
#include <vector>

#include "engine.h"
#include "fixed.cuh"
#include "libm_compat.cuh"

using namespace ftl;

struct ftl_plotter {
    Engine eng;
    Geometry geo;
    void *raster = nullptr;
    float e[6] = {1, 0, 0, 0, 1, 0};  // Transform::default (plotter.rs:110)
    float tol_sq = 0.3f * 0.3f;       // plotter.rs:97,111
    float s_width = 1.0f;             // plotter.rs:112
    int join = FTL_JOIN_MITER;        // JoinStyle::Miter(4.0) (plotter.rs:113)
    float miter_limit = 4.0f;
    bool strict_vid = false;          // reproduce Fig::add_point's 65 535-point cap (fig.rs:430)
    explicit ftl_plotter(int device) : eng(device) {}
};

struct ftl_batch {
    Engine eng;
    Geometry geo;
    uint32_t capacity = 0;
    void *rasters = nullptr;
    float tol_sq = 0.3f * 0.3f;
    int join = FTL_JOIN_MITER;
    float miter_limit = 4.0f;
    explicit ftl_batch(int device) : eng(device) {}
};

#define GUARD_BEGIN try {
#define GUARD_END                                       \
    }                                                   \
    catch (const std::bad_alloc &) {                    \
        set_error("host allocation failed");            \
        return FTL_ERR_NOMEM;                           \
    }                                                   \
    catch (...) {                                       \
        set_error("unexpected host exception");         \
        return FTL_ERR_INVALID;                         \
    }

static int bad(const char *msg) {
    set_error(msg);
    return FTL_ERR_INVALID;
}
static bool fmt_ok(int f) { return f == FTL_MATTE8 || f == FTL_GRAYA8P || f == FTL_RGBA8P; }

extern "C" {

int ftl_abi_version(void) { return FTL_ABI_VERSION; }
const char *ftl_last_error(void) { return last_error(); }

int ftl_device_count(int *count) {
    GUARD_BEGIN
    if (!count) return bad("count is null");
    return Engine::device_count(count);
    GUARD_END
}

int ftl_plotter_new_band(uint32_t width, uint32_t height, uint32_t row_begin, uint32_t row_end, int format, const void *init_pixels,
                         int device, ftl_plotter **out) {
    GUARD_BEGIN
    if (!out) return bad("out is null");
    *out = nullptr;
    if (!fmt_ok(format)) return bad("unknown pixel format");
    if (row_begin > row_end || row_end > height) return bad("row band outside the raster");
    if (width > (1u << 24) || height > (1u << 24)) return bad("raster too large");
    ftl_plotter *p = new ftl_plotter(device);
    p->geo.width = width; p->geo.height = height; p->geo.row_begin = row_begin; p->geo.row_end = row_end; p->geo.format = format;
    int rc = p->eng.alloc_raster(p->geo.bytes(), &p->raster);
    if (!rc && p->geo.bytes()) {
        if (init_pixels) rc = p->eng.copy_in(p->raster, init_pixels, p->geo.bytes());
        else rc = p->eng.memset_async(p->raster, 0, p->geo.bytes());
    }
    if (rc) {
        delete p;
        return rc;
    }
    *out = p;
    return FTL_OK;
    GUARD_END
}

int ftl_plotter_new(uint32_t width, uint32_t height, int format, const void *init_pixels, int device, ftl_plotter **out) {
    return ftl_plotter_new_band(width, height, 0, height, format, init_pixels, device, out);
}

int ftl_plotter_free(ftl_plotter *p) {
    GUARD_BEGIN
    if (!p) return FTL_OK;
    p->eng.free_raster(p->raster);
    delete p;
    return FTL_OK;
    GUARD_END
}

uint32_t ftl_width(const ftl_plotter *p) { return p ? p->geo.width : 0; }
uint32_t ftl_height(const ftl_plotter *p) { return p ? p->geo.height : 0; }
float ftl_pen_width(const ftl_plotter *p) { return p ? p->s_width : 0.0f; }

int ftl_set_tolerance(ftl_plotter *p, float t) {
    if (!p) return bad("null plotter");
    float tol = t > 0.01f ? t : 0.01f;  // t.max(0.01) (plotter.rs:134); NaN.max(0.01) = 0.01 as well
    p->tol_sq = tol * tol;
    return FTL_OK;
}
int ftl_set_transform(ftl_plotter *p, const float e[6]) {
    if (!p || !e) return bad("null argument");
    for (int k = 0; k < 6; k++)
        if (!(e[k] - e[k] == 0.0f)) {
            set_error("non-finite transform");
            return FTL_ERR_NONFINITE;
        }
    memcpy(p->e, e, sizeof(p->e));
    return FTL_OK;
}
int ftl_set_strict_vid(ftl_plotter *p, int enabled) {
    if (!p) return bad("null plotter");
    p->strict_vid = enabled != 0;
    return FTL_OK;
}
int ftl_set_join(ftl_plotter *p, int join, float miter_limit) {
    if (!p) return bad("null plotter");
    if (join < FTL_JOIN_MITER || join > FTL_JOIN_ROUND) return bad("unknown join style");
    p->join = join;
    p->miter_limit = miter_limit;
    return FTL_OK;
}

static int check_finite_ops(const ftl_path_op *ops, size_t n_ops);

// Strict Vid(u16) mode (ftl_set_strict_vid): the reference keeps vertex ids in a u16 and Fig::add_point ignores every
// point while 65 535 are stored (fig.rs:428-442, vid.rs:20-24) - a sequential rule, because Fig::close pops a closing
// point (fig.rs:373-383) and so makes room for one more.  A fill that can reach the cap is therefore taken through
// the intake on the host: the path is flattened here (the same f32 ops as the device kernel, stroker.cpp), add_point /
// close are replayed with the cap, and the points that were stored travel on as Move / Line ops under the identity
// transform (1*x + 0*y + 0 is x), one Move per sub-figure.  Returns false when the cap is never reached: the fill
// then takes the ordinary path unchanged.
static bool strict_intake(const float e[6], float tol_sq, const ftl_path_op *ops, size_t n_ops, std::vector<ftl_path_op> *out) {
    bool curves = false;
    for (size_t i = 0; i < n_ops && !curves; i++) curves = ops[i].tag == FTL_OP_QUAD || ops[i].tag == FTL_OP_CUBIC;
    if (!curves && n_ops < 65535) return false;  // one point per Move / Line at most
    std::vector<float> opw(2 * n_ops, 0.0f);
    WideFlat flat;
    flatten_wide_host(e, tol_sq, ops, n_ops, opw.data(), &flat);
    if (flat.xyw.size() / 3 < 65535) return false;
    out->clear();
    size_t stored = 0, at = 0;              // points.len()
    std::vector<std::pair<fx_t, fx_t>> sub;  // the points of the current sub-figure
    bool done = false, capped = false;
    auto close = [&]() {  // Fig::close (fig.rs:457-461) -> sub_set_done (fig.rs:373-383)
        if (stored == 0 || sub.empty()) return;
        if (sub.back() == sub.front()) {  // the closing point is popped (also the only point of a one-point sub-figure)
            sub.pop_back();
            stored--;
        }
        done = true;
    };
    for (size_t i = 0; i < n_ops; i++) {
        if (ops[i].tag == FTL_OP_CLOSE || ops[i].tag == FTL_OP_MOVE) close();
        for (uint32_t k = 0; k < flat.counts[i]; k++, at++) {
            if (stored >= 65535) {  // fig.rs:430
                capped = true;
                continue;
            }
            const std::pair<fx_t, fx_t> q(fx_from_f32(flat.xyw[3 * at]), fx_from_f32(flat.xyw[3 * at + 1]));
            const bool start = done || stored == 0;
            // is_coincident compares with points.last(): the current sub-figure's last point (a sub-figure emptied by a
            // pop is always `done`, so `start` covers it)
            if (start || q != sub.back()) {
                if (start) {
                    sub.clear();
                    done = false;
                }
                ftl_path_op o{};
                o.tag = start ? FTL_OP_MOVE : FTL_OP_LINE;
                o.v[0] = flat.xyw[3 * at]; o.v[1] = flat.xyw[3 * at + 1];
                out->push_back(o);
                sub.push_back(q);
                stored++;
            }
        }
    }
    return capped;
}

static int plot_fill(ftl_plotter *p, int rule, const ftl_path_op *ops, size_t n_ops, const uint8_t *color, bool upload_only = false) {
    if (rule != FTL_NONZERO && rule != FTL_EVENODD) return bad("unknown fill rule");
    if (n_ops && !ops) return bad("ops is null");
    std::vector<HostJob> jobs(1);
    HostJob &j = jobs[0];
    j.op_begin = 0; j.op_end = (uint32_t)n_ops;
    memcpy(j.e, p->e, sizeof(j.e));
    std::vector<ftl_path_op> strict_ops;
    if (p->strict_vid && n_ops) {
        int rc0 = check_finite_ops(ops, n_ops);
        if (rc0) return rc0;
        if (strict_intake(p->e, p->tol_sq, ops, n_ops, &strict_ops)) {
            for (size_t i = 0; i < n_ops; i++)  // PenWidth persists on the plotter (plotter.rs:151-153)
                if (ops[i].tag == FTL_OP_PENWIDTH) p->s_width = ops[i].v[0];
            ops = strict_ops.data();
            n_ops = strict_ops.size();
            j.op_end = (uint32_t)n_ops;
            const float ident[6] = {1, 0, 0, 0, 1, 0};
            memcpy(j.e, ident, sizeof(j.e));
        }
    }
    j.tol_sq = p->tol_sq;
    j.rule = rule;
    if (color) memcpy(j.color, color, p->geo.bpp());
    j.raster = p->raster;
    int rc = upload_only ? p->eng.upload(p->geo, jobs, ops, n_ops) : p->eng.fill(p->geo, jobs, ops, n_ops, true);
    if (rc) return rc;  // a rejected call leaves the plotter state alone
    // PenWidth persists on the plotter across calls (plotter.rs:151-153)
    for (size_t i = 0; i < n_ops; i++)
        if (ops[i].tag == FTL_OP_PENWIDTH) p->s_width = ops[i].v[0];
    return FTL_OK;
}

int ftl_fill(ftl_plotter *p, int rule, const ftl_path_op *ops, size_t n_ops, const uint8_t *color) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    return plot_fill(p, rule, ops, n_ops, color);
    GUARD_END
}

int ftl_fill_upload(ftl_plotter *p, int rule, const ftl_path_op *ops, size_t n_ops, const uint8_t *color) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    return plot_fill(p, rule, ops, n_ops, color, true);
    GUARD_END
}
int ftl_fill_replay(ftl_plotter *p) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    return p->eng.replay();
    GUARD_END
}

int ftl_fill_layers(ftl_plotter *p, uint32_t n_layers, const ftl_path_op *ops, const uint64_t *op_offsets, const uint8_t *rules,
                    const uint8_t *colors) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    if (n_layers == 0) return FTL_OK;
    if (!op_offsets) return bad("op_offsets is null");
    const size_t n_ops = (size_t)op_offsets[n_layers];
    if (n_ops && !ops) return bad("ops is null");
    std::vector<HostJob> jobs(n_layers);
    for (uint32_t l = 0; l < n_layers; l++) {
        HostJob &j = jobs[l];
        if (op_offsets[l + 1] < op_offsets[l] || op_offsets[l + 1] > 0x7FFFFFFFull) return bad("op_offsets must be non-decreasing");
        j.op_begin = (uint32_t)op_offsets[l];
        j.op_end = (uint32_t)op_offsets[l + 1];
        memcpy(j.e, p->e, sizeof(j.e));
        j.tol_sq = p->tol_sq;
        j.rule = rules ? rules[l] : FTL_NONZERO;
        if (j.rule != FTL_NONZERO && j.rule != FTL_EVENODD) return bad("unknown fill rule");
        if (colors) memcpy(j.color, colors + 4 * (size_t)l, 4);
        j.raster = p->raster;
    }
    int rc = p->eng.fill_layers(p->geo, jobs, ops, n_ops);
    if (rc) return rc;
    for (size_t i = 0; i < n_ops; i++)  // PenWidth persists on the plotter (plotter.rs:151-153)
        if (ops[i].tag == FTL_OP_PENWIDTH) p->s_width = ops[i].v[0];
    return FTL_OK;
    GUARD_END
}

// The stroke-side flatten runs on the host unless the path is huge (or FTL_DEVICE_STROKE_FLATTEN=1 asks for the device
// kernel: the parity tests compare the two): a device round trip costs more than flattening a few thousand ops here.
static bool stroke_flatten_on_host(size_t n_ops) {
    const char *ev = getenv("FTL_DEVICE_STROKE_FLATTEN");
    if (ev && atoi(ev) != 0) return false;
    return n_ops <= (1u << 16);
}
static int check_finite_ops(const ftl_path_op *ops, size_t n_ops) {
    for (size_t i = 0; i < n_ops; i++) {
        if (ops[i].tag > FTL_OP_PENWIDTH) return bad("unknown path op tag");
        const int nv = ops[i].tag == FTL_OP_CLOSE ? 0 : (ops[i].tag == FTL_OP_QUAD ? 4 : (ops[i].tag == FTL_OP_CUBIC ? 6 : (ops[i].tag == FTL_OP_PENWIDTH ? 1 : 2)));
        for (int k = 0; k < nv; k++)
            if (!(ops[i].v[k] - ops[i].v[k] == 0.0f)) {
                set_error("non-finite coordinate in path op");
                return FTL_ERR_NONFINITE;
            }
    }
    return FTL_OK;
}

static int stroke_ops(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, std::vector<ftl_path_op> *outline) {
    std::vector<float> opw;
    float final_w = stroke_widths(p->s_width, ops, n_ops, &opw);
    WideFlat flat;
    int rc = FTL_OK;
    if (stroke_flatten_on_host(n_ops)) {
        if ((rc = check_finite_ops(ops, n_ops))) return rc;
        flatten_wide_host(p->e, p->tol_sq, ops, n_ops, opw.data(), &flat);
    } else rc = p->eng.flatten_wide(p->e, p->tol_sq, ops, n_ops, opw.data(), &flat);
    if (rc) return rc;
    p->s_width = final_w;
    StrokeParams sp;
    sp.join = p->join; sp.miter_limit = p->miter_limit; sp.tol_sq = p->tol_sq;
    stroke_outline(sp, ops, n_ops, flat, outline);
    return FTL_OK;
}

// Where the stroker runs.  FTL_DEVICE_STROKE=1 / 0 forces the device / the host; otherwise a stroke of a few hundred ops
// is outlined on the host in microseconds (less than the device path's launches and its one synchronisation, and it
// overlaps the device's work on the previous call), larger strokes and batches go to the device (stroke_kernels.cuh).
static bool stroke_on_device(size_t n_ops, bool batch) {
    if (const char *ev = getenv("FTL_DEVICE_STROKE")) return atoi(ev) != 0;
    return n_ops >= (batch ? 2048u : 512u);
}

// Plotter::stroke with the stroker on the device; *done = false when the host stroker has to take the call.
static int stroke_device(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, const uint8_t *color, bool *done) {
    *done = false;
    std::vector<float> opw;
    const float final_w = stroke_widths(p->s_width, ops, n_ops, &opw);
    std::vector<HostJob> jobs(1);
    HostJob &j = jobs[0];
    j.op_begin = 0; j.op_end = (uint32_t)n_ops;
    memcpy(j.e, p->e, sizeof(j.e));
    j.tol_sq = p->tol_sq;
    j.rule = FTL_NONZERO;
    if (color) memcpy(j.color, color, p->geo.bpp());
    j.raster = p->raster;
    bool needs_host = false;
    int rc = p->eng.stroke(p->geo, jobs, ops, n_ops, opw.data(), p->join, p->miter_limit, &needs_host);
    if (rc) return rc;
    if (needs_host) return FTL_OK;
    p->s_width = final_w;
    *done = true;
    return FTL_OK;
}

int ftl_stroke(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, const uint8_t *color) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    if (n_ops && !ops) return bad("ops is null");
    if (n_ops && n_ops < 0x7FFFFFFFull && stroke_on_device(n_ops, false)) {
        bool done = false;
        int rc = stroke_device(p, ops, n_ops, color, &done);
        if (rc || done) return rc;
    }
    std::vector<ftl_path_op> outline;
    int rc = stroke_ops(p, ops, n_ops, &outline);
    if (rc) return rc;
    // self.fill(FillRule::NonZero, ops.iter(), clr) (plotter.rs:364): the outline goes through the
    // plotter's transform a second time, exactly as in the reference.
    return plot_fill(p, FTL_NONZERO, outline.data(), outline.size(), color);
    GUARD_END
}

int ftl_stroke_outline(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, ftl_path_op *out, size_t cap, size_t *n_out) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    if (n_ops && !ops) return bad("ops is null");
    std::vector<ftl_path_op> outline;
    int rc = stroke_ops(p, ops, n_ops, &outline);  // updates the persistent pen width like ftl_stroke
    if (rc) return rc;
    if (n_out) *n_out = outline.size();
    if (out) memcpy(out, outline.data(), sizeof(ftl_path_op) * (outline.size() < cap ? outline.size() : cap));
    return FTL_OK;
    GUARD_END
}

int ftl_read_raster(ftl_plotter *p, void *dst, size_t nbytes) {
    GUARD_BEGIN
    if (!p || (!dst && nbytes)) return bad("null argument");
    if (nbytes != p->geo.bytes()) return bad("nbytes does not match rows*width*bpp");
    if (!nbytes) return p->eng.sync();
    return p->eng.copy_out(dst, p->raster, nbytes);
    GUARD_END
}
int ftl_read_raster_srgb(ftl_plotter *p, void *dst, size_t nbytes) {
    GUARD_BEGIN
    if (!p || (!dst && nbytes)) return bad("null argument");
    if (nbytes != p->geo.bytes()) return bad("nbytes does not match rows*width*bpp");
    if (!nbytes) return p->eng.sync();
    return p->eng.copy_out_srgb(dst, p->raster, nbytes, p->geo.format);
    GUARD_END
}
int ftl_write_raster(ftl_plotter *p, const void *src, size_t nbytes) {
    GUARD_BEGIN
    if (!p || (!src && nbytes)) return bad("null argument");
    if (nbytes != p->geo.bytes()) return bad("nbytes does not match rows*width*bpp");
    if (!nbytes) return FTL_OK;
    return p->eng.copy_in(p->raster, src, nbytes);
    GUARD_END
}
int ftl_sync(ftl_plotter *p) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    return p->eng.sync();
    GUARD_END
}
int ftl_raster_device_ptr(ftl_plotter *p, void **dptr, size_t *nbytes) {
    if (!p || !dptr) return bad("null argument");
    int rc = p->eng.sync();  // fills issued so far (and any deferred repeat of an overflowed replay) are complete
    if (rc) return rc;
    *dptr = p->raster;
    if (nbytes) *nbytes = p->geo.bytes();
    return FTL_OK;
}

// ---- batch ----
int ftl_batch_new(uint32_t width, uint32_t height, int format, uint32_t capacity, int device, ftl_batch **out) {
    GUARD_BEGIN
    if (!out) return bad("out is null");
    *out = nullptr;
    if (!fmt_ok(format)) return bad("unknown pixel format");
    if (capacity == 0) return bad("capacity is zero");
    if (width > (1u << 24) || height > (1u << 24)) return bad("raster too large");
    ftl_batch *b = new ftl_batch(device);
    b->geo.width = width; b->geo.height = height; b->geo.row_begin = 0; b->geo.row_end = height; b->geo.format = format;
    b->capacity = capacity;
    int rc = b->eng.alloc_raster(b->geo.bytes() * capacity, &b->rasters);
    if (!rc && b->geo.bytes()) rc = b->eng.memset_async(b->rasters, 0, b->geo.bytes() * capacity);
    if (rc) {
        delete b;
        return rc;
    }
    *out = b;
    return FTL_OK;
    GUARD_END
}
int ftl_batch_free(ftl_batch *b) {
    GUARD_BEGIN
    if (!b) return FTL_OK;
    b->eng.free_raster(b->rasters);
    delete b;
    return FTL_OK;
    GUARD_END
}
int ftl_batch_set_tolerance(ftl_batch *b, float t) {
    if (!b) return bad("null batch");
    float tol = t > 0.01f ? t : 0.01f;
    b->tol_sq = tol * tol;
    return FTL_OK;
}
int ftl_batch_clear(ftl_batch *b, uint32_t first, uint32_t count) {
    GUARD_BEGIN
    if (!b) return bad("null batch");
    if ((uint64_t)first + count > b->capacity) return bad("raster range outside the batch");
    if (!count || !b->geo.bytes()) return FTL_OK;
    return b->eng.memset_async((uint8_t *)b->rasters + b->geo.bytes() * first, 0, b->geo.bytes() * count);
    GUARD_END
}

static int batch_jobs(ftl_batch *b, uint32_t n_jobs, const uint64_t *op_offsets, const uint8_t *rules, const float *transforms,
                      const uint8_t *colors, std::vector<HostJob> *jobs) {
    if (n_jobs > b->capacity) return bad("more jobs than rasters in the batch");
    if (n_jobs && !op_offsets) return bad("op_offsets is null");
    jobs->resize(n_jobs);
    for (uint32_t j = 0; j < n_jobs; j++) {
        HostJob &h = (*jobs)[j];
        if (op_offsets[j + 1] < op_offsets[j] || op_offsets[j + 1] > 0x7FFFFFFFull) return bad("op_offsets must be non-decreasing");
        h.op_begin = (uint32_t)op_offsets[j];
        h.op_end = (uint32_t)op_offsets[j + 1];
        if (transforms) memcpy(h.e, transforms + 6 * (size_t)j, sizeof(h.e));
        h.tol_sq = b->tol_sq;
        h.rule = rules ? rules[j] : FTL_NONZERO;
        if (h.rule != FTL_NONZERO && h.rule != FTL_EVENODD) return bad("unknown fill rule");
        if (colors) memcpy(h.color, colors + 4 * (size_t)j, 4);
        h.raster = (uint8_t *)b->rasters + b->geo.bytes() * j;
    }
    return FTL_OK;
}

int ftl_batch_fill(ftl_batch *b, uint32_t n_jobs, const ftl_path_op *ops, const uint64_t *op_offsets, const uint8_t *rules,
                   const float *transforms, const uint8_t *colors) {
    GUARD_BEGIN
    if (!b) return bad("null batch");
    if (n_jobs == 0) return FTL_OK;
    std::vector<HostJob> jobs;
    int rc = batch_jobs(b, n_jobs, op_offsets, rules, transforms, colors, &jobs);
    if (rc) return rc;
    size_t n_ops = (size_t)op_offsets[n_jobs];
    if (n_ops && !ops) return bad("ops is null");
    return b->eng.fill(b->geo, jobs, ops, n_ops);
    GUARD_END
}
int ftl_batch_set_join(ftl_batch *b, int join, float miter_limit) {
    if (!b) return bad("null batch");
    if (join < FTL_JOIN_MITER || join > FTL_JOIN_ROUND) return bad("unknown join style");
    b->join = join;
    b->miter_limit = miter_limit;
    return FTL_OK;
}

// n_jobs strokes in one pass: every job is flattened and outlined on host threads (stroker.rs:204-416 is sequential f32
// arithmetic over libm, a few microseconds per path), the outlines are concatenated and filled NonZero by ONE pass of
// the device pipeline - with transforms[j] applied a second time, exactly as Plotter::stroke does (plotter.rs:361-364).
int ftl_batch_stroke(ftl_batch *b, uint32_t n_jobs, const ftl_path_op *ops, const uint64_t *op_offsets, const float *transforms, const uint8_t *colors) {
    GUARD_BEGIN
    if (!b) return bad("null batch");
    if (n_jobs == 0) return FTL_OK;
    if (n_jobs > b->capacity) return bad("more jobs than rasters in the batch");
    if (!op_offsets) return bad("op_offsets is null");
    const size_t n_ops = (size_t)op_offsets[n_jobs];
    if (n_ops && !ops) return bad("ops is null");
    for (uint32_t j = 0; j < n_jobs; j++)
        if (op_offsets[j + 1] < op_offsets[j] || op_offsets[j + 1] > 0x7FFFFFFFull) return bad("op_offsets must be non-decreasing");
    int rc = check_finite_ops(ops, n_ops);
    if (rc) return rc;
    if (n_ops && stroke_on_device(n_ops, true)) {
        std::vector<float> opw(2 * n_ops), w1;
        std::vector<HostJob> jobs;
        if ((rc = batch_jobs(b, n_jobs, op_offsets, nullptr, transforms, colors, &jobs))) return rc;
        for (uint32_t j = 0; j < n_jobs; j++) {
            const size_t o0 = (size_t)op_offsets[j], jn = (size_t)(op_offsets[j + 1] - op_offsets[j]);
            stroke_widths(1.0f, ops + o0, jn, &w1);  // a new Plotter starts with pen width 1 (plotter.rs:112)
            if (jn) memcpy(opw.data() + 2 * o0, w1.data(), 2 * jn * sizeof(float));
        }
        bool needs_host = false;
        if ((rc = b->eng.stroke(b->geo, jobs, ops, n_ops, opw.data(), b->join, b->miter_limit, &needs_host))) return rc;
        if (!needs_host) return FTL_OK;
    }
    std::vector<std::vector<ftl_path_op>> outlines(n_jobs);
    StrokeParams sp;
    sp.join = b->join; sp.miter_limit = b->miter_limit; sp.tol_sq = b->tol_sq;
    auto work = [&](uint32_t j0, uint32_t j1) {
        std::vector<float> opw;
        WideFlat flat;
        static const float ident[6] = {1, 0, 0, 0, 1, 0};
        for (uint32_t j = j0; j < j1; j++) {
            const ftl_path_op *jo = ops + op_offsets[j];
            const size_t jn = (size_t)(op_offsets[j + 1] - op_offsets[j]);
            stroke_widths(1.0f, jo, jn, &opw);  // a new Plotter starts with pen width 1 (plotter.rs:112)
            flatten_wide_host(transforms ? transforms + 6 * (size_t)j : ident, b->tol_sq, jo, jn, opw.data(), &flat);
            stroke_outline(sp, jo, jn, flat, &outlines[j]);
        }
    };
    unsigned nt = n_ops < 4096 ? 1u : std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    nt = std::min<unsigned>(nt, n_jobs);
    if (nt <= 1) work(0, n_jobs);
    else {
        std::vector<std::thread> th;
        const uint32_t per = (n_jobs + nt - 1) / nt;
        for (unsigned t = 0; t < nt; t++) {
            const uint32_t j0 = std::min(n_jobs, t * per), j1 = std::min(n_jobs, j0 + per);
            if (j0 < j1) th.emplace_back(work, j0, j1);
        }
        for (std::thread &t : th) t.join();
    }
    std::vector<uint64_t> offs(n_jobs + 1, 0);
    for (uint32_t j = 0; j < n_jobs; j++) offs[j + 1] = offs[j] + outlines[j].size();
    std::vector<ftl_path_op> all((size_t)offs[n_jobs]);
    for (uint32_t j = 0; j < n_jobs; j++)
        if (!outlines[j].empty()) memcpy(all.data() + offs[j], outlines[j].data(), outlines[j].size() * sizeof(ftl_path_op));
    std::vector<HostJob> jobs;
    if ((rc = batch_jobs(b, n_jobs, offs.data(), nullptr, transforms, colors, &jobs))) return rc;
    return b->eng.fill(b->geo, jobs, all.data(), all.size());
    GUARD_END
}

int ftl_batch_upload(ftl_batch *b, uint32_t n_jobs, const ftl_path_op *ops, const uint64_t *op_offsets, const uint8_t *rules,
                     const float *transforms, const uint8_t *colors) {
    GUARD_BEGIN
    if (!b) return bad("null batch");
    std::vector<HostJob> jobs;
    int rc = batch_jobs(b, n_jobs, op_offsets, rules, transforms, colors, &jobs);
    if (rc) return rc;
    size_t n_ops = n_jobs ? (size_t)op_offsets[n_jobs] : 0;
    if (n_ops && !ops) return bad("ops is null");
    return b->eng.upload(b->geo, jobs, ops, n_ops);
    GUARD_END
}
int ftl_batch_run(ftl_batch *b) {
    GUARD_BEGIN
    if (!b) return bad("null batch");
    return b->eng.replay();
    GUARD_END
}
int ftl_batch_read(ftl_batch *b, uint32_t first, uint32_t count, void *dst, size_t nbytes) {
    GUARD_BEGIN
    if (!b || (!dst && nbytes)) return bad("null argument");
    if ((uint64_t)first + count > b->capacity) return bad("raster range outside the batch");
    if (nbytes != b->geo.bytes() * count) return bad("nbytes does not match count*height*width*bpp");
    if (!nbytes) return b->eng.sync();
    return b->eng.copy_out(dst, (uint8_t *)b->rasters + b->geo.bytes() * first, nbytes);
    GUARD_END
}
int ftl_batch_checksums(ftl_batch *b, uint32_t first, uint32_t count, uint64_t *out) {
    GUARD_BEGIN
    if (!b || (!out && count)) return bad("null argument");
    if ((uint64_t)first + count > b->capacity) return bad("raster range outside the batch");
    return b->eng.checksums((uint8_t *)b->rasters + b->geo.bytes() * first, b->geo.bytes(), count, out);
    GUARD_END
}
int ftl_batch_sync(ftl_batch *b) {
    GUARD_BEGIN
    if (!b) return bad("null batch");
    return b->eng.sync();
    GUARD_END
}
int ftl_batch_device_ptr(ftl_batch *b, void **dptr, size_t *nbytes) {
    if (!b || !dptr) return bad("null argument");
    int rc = b->eng.sync();  // fills issued so far (and any deferred repeat of an overflowed replay) are complete
    if (rc) return rc;
    *dptr = b->rasters;
    if (nbytes) *nbytes = b->geo.bytes() * b->capacity;
    return FTL_OK;
}

int ftl_batch_stream(ftl_batch *b, void **stream) {
    GUARD_BEGIN
    if (!b || !stream) return bad("null argument");
    int rc = b->eng.sync();  // forces lazy stream creation
    if (rc) return rc;
    *stream = b->eng.stream();
    return FTL_OK;
    GUARD_END
}
int ftl_stream(ftl_plotter *p, void **stream) {
    GUARD_BEGIN
    if (!p || !stream) return bad("null argument");
    int rc = p->eng.sync();
    if (rc) return rc;
    *stream = p->eng.stream();
    return FTL_OK;
    GUARD_END
}

// ---- instrumentation ----
uint64_t ftl_launch_count(void) { return Engine::launch_count(); }
int ftl_transfer_bytes(int reset, uint64_t *h2d, uint64_t *d2h) {
    Engine::transfer_bytes(reset != 0, h2d, d2h);
    return FTL_OK;
}
int ftl_set_profiling(int enabled) {
    Engine::set_profiling(enabled != 0);
    return FTL_OK;
}
int ftl_tile_kernel_time(int reset, double *ms, uint64_t *launches) {
    Engine::tile_kernel_time(reset != 0, ms, launches);
    return FTL_OK;
}
int ftl_plotter_tile_kernel_time(ftl_plotter *p, int reset, double *ms, uint64_t *launches) {
    if (!p) return bad("null plotter");
    p->eng.tile_time(reset != 0, ms, launches);
    return FTL_OK;
}
int ftl_batch_tile_kernel_time(ftl_batch *b, int reset, double *ms, uint64_t *launches) {
    if (!b) return bad("null batch");
    b->eng.tile_time(reset != 0, ms, launches);
    return FTL_OK;
}

// Per-call latency of Plotter::fill through this ABI, timed inside the library so that no binding overhead is counted:
// `iters` calls of ftl_fill (+ ftl_sync after each one when sync_each != 0, one ftl_sync at the end otherwise).
int ftl_time_fills(ftl_plotter *p, int rule, const ftl_path_op *ops, size_t n_ops, const uint8_t *color, uint32_t iters, int sync_each,
                   double *us_per_call) {
    GUARD_BEGIN
    if (!p || !us_per_call) return bad("null argument");
    int rc = p->eng.sync();
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < iters; i++) {
        if ((rc = plot_fill(p, rule, ops, n_ops, color))) return rc;
        if (sync_each && (rc = p->eng.sync())) return rc;
    }
    if ((rc = p->eng.sync())) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    *us_per_call = iters ? std::chrono::duration<double, std::micro>(t1 - t0).count() / iters : 0.0;
    return FTL_OK;
    GUARD_END
}

int ftl_batch_debug_top_rows(ftl_batch *b, uint32_t first, uint32_t count, int32_t *top_rows) {
    GUARD_BEGIN
    if (!b || (!top_rows && count)) return bad("null argument");
    if (!count) return FTL_OK;
    return b->eng.job_top_rows(first, count, top_rows);
    GUARD_END
}
int ftl_debug_small_profile(ftl_plotter *p, int64_t stamps[9]) {
    GUARD_BEGIN
    if (!p || !stamps) return bad("null argument");
    long long t[9];
    int rc = p->eng.small_profile(t);
    for (int k = 0; k < 9; k++) stamps[k] = t[k];
    return rc;
    GUARD_END
}

// ---- parity probes ----
int ftl_debug_flatten(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, int32_t *xy, size_t cap, size_t *n_points, uint32_t *subs,
                      size_t sub_cap, size_t *n_subs) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    std::vector<int32_t> v;
    std::vector<uint32_t> s;
    int rc = p->eng.debug_flatten(p->e, p->tol_sq, ops, n_ops, &v, &s);
    if (rc) return rc;
    size_t np = v.size() / 2, ns = s.size() / 2;
    if (n_points) *n_points = np;
    if (n_subs) *n_subs = ns;
    if (xy) memcpy(xy, v.data(), sizeof(int32_t) * 2 * (np < cap ? np : cap));
    if (subs) memcpy(subs, s.data(), sizeof(uint32_t) * 2 * (ns < sub_cap ? ns : sub_cap));
    return FTL_OK;
    GUARD_END
}
int ftl_debug_edges(ftl_plotter *p, int32_t *rec, size_t cap, size_t *n_edges) {
    GUARD_BEGIN
    if (!p || !n_edges || (!rec && cap)) return bad("null argument");
    std::vector<int32_t> v;
    int rc = p->eng.debug_edges(&v);
    if (rc) return rc;
    *n_edges = v.size() / 6;
    memcpy(rec, v.data(), std::min(v.size(), cap * 6) * sizeof(int32_t));
    return FTL_OK;
    GUARD_END
}
int ftl_debug_area(ftl_plotter *p, int32_t row, int16_t *area, size_t width) {
    GUARD_BEGIN
    if (!p || (!area && width)) return bad("null argument");
    if (width != p->geo.width) return bad("width does not match the raster");
    std::vector<int16_t> a;
    int rc = p->eng.debug_area(row, (uint32_t)width, &a);
    if (rc) return rc;
    if (width) memcpy(area, a.data(), width * sizeof(int16_t));
    return FTL_OK;
    GUARD_END
}
int ftl_debug_last_fill(ftl_plotter *p, int32_t info[3]) {
    GUARD_BEGIN
    if (!p || !info) return bad("null argument");
    FillInfo fi;
    int rc = p->eng.last_fill_info(&fi);
    if (rc) return rc;
    info[0] = fi.dir; info[1] = fi.top_row; info[2] = (int32_t)fi.n_points;
    return FTL_OK;
    GUARD_END
}
int ftl_debug_stroke_ops(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, ftl_path_op *out, size_t cap, size_t *n_out) {
    GUARD_BEGIN
    if (!p) return bad("null plotter");
    float keep = p->s_width;
    std::vector<ftl_path_op> outline;
    int rc = stroke_ops(p, ops, n_ops, &outline);
    p->s_width = keep;  // a probe must not change plotter state
    if (rc) return rc;
    if (n_out) *n_out = outline.size();
    if (out) memcpy(out, outline.data(), sizeof(ftl_path_op) * (outline.size() < cap ? outline.size() : cap));
    return FTL_OK;
    GUARD_END
}
// The outline as the DEVICE stroker builds it (stroke_kernels.cuh); *fell_back = 1 when it declined (the host stroker
// would take the call) and nothing was written.
int ftl_debug_stroke_ops_device(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, ftl_path_op *out, size_t cap, size_t *n_out, int *fell_back) {
    GUARD_BEGIN
    if (!p || (n_ops && !ops)) return bad("null argument");
    std::vector<float> opw;
    stroke_widths(p->s_width, ops, n_ops, &opw);
    std::vector<HostJob> jobs(1);
    jobs[0].op_begin = 0; jobs[0].op_end = (uint32_t)n_ops;
    memcpy(jobs[0].e, p->e, sizeof(jobs[0].e));
    jobs[0].tol_sq = p->tol_sq;
    jobs[0].raster = p->raster;
    bool needs_host = false;
    std::vector<ftl_path_op> outline;
    int rc = p->eng.stroke(p->geo, jobs, ops, n_ops, opw.data(), p->join, p->miter_limit, &needs_host, &outline);
    if (rc) return rc;
    if (fell_back) *fell_back = needs_host ? 1 : 0;
    if (n_out) *n_out = outline.size();
    if (out) memcpy(out, outline.data(), sizeof(ftl_path_op) * (outline.size() < cap ? outline.size() : cap));
    return FTL_OK;
    GUARD_END
}
// The sub-stroke table the device stroker starts from (stroke_sub_table, stroker.cpp): 4 words per sub-stroke = first
// drawing op, one past the last, joined, job (0).  Pure host code.
int ftl_debug_stroke_subs(const ftl_path_op *ops, size_t n_ops, uint32_t *out, size_t cap, size_t *n_subs) {
    GUARD_BEGIN
    if ((n_ops && !ops) || !n_subs) return bad("null argument");
    if (n_ops >= 0x7FFFFFFFull) return bad("too many ops");
    std::vector<uint32_t> subs, op_sub(n_ops ? n_ops : 1);
    stroke_sub_table(ops, 0, (uint32_t)n_ops, 0, &subs, op_sub.data());
    *n_subs = subs.size() / 4;
    if (out) memcpy(out, subs.data(), sizeof(uint32_t) * 4 * (*n_subs < cap ? *n_subs : cap));
    return FTL_OK;
    GUARD_END
}
// The strict Vid(u16) intake alone (ftl_set_strict_vid): the Move / Line ops, under the identity transform, that the device
// would rasterise for this fill; *capped = 0 (and nothing written) when the fill stays below the 65 535-point cap and takes
// the ordinary path.  Pure host code.
int ftl_debug_strict_intake(const float e[6], float tolerance, const ftl_path_op *ops, size_t n_ops, ftl_path_op *out, size_t cap, size_t *n_out,
                            int *capped) {
    GUARD_BEGIN
    if (!e || (n_ops && !ops) || !n_out || !capped) return bad("null argument");
    int rc = check_finite_ops(ops, n_ops);
    if (rc) return rc;
    const float tol = tolerance > 0.01f ? tolerance : 0.01f;
    std::vector<ftl_path_op> v;
    *capped = strict_intake(e, tol * tol, ops, n_ops, &v) ? 1 : 0;
    *n_out = *capped ? v.size() : 0;
    if (*capped && out) memcpy(out, v.data(), sizeof(ftl_path_op) * (v.size() < cap ? v.size() : cap));
    return FTL_OK;
    GUARD_END
}
// Pin of libm_compat.cuh: n random (y, x) pairs per input class through hypotf_glibc / atan2f_glibc and through this
// host's libm; counts the results that differ in any bit.  Runs on the CPU.
int ftl_debug_libm_selftest(uint64_t n, uint64_t seed, uint64_t *hypot_mismatches, uint64_t *atan2_mismatches, uint64_t *sin_mismatches,
                            uint64_t *sin_undecided) {
    GUARD_BEGIN
    uint64_t s = seed ? seed : 88172645463325252ull, bad_h = 0, bad_a = 0;
    auto rnd = [&s]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    auto draw = [&rnd](int mode) -> float {
        if (mode == 0) return (float)((int64_t)(rnd() % 2000000) - 1000000) / 37.0f;            // pixel-scale coordinates
        if (mode == 1) return (float)((int64_t)(rnd() % 4000) - 2000) / 8.0f;                    // many exact ties / axis-aligned
        uint32_t u = (uint32_t)rnd();                                                            // any finite bit pattern
        if (mode == 2) u = (u & 0x807FFFFFu) | ((100u + (uint32_t)(rnd() % 56)) << 23);          // mid exponents
        float f;
        memcpy(&f, &u, 4);
        return f;
    };
    for (int mode = 0; mode < 4; mode++)
        for (uint64_t i = 0; i < n; i++) {
            const float y = draw(mode), x = draw(mode);
            const float h0 = hypotf(x, y), h1 = libm::hypotf_glibc(x, y);
            const float a0 = atan2f(y, x), a1 = libm::atan2f_glibc(y, x);
            if (libm::f2u(h0) != libm::f2u(h1) && !(h0 != h0 && h1 != h1)) bad_h++;
            if (libm::f2u(a0) != libm::f2u(a1) && !(a0 != a0 && a1 != a1)) bad_a++;
        }
    // the two |sin| comparisons of the miter join: every float within 2e-3 of pi/2 (both signs) for `sm < 1`, random
    // angles and thresholds for `sm >= sm_min`; a prediction of 0 / 1 must agree with the host's sinf
    uint64_t bad_s = 0, undecided = 0;
    {
        float lo = 1.5707963705e+00f - 2.0e-3f, hi = 1.5707963705e+00f + 2.0e-3f;
        for (float x = lo; x <= hi; x = nextafterf(x, 4.0f))
            for (int sg = 0; sg < 2; sg++) {
                const float v = sg ? -x : x;
                const int p = libm::abs_sin_lt_one(v);
                if (p < 0) undecided++;
                else if ((fabsf(sinf(v)) < 1.0f) != (p == 1)) bad_s++;
            }
        for (uint64_t i = 0; i < n; i++) {
            const float x = (float)((double)(rnd() % 2000001) / 1000000.0 - 1.0) * 1.5707964f;
            const float t = i & 1 ? (float)((double)(rnd() % 1000001) / 1000000.0) : fabsf(sinf(x)) + (float)((int)(rnd() % 9) - 4) * 3.0e-8f;
            const int p = libm::abs_sin_ge(x, t);
            if (p < 0) undecided++;
            else if ((fabsf(sinf(x)) >= t) != (p == 1)) bad_s++;
        }
    }
    if (hypot_mismatches) *hypot_mismatches = bad_h;
    if (atan2_mismatches) *atan2_mismatches = bad_a;
    if (sin_mismatches) *sin_mismatches = bad_s;
    if (sin_undecided) *sin_undecided = undecided;
    return FTL_OK;
    GUARD_END
}
int ftl_debug_stroke_outline(int join, float miter_limit, float tol_sq, const ftl_path_op *ops, size_t n_ops, const uint32_t *counts,
                             const float *xyw, ftl_path_op *out, size_t cap, size_t *n_out) {
    GUARD_BEGIN
    if (n_ops && (!ops || !counts)) return bad("null argument");
    WideFlat flat;
    flat.counts.assign(counts, counts + n_ops);
    size_t np = 0;
    for (size_t i = 0; i < n_ops; i++) np += counts[i];
    if (np && !xyw) return bad("xyw is null");
    flat.xyw.assign(xyw, xyw + 3 * np);
    StrokeParams sp;
    sp.join = join; sp.miter_limit = miter_limit; sp.tol_sq = tol_sq;
    std::vector<ftl_path_op> outline;
    stroke_outline(sp, ops, n_ops, flat, &outline);
    if (n_out) *n_out = outline.size();
    if (out) memcpy(out, outline.data(), sizeof(ftl_path_op) * (outline.size() < cap ? outline.size() : cap));
    return FTL_OK;
    GUARD_END
}
int ftl_debug_accumulate(int rule, const int16_t *src, uint8_t *dst, size_t n, size_t rows, int device) {
    GUARD_BEGIN
    if ((!src || !dst) && n && rows) return bad("null argument");
    Engine eng(device);
    return eng.accumulate_rows(rule, src, dst, n, rows);
    GUARD_END
}

}  // extern "C"
