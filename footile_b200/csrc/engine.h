// engine.h — host-side interface of the device pipeline (engine.cu).
// One Engine per handle: it owns a CUDA stream, a pinned staging area and a
// grow-only scratch arena in HBM, and runs the four stages of the hot path
//   (a) flatten   (b) edge prep + band binning   (c)+(d) tile raster kernel
// for a set of jobs that share one raster geometry.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/footile_b200.h"

namespace ftl {

struct Geometry {
    uint32_t width = 0, height = 0;      // full raster size
    uint32_t row_begin = 0, row_end = 0; // rows owned by this device
    int format = FTL_MATTE8;
    uint32_t bpp() const { return format == FTL_MATTE8 ? 1u : (format == FTL_GRAYA8P ? 2u : 4u); }
    uint32_t rows() const { return row_end - row_begin; }
    size_t pitch() const { return (size_t)width * bpp(); }
    size_t bytes() const { return pitch() * rows(); }
};

// One fill: ops[op_begin, op_end) drawn into the raster whose owned rows start
// at device address `raster`.
struct HostJob {
    uint32_t op_begin = 0, op_end = 0;
    float e[6] = {1, 0, 0, 0, 1, 0};
    float tol_sq = 0.09f;
    int rule = 0;
    uint8_t color[4] = {255, 255, 255, 255};
    void *raster = nullptr;
};

struct FillInfo {
    int32_t dir = 0, top_row = 0;
    uint32_t n_points = 0;
};

struct WideFlat {                 // output of the stroke-side flatten
    std::vector<uint32_t> counts; // points emitted per op
    std::vector<float> xyw;       // (x, y, w) per point
};

class Engine {
public:
    explicit Engine(int device);
    ~Engine();
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    int device() const { return device_; }
    void *stream() const { return stream_; }

    // Fill jobs.  If `ops` is null the ops uploaded by the previous call (or by
    // upload()) are reused from HBM.
    // allow_small: a single job of a few ops on a small raster may take the one-launch path (small_kernel.cuh).
    int fill(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, bool allow_small = false);
    // The jobs are layers drawn in order onto ONE raster (all jobs carry the same raster pointer):
    // flatten / edge prep / binning run once for all layers, the tile kernel once per layer.
    int fill_layers(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops);
    // Plotter::stroke for a set of jobs with the stroker on the device (stroke_kernels.cuh): flatten with widths, outline
    // and fill without the outline leaving HBM; one host synchronisation (the outline's size).  opw: per op (pen_w, s_width)
    // as stroke_widths() computes them per job.  *needs_host is set, and nothing is drawn, when the device could not
    // reproduce the host stroker's decisions (libm_compat.cuh): the caller then outlines on the host.
    // outline (optional probe): the outline ops and their per-job offsets are copied back instead of being filled.
    int stroke(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, const float *opw, int join,
               float miter_limit, bool *needs_host, std::vector<ftl_path_op> *outline = nullptr, std::vector<uint32_t> *outline_offsets = nullptr);
    // Upload only (device-resident replay).
    int upload(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, bool layered = false);
    int replay();

    // Stroke-side flatten: raw f32 points with widths, per op (blocking).
    // opw: per op (pen_w, s_width) computed by the host from the PenWidth ops.
    int flatten_wide(const float e[6], float tol_sq, const ftl_path_op *ops, size_t n_ops, const float *opw,
                     WideFlat *out);
    // Parity probes (blocking).
    int debug_flatten(const float e[6], float tol_sq, const ftl_path_op *ops, size_t n_ops,
                      std::vector<int32_t> *xy, std::vector<uint32_t> *subs);
    int last_fill_info(FillInfo *info);
    // top_row (fig.rs:496) of jobs [first, first + count) of the resident job set after its last run; INT32_MAX for an empty job.
    int job_top_rows(uint32_t first, uint32_t count, int32_t *out);
    // Probe: the edge records stage (b) built for job 0 of the last fill, 6 values per edge:
    // x_bot, inv_slope, step_pix, y_upper, y_lower (Fixed), sign (+1 / -1) (fig.rs:47-66,179-210,286).
    int debug_edges(std::vector<int32_t> *out);
    // Probe: the i16 signed-area deltas geometry row `row` receives from the edges of job 0 of the last fill (stage (c) alone).
    int debug_area(int32_t row, uint32_t width, std::vector<int16_t> *out);
    int accumulate_rows(int rule, const int16_t *src, uint8_t *dst, size_t n, size_t rows);
    int checksums(const void *rasters, size_t raster_bytes, uint32_t count, uint64_t *out);

    int sync();
    int small_profile(long long out[9]);  // clock64() stamps of the last small fill's phases (FTL_SMALL_PROF=1)
    int alloc_raster(size_t bytes, void **dptr);
    int free_raster(void *dptr);
    int memset_async(void *dptr, int value, size_t bytes);
    int copy_in(void *dptr, const void *src, size_t bytes);   // blocking
    int copy_out(void *dst, const void *dptr, size_t bytes);  // blocking
    int copy_out_srgb(void *dst, const void *dptr, size_t bytes, int format);  // blocking; converted for output (pack_kernels.cuh)
    static void srgb_encode_table(uint8_t t[256]);

    static int device_count(int *count);
    static uint64_t launch_count();
    static void transfer_bytes(bool reset, uint64_t *h2d, uint64_t *d2h);
    static void set_profiling(bool on);
    static void tile_kernel_time(bool reset, double *ms, uint64_t *launches);  // summed over the live engines of the process
    void tile_time(bool reset, double *ms, uint64_t *launches);                // this engine's tile launches only

    struct Impl;

private:
    Impl *impl_;
    int device_;
    void *stream_;
};

void set_error(const std::string &msg);
const char *last_error();

// Host-side stroker (stroker.cpp): outline ops of a flattened wide polyline.
// Mirrors Stroke::add_point / close / path_ops (src/stroker.rs:204-262).
struct StrokeParams {
    int join = FTL_JOIN_MITER;
    float miter_limit = 4.0f;
    float tol_sq = 0.09f;
};
void stroke_outline(const StrokeParams &sp, const ftl_path_op *ops, size_t n_ops, const WideFlat &flat,
                    std::vector<ftl_path_op> *out);
// The stroke-side flatten on the host (same bits as Engine::flatten_wide; stroker.cpp).
void flatten_wide_host(const float e[6], float tol_sq, const ftl_path_op *ops, size_t n_ops, const float *opw, WideFlat *out);
// Sub-strokes of ops[op_begin, op_end) as Stroke::add_point / close form them (stroker.rs:204-236): appends 4 words per
// sub-stroke to `subs` (first drawing op, one past the last, joined, job) and sets op_sub[i] for every drawing op.
void stroke_sub_table(const ftl_path_op *ops, uint32_t op_begin, uint32_t op_end, uint32_t job, std::vector<uint32_t> *subs, uint32_t *op_sub);
// Per-op (pen_w, s_width) from the PenWidth ops; returns the final s_width
// (plotter.rs:128-130,151-153,233-236,293-298).
float stroke_widths(float s_width, const ftl_path_op *ops, size_t n_ops, std::vector<float> *opw);

}  // namespace ftl
