/* footile_b200.h — C ABI of the B200-native footile hot path.
 *
 * The reference (DougLau/footile, Rust) has no FFI: its seams are Rust method
 * calls.  Every entry point below names the reference interface it replaces
 * (file:line under the reference tree).  A Rust `extern "C"` shim binding
 * these is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this ABI;
 *   - every function returns an ftl_status (0 = OK) and never unwinds;
 *     ftl_last_error() returns a thread-local message for the last failure;
 *   - the raster lives in device memory (HBM) owned by the handle; host
 *     pixels are copied in at creation / ftl_write_raster and out at
 *     ftl_read_raster (the only blocking points besides ftl_sync);
 *   - a handle is used from one thread at a time (the reference takes
 *     `&mut self` on every drawing call); distinct handles are independent;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with FTL_ERR_NO_DEVICE.
 *
 * Deliberate deviations from the reference (all towards "draw what was asked"):
 *   - vertex ids are 32-bit: the reference stores them in a u16 (src/vid.rs:20-24) and
 *     Fig::add_point silently drops every point once 65 535 are stored (src/fig.rs:430); fills here
 *     render all points (up to 2^31 - 1 per call).  The stroker keeps the reference's cap per
 *     stroke (src/stroker.rs:63).  Inputs that stay below 65 535 flattened points per fill - every
 *     example and benchmark of the reference - are unaffected; ftl_set_strict_vid(p, 1) restores the cap;
 *   - curve subdivision is capped at depth 16 (the reference recurses without bound, README.md:32-33);
 *   - NaN/Inf coordinates are rejected with FTL_ERR_NONFINITE instead of recursing forever.
 */
#ifndef FOOTILE_B200_H
#define FOOTILE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTL_ABI_VERSION 1

/* ---- vocabulary -------------------------------------------------------- */

/* PathOp (src/path.rs:18-31). v holds up to three points (x,y) or the pen width:
 *   Close: -, Move/Line: v[0..1], Quad: v[0..3], Cubic: v[0..5], PenWidth: v[0]. */
typedef struct ftl_path_op {
    uint32_t tag;
    float v[6];
} ftl_path_op;

enum ftl_op_tag { FTL_OP_CLOSE = 0, FTL_OP_MOVE = 1, FTL_OP_LINE = 2, FTL_OP_QUAD = 3, FTL_OP_CUBIC = 4, FTL_OP_PENWIDTH = 5 };

/* FillRule (src/path.rs:9-14) */
enum ftl_fill_rule { FTL_NONZERO = 0, FTL_EVENODD = 1 };

/* JoinStyle (src/stroker.rs:13-20); the miter limit travels beside it. */
enum ftl_join { FTL_JOIN_MITER = 0, FTL_JOIN_BEVEL = 1, FTL_JOIN_ROUND = 2 };

/* Pixel formats of the generic bound P: Pixel<Chan=Ch8, Alpha=Premultiplied,
 * Gamma=Linear> (src/plotter.rs:38-41) that the reference is used with:
 * pix::matte::Matte8 (1 B), pix::gray::Graya8p (2 B: gray, alpha),
 * pix::rgb::Rgba8p (4 B: r, g, b, a).  Row-major, pitch = width * bpp. */
enum ftl_format { FTL_MATTE8 = 0, FTL_GRAYA8P = 1, FTL_RGBA8P = 2 };

typedef enum ftl_status {
    FTL_OK = 0,
    FTL_ERR_INVALID = 1,     /* bad argument */
    FTL_ERR_NO_DEVICE = 2,   /* no CUDA device / driver: there is no CPU fallback */
    FTL_ERR_CUDA = 3,        /* CUDA runtime failure, see ftl_last_error() */
    FTL_ERR_NOMEM = 4,
    FTL_ERR_NONFINITE = 5,   /* NaN/Inf in a path op or transform (the reference recurses without bound: README.md:32-33) */
    FTL_ERR_TOO_WIDE = 6     /* raster row does not fit the shared-memory row tile */
} ftl_status;

typedef struct ftl_plotter ftl_plotter; /* Plotter<P> (src/plotter.rs:38-56) */
typedef struct ftl_batch ftl_batch;     /* N independent Plotter<P>s of one size, driven together */

/* ---- library ----------------------------------------------------------- */
int ftl_abi_version(void);
const char *ftl_last_error(void);
int ftl_device_count(int *count);

/* ---- Plotter (src/plotter.rs:96-380) ------------------------------------ */

/* Plotter::new(raster) (plotter.rs:96-115).  init_pixels: width*height*bpp host
 * bytes (Raster::with_pixels / with_color), or NULL for Raster::with_clear. */
int ftl_plotter_new(uint32_t width, uint32_t height, int format, const void *init_pixels, int device,
                    ftl_plotter **out);
/* Row-band variant for one raster split across GPUs: this handle owns rows
 * [row_begin, row_end) of a width x height raster; init_pixels / read_raster
 * cover only those rows.  Extension beyond the reference (which only loops
 * rows: fig.rs:539); results equal the same rows of the unsplit fill. */
int ftl_plotter_new_band(uint32_t width, uint32_t height, uint32_t row_begin, uint32_t row_end, int format,
                         const void *init_pixels, int device, ftl_plotter **out);
int ftl_plotter_free(ftl_plotter *p);                       /* drop(Plotter) */
uint32_t ftl_width(const ftl_plotter *p);                   /* Plotter::width  (plotter.rs:118-120) */
uint32_t ftl_height(const ftl_plotter *p);                  /* Plotter::height (plotter.rs:123-125) */
int ftl_set_tolerance(ftl_plotter *p, float t);             /* Plotter::set_tolerance, clamped >= 0.01 (plotter.rs:133-137) */
int ftl_set_transform(ftl_plotter *p, const float e[6]);    /* Plotter::set_transform (plotter.rs:140-143); e = pointy Transform rows [a b tx; c d ty] */
int ftl_set_join(ftl_plotter *p, int join, float miter_limit); /* Plotter::set_join (plotter.rs:158-161) */
/* Strict Vid(u16) mode, off by default: ftl_fill then reproduces the reference's vertex cap - Fig::add_point ignores
 * points while 65 535 are stored (fig.rs:428-442, vid.rs:20-24).  Fills that stay below the cap are unaffected; one
 * that reaches it has its point intake replayed on the host (sequential by nature: a Fig::close can pop a point and
 * make room again) before the device rasterises the surviving points.  Applies to ftl_fill / ftl_fill_upload. */
int ftl_set_strict_vid(ftl_plotter *p, int enabled);
float ftl_pen_width(const ftl_plotter *p);                  /* the persistent s_width (plotter.rs:53,151-153) */

/* Plotter::fill(rule, ops, clr) (plotter.rs:339-350).  color: bpp bytes
 * (premultiplied); ignored for FTL_MATTE8 exactly as the reference ignores it
 * (fig.rs:632-636).  Asynchronous on the handle's stream. */
int ftl_fill(ftl_plotter *p, int rule, const ftl_path_op *ops, size_t n_ops, const uint8_t *color);
/* Plotter::stroke(ops, clr) (plotter.rs:356-365).  The stroker (stroker.rs:204-416) runs on the device for strokes of
 * 512 ops or more (flatten with widths -> outline -> fill, the outline never leaves HBM; one host synchronisation) and
 * on the host for smaller ones, where it costs microseconds; both produce the same outline bit for bit (the device
 * restates glibc's hypotf / atan2f exactly and hands the call to the host stroker when a |sin| comparison of a miter
 * join is too close to call).  FTL_DEVICE_STROKE=1 / 0 in the environment forces one or the other. */
int ftl_stroke(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, const uint8_t *color);

/* A scene: n_layers fills drawn IN ORDER onto the plotter's raster by one pass of the device
 * pipeline — the same pixels as n_layers ftl_fill calls (callers such as examples/fishy.rs:29-31 issue
 * several fill/stroke calls per raster).  Layer l = ops[op_offsets[l] .. op_offsets[l+1]) with
 * rules[l] (NonZero if NULL) and colors[4*l ..] (opaque white if NULL).  A stroke becomes a layer
 * through ftl_stroke_outline (its outline, filled NonZero). */
int ftl_fill_layers(ftl_plotter *p, uint32_t n_layers, const ftl_path_op *ops, const uint64_t *op_offsets,
                    const uint8_t *rules, const uint8_t *colors);
/* The outline Plotter::stroke would fill (stroker.rs:239-247 after plotter.rs:361-363): writes up to
 * cap ops, returns the count in *n_out; updates the persistent pen width exactly like ftl_stroke. */
int ftl_stroke_outline(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, ftl_path_op *out, size_t cap,
                       size_t *n_out);

/* Plotter::raster() / into_raster() (plotter.rs:368-380): synchronise and copy
 * the owned rows to host memory.  nbytes must equal rows*width*bpp. */
int ftl_read_raster(ftl_plotter *p, void *dst, size_t nbytes);
/* The owned rows converted for output, as the reference's examples do before encoding a PNG:
 * Raster::<SRgba8>::with_raster(&p.raster()) (examples/fishy.rs:33) for FTL_RGBA8P - colour / alpha, then the sRGB
 * transfer function, alpha copied -, SGraya8 likewise for FTL_GRAYA8P, and the byte-for-byte SGray8 view of a
 * FTL_MATTE8 raster (examples/png/mod.rs:22-27).  The conversion runs on the device; same size as ftl_read_raster. */
int ftl_read_raster_srgb(ftl_plotter *p, void *dst, size_t nbytes);
/* Plotter::raster_mut() (plotter.rs:373-375): replace the owned rows. */
int ftl_write_raster(ftl_plotter *p, const void *src, size_t nbytes);
int ftl_sync(ftl_plotter *p);
/* Zero-copy interop: device pointer of the owned rows (valid until free). */
int ftl_raster_device_ptr(ftl_plotter *p, void **dptr, size_t *nbytes);

/* ---- Batch: many fills per launch -------------------------------------- */
/* The reference draws one path per Plotter::fill call on one core; callers
 * that draw many independent paths (benches/fishyb.rs:18-20 allocates a raster
 * and a plotter per iteration) loop.  A batch is `capacity` rasters of one
 * size and format, resident in HBM, filled by ONE pass of the device pipeline.
 * Job j = ops[op_offsets[j] .. op_offsets[j+1]) drawn into raster j with
 * rules[j] (or rule 0 if NULL), transforms[6*j..] (or the identity if NULL),
 * colors[4*j..] (or opaque white if NULL).  Equivalent to n_jobs independent
 * `Plotter::new(raster_j).set_transform(..).fill(..)` calls. */
int ftl_batch_new(uint32_t width, uint32_t height, int format, uint32_t capacity, int device, ftl_batch **out);
int ftl_batch_free(ftl_batch *b);
int ftl_batch_set_tolerance(ftl_batch *b, float t);
/* Raster::with_clear for rasters [first, first+count) (zero them). */
int ftl_batch_clear(ftl_batch *b, uint32_t first, uint32_t count);
int ftl_batch_fill(ftl_batch *b, uint32_t n_jobs, const ftl_path_op *ops, const uint64_t *op_offsets,
                   const uint8_t *rules, const float *transforms, const uint8_t *colors);
/* Plotter::set_join for every stroke of the batch (plotter.rs:158-161). */
int ftl_batch_set_join(ftl_batch *b, int join, float miter_limit);
/* n_jobs independent `Plotter::new(raster_j).set_transform(..).stroke(ops_j, color_j)` calls (plotter.rs:356-365) in one
 * pass: flatten with widths, outline (stroker.rs:204-416) and NonZero fill all run on the device, one thread per
 * (point, side) of the outline; the host stroker (threads) takes batches of fewer than 2048 ops (microseconds of work
 * that overlap the device), and any batch the device declines (see ftl_stroke) or when FTL_DEVICE_STROKE=0.  Every job
 * starts with pen width 1 like a new Plotter. */
int ftl_batch_stroke(ftl_batch *b, uint32_t n_jobs, const ftl_path_op *ops, const uint64_t *op_offsets, const float *transforms,
                     const uint8_t *colors);
/* Copy rasters [first, first+count) to host (blocking). */
int ftl_batch_read(ftl_batch *b, uint32_t first, uint32_t count, void *dst, size_t nbytes);
/* 64-bit FNV-1a of each raster's bytes, computed on the device (blocking). */
int ftl_batch_checksums(ftl_batch *b, uint32_t first, uint32_t count, uint64_t *out);
int ftl_batch_sync(ftl_batch *b);
int ftl_batch_device_ptr(ftl_batch *b, void **dptr, size_t *nbytes);
/* The cudaStream_t every call on this handle is issued on (for event timing
 * and for ordering a caller's own kernels after a fill). */
int ftl_batch_stream(ftl_batch *b, void **stream);
int ftl_stream(ftl_plotter *p, void **stream);

/* ---- Several GPUs behind one handle (one host thread per device, no data-path collective) ---------------- */
/* The two ways the path shards (BASELINE north_star): independent paths / rasters in contiguous blocks of jobs over the
 * devices, and ONE raster split into row bands with every device culling the sub-figures outside its band.  One process
 * per GPU (how bench.py runs) needs only ftl_shard_range / ftl_band_rows and the per-device entry points above. */
typedef struct ftl_ctx ftl_ctx;
/* devices: n_devices CUDA device indices (NULL: 0..n_devices-1; n_devices <= 0: every device).  An index may repeat. */
int ftl_ctx_new(int n_devices, const int *devices, ftl_ctx **out);
int ftl_ctx_free(ftl_ctx *c);
int ftl_ctx_size(const ftl_ctx *c);
/* Contiguous block of range(n) owned by `rank` of `world`. */
int ftl_shard_range(uint32_t n, uint32_t rank, uint32_t world, uint32_t *first, uint32_t *count);
/* Row band [row_begin, row_end) of `rank`; band boundaries are multiples of `align` rows (0: 32, the binned kernel's band). */
int ftl_band_rows(uint32_t height, uint32_t rank, uint32_t world, uint32_t align, uint32_t *row_begin, uint32_t *row_end);
/* n_jobs independent fills (the arguments of ftl_batch_fill) sharded over the devices; every raster (cleared first) comes
 * back in dst, job after job.  tolerance <= 0: the default 0.3. */
int ftl_ctx_fill_batch(ftl_ctx *c, uint32_t width, uint32_t height, int format, float tolerance, uint32_t n_jobs, const ftl_path_op *ops,
                       const uint64_t *op_offsets, const uint8_t *rules, const float *transforms, const uint8_t *colors, void *dst, size_t nbytes);
/* One fill of one width x height raster, rows split into one band per device; init_pixels (NULL: clear) / dst cover the
 * whole raster.  Equal, byte for byte, to the unsplit Plotter::fill (fig.rs:497,539 only loops rows). */
int ftl_ctx_fill_bands(ftl_ctx *c, uint32_t width, uint32_t height, int format, int rule, const ftl_path_op *ops, size_t n_ops, const float transform[6],
                       float tolerance, const uint8_t *color, const void *init_pixels, void *dst, size_t nbytes);

/* ---- Device-resident replay (bench `value` leg: inputs already in HBM) ---- */
/* Upload the jobs of a batch once; ftl_batch_run() then repeats the device
 * pipeline on the resident ops without touching host memory. */
int ftl_batch_upload(ftl_batch *b, uint32_t n_jobs, const ftl_path_op *ops, const uint64_t *op_offsets,
                     const uint8_t *rules, const float *transforms, const uint8_t *colors);
int ftl_batch_run(ftl_batch *b);

/* The same for one Plotter: ftl_fill_upload validates the ops and makes the fill resident (transform, tolerance,
 * rule and colour as set at this call) without drawing; every ftl_fill_replay() then draws it again with no host
 * traffic at all - Plotter::fill (plotter.rs:339-350) of a path that already lives in HBM.  Config 5 sends 291 MB
 * of ops per fill otherwise. */
int ftl_fill_upload(ftl_plotter *p, int rule, const ftl_path_op *ops, size_t n_ops, const uint8_t *color);
int ftl_fill_replay(ftl_plotter *p);

/* ---- Instrumentation ---------------------------------------------------- */
/* Number of kernel launches issued by this library since load (all handles). */
uint64_t ftl_launch_count(void);
/* Bytes of path data and pixels this library has moved over PCIe since load / the last reset (all handles): what
 * really crosses the wire - the packed read-back of ftl_read_raster / ftl_batch_read moves far less than the raster. */
int ftl_transfer_bytes(int reset, uint64_t *h2d, uint64_t *d2h);
/* Device time (ms) and launch count of the raster-tile kernel (scatter + row
 * scan + fill rule + store/blend) accumulated since the last call with
 * reset != 0.  Measured with CUDA events on the handle's stream when
 * profiling is enabled via ftl_set_profiling(1) (adds two event records per
 * launch; off by default). */
int ftl_set_profiling(int enabled);
int ftl_tile_kernel_time(int reset, double *ms, uint64_t *launches);          /* summed over the live handles of the process */
/* The same for one handle: the timing state lives in the handle, so handles profiled at the same time do not mix. */
int ftl_plotter_tile_kernel_time(ftl_plotter *p, int reset, double *ms, uint64_t *launches);
int ftl_batch_tile_kernel_time(ftl_batch *b, int reset, double *ms, uint64_t *launches);

/* Per-call latency of ftl_fill (benches/fishyb.rs:10-39 times exactly this call), measured inside the library so that
 * no binding overhead is counted: iters calls, each followed by ftl_sync when sync_each != 0 (otherwise one at the end). */
int ftl_time_fills(ftl_plotter *p, int rule, const ftl_path_op *ops, size_t n_ops, const uint8_t *color, uint32_t iters, int sync_each,
                   double *us_per_call);

/* ---- Parity probes (used by tests only) --------------------------------- */
/* Flattened Fixed points of a fill after point intake (fig.rs:428-461):
 * returns the count via *n_points and writes up to cap (x,y) i32 pairs; subs
 * receives up to sub_cap (start,len) pairs.  Blocking. */
int ftl_debug_flatten(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, int32_t *xy, size_t cap,
                      size_t *n_points, uint32_t *subs, size_t sub_cap, size_t *n_subs);
/* Stage (c) probe: the signed-area deltas row `row` of the last ftl_fill receives from its edges BEFORE the prefix sum -
 * the contents of the reference's i16 area buffer when Fig::fill reaches accumulate for that row (fig.rs:285-302,
 * 536-573; plotter.rs:45).  `row` counts as the reference's scan loop does (row 0 = top_row when top_row < 0). */
int ftl_debug_area(ftl_plotter *p, int32_t row, int16_t *area, size_t width);
/* (dir, top_row, n_points) of the last fill (fig.rs:495-496); dir 0 = Forward. */
int ftl_debug_last_fill(ftl_plotter *p, int32_t info[3]);
/* Stage (b) probe: the edges of the last ftl_fill (Edge::new, fig.rs:179-210, with the winding sign of fig.rs:286), in no
 * particular order: 6 int32 per edge = x_bot, inv_slope, step_pix, y_upper, y_lower (Fixed 16.16), sign (+1/-1).
 * Writes min(*n_edges, cap) records. */
int ftl_debug_edges(ftl_plotter *p, int32_t *rec, size_t cap, size_t *n_edges);
/* top_row (fig.rs:496) of jobs [first, first + count) of the batch's last fill / run: the first row the reference would
 * resolve is max(top_row, 0).  INT32_MAX for a job that drew nothing. */
int ftl_batch_debug_top_rows(ftl_batch *b, uint32_t first, uint32_t count, int32_t *top_rows);
/* Development aid: with FTL_SMALL_PROF=1 in the environment, the SM clock (clock64) at the nine phase boundaries of the
 * last one-launch small fill: start, flatten, vertices, top vertex, edges, staged, scatter, resolve, end. */
int ftl_debug_small_profile(ftl_plotter *p, int64_t stamps[9]);
/* The outline ops Plotter::stroke hands to fill (stroker.rs:239-247). */
int ftl_debug_stroke_ops(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, ftl_path_op *out, size_t cap,
                         size_t *n_out);
/* The same outline as the DEVICE stroker builds it (stroke_kernels.cuh); *fell_back = 1 (and *n_out = 0) when it
 * declined and the host stroker would take the call.  The parity tests compare the two bit for bit. */
int ftl_debug_stroke_ops_device(ftl_plotter *p, const ftl_path_op *ops, size_t n_ops, ftl_path_op *out, size_t cap,
                                size_t *n_out, int *fell_back);
/* The sub-strokes of a path as Stroke::add_point / close form them (stroker.rs:204-236), which the device stroker takes
 * from the ops: 4 uint32 per sub-stroke = first drawing op, one past its last drawing op, joined (the last close() applied
 * to it was close(true)), 0.  Pure host code. */
int ftl_debug_stroke_subs(const ftl_path_op *ops, size_t n_ops, uint32_t *out, size_t cap, size_t *n_subs);
/* The host-side point intake of strict Vid(u16) mode alone (ftl_set_strict_vid; fig.rs:428-442,373-383): the Move / Line
 * ops, under the identity transform, that a fill reaching the 65 535-point cap hands to the device.  *capped = 0 and
 * *n_out = 0 when the fill stays below the cap.  Pure host code; needs no device. */
int ftl_debug_strict_intake(const float e[6], float tolerance, const ftl_path_op *ops, size_t n_ops, ftl_path_op *out,
                            size_t cap, size_t *n_out, int *capped);
/* Pin of the libm restatement the device stroker uses (csrc/libm_compat.cuh: glibc's hypotf and atan2f): 4 * n random
 * inputs through it and through this host's libm, counting results that differ in any bit; and the two three-valued
 * |sin| comparisons of the miter join against the host's sinf (wrong predictions, and how many were left undecided -
 * about half of the n `>=` probes sit within 1.2e-7 of their threshold on purpose).  Pure host code. */
int ftl_debug_libm_selftest(uint64_t n, uint64_t seed, uint64_t *hypot_mismatches, uint64_t *atan2_mismatches,
                            uint64_t *sin_mismatches, uint64_t *sin_undecided);
/* The host stroker alone (stroker.rs:204-416) on an already flattened wide
 * polyline: counts[i] points of op i, xyw = (x, y, width) per point.  Pure
 * host code; needs no device. */
int ftl_debug_stroke_outline(int join, float miter_limit, float tol_sq, const ftl_path_op *ops, size_t n_ops,
                             const uint32_t *counts, const float *xyw, ftl_path_op *out, size_t cap, size_t *n_out);
/* Row accumulate alone (imgbuf.rs:38-51,141-154): dst[i] = rule(prefix sum of
 * src[0..i]) over n i16 cells per row, rows independent.  Host buffers. */
int ftl_debug_accumulate(int rule, const int16_t *src, uint8_t *dst, size_t n, size_t rows, int device);

#ifdef __cplusplus
}
#endif
#endif /* FOOTILE_B200_H */
