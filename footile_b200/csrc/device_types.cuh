// device_types.cuh — device data layout and per-call parameters
// Included by engine.cu inside namespace ftl (one translation unit: the kernels share Params / EdgeRec / ...).
#pragma once

// ---------------------------------------------------------------------------
// device data layout (all arrays live in the engine's scratch arena in HBM)
// ---------------------------------------------------------------------------
constexpr uint32_t NONE32 = 0xFFFFFFFFu;
constexpr int MAX_DEPTH = 16;  // subdivision depth cap (the reference recurses without bound); 4^16 covers any in-range curve at tol 0.01

struct __align__(16) JobDesc {  // 64 B, host-filled
    uint32_t op_begin, op_end;
    float e[6];
    float tol_sq;
    uint32_t rule;
    uint32_t color;  // bytes r,g,b,a little-endian (or gray,alpha / alpha)
    uint32_t pad0;
    unsigned long long raster;  // device address of row `row_begin`
    unsigned long long pad1;
};
static_assert(sizeof(JobDesc) == 64, "JobDesc layout");

struct __align__(16) JobState {  // 48 B, device-written
    unsigned long long top_key;  // min over vertices of (y,x), sign-biased
    uint32_t top_vid;
    int32_t dir;        // 0 Forward, 1 Reverse (fig.rs:402-411)
    int32_t top_row;    // row_of(y of top-left vertex) (fig.rs:496)
    int32_t first_row;  // max(top_row, 0): first raster row the fill touches (fig.rs:497)
    int32_t shift;      // min(top_row, 0): geometry row r lands on raster row r - shift
    uint32_t vtx_begin, vtx_end;  // this job's vertex (= edge slot) range
    uint32_t pad[3];
};
static_assert(sizeof(JobState) == 48, "JobState layout");

struct __align__(16) Vtx {  // 16 B
    int32_t x, y;   // Fixed 16.16
    uint32_t sub;   // index of the first vertex of this vertex's sub-figure
    uint32_t job;
};

struct __align__(16) EdgeRec {  // 32 B: one per ring segment whose end points differ in y (fig.rs:47-66,179-201)
    int32_t x_bot0;     // X at the bottom of the edge's first row
    int32_t inv_slope;  // dx/dy
    int32_t step_pix;   // min(|dy/dx|, 1), 0 when vertical
    int32_t ry0, ry1;   // raster rows of the upper / lower vertex (geometry row - shift)
    uint32_t fr;        // fract(y_upper) | fract(y_lower) << 16
    uint32_t job;
    uint32_t flags;     // bit0 valid, bit1 set when the edge runs against the figure direction (sign -1, fig.rs:286)
};
constexpr uint32_t DIRECT_MAX = 64;  // upper bound of Params::direct_max: jobs with at most direct_max edge slots skip binning (raster_tiles scans the job's own edges); larger jobs go to raster_bins

struct __align__(8) SumHead {  // scan element over ops: vertex count + position of the last sub-figure head
    uint32_t sum, head;
};

struct Counters {
    uint32_t nv;         // vertices after intake
    uint32_t n_entries;  // (edge,band) pairs after binning
    uint32_t n_popped;   // closing vertices dropped by the sub-figure close rule (fig.rs:376-380)
    uint32_t overflow;   // a speculatively sized scratch buffer was too small: nothing was drawn, the host re-runs
    uint32_t need_v;     // vertices the call needed when it overflowed
    uint32_t need_e;     // bin entries the call needed when it overflowed
    uint32_t n_big;      // jobs with more than Params::direct_max edge slots (drawn by raster_bins)
    uint32_t pad;
};

struct Params {  // per-call constants, passed by value
    uint32_t W, H, row_begin, row_end;
    uint32_t fmt, bpp, pitch;
    uint32_t log2R, R, n_bands, WP, chunks;  // rows per tile, bands per job, smem row stride (cells), 512-cell chunks per row
    uint32_t n_jobs, n_ops, n_tiles;
    uint32_t win_chunks, warp_words, cta_warps;  // chunks per row window, smem words per warp, warps per CTA
    uint32_t win_rows;                           // rows of a narrow raster (one window per row) a warp holds at once
    uint32_t n_win;                              // windows per row of the direct kernel
    uint32_t all_direct;                         // host-proven: every job has <= DIRECT_MAX vertices (no binning needed)
    uint32_t all_tiny;                           // host-proven: every job has <= 8 vertices (every tile can take the analytic rows)
    uint32_t tile_begin, tile_end;               // tiles this launch of the direct tile kernel covers
    // binned tiles (raster_bins): bands of 32 rows x windows of b_wc columns
    uint32_t b_nbands, b_nwin, b_wc, n_bins;     // bins = n_jobs * b_nbands * b_nwin
    uint32_t b_lookback;                         // 1: one (band, window) tile per ticket, row sums passed through `look`; 0: a ticket walks all windows of a band
    uint32_t job_begin, job_end;                 // jobs this launch of the binned kernel covers
    uint32_t has_curves;                         // the job set holds Quad / Cubic ops (flatten parks their points between its two passes)
    uint32_t direct_max;                         // jobs with at most this many edge slots (<= DIRECT_MAX) are drawn by raster_tiles, larger ones by raster_bins
    uint32_t cull;                               // 1: flatten only the sub-figures that can reach rows [row_begin, row_end) or hold the top vertex
};
