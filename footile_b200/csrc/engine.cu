// engine.cu — the device pipeline of the footile hot path for sm_100a.
//
// Stages (reference functions they replace are cited per kernel):
//   (a) flatten_ops      per-PathOp adaptive De Casteljau -> Fixed vertices
//                        (plotter.rs:175-332, fig.rs:428-461)
//   (b) vtx_topkey / edge_build / job_finalize / bin_count / bin_fill
//                        ring edges, Edge::new, global winding + top row,
//                        counting sort of edges by row band
//                        (fig.rs:143-210,402-411,464-502,576-617)
//   (c)+(d) raster_tiles per (raster,row band) tile: scatter signed coverage
//                        deltas of every (edge,row) into a shared-memory row
//                        tile, warp-scan each row, apply the fill rule, store
//                        the matte / blend the colour (fig.rs:238-321,536-573,
//                        621-682; imgbuf.rs:22-199)
//
// Files (one translation unit; the .cuh parts are included below, inside namespace ftl):
//   device_types.cuh   records in HBM and per-call Params
//   scan.cuh           generic 3-phase exclusive scan
//   front_kernels.cuh  stages (a)+(b)
//   tile_kernel.cuh    stages (c)+(d) for jobs of at most 8 edges: analytic rows, scatter rows, resolve, composite
//   bin_kernel.cuh     stages (c)+(d) for jobs of more than 8 edges: (32 rows x 64 / 128 columns) tiles, lanes = rows or (edge, row) items
//   small_kernel.cuh   one small fill (a few ops, a small raster) in one launch: all four stages in a CTA
//   stroke_kernels.cuh the stroker (stroker.rs:204-416): outline ops of flattened wide polylines, one thread per (point, side)
//   libm_compat.cuh    glibc's hypotf / atan2f restated bit for bit, three-valued sin comparisons (included outside the namespace)
//   pack_kernels.cuh   packed read-back, raster checksums
//   engine.cu          host side: scratch arena, graph replay, launches, read-back
//
// There is no active-edge list and no per-row serial dependency: every
// (edge,row) contribution is evaluated in closed form,
// which tests/test_oracle_orderfree.py proves equal to the reference's scan.
//
// Build flags: -fmad=false (Rust never fuses a*b+c), no fast-math.
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "engine.h"
#include "fixed.cuh"
#include "libm_compat.cuh"
#include "pix_compat.cuh"
#include "pointy_compat.cuh"

namespace ftl {

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
const char *last_error() { return g_err.c_str(); }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                         \
            return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? FTL_ERR_NO_DEVICE \
                   : (e_ == cudaErrorMemoryAllocation ? FTL_ERR_NOMEM : FTL_ERR_CUDA);             \
        }                                                                                          \
    } while (0)

static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0};  // bytes this library moved over PCIe (rasters and path data; not the few counter words)
static std::atomic<bool> g_profiling{false};
#define LAUNCHED() g_launches.fetch_add(1, std::memory_order_relaxed)

#include "device_types.cuh"
#include "scan.cuh"
#include "front_kernels.cuh"
#include "tile_kernel.cuh"
#include "bin_kernel.cuh"
#include "small_kernel.cuh"
#include "stroke_kernels.cuh"
#include "pack_kernels.cuh"

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, cudaStream_t st) {
        if (bytes <= cap) return FTL_OK;
        if (p) {
            CK(cudaStreamSynchronize(st));
            CK(cudaFree(p));
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 4 + 256;
        CK(cudaMalloc(&p, want));
        cap = want;
        return FTL_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return FTL_OK;
        if (p) CK(cudaFreeHost(p));
        p = nullptr;
        size_t want = bytes + bytes / 4 + 256;
        CK(cudaMallocHost(&p, want));
        cap = want;
        return FTL_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Tile-kernel timing (CUDA events around every tile launch while profiling is on) is kept PER ENGINE, so that two handles
// profiling at the same time do not mix their numbers; the process-wide query sums over the live engines.
struct ProfSpan {
    cudaEvent_t a, b;
};
static std::vector<Engine::Impl *> g_engines;  // live engines, for the process-wide sum
static std::mutex g_engines_mu;

struct Engine::Impl {
    cudaStream_t st = nullptr;
    int n_sms = 148;
    size_t max_smem = 0;
    DevBuf ops, jobs, jstate, cnt, off, partials, vtx, edges, sub_last, tcount, toff, tpart, entries, counters, opw, wide, misc, tickets, look;
    DevBuf cull_mark, cull_head, cull_part, cull_lo, cull_hi, cull_job;  // row-band culling (front_kernels.cuh)
    DevBuf slabs;                         // points of curve ops parked by the counting pass of flatten_ops
    uint32_t look_epoch = 0;              // launch epoch of the look-back words (0: the buffer must be cleared first)
    std::vector<uint8_t> host_direct;    // per job: provably at most Params::direct_max edge slots (line-only, few ops)
    // small fills (small_kernel.cuh): one launch each, issued without waiting; a fill that did not fit raises its flag
    // in mapped host memory and is repeated through the general pipeline at the next blocking call
    struct SmallSaved {
        SmallArgs args;
        Geometry geo;
    };
    std::vector<SmallSaved> small_saved;
    uint32_t small_n = 0;                 // small fills issued since the flags were last checked
    bool last_small = false;              // the last fill went through small_fill (probes read its records)
    bool small_tail = false;              // the last operation issued on the stream is a small fill: its completion word tells when the stream is idle
    uint32_t small_seq = 0;               // sequence number of the last small fill
    PinBuf small_flags;                   // SMALL_RING overflow flags + the completion word
    DevBuf small_poison;                  // [0] poison, [1] CTA completion counter
    DevBuf srgb_tmp;                      // converted copy of a raster on its way to the host
    // device stroker (stroke_kernels.cuh): inputs in one blob, then the intermediate arrays; capacities persist
    DevBuf s_in, s_wop, s_keep, s_kidx, s_kp, s_info, s_cnt2, s_off2, s_counters, s_part;
    PinBuf pin_stroke;
    std::vector<ProfSpan> spans;          // tile launches timed but not yet collected
    std::mutex spans_mu;
    double tile_ms = 0.0;
    uint64_t tile_launches = 0;
    // fold the finished spans into (tile_ms, tile_launches); returns them, optionally resetting
    void collect_tile_time(bool reset, double *ms, uint64_t *launches) {
        std::lock_guard<std::mutex> lock(spans_mu);
        for (ProfSpan &s : spans) {
            float t = 0.f;
            if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
                tile_ms += t;
                tile_launches++;
            }
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        spans.clear();
        *ms += tile_ms;
        *launches += tile_launches;
        if (reset) {
            tile_ms = 0.0;
            tile_launches = 0;
        }
    }
    bool skip_graph = false;              // the next pipeline pass runs once with these sizes: do not capture a graph for it
    PinBuf pin_ops, pin_jobs, pin_small, pin_misc;
    PinBuf pin_ring, pin_pack[2], pin_lit[2];
    DevBuf pack_fixed, pack_cnt, pack_lit;
    // resident job set
    Params P{};
    bool have_jobs = false;
    bool layered = false;  // the jobs are layers of ONE raster: the tile kernel runs once per job, in order
    int smem_bytes = 0;
    // replays issued without a host round trip whose counters have not been checked yet
    uint32_t pending = 0;
    bool use_graph = true;
    cudaGraphExec_t graph = nullptr;
    uint64_t graph_kernels = 0;  // kernel launches one replay of the graph stands for
    std::vector<uint64_t> graph_key;
    void drop_graph() {
        if (graph) cudaGraphExecDestroy(graph);
        graph = nullptr;
        graph_key.clear();
    }
};
constexpr uint32_t RING = 64;
constexpr uint32_t SMALL_RING = 64;

int Engine::device_count(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        *count = 0;
        set_error(std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
        return FTL_ERR_NO_DEVICE;
    }
    *count = n;
    return FTL_OK;
}
uint64_t Engine::launch_count() { return g_launches.load(); }
void Engine::transfer_bytes(bool reset, uint64_t *h2d, uint64_t *d2h) {
    if (h2d) *h2d = g_h2d_bytes.load();
    if (d2h) *d2h = g_d2h_bytes.load();
    if (reset) {
        g_h2d_bytes.store(0);
        g_d2h_bytes.store(0);
    }
}
void Engine::set_profiling(bool on) { g_profiling.store(on); }
void Engine::tile_kernel_time(bool reset, double *ms, uint64_t *launches) {
    double t = 0.0;
    uint64_t n = 0;
    std::lock_guard<std::mutex> lock(g_engines_mu);
    for (Engine::Impl *m : g_engines) m->collect_tile_time(reset, &t, &n);
    if (ms) *ms = t;
    if (launches) *launches = n;
}
void Engine::tile_time(bool reset, double *ms, uint64_t *launches) {
    double t = 0.0;
    uint64_t n = 0;
    cudaSetDevice(device_);
    impl_->collect_tile_time(reset, &t, &n);
    if (ms) *ms = t;
    if (launches) *launches = n;
}

Engine::Engine(int device) : impl_(new Impl()), device_(device), stream_(nullptr) {
    std::lock_guard<std::mutex> lock(g_engines_mu);
    g_engines.push_back(impl_);
}

Engine::~Engine() {
    if (impl_) {
        {
            std::lock_guard<std::mutex> lock(g_engines_mu);
            g_engines.erase(std::remove(g_engines.begin(), g_engines.end(), impl_), g_engines.end());
        }
        {
            double t = 0.0;
            uint64_t n = 0;
            if (impl_->st) cudaSetDevice(device_);
            impl_->collect_tile_time(false, &t, &n);  // destroys the events of spans nobody collected
        }
        if (impl_->st) {
            cudaSetDevice(device_);
            cudaStreamSynchronize(impl_->st);
            Impl &m = *impl_;
            for (DevBuf *b : {&m.ops, &m.jobs, &m.jstate, &m.cnt, &m.off, &m.partials, &m.vtx, &m.edges, &m.sub_last, &m.tcount, &m.toff,
                              &m.tpart, &m.entries, &m.counters, &m.opw, &m.wide, &m.misc, &m.tickets, &m.look, &m.cull_mark, &m.cull_head, &m.cull_part, &m.cull_lo,
                              &m.cull_hi, &m.cull_job, &m.slabs, &m.small_poison, &m.srgb_tmp, &m.pack_fixed, &m.pack_cnt, &m.pack_lit, &m.s_in, &m.s_wop, &m.s_keep, &m.s_kidx, &m.s_kp,
                              &m.s_info, &m.s_cnt2, &m.s_off2, &m.s_counters, &m.s_part})
                b->release();
            m.drop_graph();
            for (PinBuf *b : {&m.pin_ops, &m.pin_jobs, &m.pin_small, &m.pin_misc, &m.pin_ring, &m.pin_pack[0], &m.pin_pack[1], &m.pin_lit[0], &m.pin_lit[1], &m.small_flags, &m.pin_stroke}) b->release();
            cudaStreamDestroy(impl_->st);
        }
        delete impl_;
    }
}

static int engine_init(Engine::Impl *m, int device, void **stream_out);
static int run_pipeline(Engine::Impl &m, bool exact);
static int resolve_pending(Engine::Impl &m);
static int upload_jobs(Engine::Impl &m, const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, bool layered, bool ops_on_device = false);

typedef void (*TileKernel)(const EdgeRec *, const JobDesc *, const JobState *, Params, const Counters *);
static TileKernel tile_kernel(int fmt, bool aligned) {
    switch (fmt) {
    case FTL_MATTE8: return aligned ? raster_tiles<FTL_MATTE8, true> : raster_tiles<FTL_MATTE8, false>;
    case FTL_GRAYA8P: return aligned ? raster_tiles<FTL_GRAYA8P, true> : raster_tiles<FTL_GRAYA8P, false>;
    default: return aligned ? raster_tiles<FTL_RGBA8P, true> : raster_tiles<FTL_RGBA8P, false>;
    }
}
typedef void (*BinKernel)(const EdgeRec *, const JobDesc *, const JobState *, Params, const uint32_t *, const uint32_t *, const Counters *, uint32_t *,
                          uint32_t *, uint32_t);
template <int WC>
static BinKernel bin_kernel_wc(int fmt, bool aligned) {
    switch (fmt) {
    case FTL_MATTE8: return aligned ? raster_bins<FTL_MATTE8, true, WC> : raster_bins<FTL_MATTE8, false, WC>;
    case FTL_GRAYA8P: return aligned ? raster_bins<FTL_GRAYA8P, true, WC> : raster_bins<FTL_GRAYA8P, false, WC>;
    default: return aligned ? raster_bins<FTL_RGBA8P, true, WC> : raster_bins<FTL_RGBA8P, false, WC>;
    }
}
typedef void (*SmallKernel)(const SmallArgs, JobState *, Counters *, EdgeRec *, uint32_t *, uint32_t *, uint32_t *, uint32_t *, long long *);
static SmallKernel small_kernel(int fmt, bool aligned) {
    switch (fmt) {
    case FTL_MATTE8: return aligned ? small_fill<FTL_MATTE8, true> : small_fill<FTL_MATTE8, false>;
    case FTL_GRAYA8P: return aligned ? small_fill<FTL_GRAYA8P, true> : small_fill<FTL_GRAYA8P, false>;
    default: return aligned ? small_fill<FTL_RGBA8P, true> : small_fill<FTL_RGBA8P, false>;
    }
}
static size_t small_smem_bytes() { return ((sizeof(SmallShared) + 15u) & ~(size_t)15u) + BIN_ROWS * (SMALL_MAX_DIM * 2 + BIN_ROW_PAD); }
static BinKernel bin_kernel(int fmt, bool aligned, uint32_t wc) { return wc == 64 ? bin_kernel_wc<64>(fmt, aligned) : (wc == 128 ? bin_kernel_wc<128>(fmt, aligned) : bin_kernel_wc<256>(fmt, aligned)); }
static size_t bin_smem_bytes(uint32_t wc) { return wc == 64 ? BinTile<64>::BYTES : (wc == 128 ? BinTile<128>::BYTES : BinTile<256>::BYTES); }

#define ENSURE_INIT()                                                        \
    do {                                                                     \
        CK(cudaSetDevice(device_));                                          \
        if (!impl_->st) {                                                    \
            int rc_ = engine_init(impl_, device_, &stream_);                 \
            if (rc_) return rc_;                                             \
        }                                                                    \
    } while (0)

// Per-device facts and kernel attributes are set up once per process, not once per handle.
struct DeviceInfo {
    bool ready = false;
    int n_sms = 0;
    size_t max_smem = 0;
};
static DeviceInfo g_dev[64];
static std::mutex g_dev_mu;

static int engine_init(Engine::Impl *m, int device, void **stream_out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
        return FTL_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n || device >= 64) {
        set_error("device index out of range");
        return FTL_ERR_INVALID;
    }
    CK(cudaSetDevice(device));
    {
        std::lock_guard<std::mutex> lock(g_dev_mu);
        DeviceInfo &di = g_dev[device];
        if (!di.ready) {
            int v = 0;
            CK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
            di.n_sms = v;
            CK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
            di.max_smem = (size_t)v;
            for (int f = 0; f < 3; f++)
                for (int a = 0; a < 2; a++) {
                    CK(cudaFuncSetAttribute(tile_kernel(f, a != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)di.max_smem));
                    CK(cudaFuncSetAttribute(tile_kernel(f, a != 0), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    CK(cudaFuncSetAttribute(small_kernel(f, a != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes()));
                    for (uint32_t wc : {64u, 128u, 256u}) {
                        CK(cudaFuncSetAttribute(bin_kernel(f, a != 0, wc), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem_bytes(wc)));
                        CK(cudaFuncSetAttribute(bin_kernel(f, a != 0, wc), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    }
                }
            CK(cudaFuncSetAttribute(accumulate_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)di.max_smem));
            CK(cudaFuncSetAttribute(flatten_ops<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM_BYTES));
            CK(cudaFuncSetAttribute(flatten_ops<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_SMEM_BYTES));
            di.ready = true;
        }
        m->n_sms = di.n_sms;
        m->max_smem = di.max_smem;
    }
    CK(cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking));
    if (const char *ev = getenv("FTL_NO_GRAPH")) m->use_graph = atoi(ev) == 0;
    *stream_out = m->st;
    return FTL_OK;
}

static inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

template <class Op>
static int run_scan(cudaStream_t st, const typename Op::T *in, uint32_t n, typename Op::T *out, DevBuf &partials) {
    uint32_t nb = div_up(n > 0 ? n : 1, SCAN_BLOCK);
    int rc = partials.ensure((size_t)nb * sizeof(typename Op::T), st);
    if (rc) return rc;
    typename Op::T *part = (typename Op::T *)partials.p;
    scan_reduce<Op><<<nb, SCAN_THREADS, 0, st>>>(in, n, part); LAUNCHED();
    scan_partials<Op><<<1, SCAN_THREADS, 0, st>>>(part, nb); LAUNCHED();
    scan_apply<Op><<<nb, SCAN_THREADS, 0, st>>>(in, n, part, out); LAUNCHED();
    CK(cudaGetLastError());
    return FTL_OK;
}

static int choose_tiling(const Geometry &g, size_t max_smem, uint32_t jobs_per_launch, int n_sms, bool all_direct, Params *P) {
    const uint32_t warp_slots = (uint32_t)n_sms * 16u;
    P->W = g.width; P->H = g.height; P->row_begin = g.row_begin; P->row_end = g.row_end;
    P->fmt = (uint32_t)g.format; P->bpp = g.bpp(); P->pitch = (uint32_t)g.pitch();
    // ---- direct tiles (raster_tiles) ----
    P->chunks = (g.width + CHUNK - 1) / CHUNK;
    P->WP = P->chunks * CHUNK;
    // each warp keeps one row window of up to 4 chunks (2048 px) in shared memory
    uint32_t win = 4;
    if (const char *ev = getenv("FTL_WIN_CHUNKS")) win = (uint32_t)std::min(32, std::max(1, atoi(ev)));  // tuning knob
    P->win_chunks = std::min(P->chunks, win);
    P->n_win = (P->chunks + P->win_chunks - 1) / P->win_chunks;
    P->win_rows = 1;
    P->cta_warps = 4;
    if (const char *ev = getenv("FTL_CTA_WARPS")) P->cta_warps = (uint32_t)std::min(4, std::max(1, atoi(ev)));  // tuning knob

    // Band height: 8 rows per warp amortise the per-tile set-up when every job is provably a polygon of a few
    // edges; jobs of up to 64 edges do better with 4 (one scatter pass per tile); fewer rows per band when one
    // launch would otherwise leave most of the GPU's warp slots empty (a single raster, a layer of a scene).
    uint32_t log2R = all_direct ? 3 : 2;
    while (log2R > 0 && (1u << log2R) >= 2 * g.rows()) log2R--;
    while (log2R > 0 && (uint64_t)jobs_per_launch * div_up(g.rows(), 1u << log2R) < 2ull * warp_slots) log2R--;
    if (const char *ev = getenv("FTL_LOG2R")) log2R = (uint32_t)std::min(5, std::max(0, atoi(ev)));  // tuning knob
    P->log2R = log2R; P->R = 1u << log2R;
    P->n_bands = div_up(g.rows(), P->R);
    // a narrow raster (one window per row) fills the window with several rows of the band
    if (P->n_win == 1) P->win_rows = std::max(1u, std::min(win / P->win_chunks, P->R));
    if (const char *ev = getenv("FTL_WIN_ROWS")) P->win_rows = P->n_win == 1 ? (uint32_t)std::min(8, std::max(1, atoi(ev))) : 1u;  // tuning knob
    P->warp_words = (P->win_rows * (P->win_chunks * CHUNK + P->win_chunks) + 3u) & ~3u;  // cells + masks, 16-byte multiple
    size_t cta_bytes = (size_t)P->warp_words * 4 * P->cta_warps;
    if (cta_bytes > max_smem) {
        set_error("row window exceeds shared memory");
        return FTL_ERR_TOO_WIDE;
    }
    // ---- binned tiles (raster_bins): bands of 32 rows x windows of b_wc columns ----
    P->b_wc = 128;  // 8.8 KB per warp: 21 resident warps per SM (256 columns: 12 warps, 20-30 % slower on every workload measured)
    P->b_nbands = div_up(g.rows(), BIN_ROWS);
    P->b_nwin = div_up(g.width, P->b_wc);
    // One ticket per (band, window) with the row sums handed to the right neighbour, unless the launch has
    // plenty of bands anyway (many narrow rasters): then a ticket walks the windows of its band serially.
    const uint32_t bin_slots = (uint32_t)n_sms * 12u;
    P->b_lookback = P->b_nwin > 1 && (uint64_t)jobs_per_launch * P->b_nbands < 8ull * bin_slots;
    // Many narrow rasters (no look-back): windows of 64 columns halve the tile's shared memory, 40 warps per SM hide
    // the per-tile load chain (ticket -> job -> bin -> entries -> edges) better than 21 do: the batched fishy fills
    // 0.415 -> 0.349 ms, config 4 0.755 -> 0.66 ms per launch.  Wide rasters lose with 64 (config 5 34.6 -> 40.5 ms).
    if (!P->b_lookback && P->b_nwin > 1) P->b_wc = 64;
    if (const char *ev = getenv("FTL_BIN_WC")) P->b_wc = atoi(ev) == 256 ? 256u : (atoi(ev) == 64 ? 64u : (atoi(ev) == 128 ? 128u : P->b_wc));  // tuning knob
    P->b_nwin = div_up(g.width, P->b_wc);
    if (const char *ev = getenv("FTL_BIN_LOOKBACK")) P->b_lookback = P->b_nwin > 1 && atoi(ev) != 0;  // tuning knob
    return FTL_OK;
}

// Jobs of at most this many edge slots are drawn by raster_tiles straight from their own edge range (no binning).
// 8 = the jobs whose rows can take the analytic path (tile_kernel.cuh); anything larger is faster through the bins:
// raster_tiles' lanes-are-edges scatter of 9..64 edges measured 2.56 ms on the batched fishy fills (33 edges, 256^2)
// against 0.42 ms for raster_bins, and 0.35 against 0.24 ms on the 4K stroke scenes.  FTL_DIRECT_MAX (1 .. 64): tuning knob.
static uint32_t direct_max_setting() {
    if (const char *ev = getenv("FTL_DIRECT_MAX")) return (uint32_t)std::min<int>((int)DIRECT_MAX, std::max(1, atoi(ev)));
    return 8u;
}

static int validate_ops(const ftl_path_op *ops, size_t n) {
    for (size_t i = 0; i < n; i++) {
        int nv = ops[i].tag == FTL_OP_CLOSE ? 0 : (ops[i].tag == FTL_OP_QUAD ? 4 : (ops[i].tag == FTL_OP_CUBIC ? 6 : (ops[i].tag == FTL_OP_PENWIDTH ? 1 : 2)));
        if (ops[i].tag > FTL_OP_PENWIDTH) {
            set_error("unknown path op tag");
            return FTL_ERR_INVALID;
        }
        for (int k = 0; k < nv; k++) {
            float f = ops[i].v[k];
            if (!(f - f == 0.0f)) {
                set_error("non-finite coordinate in path op");
                return FTL_ERR_NONFINITE;
            }
        }
    }
    return FTL_OK;
}

// Validate ops[0, n) and copy them into the pinned staging buffer in one pass over the caller's memory; large
// arrays (config 5 sends 291 MB per fill) are split over host threads.  Returns the first failure in op order.
static int stage_ops(ftl_path_op *dst, const ftl_path_op *ops, size_t n, std::atomic<bool> *has_curves) {
    auto one = [has_curves](ftl_path_op *d, const ftl_path_op *o, size_t cnt, int *status) {
        int st = FTL_OK;
        bool curves = false;
        for (size_t i = 0; i < cnt && st == FTL_OK; i++) {
            const ftl_path_op op = o[i];
            if (op.tag > FTL_OP_PENWIDTH) st = FTL_ERR_INVALID;
            curves |= op.tag == FTL_OP_QUAD || op.tag == FTL_OP_CUBIC;
            const int nv = op.tag == FTL_OP_CLOSE ? 0 : (op.tag == FTL_OP_QUAD ? 4 : (op.tag == FTL_OP_CUBIC ? 6 : (op.tag == FTL_OP_PENWIDTH ? 1 : 2)));
            for (int k = 0; k < nv; k++)
                if (!(op.v[k] - op.v[k] == 0.0f) && st == FTL_OK) st = FTL_ERR_NONFINITE;
            d[i] = op;
        }
        if (curves) has_curves->store(true, std::memory_order_relaxed);
        *status = st;
    };
    unsigned nt = n < (1u << 16) ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<int> status(nt, FTL_OK);
    if (nt == 1) one(dst, ops, n, &status[0]);
    else {
        const size_t per = (n + nt - 1) / nt;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) {
            const size_t b = std::min(n, t * per), e = std::min(n, b + per);
            th.emplace_back(one, dst + b, ops + b, e - b, &status[t]);
        }
        for (std::thread &t : th) t.join();
    }
    for (int st : status)
        if (st != FTL_OK) {
            set_error(st == FTL_ERR_INVALID ? "unknown path op tag" : "non-finite coordinate in path op");
            return st;
        }
    return FTL_OK;
}

int Engine::upload(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, bool layered) {
    ENSURE_INIT();
    Impl &m = *impl_;
    {
        int rc0 = resolve_pending(m);  // earlier replays refer to the job set that is about to be replaced
        if (rc0) return rc0;
    }
    return upload_jobs(m, g, jobs, ops, n_ops, layered);
}

// ops_on_device: the ops are written into m.ops by kernels after this call (the device stroker's outline: Line / Close
// ops only), and so are the op ranges of the job descriptors; nothing is staged or validated here.
static int upload_jobs(Engine::Impl &m, const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, bool layered, bool ops_on_device) {
    m.have_jobs = false;  // stays false if anything below fails
    m.last_small = false;
    m.small_tail = false;
    m.layered = layered;
    if (jobs.empty() || g.rows() == 0 || g.width == 0) {
        m.have_jobs = false;
        return FTL_OK;
    }
    if (n_ops >= 0x7FFFFFFFull || jobs.size() >= 0x7FFFFFFFull) {
        set_error("too many ops/jobs for one call");
        return FTL_ERR_INVALID;
    }
    int rc = FTL_OK;
    Params P{};
    // Line-only jobs of at most direct_max ops cannot exceed direct_max vertices: no binning at all.
    P.direct_max = direct_max_setting();
    P.all_direct = 1;
    P.all_tiny = 1;
    P.has_curves = 0;
    m.host_direct.assign(jobs.size(), 1);
    for (size_t jx = 0; jx < jobs.size(); jx++) {
        const HostJob &h = jobs[jx];
        if (h.op_end - h.op_begin > 8u) P.all_tiny = 0;
        bool direct = !ops_on_device && h.op_end - h.op_begin <= P.direct_max;
        if (ops_on_device) P.all_tiny = 0;
        for (uint32_t i = h.op_begin; i < h.op_end && direct; i++)
            if (ops[i].tag == FTL_OP_QUAD || ops[i].tag == FTL_OP_CUBIC) direct = false;
        if (!direct) {
            m.host_direct[jx] = 0;
            P.all_direct = 0;
            if (!layered) break;  // only layered launches look at the per-job flags
        }
    }
    rc = choose_tiling(g, m.max_smem, layered ? 1u : (uint32_t)jobs.size(), m.n_sms, P.all_direct != 0, &P);
    if (rc) return rc;
    P.n_jobs = (uint32_t)jobs.size();
    P.n_ops = (uint32_t)n_ops;
    P.cull = (g.row_begin > 0 || g.row_end < g.height) && n_ops >= 4096;
    if (const char *ev = getenv("FTL_CULL")) P.cull = atoi(ev) != 0 && n_ops > 0;  // tuning / test knob
    uint64_t nt = (uint64_t)P.n_jobs * P.n_bands;
    if (nt >= 0x7FFFFFFFull) {
        set_error("too many tiles for one call");
        return FTL_ERR_INVALID;
    }
    P.n_tiles = (uint32_t)nt;
    const uint64_t nb = P.all_direct ? 0ull : (uint64_t)P.n_jobs * P.b_nbands * P.b_nwin;
    if (nb >= 0x7FFFFFFFull) {
        set_error("too many tiles for one call");
        return FTL_ERR_INVALID;
    }
    P.n_bins = (uint32_t)nb;
    // stage + upload ops and job descriptors
    size_t ops_bytes = n_ops * sizeof(ftl_path_op), jobs_bytes = jobs.size() * sizeof(JobDesc);
    // the previous call's async copies out of the pinned staging must be done before it is overwritten or freed
    CK(cudaStreamSynchronize(m.st));
    if (!ops_on_device && (rc = m.pin_ops.ensure(ops_bytes ? ops_bytes : 1))) return rc;
    if ((rc = m.pin_jobs.ensure(jobs_bytes))) return rc;
    if ((rc = m.ops.ensure(ops_bytes ? ops_bytes : 1, m.st))) return rc;
    if ((rc = m.jobs.ensure(jobs_bytes, m.st))) return rc;
    // validate + stage + copy in pieces: the DMA of one piece runs while the host threads stage the next
    if (!ops_on_device) {
        const size_t PIECE = 1u << 20;  // ops per piece (28 MB)
        std::atomic<bool> curves{false};
        for (size_t at = 0; at < n_ops; at += PIECE) {
            const size_t cnt = std::min(PIECE, n_ops - at);
            if ((rc = stage_ops((ftl_path_op *)m.pin_ops.p + at, ops + at, cnt, &curves))) {
                cudaStreamSynchronize(m.st);
                return rc;
            }
            CK(cudaMemcpyAsync((ftl_path_op *)m.ops.p + at, (const ftl_path_op *)m.pin_ops.p + at, cnt * sizeof(ftl_path_op), cudaMemcpyHostToDevice, m.st));
            g_h2d_bytes.fetch_add(cnt * sizeof(ftl_path_op), std::memory_order_relaxed);
        }
        P.has_curves = curves.load() ? 1u : 0u;
    }
    JobDesc *jd = (JobDesc *)m.pin_jobs.p;
    for (size_t j = 0; j < jobs.size(); j++) {
        const HostJob &h = jobs[j];
        JobDesc d{};
        d.op_begin = h.op_begin; d.op_end = h.op_end;
        for (int k = 0; k < 6; k++) {
            d.e[k] = h.e[k];
            if (!(h.e[k] - h.e[k] == 0.0f)) {
                set_error("non-finite transform");
                return FTL_ERR_NONFINITE;
            }
        }
        d.tol_sq = h.tol_sq;
        d.rule = (uint32_t)h.rule;
        d.color = (uint32_t)h.color[0] | ((uint32_t)h.color[1] << 8) | ((uint32_t)h.color[2] << 16) | ((uint32_t)h.color[3] << 24);
        d.raster = (unsigned long long)(uintptr_t)h.raster;
        jd[j] = d;
    }
    CK(cudaMemcpyAsync(m.jobs.p, m.pin_jobs.p, jobs_bytes, cudaMemcpyHostToDevice, m.st));
    g_h2d_bytes.fetch_add(jobs_bytes, std::memory_order_relaxed);
    m.P = P;
    m.smem_bytes = (int)((size_t)P.warp_words * 4 * P.cta_warps);
    m.have_jobs = true;
    return FTL_OK;
}

// FTL_SMALL_PROF=1: block 0 of every small fill stamps clock64() at its phase boundaries into 16 words behind the poison
// word pair (read back with Engine::small_profile); a development aid, off by default.
static long long *small_prof(Engine::Impl &m) {
    static const bool on = getenv("FTL_SMALL_PROF") && atoi(getenv("FTL_SMALL_PROF")) != 0;
    return on ? (long long *)((uint32_t *)m.small_poison.p + 2) : nullptr;
}
int Engine::small_profile(long long out[9]) {
    ENSURE_INIT();
    Impl &m = *impl_;
    for (int k = 0; k < 9; k++) out[k] = 0;
    if (!m.small_poison.p) return FTL_OK;
    CK(cudaStreamSynchronize(m.st));
    CK(cudaMemcpy(out, (uint32_t *)m.small_poison.p + 2, 9 * sizeof(long long), cudaMemcpyDeviceToHost));
    return FTL_OK;
}

// One small fill in one launch (small_kernel.cuh).  Returns -1 when the call is not eligible.
static int fill_small(Engine::Impl &m, const Geometry &g, const HostJob &h, const ftl_path_op *ops, size_t n_ops) {
    const char *no_small = getenv("FTL_NO_SMALL");  // read per call: the parity tests draw the same input through both paths
    if ((no_small && atoi(no_small) != 0) || n_ops == 0 || n_ops > SMALL_MAX_OPS || g.width == 0 || g.width > SMALL_MAX_DIM || g.rows() == 0 || g.rows() > SMALL_MAX_DIM) return -1;
    if (small_smem_bytes() > m.max_smem) return -1;
    int rc;
    if (m.pending && (rc = resolve_pending(m))) return rc;  // an earlier general replay may still have to be repeated: keep the order
    if (m.small_n >= SMALL_RING && (rc = resolve_pending(m))) return rc;
    if ((rc = validate_ops(ops, n_ops))) return rc;
    for (int k = 0; k < 6; k++)
        if (!(h.e[k] - h.e[k] == 0.0f)) {
            set_error("non-finite transform");
            return FTL_ERR_NONFINITE;
        }
    cudaStream_t st = m.st;
    if (!m.small_flags.p) {
        if ((rc = m.small_flags.ensure((SMALL_RING + 1) * sizeof(uint32_t)))) return rc;
        memset(m.small_flags.p, 0, (SMALL_RING + 1) * sizeof(uint32_t));
        if ((rc = m.small_poison.ensure(2 * sizeof(uint32_t) + 16 * sizeof(long long), st))) return rc;
        CK(cudaMemsetAsync(m.small_poison.p, 0, 2 * sizeof(uint32_t) + 16 * sizeof(long long), st));
        m.small_saved.resize(SMALL_RING);
    }
    if ((rc = m.counters.ensure(sizeof(Counters), st))) return rc;
    if ((rc = m.jstate.ensure(sizeof(JobState), st))) return rc;
    if ((rc = m.edges.ensure((size_t)SMALL_MAX_V * sizeof(EdgeRec), st))) return rc;
    Engine::Impl::SmallSaved &sv = m.small_saved[m.small_n];
    SmallArgs &A = sv.args;
    sv.geo = g;
    A.job = JobDesc{};
    A.job.op_begin = 0; A.job.op_end = (uint32_t)n_ops;
    memcpy(A.job.e, h.e, sizeof(A.job.e));
    A.job.tol_sq = h.tol_sq;
    A.job.rule = (uint32_t)h.rule;
    A.job.color = (uint32_t)h.color[0] | ((uint32_t)h.color[1] << 8) | ((uint32_t)h.color[2] << 16) | ((uint32_t)h.color[3] << 24);
    A.job.raster = (unsigned long long)(uintptr_t)h.raster;
    A.n_ops = (uint32_t)n_ops; A.W = g.width; A.H = g.height; A.row_begin = g.row_begin; A.row_end = g.row_end;
    A.pitch = (uint32_t)g.pitch(); A.bpp = g.bpp(); A.seq = ++m.small_seq;
    memcpy(A.ops, ops, n_ops * sizeof(ftl_path_op));
    uint32_t *flags = (uint32_t *)m.small_flags.p;
    flags[m.small_n] = 0;
    const bool aligned = g.width % (16u / g.bpp()) == 0;
    const uint32_t n_bands = div_up(g.rows(), BIN_ROWS);
    const size_t tile_bytes = (size_t)BIN_ROWS * (div_up(g.width, g.width <= 256u ? 256u : 512u) * (g.width <= 256u ? 256u : 512u) * 2 + BIN_ROW_PAD);
    small_kernel(g.format, aligned)<<<n_bands, SMALL_WARPS * 32, ((sizeof(SmallShared) + 15u) & ~(size_t)15u) + tile_bytes, st>>>(
        A, (JobState *)m.jstate.p, (Counters *)m.counters.p, (EdgeRec *)m.edges.p, (uint32_t *)m.small_poison.p, flags + m.small_n,
        (uint32_t *)m.small_poison.p + 1, flags + SMALL_RING, small_prof(m)); LAUNCHED();
    CK(cudaGetLastError());
    g_h2d_bytes.fetch_add(sizeof(JobDesc) + n_ops * sizeof(ftl_path_op), std::memory_order_relaxed);  // as kernel arguments
    m.small_n++;
    m.small_tail = true;
    m.have_jobs = false;
    m.last_small = true;
    return FTL_OK;
}

int Engine::fill(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, bool allow_small) {
    if (allow_small && jobs.size() == 1) {
        ENSURE_INIT();
        int rs = fill_small(*impl_, g, jobs[0], ops, n_ops);
        if (rs != -1) return rs;
    }
    int rc = upload(g, jobs, ops, n_ops);
    if (rc) return rc;
    return replay();
}

int Engine::fill_layers(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops) {
    int rc = upload(g, jobs, ops, n_ops, true);
    if (rc) return rc;
    return replay();
}

// Plotter::stroke on the device (plotter.rs:356-365 with stroker.rs:204-416 as stroke_kernels.cuh).
int Engine::stroke(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, const float *opw, int join,
                   float miter_limit, bool *needs_host, std::vector<ftl_path_op> *outline, std::vector<uint32_t> *outline_offsets) {
    ENSURE_INIT();
    Impl &m = *impl_;
    *needs_host = false;
    if (outline) outline->clear();
    if (outline_offsets) outline_offsets->assign(jobs.size() + 1, 0u);
    int rc = resolve_pending(m);
    if (rc) return rc;
    if (jobs.empty() || n_ops == 0) return FTL_OK;
    if (n_ops >= 0x7FFFFFFFull || jobs.size() >= 0x7FFFFFFFull) {
        set_error("too many ops/jobs for one call");
        return FTL_ERR_INVALID;
    }
    if ((rc = validate_ops(ops, n_ops))) return rc;
    cudaStream_t st = m.st;
    m.small_tail = false;
    m.have_jobs = false;
    m.last_small = false;
    const uint32_t n_jobs = (uint32_t)jobs.size(), no = (uint32_t)n_ops;
    // ---- host: sub-strokes from the ops, job descriptors of the flatten ----
    std::vector<uint32_t> subs, op_sub(n_ops, NONE32), jfs(n_jobs + 1, 0u);
    for (uint32_t j = 0; j < n_jobs; j++) {
        jfs[j] = (uint32_t)(subs.size() / 4);
        stroke_sub_table(ops, jobs[j].op_begin, jobs[j].op_end, j, &subs, op_sub.data());
    }
    const uint32_t n_subs = (uint32_t)(subs.size() / 4);
    jfs[n_jobs] = n_subs;
    if (n_subs == 0) return FTL_OK;  // no drawing op: Stroke::path_ops is empty and the fill draws nothing (fig.rs:491)
    auto up16 = [](size_t v) { return (v + 15u) & ~(size_t)15u; };
    const size_t o_ops = 0, o_jobs = up16(o_ops + n_ops * sizeof(ftl_path_op)), o_opw = up16(o_jobs + (size_t)n_jobs * sizeof(JobDesc)),
                 o_opsub = up16(o_opw + n_ops * 2 * sizeof(float)), o_subs = up16(o_opsub + n_ops * sizeof(uint32_t)),
                 o_jfs = up16(o_subs + subs.size() * sizeof(uint32_t)), in_bytes = up16(o_jfs + jfs.size() * sizeof(uint32_t));
    CK(cudaStreamSynchronize(st));  // the previous call's copy out of the pinned staging is done
    if ((rc = m.pin_stroke.ensure(in_bytes + sizeof(StrokeCounters)))) return rc;
    if ((rc = m.s_in.ensure(in_bytes, st))) return rc;
    uint8_t *hp = (uint8_t *)m.pin_stroke.p;
    memcpy(hp + o_ops, ops, n_ops * sizeof(ftl_path_op));
    JobDesc *jd = (JobDesc *)(hp + o_jobs);
    for (uint32_t j = 0; j < n_jobs; j++) {
        JobDesc d{};
        d.op_begin = jobs[j].op_begin; d.op_end = jobs[j].op_end;
        for (int k = 0; k < 6; k++) {
            d.e[k] = jobs[j].e[k];
            if (!(d.e[k] - d.e[k] == 0.0f)) {
                set_error("non-finite transform");
                return FTL_ERR_NONFINITE;
            }
        }
        d.tol_sq = jobs[j].tol_sq;
        jd[j] = d;
    }
    memcpy(hp + o_opw, opw, n_ops * 2 * sizeof(float));
    memcpy(hp + o_opsub, op_sub.data(), n_ops * sizeof(uint32_t));
    memcpy(hp + o_subs, subs.data(), subs.size() * sizeof(uint32_t));
    memcpy(hp + o_jfs, jfs.data(), jfs.size() * sizeof(uint32_t));
    CK(cudaMemcpyAsync(m.s_in.p, hp, in_bytes, cudaMemcpyHostToDevice, st));
    g_h2d_bytes.fetch_add(in_bytes, std::memory_order_relaxed);
    const uint8_t *dp = (const uint8_t *)m.s_in.p;
    const ftl_path_op *d_ops = (const ftl_path_op *)(dp + o_ops);
    const JobDesc *d_jobs = (const JobDesc *)(dp + o_jobs);
    const float *d_opw = (const float *)(dp + o_opw);
    const uint32_t *d_opsub = (const uint32_t *)(dp + o_opsub);
    const StrokeSub *d_subs = (const StrokeSub *)(dp + o_subs);
    const uint32_t *d_jfs = (const uint32_t *)(dp + o_jfs);

    if ((rc = m.cnt.ensure((n_ops + 1) * sizeof(SumHead), st))) return rc;
    if ((rc = m.off.ensure((n_ops + 1) * sizeof(SumHead), st))) return rc;
    if ((rc = m.s_counters.ensure(sizeof(StrokeCounters), st))) return rc;
    if ((rc = m.s_info.ensure((size_t)n_subs * sizeof(StrokeSubInfo), st))) return rc;
    if ((rc = m.wide.ensure(3 * sizeof(float), st))) return rc;
    StrokeCounters *d_sc = (StrokeCounters *)m.s_counters.p;
    StrokeCounters *h_sc = (StrokeCounters *)(hp + in_bytes);
    Params P{};
    P.n_jobs = n_jobs;
    P.n_ops = no;
    const StrokeJoin sj{join, miter_limit, jobs[0].tol_sq};
    const uint32_t fb = std::min<uint32_t>(div_up(no, FLAT_THREADS), (uint32_t)m.n_sms * 16);
    const CullBufs no_cull{nullptr, nullptr, nullptr, nullptr};
    uint32_t cap_raw = 0, n_slots = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        // capacity of the point arrays: what earlier calls needed (the first call learns it from its overflow)
        cap_raw = (uint32_t)std::min<size_t>(m.wide.cap / (3 * sizeof(float)), (size_t)0x3FFFFFFF);
        if ((rc = m.s_wop.ensure((size_t)cap_raw * sizeof(uint32_t), st))) return rc;
        if ((rc = m.s_keep.ensure((size_t)cap_raw * sizeof(uint32_t), st))) return rc;
        if ((rc = m.s_kidx.ensure(((size_t)cap_raw + 1) * sizeof(uint32_t), st))) return rc;
        if ((rc = m.s_kp.ensure((size_t)cap_raw * sizeof(float4), st))) return rc;
        n_slots = 2u * (cap_raw + n_subs);
        if ((rc = m.s_cnt2.ensure((size_t)n_slots * sizeof(uint32_t), st))) return rc;
        if ((rc = m.s_off2.ensure(((size_t)n_slots + 1) * sizeof(uint32_t), st))) return rc;
        CK(cudaMemsetAsync(d_sc, 0, sizeof(StrokeCounters), st));
        flatten_ops<true, false><<<fb, FLAT_THREADS, 0, st>>>(d_ops, d_jobs, P, d_opw, (SumHead *)m.cnt.p, nullptr, nullptr, nullptr, nullptr, no_cull); LAUNCHED();
        if ((rc = run_scan<SumHeadOp>(st, (const SumHead *)m.cnt.p, no, (SumHead *)m.off.p, m.partials))) return rc;
        stroke_set_raw<<<div_up(n_jobs, 256), 256, 0, st>>>(d_sc, (const SumHead *)m.off.p, d_jobs, n_jobs, no, cap_raw); LAUNCHED();
        flatten_ops<true, true><<<fb, FLAT_THREADS, 0, st>>>(d_ops, d_jobs, P, d_opw, nullptr, (const SumHead *)m.off.p, nullptr, (float *)m.wide.p, &d_sc->overflow,
                                                           no_cull, nullptr, 0, (uint32_t *)m.s_wop.p); LAUNCHED();
        const uint32_t pb = std::max<uint32_t>(1u, std::min<uint32_t>(div_up(cap_raw, 256), (uint32_t)m.n_sms * 8));
        stroke_keep<<<pb, 256, 0, st>>>((const float *)m.wide.p, (const uint32_t *)m.s_wop.p, (const SumHead *)m.off.p, d_opsub, d_subs, d_sc, cap_raw,
                                       (uint32_t *)m.s_keep.p); LAUNCHED();
        if ((rc = run_scan<AddU32>(st, (const uint32_t *)m.s_keep.p, cap_raw, (uint32_t *)m.s_kidx.p, m.s_part))) return rc;
        stroke_compact<<<pb, 256, 0, st>>>((const float *)m.wide.p, (const uint32_t *)m.s_wop.p, d_opsub, (const uint32_t *)m.s_keep.p,
                                          (const uint32_t *)m.s_kidx.p, d_sc, cap_raw, (float4 *)m.s_kp.p); LAUNCHED();
        stroke_sub_info<<<div_up(n_subs, 128), 128, 0, st>>>(d_subs, n_subs, (const SumHead *)m.off.p, (const uint32_t *)m.s_kidx.p, d_sc,
                                                            (StrokeSubInfo *)m.s_info.p); LAUNCHED();
        CK(cudaMemsetAsync(m.s_cnt2.p, 0, (size_t)n_slots * sizeof(uint32_t), st));
        const uint32_t sb = std::max<uint32_t>(1u, std::min<uint32_t>(div_up(2ull * cap_raw, 128), (uint32_t)m.n_sms * 16));
        stroke_segments<false><<<sb, 128, 0, st>>>(sj, (const float4 *)m.s_kp.p, (const StrokeSubInfo *)m.s_info.p, d_sc, (uint32_t *)m.s_cnt2.p, nullptr, nullptr); LAUNCHED();
        if ((rc = run_scan<AddU32>(st, (const uint32_t *)m.s_cnt2.p, n_slots, (uint32_t *)m.s_off2.p, m.s_part))) return rc;
        stroke_set_out<<<1, 1, 0, st>>>(d_sc, (const uint32_t *)m.s_off2.p, n_slots); LAUNCHED();
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_sc, d_sc, sizeof(StrokeCounters), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));  // the one host round trip of a device stroke: the outline's size
        if (!h_sc->overflow) break;
        if (attempt == 1) {
            set_error("device stroker: the point capacity guard tripped twice");
            return FTL_ERR_CUDA;
        }
        if ((rc = m.wide.ensure((size_t)h_sc->need_raw * 3 * sizeof(float), st))) return rc;
    }
    if (h_sc->fallback) {
        *needs_host = true;
        return FTL_OK;
    }
    const uint32_t n_out = h_sc->n_out;
    const uint32_t sb = std::max<uint32_t>(1u, std::min<uint32_t>(div_up(2ull * h_sc->nk, 128), (uint32_t)m.n_sms * 16));
    if (outline) {  // probe: hand the outline back instead of filling it
        if ((rc = m.ops.ensure((size_t)std::max(n_out, 1u) * sizeof(ftl_path_op), st))) return rc;
        if ((rc = m.jobs.ensure((size_t)n_jobs * sizeof(JobDesc), st))) return rc;
        stroke_segments<true><<<sb, 128, 0, st>>>(sj, (const float4 *)m.s_kp.p, (const StrokeSubInfo *)m.s_info.p, d_sc, nullptr, (const uint32_t *)m.s_off2.p,
                                                 (ftl_path_op *)m.ops.p); LAUNCHED();
        stroke_patch_jobs<<<div_up(n_jobs, 128), 128, 0, st>>>((JobDesc *)m.jobs.p, n_jobs, d_jfs, n_subs, (const StrokeSubInfo *)m.s_info.p,
                                                              (const uint32_t *)m.s_off2.p, d_sc); LAUNCHED();
        CK(cudaGetLastError());
        outline->resize(n_out);
        std::vector<JobDesc> hj(n_jobs);
        if (n_out) CK(cudaMemcpyAsync(outline->data(), m.ops.p, (size_t)n_out * sizeof(ftl_path_op), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hj.data(), m.jobs.p, (size_t)n_jobs * sizeof(JobDesc), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (outline_offsets) {
            for (uint32_t j = 0; j < n_jobs; j++) (*outline_offsets)[j] = hj[j].op_begin;
            (*outline_offsets)[n_jobs] = n_out;
        }
        return FTL_OK;
    }
    if (n_out == 0) return FTL_OK;
    // ---- the fill of the outline (NonZero, plotter.rs:364): its ops and op ranges are written by the two kernels below ----
    std::vector<HostJob> fjobs(jobs);
    for (HostJob &h : fjobs) {
        h.op_begin = h.op_end = 0;
        h.rule = FTL_NONZERO;
    }
    if ((rc = upload_jobs(m, g, fjobs, nullptr, n_out, false, true))) return rc;
    if (!m.have_jobs) return FTL_OK;
    stroke_segments<true><<<sb, 128, 0, st>>>(sj, (const float4 *)m.s_kp.p, (const StrokeSubInfo *)m.s_info.p, d_sc, nullptr, (const uint32_t *)m.s_off2.p,
                                             (ftl_path_op *)m.ops.p); LAUNCHED();
    stroke_patch_jobs<<<div_up(n_jobs, 128), 128, 0, st>>>((JobDesc *)m.jobs.p, n_jobs, d_jfs, n_subs, (const StrokeSubInfo *)m.s_info.p,
                                                          (const uint32_t *)m.s_off2.p, d_sc); LAUNCHED();
    CK(cudaGetLastError());
    m.skip_graph = true;
    const bool exact = m.vtx.cap == 0 || m.edges.cap == 0 || m.entries.cap == 0;
    rc = run_pipeline(m, exact);
    m.skip_graph = false;
    return rc;
}

// One pass of the device pipeline over the resident job set.
//   exact = true : sizes are read back after each scan (two host round trips) and the scratch
//                  buffers are grown to fit; used for the first call and after an overflow.
//   exact = false: buffers keep the capacity of earlier calls, every kernel takes its counts
//                  from device memory and nothing waits for the host; the stages before the tile
//                  kernel are replayed from a CUDA graph.
static int run_pipeline(Engine::Impl &m, bool exact) {
    m.small_tail = false;  // other work follows on the stream: its end is no longer the last small fill's completion word
    const Params P = m.P;
    cudaStream_t st = m.st;
    int rc;
    const ftl_path_op *d_ops = (const ftl_path_op *)m.ops.p;
    const JobDesc *d_jobs = (const JobDesc *)m.jobs.p;
    if ((rc = m.counters.ensure(sizeof(Counters), st))) return rc;
    if ((rc = m.pin_small.ensure(64))) return rc;
    if ((rc = m.pin_ring.ensure(RING * sizeof(Counters)))) return rc;
    if ((rc = m.jstate.ensure((size_t)P.n_jobs * sizeof(JobState), st))) return rc;
    if ((rc = m.cnt.ensure((size_t)(P.n_ops + 1) * sizeof(SumHead), st))) return rc;
    if ((rc = m.off.ensure(((size_t)P.n_ops + 1) * sizeof(SumHead), st))) return rc;
    if ((rc = m.tcount.ensure(((size_t)P.n_bins + 1) * sizeof(uint32_t), st))) return rc;
    if ((rc = m.toff.ensure(((size_t)P.n_bins + 1) * sizeof(uint32_t), st))) return rc;
    if ((rc = m.partials.ensure((size_t)div_up(P.n_ops + 1, SCAN_BLOCK) * sizeof(SumHead), st))) return rc;
    if ((rc = m.tpart.ensure((size_t)div_up(P.n_bins + 1, SCAN_BLOCK) * sizeof(uint32_t), st))) return rc;
    if ((rc = m.entries.ensure(sizeof(uint32_t), st))) return rc;
    if (!exact) {
        if ((rc = m.vtx.ensure(sizeof(Vtx), st))) return rc;
        if ((rc = m.edges.ensure(sizeof(EdgeRec), st))) return rc;
        if ((rc = m.sub_last.ensure(sizeof(uint32_t), st))) return rc;
        if ((rc = m.entries.ensure(sizeof(uint32_t), st))) return rc;
    }
    Counters *d_cnt = (Counters *)m.counters.p;
    JobState *d_js = (JobState *)m.jstate.p;
    uint32_t cap_v = (uint32_t)std::min<size_t>({m.vtx.cap / sizeof(Vtx), m.edges.cap / sizeof(EdgeRec), m.sub_last.cap / sizeof(uint32_t), (size_t)0x7FFFFFFF});
    uint32_t cap_e = (uint32_t)std::min<size_t>(m.entries.cap / sizeof(uint32_t), (size_t)0x7FFFFFFF);
    const uint32_t fb = std::min<uint32_t>(div_up(P.n_ops ? P.n_ops : 1, 128), (uint32_t)m.n_sms * 16);

    // ---- stages (a) and (b), as a replayable sequence ----
    // A handle that owns a row band of a raster flattens only the sub-figures that can reach its rows (or hold the
    // figure's top-left vertex): three light passes over the ops make the rest of the front end proportional to the band.
    const bool cull_on = P.cull != 0;
    CullBufs cull{nullptr, nullptr, nullptr, nullptr};
    if (cull_on) {
        if ((rc = m.cull_mark.ensure(((size_t)P.n_ops + 1) * sizeof(uint32_t), st))) return rc;
        if ((rc = m.cull_head.ensure(((size_t)P.n_ops + 1) * sizeof(uint32_t), st))) return rc;
        if ((rc = m.cull_part.ensure((size_t)div_up(P.n_ops + 1, SCAN_BLOCK) * sizeof(uint32_t), st))) return rc;
        if ((rc = m.cull_lo.ensure((size_t)P.n_ops * sizeof(int32_t), st))) return rc;
        if ((rc = m.cull_hi.ensure((size_t)P.n_ops * sizeof(int32_t), st))) return rc;
        if ((rc = m.cull_job.ensure((size_t)P.n_jobs * 2 * sizeof(int32_t), st))) return rc;
        cull = CullBufs{(const uint32_t *)m.cull_head.p, (const int32_t *)m.cull_lo.p, (const int32_t *)m.cull_hi.p, (const int32_t *)m.cull_job.p};
    }
    // Curves are subdivided once: the counting pass parks up to slab_pts points per op (8 bytes each), the emitting pass
    // copies them.  Only when the job set has curves, and within 256 MB of scratch.
    uint32_t slab_pts = 0;
    if (P.has_curves && P.n_ops > 0) {
        slab_pts = (uint32_t)std::min<size_t>(32, ((size_t)256 << 20) / ((size_t)P.n_ops * sizeof(int2)));
        if (slab_pts < 8) slab_pts = 0;
        if (const char *ev = getenv("FTL_SLAB_PTS")) slab_pts = (uint32_t)std::max(0, std::min(64, atoi(ev)));  // tuning knob
        if (slab_pts && (rc = m.slabs.ensure((size_t)P.n_ops * slab_pts * sizeof(int2), st))) return rc;
    }
    int2 *d_slabs = slab_pts ? (int2 *)m.slabs.p : nullptr;
    auto front = [&](bool sync_sizes) -> int {
        CK(cudaMemsetAsync(d_cnt, 0, sizeof(Counters), st));
        uint32_t nv_hint = cap_v;
        if (P.n_ops > 0) {
            if (cull_on) {
                const uint32_t cb = std::min<uint32_t>(div_up(P.n_ops, 256), (uint32_t)m.n_sms * 8);
                cull_init_jobs<<<div_up(2 * P.n_jobs, 256), 256, 0, st>>>((int32_t *)m.cull_job.p, P.n_jobs); LAUNCHED();
                cull_op_extents<<<cb, 256, 0, st>>>(d_ops, d_jobs, P, (uint32_t *)m.cull_mark.p, (int32_t *)m.cull_lo.p, (int32_t *)m.cull_hi.p, (int32_t *)m.cull_job.p); LAUNCHED();
                int r1 = run_scan<MaxU32>(st, (const uint32_t *)m.cull_mark.p, P.n_ops, (uint32_t *)m.cull_head.p, m.cull_part);
                if (r1) return r1;
                cull_sub_extents<<<cb, 256, 0, st>>>(d_ops, d_jobs, P, (const uint32_t *)m.cull_head.p, (int32_t *)m.cull_lo.p, (int32_t *)m.cull_hi.p); LAUNCHED();
            }
            flatten_ops<false, false><<<fb, FLAT_THREADS, P.has_curves ? FLAT_SMEM_BYTES : 0, st>>>(d_ops, d_jobs, P, nullptr, (SumHead *)m.cnt.p, nullptr, nullptr, nullptr, nullptr, cull, d_slabs, slab_pts); LAUNCHED();
            const bool small_ops = !sync_sizes && P.n_ops <= SCAN_SMALL_MAX;  // one-block scan + the capacity guard in one launch
            int r2 = FTL_OK;
            if (small_ops) {
                scan_small<SumHeadOp><<<1, SCAN_THREADS, 0, st>>>((const SumHead *)m.cnt.p, P.n_ops, (SumHead *)m.off.p, FinVertexCount{d_cnt, cap_v}); LAUNCHED();
            } else r2 = run_scan<SumHeadOp>(st, (const SumHead *)m.cnt.p, P.n_ops, (SumHead *)m.off.p, m.partials);
            if (r2) return r2;
            if (sync_sizes) {
                CK(cudaMemcpyAsync(m.pin_small.p, &((SumHead *)m.off.p)[P.n_ops].sum, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                uint32_t nv = *(uint32_t *)m.pin_small.p;
                if (nv >= 0x7FFFFFFFu) {
                    set_error("too many vertices for one call (2^31 - 1 after flattening)");
                    return FTL_ERR_INVALID;
                }
                size_t want = nv ? nv : 1;
                if ((r2 = m.vtx.ensure(want * sizeof(Vtx), st))) return r2;
                if ((r2 = m.edges.ensure(want * sizeof(EdgeRec), st))) return r2;
                if ((r2 = m.sub_last.ensure(want * sizeof(uint32_t), st))) return r2;
                cap_v = (uint32_t)std::min<size_t>({m.vtx.cap / sizeof(Vtx), m.edges.cap / sizeof(EdgeRec), m.sub_last.cap / sizeof(uint32_t), (size_t)0x7FFFFFFF});
                nv_hint = nv;
            }
            if (!small_ops) { set_vertex_count<<<1, 1, 0, st>>>(d_cnt, (const SumHead *)m.off.p, P.n_ops, cap_v); LAUNCHED(); }
            flatten_ops<false, true><<<fb, FLAT_THREADS, P.has_curves ? FLAT_SMEM_BYTES : 0, st>>>(d_ops, d_jobs, P, nullptr, nullptr, (const SumHead *)m.off.p, (Vtx *)m.vtx.p, nullptr, &d_cnt->overflow, cull, d_slabs, slab_pts); LAUNCHED();
        }
        init_job_state<<<div_up(P.n_jobs, 256), 256, 0, st>>>(d_js, d_jobs, P.n_ops > 0 ? (const SumHead *)m.off.p : nullptr, P.n_jobs, d_cnt, P.direct_max); LAUNCHED();
        const uint32_t vb = std::max<uint32_t>(1u, std::min<uint32_t>(div_up(nv_hint, 256), (uint32_t)m.n_sms * 8));
        vtx_topkey<<<vb, 256, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js); LAUNCHED();
        vtx_topvid<<<vb, 256, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js, (uint32_t *)m.sub_last.p); LAUNCHED();
        job_finalize<<<div_up(P.n_jobs, 128), 128, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js, (const uint32_t *)m.sub_last.p, P.n_jobs); LAUNCHED();
        if (P.all_direct) {  // every tile scans its job's own edges: nothing to bin
            edge_build<false><<<vb, 256, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js, (EdgeRec *)m.edges.p, P, nullptr); LAUNCHED();
            return FTL_OK;
        }
        CK(cudaMemsetAsync(m.tcount.p, 0, (size_t)P.n_bins * sizeof(uint32_t), st));
        edge_build<true><<<vb, 256, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js, (EdgeRec *)m.edges.p, P, (uint32_t *)m.tcount.p); LAUNCHED();
        const bool small_bins = !sync_sizes && P.n_bins <= SCAN_SMALL_MAX;
        int r3 = FTL_OK;
        if (small_bins) {
            scan_small<AddU32><<<1, SCAN_THREADS, 0, st>>>((const uint32_t *)m.tcount.p, P.n_bins, (uint32_t *)m.toff.p, FinEntryCount{d_cnt, cap_e}); LAUNCHED();
        } else r3 = run_scan<AddU32>(st, (const uint32_t *)m.tcount.p, P.n_bins, (uint32_t *)m.toff.p, m.tpart);
        if (r3) return r3;
        if (sync_sizes) {
            CK(cudaMemcpyAsync(m.pin_small.p, &((uint32_t *)m.toff.p)[P.n_bins], sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            uint32_t n_entries = *(uint32_t *)m.pin_small.p;
            if ((r3 = m.entries.ensure((size_t)(n_entries ? n_entries : 1) * sizeof(uint32_t), st))) return r3;
            cap_e = (uint32_t)std::min<size_t>(m.entries.cap / sizeof(uint32_t), (size_t)0x7FFFFFFF);
        }
        if (!small_bins) { set_entry_count<<<1, 1, 0, st>>>(d_cnt, (const uint32_t *)m.toff.p, P.n_bins, cap_e); LAUNCHED(); }
        CK(cudaMemsetAsync(m.tcount.p, 0, (size_t)P.n_bins * sizeof(uint32_t), st));
        bin_fill<<<vb, 256, 0, st>>>((const EdgeRec *)m.edges.p, d_cnt, d_js, P, (uint32_t *)m.tcount.p, (const uint32_t *)m.toff.p,
                                     (uint32_t *)m.entries.p); LAUNCHED();
        CK(cudaGetLastError());
        return FTL_OK;
    };

    if (exact) {
        m.drop_graph();
        if ((rc = front(true))) return rc;
    } else if (!m.use_graph || m.skip_graph) {
        if ((rc = front(false))) return rc;
    } else {
        // the sequence depends only on these values: replay the captured graph while they are unchanged
        std::vector<uint64_t> key = {(uint64_t)(uintptr_t)m.ops.p, (uint64_t)(uintptr_t)m.jobs.p, (uint64_t)(uintptr_t)m.cnt.p, (uint64_t)(uintptr_t)m.off.p,
                                     (uint64_t)(uintptr_t)m.partials.p, (uint64_t)(uintptr_t)m.vtx.p, (uint64_t)(uintptr_t)m.edges.p,
                                     (uint64_t)(uintptr_t)m.sub_last.p, (uint64_t)(uintptr_t)m.tcount.p, (uint64_t)(uintptr_t)m.toff.p,
                                     (uint64_t)(uintptr_t)m.tpart.p, (uint64_t)(uintptr_t)m.entries.p, (uint64_t)(uintptr_t)m.counters.p,
                                     (uint64_t)(uintptr_t)m.jstate.p, cap_v, cap_e, P.W, P.H, P.row_begin, P.row_end, P.fmt, P.log2R, P.n_jobs, P.n_ops,
                                     P.n_tiles, P.win_chunks, P.n_bins, P.all_direct, P.direct_max, P.b_wc, P.b_nwin, P.b_nbands, P.cull, slab_pts, (uint64_t)(uintptr_t)m.slabs.p,
                                     (uint64_t)(uintptr_t)m.cull_mark.p, (uint64_t)(uintptr_t)m.cull_head.p, (uint64_t)(uintptr_t)m.cull_part.p,
                                     (uint64_t)(uintptr_t)m.cull_lo.p, (uint64_t)(uintptr_t)m.cull_hi.p, (uint64_t)(uintptr_t)m.cull_job.p};
        if (!m.graph || key != m.graph_key) {
            m.drop_graph();
            cudaGraph_t g = nullptr;
            const uint64_t l0 = g_launches.load();
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = front(false);
            cudaError_t ce = cudaStreamEndCapture(st, &g);
            m.graph_kernels = g_launches.load() - l0;
            g_launches.fetch_sub(m.graph_kernels);  // captured, not launched
            if (rc) {
                if (g) cudaGraphDestroy(g);
                return rc;
            }
            CK(ce);
            CK(cudaGraphInstantiate(&m.graph, g, 0));
            CK(cudaGraphDestroy(g));
            m.graph_key = key;
        }
        CK(cudaGraphLaunch(m.graph, st));
        g_launches.fetch_add(m.graph_kernels, std::memory_order_relaxed);
    }

    // ---- (c)+(d) tiles ----
    int occ = 1;
    const bool aligned = P.W % (16u / P.bpp) == 0;  // rows start on 16-byte boundaries
    TileKernel tk = tile_kernel((int)P.fmt, aligned);
    const int tile_threads = (int)P.cta_warps * 32;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tk, tile_threads, m.smem_bytes));
    if (occ < 1) occ = 1;
    // A launch whose tiles all take the analytic rows into Matte8 rasters is a stream of stores with little
    // else to hide, and the wider the rows the fewer resident warps it wants: 1024/2048-px rows are fastest
    // at 5 CTAs per SM, 4096-px rows at 3 (13 % faster than 5 at every batch size tried), 8192-px rows at 2
    // (tools/occ_probe.py).  Everything that reads or scatters wants all the warps it can get.
    if (P.fmt == FTL_MATTE8 && P.all_direct && P.all_tiny && aligned && (P.W & 15u) == 0) occ = std::min(occ, P.W >= 6144u ? 2 : (P.W >= 3072u ? 3 : 5));
    if (const char *ev = getenv("FTL_OCC")) occ = std::max(1, std::min(16, atoi(ev)));  // tuning knob
    // binned jobs: one-warp CTAs of raster_bins, as many as fit
    BinKernel bk = bin_kernel((int)P.fmt, aligned, P.b_wc);
    const int bin_smem = (int)bin_smem_bytes(P.b_wc);
    int bocc = 1;
    if (!P.all_direct) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bocc, bk, 32, bin_smem));
        if (bocc < 1) bocc = 1;
        if (const char *ev = getenv("FTL_BIN_OCC")) bocc = std::max(1, std::min(32, atoi(ev)));  // tuning knob
    }
    // Independent rasters: one launch over all tiles.  Layers of one raster: one launch per job, in
    // order, each over that job's tiles (the stages before ran once for all layers).
    const uint32_t n_launches = m.layered ? P.n_jobs : 1u;
    uint32_t *d_tickets = nullptr, *d_look = nullptr;
    if (!P.all_direct) {
        if ((rc = m.tickets.ensure((size_t)n_launches * sizeof(uint32_t), st))) return rc;
        d_tickets = (uint32_t *)m.tickets.p;
        CK(cudaMemsetAsync(d_tickets, 0, (size_t)n_launches * sizeof(uint32_t), st));
        if (P.b_lookback) {
            const size_t look_bytes = (size_t)P.n_bins * 32 * sizeof(uint32_t);
            const void *before = m.look.p;
            if ((rc = m.look.ensure(look_bytes, st))) return rc;
            if (m.look.p != before || m.look_epoch + n_launches >= 0xFFFFu) m.look_epoch = 0;
            if (m.look_epoch == 0) CK(cudaMemsetAsync(m.look.p, 0, m.look.cap, st));
            d_look = (uint32_t *)m.look.p;
        }
    }
    for (uint32_t l = 0; l < n_launches; l++) {
        Params PL = P;
        PL.tile_begin = m.layered ? l * P.n_bands : 0u;
        PL.tile_end = m.layered ? (l + 1) * P.n_bands : P.n_tiles;
        PL.job_begin = m.layered ? l : 0u;
        PL.job_end = m.layered ? l + 1 : P.n_jobs;
        ProfSpan span{};
        const bool prof = g_profiling.load();
        if (prof) {
            CK(cudaEventCreate(&span.a));
            CK(cudaEventCreate(&span.b));
            CK(cudaEventRecord(span.a, st));
        }
        // jobs of at most direct_max edge slots (skipped when the host knows this layer is a binned job... it cannot: curves may flatten to few edges)
        {
            uint32_t grid = std::min<uint32_t>(div_up(PL.tile_end - PL.tile_begin, P.cta_warps), (uint32_t)(m.n_sms * occ));
            tk<<<grid, tile_threads, m.smem_bytes, st>>>((const EdgeRec *)m.edges.p, d_jobs, d_js, PL, d_cnt); LAUNCHED();
        }
        // larger jobs (skipped when the host has proven there are none in this launch)
        if (!P.all_direct && !(m.layered && m.host_direct[l])) {
            const uint64_t tasks = (uint64_t)(PL.job_end - PL.job_begin) * P.b_nbands * (P.b_lookback ? P.b_nwin : 1u);
            uint32_t grid = (uint32_t)std::min<uint64_t>(tasks, (uint64_t)m.n_sms * bocc);
            bk<<<grid, 32, bin_smem, st>>>((const EdgeRec *)m.edges.p, d_jobs, d_js, PL, (const uint32_t *)m.toff.p, (const uint32_t *)m.entries.p, d_cnt,
                                          d_tickets + l, d_look, P.b_lookback ? ++m.look_epoch : 0u); LAUNCHED();
        }
        if (prof) {
            CK(cudaEventRecord(span.b, st));
            std::lock_guard<std::mutex> lock(m.spans_mu);
            m.spans.push_back(span);
        }
    }
    CK(cudaGetLastError());
    if (!exact) {
        CK(cudaMemcpyAsync((Counters *)m.pin_ring.p + m.pending, d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        m.pending++;
    }
    return FTL_OK;
}

// Check the counters of the replays issued since the last check; repeat, with exact sizes, those
// that found a scratch buffer too small (they drew nothing).
// The last operation on the stream is a small fill: wait for its completion word in mapped host memory (a few hundred
// nanoseconds after the kernel ends; cudaStreamSynchronize needs several microseconds).  False: the caller must synchronise.
static bool small_tail_done(Engine::Impl &m) {
    if (!m.small_tail || !m.small_flags.p) return false;
    const volatile uint32_t *done = (const volatile uint32_t *)m.small_flags.p + SMALL_RING;
    for (int spin = 0; spin < 200000; spin++) {
        if (*done == m.small_seq) {
            std::atomic_thread_fence(std::memory_order_acquire);
            return true;
        }
        _mm_pause();
    }
    return false;
}
static int stream_idle(Engine::Impl &m) {
    if (small_tail_done(m)) return FTL_OK;
    CK(cudaStreamSynchronize(m.st));
    return FTL_OK;
}

static int resolve_pending(Engine::Impl &m) {
    if (m.pending == 0 && m.small_n == 0) return FTL_OK;
    {
        int rc0 = stream_idle(m);
        if (rc0) return rc0;
    }
    uint32_t redo = 0;
    const Counters *ring = (const Counters *)m.pin_ring.p;
    for (uint32_t i = 0; i < m.pending; i++) redo += ring[i].overflow ? 1u : 0u;
    m.pending = 0;
    for (uint32_t i = 0; i < redo; i++) {
        int rc = run_pipeline(m, true);
        if (rc) return rc;
    }
    // small fills that did not fit their kernel (they, and every small fill after them, drew nothing): repeat them in
    // order through the general pipeline, then lift the poison
    const uint32_t n_small = m.small_n;
    m.small_n = 0;
    bool any = false;
    for (uint32_t i = 0; i < n_small; i++) {
        if (!((const uint32_t *)m.small_flags.p)[i]) continue;
        any = true;
        const Engine::Impl::SmallSaved sv = m.small_saved[i];
        std::vector<HostJob> jobs(1);
        HostJob &h = jobs[0];
        h.op_begin = 0; h.op_end = sv.args.n_ops;
        memcpy(h.e, sv.args.job.e, sizeof(h.e));
        h.tol_sq = sv.args.job.tol_sq;
        h.rule = (int)sv.args.job.rule;
        for (int k = 0; k < 4; k++) h.color[k] = (uint8_t)(sv.args.job.color >> (8 * k));
        h.raster = (void *)(uintptr_t)sv.args.job.raster;
        int rc = upload_jobs(m, sv.geo, jobs, sv.args.ops, sv.args.n_ops, false);
        if (!rc) rc = run_pipeline(m, true);
        if (rc) return rc;
    }
    if (any) {
        CK(cudaMemsetAsync(m.small_poison.p, 0, sizeof(uint32_t), m.st));
        CK(cudaStreamSynchronize(m.st));
    }
    return FTL_OK;
}

int Engine::replay() {
    ENSURE_INIT();
    Impl &m = *impl_;
    if (!m.have_jobs) return FTL_OK;
    if (m.pending >= RING) {
        int rc = resolve_pending(m);
        if (rc) return rc;
    }
    // speculate once every scratch buffer has a capacity from an earlier call
    const bool exact = m.vtx.cap == 0 || m.edges.cap == 0 || m.entries.cap == 0;
    return run_pipeline(m, exact);
}

int Engine::last_fill_info(FillInfo *info) {
    ENSURE_INIT();
    Impl &m = *impl_;
    *info = FillInfo();
    if (!m.have_jobs && !m.last_small) return FTL_OK;
    {
        int rc = resolve_pending(m);
        if (rc) return rc;
    }
    JobState js;
    Counters c;
    CK(cudaStreamSynchronize(m.st));
    CK(cudaMemcpy(&js, m.jstate.p, sizeof(js), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&c, m.counters.p, sizeof(c), cudaMemcpyDeviceToHost));
    info->dir = js.dir;
    info->top_row = js.top_row;
    info->n_points = c.nv - c.n_popped;
    return FTL_OK;
}

int Engine::job_top_rows(uint32_t first, uint32_t count, int32_t *out) {
    ENSURE_INIT();
    Impl &m = *impl_;
    if (!m.have_jobs || (uint64_t)first + count > m.P.n_jobs) {
        set_error("no resident job set covers that range");
        return FTL_ERR_INVALID;
    }
    int rc = resolve_pending(m);
    if (rc) return rc;
    CK(cudaStreamSynchronize(m.st));
    std::vector<JobState> js(count);
    CK(cudaMemcpy(js.data(), (const JobState *)m.jstate.p + first, (size_t)count * sizeof(JobState), cudaMemcpyDeviceToHost));
    for (uint32_t j = 0; j < count; j++) out[j] = js[j].top_vid == NONE32 ? INT32_MAX : js[j].top_row;  // INT32_MAX: the job drew nothing
    return FTL_OK;
}

int Engine::debug_area(int32_t row, uint32_t width, std::vector<int16_t> *out) {
    ENSURE_INIT();
    Impl &m = *impl_;
    out->assign(width, 0);
    if ((!m.have_jobs && !m.last_small) || width == 0) return FTL_OK;
    int rc = resolve_pending(m);
    if (rc) return rc;
    JobState js;
    CK(cudaStreamSynchronize(m.st));
    CK(cudaMemcpy(&js, m.jstate.p, sizeof(js), cudaMemcpyDeviceToHost));
    const uint32_t n = js.vtx_end - js.vtx_begin;
    if (n == 0 || js.top_vid == NONE32) return FTL_OK;
    if ((rc = m.misc.ensure((size_t)width * sizeof(int32_t), m.st))) return rc;
    m.small_tail = false;
    CK(cudaMemsetAsync(m.misc.p, 0, (size_t)width * sizeof(int32_t), m.st));
    // `row` is a raster row, like the rows of the edge records (geometry row - shift)
    area_row_probe<<<div_up(n, 256), 256, 0, m.st>>>((const EdgeRec *)m.edges.p, js.vtx_begin, js.vtx_end, row, (int32_t)width, (int32_t *)m.misc.p); LAUNCHED();
    CK(cudaGetLastError());
    std::vector<int32_t> a(width);
    CK(cudaMemcpyAsync(a.data(), m.misc.p, (size_t)width * sizeof(int32_t), cudaMemcpyDeviceToHost, m.st));
    CK(cudaStreamSynchronize(m.st));
    for (uint32_t i = 0; i < width; i++) (*out)[i] = (int16_t)a[i];  // the reference's cells are wrapping i16 (plotter.rs:45)
    return FTL_OK;
}

int Engine::debug_edges(std::vector<int32_t> *out) {
    ENSURE_INIT();
    Impl &m = *impl_;
    out->clear();
    if (!m.have_jobs && !m.last_small) return FTL_OK;
    {
        int rc = resolve_pending(m);
        if (rc) return rc;
    }
    JobState js;
    CK(cudaStreamSynchronize(m.st));
    CK(cudaMemcpy(&js, m.jstate.p, sizeof(js), cudaMemcpyDeviceToHost));
    const uint32_t n = js.vtx_end - js.vtx_begin;
    if (n == 0 || js.top_vid == NONE32) return FTL_OK;  // nothing was drawn: no edges were built
    std::vector<EdgeRec> e(n);
    CK(cudaMemcpy(e.data(), (const EdgeRec *)m.edges.p + js.vtx_begin, (size_t)n * sizeof(EdgeRec), cudaMemcpyDeviceToHost));
    for (const EdgeRec &r : e) {
        if (!(r.flags & 1u)) continue;
        // raster row -> geometry row: + shift; the fractions are kept beside the rows
        const int32_t y_upper = (int32_t)(((uint32_t)(r.ry0 + js.shift) << 16) | (r.fr & 0xFFFFu));
        const int32_t y_lower = (int32_t)(((uint32_t)(r.ry1 + js.shift) << 16) | (r.fr >> 16));
        const int32_t rec[6] = {r.x_bot0, r.inv_slope, r.step_pix, y_upper, y_lower, (r.flags & 2u) ? -1 : 1};
        out->insert(out->end(), rec, rec + 6);
    }
    return FTL_OK;
}

int Engine::debug_flatten(const float e[6], float tol_sq, const ftl_path_op *ops, size_t n_ops, std::vector<int32_t> *xy,
                          std::vector<uint32_t> *subs) {
    ENSURE_INIT();
    Impl &m = *impl_;
    xy->clear();
    subs->clear();
    if (n_ops == 0) return FTL_OK;
    int rc = resolve_pending(m);
    if (rc) return rc;
    if ((rc = validate_ops(ops, n_ops))) return rc;
    cudaStream_t st = m.st;
    m.small_tail = false;
    Params P{};
    P.n_jobs = 1;
    P.n_ops = (uint32_t)n_ops;
    JobDesc jd{};
    jd.op_begin = 0; jd.op_end = (uint32_t)n_ops;
    memcpy(jd.e, e, sizeof(jd.e));
    jd.tol_sq = tol_sq;
    CK(cudaStreamSynchronize(st));
    if ((rc = m.ops.ensure(n_ops * sizeof(ftl_path_op), st))) return rc;
    if ((rc = m.jobs.ensure(sizeof(JobDesc), st))) return rc;
    m.have_jobs = false;  // the resident job set is clobbered
    CK(cudaMemcpy(m.ops.p, ops, n_ops * sizeof(ftl_path_op), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.jobs.p, &jd, sizeof(jd), cudaMemcpyHostToDevice));
    if ((rc = m.cnt.ensure(n_ops * sizeof(SumHead), st))) return rc;
    if ((rc = m.off.ensure((n_ops + 1) * sizeof(SumHead), st))) return rc;
    uint32_t fb = div_up(P.n_ops, 128);
    flatten_ops<false, false><<<fb, FLAT_THREADS, FLAT_SMEM_BYTES, st>>>((const ftl_path_op *)m.ops.p, (const JobDesc *)m.jobs.p, P, nullptr, (SumHead *)m.cnt.p, nullptr, nullptr, nullptr, nullptr, CullBufs{nullptr, nullptr, nullptr, nullptr}); LAUNCHED();
    if ((rc = run_scan<SumHeadOp>(st, (const SumHead *)m.cnt.p, P.n_ops, (SumHead *)m.off.p, m.partials))) return rc;
    SumHead tot;
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(&tot, &((SumHead *)m.off.p)[n_ops], sizeof(tot), cudaMemcpyDeviceToHost));
    uint32_t nv = tot.sum;
    if (nv == 0) return FTL_OK;
    if ((rc = m.vtx.ensure((size_t)nv * sizeof(Vtx), st))) return rc;
    flatten_ops<false, true><<<fb, FLAT_THREADS, FLAT_SMEM_BYTES, st>>>((const ftl_path_op *)m.ops.p, (const JobDesc *)m.jobs.p, P, nullptr, nullptr, (const SumHead *)m.off.p, (Vtx *)m.vtx.p, nullptr, nullptr, CullBufs{nullptr, nullptr, nullptr, nullptr}); LAUNCHED();
    CK(cudaStreamSynchronize(st));
    std::vector<Vtx> v(nv);
    CK(cudaMemcpy(v.data(), m.vtx.p, (size_t)nv * sizeof(Vtx), cudaMemcpyDeviceToHost));
    // Apply the sub-figure closing rule (fig.rs:373-383) the way edge_build sees it.
    for (uint32_t k = 0; k < nv;) {
        uint32_t s = k, e2 = k;
        while (e2 + 1 < nv && v[e2 + 1].sub == s) e2++;
        uint32_t n = e2 - s + 1;
        if (v[e2].x == v[s].x && v[e2].y == v[s].y) n--;
        if (n > 0) {
            subs->push_back((uint32_t)(xy->size() / 2));
            subs->push_back(n);
            for (uint32_t i = 0; i < n; i++) {
                xy->push_back(v[s + i].x);
                xy->push_back(v[s + i].y);
            }
        }
        k = e2 + 1;
    }
    return FTL_OK;
}

int Engine::flatten_wide(const float e[6], float tol_sq, const ftl_path_op *ops, size_t n_ops, const float *opw, WideFlat *out) {
    ENSURE_INIT();
    Impl &m = *impl_;
    out->counts.assign(n_ops, 0);
    out->xyw.clear();
    if (n_ops == 0) return FTL_OK;
    int rc = resolve_pending(m);
    if (rc) return rc;
    if ((rc = validate_ops(ops, n_ops))) return rc;
    cudaStream_t st = m.st;
    m.small_tail = false;
    Params P{};
    P.n_jobs = 1;
    P.n_ops = (uint32_t)n_ops;
    JobDesc jd{};
    jd.op_begin = 0; jd.op_end = (uint32_t)n_ops;
    memcpy(jd.e, e, sizeof(jd.e));
    jd.tol_sq = tol_sq;
    CK(cudaStreamSynchronize(st));
    if ((rc = m.ops.ensure(n_ops * sizeof(ftl_path_op), st))) return rc;
    if ((rc = m.jobs.ensure(sizeof(JobDesc), st))) return rc;
    if ((rc = m.opw.ensure(n_ops * 2 * sizeof(float), st))) return rc;
    m.have_jobs = false;
    CK(cudaMemcpy(m.ops.p, ops, n_ops * sizeof(ftl_path_op), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.jobs.p, &jd, sizeof(jd), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.opw.p, opw, n_ops * 2 * sizeof(float), cudaMemcpyHostToDevice));
    if ((rc = m.cnt.ensure(n_ops * sizeof(SumHead), st))) return rc;
    if ((rc = m.off.ensure((n_ops + 1) * sizeof(SumHead), st))) return rc;
    uint32_t fb = div_up(P.n_ops, 128);
    flatten_ops<true, false><<<fb, FLAT_THREADS, 0, st>>>((const ftl_path_op *)m.ops.p, (const JobDesc *)m.jobs.p, P, (const float *)m.opw.p, (SumHead *)m.cnt.p, nullptr, nullptr, nullptr, nullptr, CullBufs{nullptr, nullptr, nullptr, nullptr}); LAUNCHED();
    if ((rc = run_scan<SumHeadOp>(st, (const SumHead *)m.cnt.p, P.n_ops, (SumHead *)m.off.p, m.partials))) return rc;
    CK(cudaStreamSynchronize(st));
    std::vector<SumHead> off(n_ops + 1);
    CK(cudaMemcpy(off.data(), m.off.p, (n_ops + 1) * sizeof(SumHead), cudaMemcpyDeviceToHost));
    uint32_t np = off[n_ops].sum;
    for (size_t i = 0; i < n_ops; i++) out->counts[i] = off[i + 1].sum - off[i].sum;
    if (np == 0) return FTL_OK;
    if ((rc = m.wide.ensure((size_t)np * 3 * sizeof(float), st))) return rc;
    flatten_ops<true, true><<<fb, FLAT_THREADS, 0, st>>>((const ftl_path_op *)m.ops.p, (const JobDesc *)m.jobs.p, P, (const float *)m.opw.p, nullptr, (const SumHead *)m.off.p, nullptr, (float *)m.wide.p, nullptr, CullBufs{nullptr, nullptr, nullptr, nullptr}); LAUNCHED();
    CK(cudaStreamSynchronize(st));
    out->xyw.resize((size_t)np * 3);
    CK(cudaMemcpy(out->xyw.data(), m.wide.p, (size_t)np * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return FTL_OK;
}

int Engine::accumulate_rows(int rule, const int16_t *src, uint8_t *dst, size_t n, size_t rows) {
    ENSURE_INIT();
    Impl &m = *impl_;
    if (n == 0 || rows == 0) return FTL_OK;
    uint32_t chunks = ((uint32_t)n + CHUNK - 1) / CHUNK;
    size_t smem = (size_t)chunks * CHUNK * 4 + (size_t)chunks * 4;
    if (smem > m.max_smem) {
        set_error("row too long for the shared-memory row tile");
        return FTL_ERR_TOO_WIDE;
    }
    int rc;
    if ((rc = m.misc.ensure(n * rows * 3, m.st))) return rc;
    m.small_tail = false;
    int16_t *ds = (int16_t *)m.misc.p;
    uint8_t *dd = (uint8_t *)m.misc.p + n * rows * 2;
    CK(cudaMemcpyAsync(ds, src, n * rows * 2, cudaMemcpyHostToDevice, m.st));
    accumulate_rows_kernel<<<(uint32_t)rows, 32, smem, m.st>>>(ds, dd, (uint32_t)n, chunks, rule == FTL_EVENODD); LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, dd, n * rows, cudaMemcpyDeviceToHost, m.st));
    CK(cudaStreamSynchronize(m.st));
    return FTL_OK;
}

int Engine::checksums(const void *rasters, size_t raster_bytes, uint32_t count, uint64_t *out) {
    ENSURE_INIT();
    Impl &m = *impl_;
    if (count == 0) return FTL_OK;
    int rc;
    if ((rc = resolve_pending(m))) return rc;
    if ((rc = m.misc.ensure((size_t)count * 8, m.st))) return rc;
    m.small_tail = false;
    fnv_rasters<<<count, 256, 0, m.st>>>((const uint8_t *)rasters, raster_bytes, (uint64_t *)m.misc.p); LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, m.misc.p, (size_t)count * 8, cudaMemcpyDeviceToHost, m.st));
    CK(cudaStreamSynchronize(m.st));
    return FTL_OK;
}

// Linear -> sRGB encode table of an 8-bit channel: round(255 * srgb(i / 255)) (pix builds the same table at compile time).
static void srgb_table(uint8_t t[256]) {
    for (int i = 0; i < 256; i++) {
        const double u = i / 255.0;
        const double e = u <= 0.0031308 ? 12.92 * u : 1.055 * pow(u, 1.0 / 2.4) - 0.055;
        t[i] = (uint8_t)std::min(255.0, std::max(0.0, floor(e * 255.0 + 0.5)));
    }
}
void Engine::srgb_encode_table(uint8_t t[256]) { srgb_table(t); }

// The owned rows converted for output (examples/fishy.rs:33): Rgba8p -> SRgba8, Graya8p -> SGraya8, Matte8 -> SGray8.
int Engine::copy_out_srgb(void *dst, const void *dptr, size_t bytes, int format) {
    ENSURE_INIT();
    Impl &m = *impl_;
    if (format == FTL_MATTE8) return copy_out(dst, dptr, bytes);  // a reinterpretation, byte for byte (png/mod.rs:22-27)
    int rc = resolve_pending(m);
    if (rc) return rc;
    if ((bytes & 15u) != 0 || ((uintptr_t)dptr & 15u) != 0) {
        set_error("sRGB read-back needs a raster whose size is a multiple of 16 bytes");
        return FTL_ERR_INVALID;
    }
    static std::mutex mu;
    static bool table_ready[64] = {};
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!table_ready[device_]) {
            uint8_t t[256];
            srgb_table(t);
            CK(cudaMemcpyToSymbol(c_srgb_encode, t, 256));
            table_ready[device_] = true;
        }
    }
    if ((rc = m.srgb_tmp.ensure(bytes, m.st))) return rc;
    m.small_tail = false;
    const size_t n_words = bytes / 16;
    const uint32_t grid = div_up(n_words, 256);
    if (format == FTL_RGBA8P) srgb_convert<FTL_RGBA8P><<<grid, 256, 0, m.st>>>((const uint4 *)dptr, (uint4 *)m.srgb_tmp.p, n_words);
    else srgb_convert<FTL_GRAYA8P><<<grid, 256, 0, m.st>>>((const uint4 *)dptr, (uint4 *)m.srgb_tmp.p, n_words);
    LAUNCHED();
    CK(cudaGetLastError());
    return copy_out(dst, m.srgb_tmp.p, bytes);
}

int Engine::sync() {
    ENSURE_INIT();
    int rc = resolve_pending(*impl_);
    if (rc) return rc;
    return stream_idle(*impl_);
}
int Engine::alloc_raster(size_t bytes, void **dptr) {
    ENSURE_INIT();
    CK(cudaMalloc(dptr, bytes ? bytes : 1));
    return FTL_OK;
}
int Engine::free_raster(void *dptr) {
    if (!dptr) return FTL_OK;
    ENSURE_INIT();
    resolve_pending(*impl_);
    CK(cudaStreamSynchronize(impl_->st));
    CK(cudaFree(dptr));
    return FTL_OK;
}
int Engine::memset_async(void *dptr, int value, size_t bytes) {
    ENSURE_INIT();
    {
        // a speculative replay that overflowed its scratch buffers drew nothing and is repeated by
        // resolve_pending(): that repeat must land BEFORE the clear, not after it
        int rc = resolve_pending(*impl_);
        if (rc) return rc;
    }
    impl_->small_tail = false;
    CK(cudaMemsetAsync(dptr, value, bytes, impl_->st));
    return FTL_OK;
}
int Engine::copy_in(void *dptr, const void *src, size_t bytes) {
    ENSURE_INIT();
    {
        int rc = resolve_pending(*impl_);
        if (rc) return rc;
    }
    impl_->small_tail = false;
    g_h2d_bytes.fetch_add(bytes, std::memory_order_relaxed);
    CK(cudaMemcpyAsync(dptr, src, bytes, cudaMemcpyHostToDevice, impl_->st));
    CK(cudaStreamSynchronize(impl_->st));
    return FTL_OK;
}
// Expand packed units [u0, u1) into dst (host).  Streaming (non-temporal) stores: the output is
// written once and far larger than the caches, so read-for-ownership traffic would halve the rate.
static void unpack_units(uint8_t *dst, const uint8_t *code, const uint32_t *bitmap, const uint32_t *off, const uint8_t *lit, size_t u0, size_t u1) {
    const bool aligned = ((uintptr_t)dst & 15u) == 0;
    for (size_t u = u0; u < u1; u++) {
        uint8_t *d = dst + u * 1024;
        const uint8_t *c = code + u * 32;
        const uint32_t m = bitmap[u];
        const uint8_t *l = lit + (size_t)off[u] * 32;
        if (!aligned) {
            for (int i = 0; i < 32; i++) {
                if ((m >> i) & 1u) {
                    memcpy(d + 32 * i, l, 32);
                    l += 32;
                } else
                    memset(d + 32 * i, c[i], 32);
            }
            continue;
        }
        __m128i *o = reinterpret_cast<__m128i *>(d);
        if (m == 0) {
            uint64_t w0, w1, w2, w3;
            memcpy(&w0, c, 8); memcpy(&w1, c + 8, 8); memcpy(&w2, c + 16, 8); memcpy(&w3, c + 24, 8);
            const uint64_t bc = (uint64_t)c[0] * 0x0101010101010101ull;
            if (w0 == bc && w1 == bc && w2 == bc && w3 == bc) {  // one constant KiB
                const __m128i v = _mm_set1_epi8((char)c[0]);
                for (int i = 0; i < 64; i++) _mm_stream_si128(o + i, v);
                continue;
            }
        }
        for (int i = 0; i < 32; i++) {
            if ((m >> i) & 1u) {
                _mm_stream_si128(o + 2 * i, _mm_loadu_si128(reinterpret_cast<const __m128i *>(l)));
                _mm_stream_si128(o + 2 * i + 1, _mm_loadu_si128(reinterpret_cast<const __m128i *>(l + 16)));
                l += 32;
            } else {
                const __m128i v = _mm_set1_epi8((char)c[i]);
                _mm_stream_si128(o + 2 * i, v);
                _mm_stream_si128(o + 2 * i + 1, v);
            }
        }
    }
    _mm_sfence();
}

static void unpack_parallel(uint8_t *dst, const uint8_t *code, const uint32_t *bitmap, const uint32_t *off, const uint8_t *lit, size_t units,
                            unsigned nt) {
    if (nt <= 1 || units < 4096) {
        unpack_units(dst, code, bitmap, off, lit, 0, units);
        return;
    }
    const size_t per = (units + nt - 1) / nt;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) {
        size_t u0 = std::min(units, t * per), u1 = std::min(units, u0 + per);
        if (u0 < u1) th.emplace_back(unpack_units, dst, code, bitmap, off, lit, u0, u1);
    }
    for (std::thread &t : th) t.join();
}

// Device -> host copy of rasters.  Large copies are packed on the device (see pack_classify),
// moved in about eight pieces per read (64-512 MiB), and expanded by host threads while the next piece is in flight.
int Engine::copy_out(void *dst, const void *dptr, size_t bytes) {
    ENSURE_INIT();
    Impl &m = *impl_;
    {
        int rc = resolve_pending(m);
        if (rc) return rc;
    }
    cudaStream_t st = m.st;
    m.small_tail = false;
    static const bool raw_only = getenv("FTL_RAW_READ") && atoi(getenv("FTL_RAW_READ")) != 0;
    const size_t PACK_MIN = 4u << 20;
    if (raw_only || bytes < PACK_MIN || (bytes & 1023u) != 0 || ((uintptr_t)dptr & 15u) != 0) {
        g_d2h_bytes.fetch_add(bytes, std::memory_order_relaxed);
        CK(cudaMemcpyAsync(dst, dptr, bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return FTL_OK;
    }
    unsigned nt = std::min(8u, std::thread::hardware_concurrency());  // the expansion is memory-bound well before 8 threads
    if (const char *ev = getenv("FTL_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(ev));
    nt = std::max(1u, std::min(nt, 64u));
    // About eight pieces per read: each piece costs two stream synchronisations and a set of host threads
    // (~0.3 ms), a pipeline of N pieces exposes 1/N of the staging + expansion at its ends (4 GiB: 256 MiB
    // pieces 128 Gpx/s, 512 MiB 150, 1 GiB 138).
    size_t PIECE = std::min<size_t>(512u << 20, std::max<size_t>(64u << 20, ((bytes / 8) + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1)));
    if (const char *ev = getenv("FTL_PIECE_MB")) PIECE = (size_t)std::max(16, std::min(1024, atoi(ev))) << 20;  // tuning knob
    std::thread worker;
    int rc = FTL_OK;
    for (size_t at = 0, k = 0; at < bytes && rc == FTL_OK; at += PIECE, k++) {
        const size_t len = std::min(PIECE, bytes - at);
        const uint8_t *src = (const uint8_t *)dptr + at;
        uint8_t *out = (uint8_t *)dst + at;
        const size_t n_blocks = len / 32, units = n_blocks / 32;
        // fixed part: [code n_blocks][bitmap units*4][off (units+1)*4]
        const size_t code_bytes = n_blocks, bm_bytes = units * 4, off_bytes = ((units + 1) * 4 + 15) & ~(size_t)15;
        const size_t fixed = code_bytes + bm_bytes + off_bytes;
        PinBuf &pin_fix = m.pin_pack[k & 1], &pin_lit = m.pin_lit[k & 1];
        auto stage = [&]() -> int {
            int r;
            if ((r = m.pack_fixed.ensure(fixed, st))) return r;
            if ((r = m.pack_cnt.ensure(units * 4 + 16, st))) return r;
            if ((r = m.tpart.ensure((size_t)div_up(units + 1, SCAN_BLOCK) * sizeof(uint32_t), st))) return r;
            if ((r = pin_fix.ensure(fixed))) return r;
            uint8_t *d_code = (uint8_t *)m.pack_fixed.p;
            uint32_t *d_bm = (uint32_t *)(d_code + code_bytes);
            uint32_t *d_off = (uint32_t *)(d_code + code_bytes + bm_bytes);
            const uint32_t grid = div_up(n_blocks, 256);
            pack_classify<<<grid, 256, 0, st>>>((const uint4 *)src, n_blocks, d_code, d_bm, (uint32_t *)m.pack_cnt.p); LAUNCHED();
            if ((r = run_scan<AddU32>(st, (const uint32_t *)m.pack_cnt.p, (uint32_t)units, d_off, m.tpart))) return r;
            CK(cudaMemcpyAsync(pin_fix.p, d_code, fixed, cudaMemcpyDeviceToHost, st));
            g_d2h_bytes.fetch_add(fixed, std::memory_order_relaxed);
            CK(cudaStreamSynchronize(st));
            const uint32_t *h_off = (const uint32_t *)((const uint8_t *)pin_fix.p + code_bytes + bm_bytes);
            const size_t n_lit = h_off[units];
            if (n_lit * 32 > len / 2) return -1;  // not compressible
            if (n_lit) {
                if ((r = m.pack_lit.ensure(n_lit * 32, st))) return r;
                if ((r = pin_lit.ensure(n_lit * 32))) return r;
                pack_literals<<<grid, 256, 0, st>>>((const uint4 *)src, n_blocks, d_bm, d_off, (uint4 *)m.pack_lit.p); LAUNCHED();
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(pin_lit.p, m.pack_lit.p, n_lit * 32, cudaMemcpyDeviceToHost, st));
                g_d2h_bytes.fetch_add(n_lit * 32, std::memory_order_relaxed);
                CK(cudaStreamSynchronize(st));
            }
            return FTL_OK;
        };
        int r = stage();  // overlaps the expansion of the previous piece
        if (worker.joinable()) worker.join();
        if (r == -1) {  // plain copy of this piece
            g_d2h_bytes.fetch_add(len, std::memory_order_relaxed);
            cudaError_t e1 = cudaMemcpyAsync(out, src, len, cudaMemcpyDeviceToHost, st);
            cudaError_t e2 = cudaStreamSynchronize(st);
            if (e1 != cudaSuccess || e2 != cudaSuccess) {
                set_error(std::string("device to host copy: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
                rc = FTL_ERR_CUDA;
            }
            continue;
        }
        if (r) {
            rc = r;
            break;
        }
        const uint8_t *h_code = (const uint8_t *)pin_fix.p;
        const uint32_t *h_bm = (const uint32_t *)(h_code + code_bytes);
        const uint32_t *h_off = (const uint32_t *)(h_code + code_bytes + bm_bytes);
        const uint8_t *h_lit = (const uint8_t *)pin_lit.p;
        worker = std::thread(unpack_parallel, out, h_code, h_bm, h_off, h_lit, units, nt);
    }
    if (worker.joinable()) worker.join();
    return rc;
}

}  // namespace ftl
