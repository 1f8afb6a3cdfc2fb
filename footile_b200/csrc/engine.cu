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
// There is no active-edge list and no per-row serial dependency: every
// (edge,row) contribution is evaluated in closed form,
// which tests/test_oracle_orderfree.py proves equal to the reference's scan.
//
// Build flags: -fmad=false (Rust never fuses a*b+c), no fast-math.
#include <cuda_runtime.h>
#include <emmintrin.h>
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
static std::atomic<bool> g_profiling{false};
#define LAUNCHED() g_launches.fetch_add(1, std::memory_order_relaxed)

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
constexpr uint32_t DIRECT_MAX = 64;  // jobs with at most this many edge slots skip binning: their tiles scan the job's edges

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
    uint32_t pad[2];
};

struct Params {  // per-call constants, passed by value
    uint32_t W, H, row_begin, row_end;
    uint32_t fmt, bpp, pitch;
    uint32_t log2R, R, n_bands, WP, chunks;  // rows per tile, bands per job, smem row stride (cells), 512-cell chunks per row
    uint32_t n_jobs, n_ops, n_tiles;
    uint32_t win_chunks, warp_words, cta_warps;  // chunks per row window, smem words per warp, warps per CTA
    uint32_t win_rows;                           // rows of a narrow raster (one window per row) a warp holds at once
    uint32_t n_win, n_bins;                      // windows per row; bins = n_tiles * n_win
    uint32_t all_direct;                         // host-proven: every job has <= DIRECT_MAX vertices (no binning needed)
    uint32_t all_tiny;                           // host-proven: every job has <= 8 vertices (every tile can take the analytic rows)
    uint32_t tile_begin, tile_end;               // tiles this launch of the tile kernel covers
};

// ---------------------------------------------------------------------------
// generic 3-phase scan (reduce / scan partials / apply), exclusive, n+1 outputs
// ---------------------------------------------------------------------------
struct AddU32 {
    typedef uint32_t T;
    static __device__ __forceinline__ T identity() { return 0u; }
    static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
    static __device__ __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
};
struct SumHeadOp {
    typedef SumHead T;
    static __device__ __forceinline__ T identity() { return {0u, NONE32}; }
    static __device__ __forceinline__ T combine(T a, T b) { return {a.sum + b.sum, b.head != NONE32 ? a.sum + b.head : a.head}; }
    static __device__ __forceinline__ T shfl_up(T v, int d) {
        return {__shfl_up_sync(0xFFFFFFFFu, v.sum, d), __shfl_up_sync(0xFFFFFFFFu, v.head, d)};
    }
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

// Exclusive scan of one value per thread across the block; returns the
// exclusive prefix and the block total (to all threads).
template <class Op>
__device__ typename Op::T block_exclusive(typename Op::T v, typename Op::T *total) {
    typedef typename Op::T T;
    __shared__ T warp_tot[SCAN_THREADS / 32];
    __shared__ T blk_tot;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = Op::shfl_up(inc, d);
        if (lane >= d) inc = Op::combine(o, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = lane < SCAN_THREADS / 32 ? warp_tot[lane] : Op::identity();
        T winc = w;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            T o = Op::shfl_up(winc, d);
            if (lane >= d) winc = Op::combine(o, winc);
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = winc;  // inclusive over warps
        if (lane == SCAN_THREADS / 32 - 1) blk_tot = winc;
    }
    __syncthreads();
    T excl_in_warp = Op::shfl_up(inc, 1);
    if (lane == 0) excl_in_warp = Op::identity();
    T base = wid > 0 ? warp_tot[wid - 1] : Op::identity();
    *total = blk_tot;
    T r = Op::combine(base, excl_in_warp);
    __syncthreads();
    return r;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce(const typename Op::T *in, uint32_t n, typename Op::T *partials) {
    typedef typename Op::T T;
    uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    T acc = Op::identity();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) acc = Op::combine(acc, in[base + i]);
    T tot;
    block_exclusive<Op>(acc, &tot);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_partials(typename Op::T *partials, uint32_t n_blocks) {
    typedef typename Op::T T;
    T carry = Op::identity();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += SCAN_THREADS) {
        uint32_t i = b0 + threadIdx.x;
        T v = i < n_blocks ? partials[i] : Op::identity();
        T tot;
        T ex = block_exclusive<Op>(v, &tot);
        if (i < n_blocks) partials[i] = Op::combine(carry, ex);
        carry = Op::combine(carry, tot);
    }
}

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const typename Op::T *in, uint32_t n, const typename Op::T *partials,
                                                           typename Op::T *out) {
    typedef typename Op::T T;
    uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = Op::identity();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = base + i < n ? in[base + i] : Op::identity();
        acc = Op::combine(acc, v[i]);
    }
    T tot;
    T ex = block_exclusive<Op>(acc, &tot);
    T run = Op::combine(partials[blockIdx.x], ex);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run = Op::combine(run, v[i]);
        if (base + i + 1 == n) out[n] = run;
    }
}

// ---------------------------------------------------------------------------
// capacity guards: scratch buffers are sized from the previous call, so the
// pipeline runs without a host round trip; if a count exceeds its buffer the
// call draws nothing and the host repeats it with exact sizes.
// ---------------------------------------------------------------------------
__global__ void set_vertex_count(Counters *C, const SumHead *__restrict__ off, uint32_t n_ops, uint32_t cap_v) {
    uint32_t nv = off[n_ops].sum;
    if (nv > cap_v) {
        C->overflow = 1;
        C->need_v = nv;
        nv = 0;
    }
    C->nv = nv;
}
__global__ void set_entry_count(Counters *C, const uint32_t *__restrict__ toff, uint32_t n_tiles, uint32_t cap_e) {
    uint32_t n = toff[n_tiles];
    C->n_entries = n;
    if (n > cap_e) {
        C->overflow = 1;
        C->need_e = n;
    }
}

// ---------------------------------------------------------------------------
// (a) flatten — plotter.rs:175-332 + fig.rs:428-461
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t job_of_op(const JobDesc *jobs, uint32_t n_jobs, uint32_t i) {
    uint32_t lo = 0, hi = n_jobs;  // last job with op_begin <= i
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (jobs[mid].op_begin <= i) lo = mid; else hi = mid;
    }
    return lo;
}

struct PenInfo {
    pointy::Pt pen;
    bool starts_sub;
};
// Pen position when op i runs: the end point of the previous drawing op, or
// the origin after Close / at the start of the ops (plotter.rs:128-130,
// 200-203); PenWidth does not move the pen.
__device__ __forceinline__ PenInfo find_pen(const ftl_path_op *ops, uint32_t op_begin, uint32_t i) {
    int64_t k = (int64_t)i - 1;
    while (k >= (int64_t)op_begin && ops[k].tag >= FTL_OP_PENWIDTH) k--;
    if (k < (int64_t)op_begin || ops[k].tag == FTL_OP_CLOSE) return {{0.0f, 0.0f}, true};
    const ftl_path_op &o = ops[k];
    int at = o.tag == FTL_OP_QUAD ? 2 : (o.tag == FTL_OP_CUBIC ? 4 : 0);
    return {{o.v[at], o.v[at + 1]}, false};
}

struct WPt {
    pointy::Pt p;
    float w;
};
template <bool WIDE>
__device__ __forceinline__ WPt wmid(WPt a, WPt b) {  // WidePt::midpoint (geom.rs:31-35)
    WPt r;
    r.p = pointy::midpoint(a.p, b.p);
    r.w = WIDE ? (a.w + b.w) / 2.0f : 0.0f;
    return r;
}

// Point sink of one op.  Fill mode converts to Fixed and drops a point equal
// to its predecessor (fig.rs:436-440); wide mode keeps raw f32 + width for the
// host stroker.
template <bool WIDE, bool EMIT>
struct OpSink {
    uint32_t n = 0;
    bool force;
    int32_t px = 0, py = 0;
    Vtx *vout = nullptr;
    float *wout = nullptr;
    uint32_t sub = 0, job = 0;
    __device__ __forceinline__ void put(WPt q) {
        if (WIDE) {
            if (EMIT) {
                wout[3 * (size_t)n] = q.p.x;
                wout[3 * (size_t)n + 1] = q.p.y;
                wout[3 * (size_t)n + 2] = q.w;
            }
            n++;
        } else {
            int32_t fx = fx_from_f32(q.p.x), fy = fx_from_f32(q.p.y);
            if (force || fx != px || fy != py) {
                if (EMIT) vout[n] = {fx, fy, sub, job};
                n++;
            }
            force = false;
            px = fx;
            py = fy;
        }
    }
};

template <bool WIDE, bool EMIT>
__device__ void flatten_quad(WPt a, WPt b, WPt c, float tol_sq, OpSink<WIDE, EMIT> &sink) {  // plotter.rs:248-265
    WPt sb[MAX_DEPTH], sc[MAX_DEPTH];
    uint8_t sd[MAX_DEPTH];
    int sp = 0, depth = 0;
    for (;;) {
        WPt ab = wmid<WIDE>(a, b), bc = wmid<WIDE>(b, c), ab_bc = wmid<WIDE>(ab, bc), ac = wmid<WIDE>(a, c);
        if (pointy::distance_sq(ab_bc.p, ac.p) <= tol_sq || depth >= MAX_DEPTH) {
            sink.put(c);
            if (sp == 0) break;
            sp--;
            a = c; b = sb[sp]; c = sc[sp]; depth = sd[sp];
        } else {
            sb[sp] = bc; sc[sp] = c; sd[sp] = (uint8_t)(depth + 1); sp++;
            b = ab; c = ab_bc; depth++;
        }
    }
}

template <bool WIDE, bool EMIT>
__device__ void flatten_cubic(WPt a, WPt b, WPt c, WPt d, float tol_sq, OpSink<WIDE, EMIT> &sink) {  // plotter.rs:311-332
    WPt sb[MAX_DEPTH], sc[MAX_DEPTH], sdd[MAX_DEPTH];
    uint8_t sd[MAX_DEPTH];
    int sp = 0, depth = 0;
    for (;;) {
        WPt ab = wmid<WIDE>(a, b), bc = wmid<WIDE>(b, c), cd = wmid<WIDE>(c, d);
        WPt ab_bc = wmid<WIDE>(ab, bc), bc_cd = wmid<WIDE>(bc, cd);
        WPt pe = wmid<WIDE>(ab_bc, bc_cd), ad = wmid<WIDE>(a, d);
        if (pointy::distance_sq(pe.p, ad.p) <= tol_sq || depth >= MAX_DEPTH) {
            sink.put(d);
            if (sp == 0) break;
            sp--;
            a = d; b = sb[sp]; c = sc[sp]; d = sdd[sp]; depth = sd[sp];
        } else {
            sb[sp] = bc_cd; sc[sp] = cd; sdd[sp] = d; sd[sp] = (uint8_t)(depth + 1); sp++;
            b = ab; c = ab_bc; d = pe; depth++;
        }
    }
}

// One thread per PathOp.  Pass 1 (EMIT=false) counts the vertices the op
// contributes; after the scan, pass 2 (EMIT=true) repeats the identical
// subdivision and writes them at the scanned offset, so the output order is
// the reference's depth-first order.
template <bool WIDE, bool EMIT>
__global__ void __launch_bounds__(128) flatten_ops(const ftl_path_op *__restrict__ ops, const JobDesc *__restrict__ jobs, Params P,
                                                   const float *__restrict__ opw, SumHead *__restrict__ cnt,
                                                   const SumHead *__restrict__ off, Vtx *__restrict__ vout,
                                                   float *__restrict__ wout, const Counters *__restrict__ C) {
    if (EMIT && C && C->overflow) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_ops; i += gridDim.x * blockDim.x) {
        const ftl_path_op op = ops[i];
        OpSink<WIDE, EMIT> sink;
        uint32_t j = job_of_op(jobs, P.n_jobs, i);
        const JobDesc &jd = jobs[j];
        bool starts = false;
        if (op.tag >= FTL_OP_MOVE && op.tag <= FTL_OP_CUBIC) {
            PenInfo pi = find_pen(ops, jd.op_begin, i);
            starts = pi.starts_sub || op.tag == FTL_OP_MOVE;  // Move closes the current sub-figure (plotter.rs:210)
            float e[6];
#pragma unroll
            for (int k = 0; k < 6; k++) e[k] = jd.e[k];
            sink.force = starts;
            if (EMIT) {
                SumHead o = off[i];
                if (WIDE) sink.wout = wout + 3 * (size_t)o.sum;
                else {
                    sink.vout = vout + o.sum;
                    sink.sub = starts ? o.sum : o.head;
                    sink.job = j;
                }
            }
            float w_pen = WIDE ? opw[2 * (size_t)i] : 0.0f, w_now = WIDE ? opw[2 * (size_t)i + 1] : 0.0f;
            WPt a = {pointy::transform(e, pi.pen), w_pen};
            if (!WIDE && !starts) {
                sink.px = fx_from_f32(a.p.x);
                sink.py = fx_from_f32(a.p.y);
            }
            if (op.tag == FTL_OP_MOVE || op.tag == FTL_OP_LINE) {  // plotter.rs:208-224
                sink.put({pointy::transform(e, {op.v[0], op.v[1]}), w_now});
            } else if (op.tag == FTL_OP_QUAD) {  // plotter.rs:233-242
                WPt b = {pointy::transform(e, {op.v[0], op.v[1]}), WIDE ? (w_pen + w_now) / 2.0f : 0.0f};
                WPt c = {pointy::transform(e, {op.v[2], op.v[3]}), w_now};
                flatten_quad<WIDE, EMIT>(a, b, c, jd.tol_sq, sink);
            } else {  // plotter.rs:286-305; float_lerp(a,b,t) = b + (a-b)*t (geom.rs:14-16)
                float w0 = WIDE ? w_now + (w_pen - w_now) * (1.0f / 3.0f) : 0.0f;
                float w1 = WIDE ? w_now + (w_pen - w_now) * (2.0f / 3.0f) : 0.0f;
                WPt b = {pointy::transform(e, {op.v[0], op.v[1]}), w0};
                WPt c = {pointy::transform(e, {op.v[2], op.v[3]}), w1};
                WPt d = {pointy::transform(e, {op.v[4], op.v[5]}), w_now};
                flatten_cubic<WIDE, EMIT>(a, b, c, d, jd.tol_sq, sink);
            }
        }
        if (!EMIT) cnt[i] = {sink.n, starts ? 0u : NONE32};
    }
}

// ---------------------------------------------------------------------------
// (b) edge prep
// ---------------------------------------------------------------------------
// Sub-figure closing (fig.rs:373-383): the last vertex of a sub-figure is
// dropped when it equals the first.  Dropped vertices stay in the array as
// holes and are skipped.
__device__ __forceinline__ bool vtx_is_last(const Vtx *V, uint32_t nv, uint32_t k) { return k + 1 >= nv || V[k + 1].sub == k + 1; }
__device__ __forceinline__ bool vtx_same(const Vtx &a, const Vtx &b) { return a.x == b.x && a.y == b.y; }
__device__ __forceinline__ unsigned long long vtx_key(const Vtx &v) {  // (y,x) order of fig.rs:464-472
    return ((unsigned long long)((uint32_t)v.y ^ 0x80000000u) << 32) | (unsigned long long)((uint32_t)v.x ^ 0x80000000u);
}
// Forward ring neighbour of a live vertex (fig.rs:143-152)
__device__ __forceinline__ uint32_t vtx_next_fwd(const Vtx *V, uint32_t nv, uint32_t k, const Vtx &v, bool last) {
    if (last) return v.sub;
    if (vtx_is_last(V, nv, k + 1) && vtx_same(V[k + 1], V[v.sub])) return v.sub;
    return k + 1;
}

__global__ void init_job_state(JobState *JS, const JobDesc *__restrict__ jobs, const SumHead *__restrict__ off, uint32_t n_jobs) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    JobState s;
    s.top_key = ~0ull; s.top_vid = NONE32; s.dir = 0; s.top_row = 0; s.first_row = 0x7FFFFFFF; s.shift = 0;
    s.vtx_begin = off ? off[jobs[j].op_begin].sum : 0u;
    s.vtx_end = off ? off[jobs[j].op_end].sum : 0u;
    s.pad[0] = s.pad[1] = s.pad[2] = 0;
    JS[j] = s;
}

// Top-left vertex, pass 1: minimum (y,x) over the live vertices of each job (fig.rs:493-494).
__global__ void __launch_bounds__(256) vtx_topkey(const Vtx *__restrict__ V, const Counters *__restrict__ C, JobState *JS) {
    const uint32_t nv = C->nv;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
        Vtx v = V[k];
        if (vtx_is_last(V, nv, k) && vtx_same(v, V[v.sub])) continue;
        atomicMin(&JS[v.job].top_key, vtx_key(v));
    }
}
// Pass 2: the stable sort keeps the lowest vertex id among equal keys; also
// records each sub-figure's last live vertex for the Reverse ring neighbour.
__global__ void __launch_bounds__(256) vtx_topvid(const Vtx *__restrict__ V, Counters *__restrict__ C, JobState *JS,
                                                  uint32_t *__restrict__ sub_last) {
    const uint32_t nv = C->nv;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
        Vtx v = V[k];
        bool last = vtx_is_last(V, nv, k);
        bool pop = last && vtx_same(v, V[v.sub]);
        if (last) sub_last[v.sub] = pop ? (k > v.sub ? k - 1 : NONE32) : k;
        if (pop) atomicAdd(&C->n_popped, 1u);
        else if (vtx_key(v) == JS[v.job].top_key) atomicMin(&JS[v.job].top_vid, k);
    }
}

// Fig::get_dir on the top-left vertex + top_row (fig.rs:402-411,495-496)
__global__ void job_finalize(const Vtx *__restrict__ V, const Counters *__restrict__ C, JobState *JS, const uint32_t *__restrict__ sub_last,
                             uint32_t n_jobs) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    uint32_t k = JS[j].top_vid;
    if (k == NONE32) return;  // no vertices: first_row stays INT_MAX, nothing is drawn (fig.rs:491)
    const uint32_t nv = C->nv;
    Vtx v = V[k];
    uint32_t f = vtx_next_fwd(V, nv, k, v, vtx_is_last(V, nv, k));
    uint32_t r = k > v.sub ? k - 1 : sub_last[v.sub];
    Vtx pf = V[f], pr = V[r];
    fx_t ax = fx_sub(pr.x, v.x), ay = fx_sub(pr.y, v.y);
    fx_t bx = fx_sub(pf.x, v.x), by = fx_sub(pf.y, v.y);
    bool widdershins = fx_mul(ax, by) > fx_mul(bx, ay);  // fig.rs:116-119
    int32_t top = fx_to_i32(v.y);
    JS[j].dir = widdershins ? 0 : 1;
    JS[j].top_row = top;
    JS[j].first_row = top > 0 ? top : 0;
    JS[j].shift = top < 0 ? top : 0;
}

// Edge::new (fig.rs:179-210), with rows already mapped to raster rows and the
// sign against the figure direction resolved.
__device__ __forceinline__ EdgeRec make_edge(const Vtx &p0, const Vtx &p1, uint32_t job, uint32_t dd, const JobState &js) {
    EdgeRec e;
    fx_t dx = fx_sub(p1.x, p0.x), dy = fx_sub(p1.y, p0.y);
    e.step_pix = dx != 0 ? fx_min(fx_abs(fx_div(dy, dx)), FX_ONE) : 0;
    e.inv_slope = fx_div(dx, dy);
    fx_t y_bot = fx_sub(fx_floor(fx_add(p0.y, FX_ONE)), p0.y);
    e.x_bot0 = fx_add(p0.x, fx_mul(e.inv_slope, y_bot));
    e.ry0 = fx_to_i32(p0.y) - js.shift;
    e.ry1 = fx_to_i32(p1.y) - js.shift;
    e.fr = (uint32_t)fx_fract(p0.y) | ((uint32_t)fx_fract(p1.y) << 16);
    e.job = job;
    e.flags = 1u | ((dd != (uint32_t)js.dir ? 1u : 0u) << 1);
    return e;
}

// One thread per vertex k: the ring segment (k, next_fwd(k)) becomes at most
// one edge, directed from its upper to its lower vertex.  This is the same
// set of edges the reference creates in update_edges/add_edge (fig.rs:576-600)
// when it visits both neighbours of every vertex.
__global__ void __launch_bounds__(256) edge_build(const Vtx *__restrict__ V, const Counters *__restrict__ C, const JobState *__restrict__ JS,
                                                  EdgeRec *__restrict__ E) {
    const uint32_t nv = C->nv;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
        Vtx v = V[k];
        bool last = vtx_is_last(V, nv, k);
        bool pop = last && vtx_same(v, V[v.sub]);
        EdgeRec e;
        e.flags = 0;
        if (!pop) {
            uint32_t w = vtx_next_fwd(V, nv, k, v, last);
            if (w != k) {
                Vtx q = V[w];
                if (q.y > v.y) e = make_edge(v, q, v.job, 0u, JS[v.job]);        // v is upper; w is v's Forward neighbour
                else if (q.y < v.y) e = make_edge(q, v, v.job, 1u, JS[v.job]);   // w is upper; v is w's Reverse neighbour
            }
        }
        if (e.flags) E[k] = e;
        else E[k].flags = 0;
    }
}

// Band range of an edge inside this device's rows; returns false if none.
__device__ __forceinline__ bool edge_bands(const EdgeRec &e, const Params &P, uint32_t *b0, uint32_t *b1) {
    int32_t lo = e.ry0, hi = e.ry1;  // ry0 >= first_row >= 0 by construction
    if (lo < (int32_t)P.row_begin) lo = (int32_t)P.row_begin;
    if (hi > (int32_t)P.row_end - 1) hi = (int32_t)P.row_end - 1;
    if (lo > hi) return false;
    *b0 = (uint32_t)(lo - (int32_t)P.row_begin) >> P.log2R;
    *b1 = (uint32_t)(hi - (int32_t)P.row_begin) >> P.log2R;
    return true;
}

// Conservative range of row windows an edge can write to on the rows [ra, rb] of one band (both
// inside the edge's own rows).  The span of a row is linear in the row, so the extremes are at the
// two end rows, evaluated as edge_row_setup does; the scatter loop can run at most |dx/dy| + 2 cells
// past the leftmost one.  Anything that would wrap 32-bit arithmetic falls back to "all windows".
__device__ __forceinline__ void edge_windows(const EdgeRec &e, int32_t ra, int32_t rb, const Params &P, uint32_t *w0, uint32_t *w1) {
    *w0 = 0;
    *w1 = P.n_win - 1;
    if (P.n_win == 1) return;
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    const int64_t slope = e.inv_slope;
    for (int t = 0; t < 2; t++) {
        const int32_t r = t ? rb : ra;
        const int64_t x_bot = (int64_t)e.x_bot0 + (int64_t)(r - e.ry0) * slope;
        const int64_t x_top = x_bot - slope;
        if (x_bot != (int32_t)x_bot || x_top != (int32_t)x_top) return;
        lo = min(lo, min(x_bot, x_top));
        hi = max(hi, max(x_bot, x_top));
    }
    const int64_t run = (slope < 0 ? -slope : slope) >> 16;
    int64_t lo_pix = (lo >> 16) - 1, hi_pix = (hi >> 16) + run + 3;
    const int64_t wmax = (int64_t)P.W - 1;
    lo_pix = lo_pix < 0 ? 0 : (lo_pix > wmax ? wmax : lo_pix);
    hi_pix = hi_pix < 0 ? 0 : (hi_pix > wmax ? wmax : hi_pix);
    const uint32_t win_cells = P.win_chunks * 512u;
    *w0 = (uint32_t)lo_pix / win_cells;
    *w1 = (uint32_t)hi_pix / win_cells;
}

// Counting sort of edges by (job, row band, row window): pass FILL=false counts, pass FILL=true
// writes edge ids at the scanned offsets.  Short edges are handled by their own thread; an edge
// crossing many bands is spread over the warp.  Jobs with at most DIRECT_MAX edge slots are not
// binned at all.
template <bool FILL>
__device__ __forceinline__ void bin_one(const EdgeRec &e, uint32_t k, uint32_t tile, uint32_t band, const Params &P, uint32_t *tile_count,
                                        const uint32_t *tile_off, uint32_t *entries) {
    const int32_t row0 = (int32_t)P.row_begin + (int32_t)(band << P.log2R);
    const int32_t ra = max(e.ry0, row0), rb = min(min(e.ry1, row0 + (int32_t)P.R - 1), (int32_t)P.row_end - 1);
    uint32_t w0, w1;
    edge_windows(e, ra, rb, P, &w0, &w1);
    for (uint32_t w = w0; w <= w1; w++) {
        const uint32_t bin = tile * P.n_win + w;
        const uint32_t slot = atomicAdd(&tile_count[bin], 1u);
        if (FILL) entries[tile_off[bin] + slot] = k;
    }
}
template <bool FILL>
__global__ void __launch_bounds__(256) bin_edges(const EdgeRec *__restrict__ E, const Counters *__restrict__ C,
                                                 const JobState *__restrict__ JS, Params P, uint32_t *__restrict__ tile_count,
                                                 const uint32_t *__restrict__ tile_off, uint32_t *__restrict__ entries) {
    if (FILL && C->overflow) return;
    const uint32_t nv = C->nv;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t span = gridDim.x * blockDim.x;
    for (uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x - lane; k0 < nv; k0 += span) {
        uint32_t k = k0 + lane;
        uint32_t b0 = 0, nb = 0, tbase = 0;
        EdgeRec e;
        e.flags = 0;
        if (k < nv) {
            e = E[k];
            uint32_t b1;
            if ((e.flags & 1u) && edge_bands(e, P, &b0, &b1)) {
                const JobState &js = JS[e.job];
                if (js.vtx_end - js.vtx_begin > DIRECT_MAX) {
                    nb = b1 - b0 + 1;
                    tbase = e.job * P.n_bands;
                }
            }
        }
        if (nb > 0 && nb <= 4)
            for (uint32_t b = b0; b < b0 + nb; b++) bin_one<FILL>(e, k, tbase + b, b, P, tile_count, tile_off, entries);
        uint32_t tall = __ballot_sync(0xFFFFFFFFu, nb > 4);
        while (tall) {
            int src = __ffs(tall) - 1;
            tall &= tall - 1;
            uint32_t sb0 = __shfl_sync(0xFFFFFFFFu, b0, src), snb = __shfl_sync(0xFFFFFFFFu, nb, src);
            uint32_t stb = __shfl_sync(0xFFFFFFFFu, tbase, src);
            const EdgeRec es = E[k0 + src];
            for (uint32_t b = lane; b < snb; b += 32) bin_one<FILL>(es, k0 + src, stb + sb0 + b, sb0 + b, P, tile_count, tile_off, entries);
        }
    }
}

// ---------------------------------------------------------------------------
// (c)+(d) tile raster kernel
// ---------------------------------------------------------------------------
// Shared-memory row tile: R rows; a row is `chunks` chunks of 512 i32 cells
// plus a 4-cell pad, followed (after all rows) by R*chunks touched-group masks
// (bit g of mask word c = some cell of the 16-cell group g of chunk c is
// non-zero).  i32 sums truncated to i16 at resolve are the reference's
// wrapping i16 sums: truncation is a ring homomorphism.
//
// Cells are XOR-swizzled at 16-byte granularity so that a lane can own 16
// CONSECUTIVE cells (4 LDS.128) without bank conflicts: quad q lives at
// q ^ ((q >> 3) & 3).
constexpr uint32_t CHUNK = 512;
__device__ __forceinline__ uint32_t cell_phys(uint32_t c) {
    return c ^ ((c >> 3) & 0xCu);  // bits 3:2 (the quad within 4 quads) ^= bits 6:5
}
// The row window is addressed through 32-bit shared-window addresses and explicit ld/st/red.shared:
// a generic pointer makes the compiler rebuild the window base (S2UR CgaCtaId + ULEA) at every access.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sred_add(uint32_t a, int32_t v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sred_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t slds(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void ssts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ int4 slds4(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void ssts4_zero(uint32_t a) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u));
}

// Signed coverage of one edge on one raster row.  Closed form of
// Scanner::scan_continuing_edges / add_edge + Edge::scan_area
// (fig.rs:238-321,557-600); SURVEY Appendix A.4.  `edge_row_setup` evaluates
// everything that does not depend on the pixel; `edge_row_scatter` then adds
// the per-cell deltas of the cells inside one column window into the shared
// row buffer and can be resumed window after window.
struct EdgeRowState {
    int32_t cov;   // coverage of the row by this edge, 1..256; 0 = nothing (left) to do
    int32_t xc;    // x_cov of the next cell, already clamped to ONE
    int32_t step;  // x_cov increment per cell
    int32_t prev;  // X of the previous cell (0 before the first)
    int32_t ed;    // +1 / -1 (fig.rs:286)
    int32_t c;     // next cell
};
__device__ __forceinline__ EdgeRowState edge_row_setup(const EdgeRec &e, int32_t ry, int32_t W, int32_t win_lo) {
    EdgeRowState st;
    const bool starting = ry == e.ry0, ending = ry == e.ry1;
    const fx_t fr0 = (fx_t)(e.fr & 0xFFFFu), fr1 = (fx_t)(e.fr >> 16);
    // continuing_cov / starting_cov (fig.rs:238-241,252-259)
    st.cov = (ending ? pixel_cov(fr1) : 256) - (starting ? pixel_cov(fr0) : 0);
    // advance_edges in closed form (fig.rs:569-573)
    fx_t x_bot = (fx_t)((uint32_t)e.x_bot0 + (uint32_t)(ry - e.ry0) * (uint32_t)e.inv_slope);
    // calculate_x_limits_* / set_x_limits (fig.rs:244-249,262-278); ceil(y)-y = (ONE - fract) & MASK
    fx_t x0 = starting ? fx_sub(x_bot, fx_mul(e.inv_slope, FX_ONE - fr0)) : fx_sub(x_bot, e.inv_slope);
    fx_t x1 = ending ? fx_sub(x_bot, fx_mul(e.inv_slope, (FX_ONE - fr1) & FX_MASK)) : x_bot;
    fx_t min_x = fx_min(x0, x1), max_x = fx_max(x0, x1);
    const int32_t min_pix = fx_to_i32(min_x), max_pix = fx_to_i32(max_x);
    if (st.cov < 0 || min_pix >= W) st.cov = 0;
    // first_cov / step_cov (fig.rs:305-321); full_cov = cov/256 in Fixed = cov << 8
    fx_t rr = min_pix == max_pix ? fx_mul(fx_sub(FX_ONE, fx_fract(fx_avg(max_x, min_x))), (fx_t)(st.cov << 8))
                                 : fx_mul(fx_sub(FX_ONE, fx_fract(min_x)), FX_HALF);
    const fx_t first = e.step_pix > 0 ? fx_mul(rr, e.step_pix) : rr;
    st.step = e.step_pix > 0 ? e.step_pix : FX_ONE;
    st.ed = (e.flags & 2u) ? -1 : 1;
    // scan_area (fig.rs:285-302): X(k) = min(pixel_cov(min(first + k*step, 1)), cov); cell min_pix+k
    // receives X(k)-X(k-1); cells left of 0 fold into cell 0, which receives X(-min_pix).
    const int32_t c0 = min_pix > 0 ? min_pix : 0;
    st.c = c0 > win_lo ? c0 : win_lo;  // a later pass of a wide row starts inside the span
    int64_t xc = (int64_t)first + (int64_t)(st.c - min_pix) * (int64_t)st.step;
    st.xc = (int32_t)(xc < (int64_t)FX_ONE ? xc : (int64_t)FX_ONE);
    st.prev = 0;
    if (st.c > c0) {
        int64_t xq = xc - (int64_t)st.step;
        int32_t xk = pixel_cov((fx_t)(xq < (int64_t)FX_ONE ? xq : (int64_t)FX_ONE));
        st.prev = xk < st.cov ? xk : st.cov;
        if (st.prev >= st.cov) st.cov = 0;
    }
    return st;
}
// The cells of [st.c, win_hi): adds each cell's delta into the shared row window (which starts at
// column win_lo) and leaves `st` ready to continue in the next window.
__device__ __forceinline__ void edge_row_scatter(EdgeRowState &st, int32_t win_lo, int32_t win_hi, uint32_t cells, uint32_t mask) {
    if (st.cov <= 0 || st.c >= win_hi) return;
    const int32_t first_rel = st.c - win_lo;
    int32_t rel = first_rel;
    const int32_t end_rel = win_hi - win_lo;
    for (;;) {
        int32_t xk = pixel_cov(st.xc);
        if (xk > st.cov) xk = st.cov;
        const int32_t d = xk - st.prev;
        if (d != 0) sred_add(cells + 4u * cell_phys((uint32_t)rel), st.ed * d);
        st.prev = xk;
        rel++;
        st.xc += st.step;  // both <= ONE: no overflow
        if (st.xc > FX_ONE) st.xc = FX_ONE;
        if (xk >= st.cov) {
            st.cov = 0;  // finished: every later cell receives 0
            break;
        }
        if (rel >= end_rel) break;
    }
    st.c = win_lo + rel;
    // mark the 16-cell groups [first_rel >> 4, (rel - 1) >> 4] of the window as touched
    for (uint32_t g = (uint32_t)first_rel >> 4, g1 = (uint32_t)(rel - 1) >> 4; g <= g1;) {
        const uint32_t top = min(g1, g | 31u);
        sred_or(mask + 4u * (g >> 5), ((2u << (top - g)) - 1u) << (g & 31u));
        g = top + 1;
    }
}

// Four consecutive pixels: wrapped-i16 sums (p_i + base) -> alpha bytes
// (fig.rs:637-664; imgbuf.rs:54-66,157-167), two pixels per 16x2 SIMD op.
template <bool EVEN_ODD>
__device__ __forceinline__ uint32_t quad_alpha(int32_t p0, int32_t p1, int32_t p2, int32_t p3, int32_t base) {
    if (!EVEN_ODD) {
        const uint32_t bp = __byte_perm((uint32_t)base, (uint32_t)base, 0x1010);
        uint32_t lo = __byte_perm((uint32_t)p0, (uint32_t)p1, 0x5410), hi = __byte_perm((uint32_t)p2, (uint32_t)p3, 0x5410);
        lo = __viaddmin_s16x2_relu(lo, bp, 0x00FF00FFu);  // clamp(i16(p + base), 0, 255) per halfword
        hi = __viaddmin_s16x2_relu(hi, bp, 0x00FF00FFu);
        return __byte_perm(lo, hi, 0x6420);
    } else {
        const uint32_t bp = __byte_perm((uint32_t)base, (uint32_t)base, 0x1010);
        uint32_t lo = __byte_perm((uint32_t)p0, (uint32_t)p1, 0x5410), hi = __byte_perm((uint32_t)p2, (uint32_t)p3, 0x5410);
        lo = __viaddmin_s16x2(lo, bp, 0x7FFF7FFFu);  // wrapping i16 add of the base, per halfword
        hi = __viaddmin_s16x2(hi, bp, 0x7FFF7FFFu);
        // |(s & 0xFF) - (s & 0x100)| = odd ? 256 - v : v, then 256 saturates to 255
        uint32_t bl = (lo >> 8) & 0x00010001u, bh = (hi >> 8) & 0x00010001u;
        lo = ((lo & 0x00FF00FFu) ^ (bl * 0xFFu)) + bl;
        hi = ((hi & 0x00FF00FFu) ^ (bh * 0xFFu)) + bh;
        lo = __vimin_s16x2_relu(lo, 0x00FF00FFu);
        hi = __vimin_s16x2_relu(hi, 0x00FF00FFu);
        return __byte_perm(lo, hi, 0x6420);
    }
}

__device__ __forceinline__ uint32_t blend_rgba_general(uint32_t px, uint32_t color, uint32_t alpha, uint32_t clr_a) {
    uint32_t sa1 = 255u - pix::ch8_mul(alpha, clr_a);
    uint32_t o = 0;
#pragma unroll
    for (int ch = 0; ch < 4; ch++) o |= pix::src_over_ch((px >> (8 * ch)) & 0xFF, (color >> (8 * ch)) & 0xFF, alpha, sa1) << (8 * ch);
    return o;
}
// Ch8 d * Ch8(255) on the four channels of a pixel at once: with pix's 12-bit multiply this is
// d - 1 for 1 <= d <= 15 and d otherwise (pix_compat.cuh; checked exhaustively in tests/test_host.py).
__device__ __forceinline__ uint32_t mul255_delta(uint32_t w) {  // w - (w * Ch8(255)) per byte: 1 for bytes 1..15
    // plain integer ops on purpose: the __vset*4 video intrinsics (emulated through inline lop3 on
    // sm_100a) were mis-scheduled under if-conversion in this kernel
    uint32_t nz = w | (w >> 4);
    nz |= nz >> 2;
    nz |= nz >> 1;  // bit 0 of each byte: the byte is non-zero
    uint32_t hi = w & 0xF0F0F0F0u;
    hi |= hi >> 2;
    hi |= hi >> 1;  // bit 4 of each byte: the byte is >= 16
    return nz & ~(hi >> 4) & 0x01010101u;
}
__device__ __forceinline__ uint32_t mul255_x4(uint32_t w) { return w - mul255_delta(w); }
// Alpha 0 over 16 bytes of pixels (d * Ch8(255) per channel): most pixels do not change (only channel
// values 1..15 do), and unchanged words are not written back - the blend then costs its read only.
__device__ __forceinline__ void mul255_rmw(uint4 *p, const uint4 t) {
    // cheap test first: bit 4 of a byte of (lo + 15) is "low nibble != 0", of (hi + 15) "high nibble != 0"
    // (nibbles spread to bytes cannot carry into the neighbour); a byte changes iff low != 0 and high == 0
    const uint32_t K = 0x0F0F0F0Fu;
    uint32_t any = ((t.x & K) + K) & ~(((t.x >> 4) & K) + K);
    any |= ((t.y & K) + K) & ~(((t.y >> 4) & K) + K);
    any |= ((t.z & K) + K) & ~(((t.z >> 4) & K) + K);
    any |= ((t.w & K) + K) & ~(((t.w >> 4) & K) + K);
    if (any & 0x10101010u) *p = make_uint4(mul255_x4(t.x), mul255_x4(t.y), mul255_x4(t.z), mul255_x4(t.w));
}
// SrcOver of one Rgba8p pixel (the 4-pixel and 512-pixel fast paths for alpha = 0 and for opaque
// full coverage live in emit16 / resolve_row).
__device__ __forceinline__ uint32_t blend_rgba(uint32_t px, uint32_t color, uint32_t alpha, uint32_t clr_a) {
    return blend_rgba_general(px, color, alpha, clr_a);
}

// alpha of one pixel from the wrapped i16 sum (fig.rs:637-664; imgbuf.rs:54-66,157-167)
template <bool EVEN_ODD>
__device__ __forceinline__ uint32_t rule_alpha(int32_t sum) {
    int32_t s = (int32_t)(int16_t)sum;
    if (EVEN_ODD) {
        int32_t c = (s & 0xFF) - (s & 0x100);
        s = c < 0 ? -c : c;
    }
    return (uint32_t)(s < 0 ? 0 : (s > 255 ? 255 : s));
}

// Output of one pixel.
template <int FMT>
__device__ __forceinline__ void emit1(uint8_t *dst, uint32_t x, uint32_t W, uint32_t alpha, uint32_t color, uint32_t clr_a) {
    if (x >= W) return;
    if (FMT == FTL_MATTE8) dst[x] = (uint8_t)alpha;
    else if (FMT == FTL_RGBA8P) {
        uint32_t *d = reinterpret_cast<uint32_t *>(dst) + x;
        *d = blend_rgba(*d, color, alpha, clr_a);
    } else {
        uint16_t *d = reinterpret_cast<uint16_t *>(dst) + x;
        uint32_t p = *d, sa1 = 255u - pix::ch8_mul(alpha, clr_a);
        *d = (uint16_t)(pix::src_over_ch(p & 0xFF, color & 0xFF, alpha, sa1) | (pix::src_over_ch(p >> 8, (color >> 8) & 0xFF, alpha, sa1) << 8));
    }
}

// Output of one lane's 16 pixels: alpha words a[0..3] (4 pixels each).
template <int FMT, bool ALIGNED>
__device__ __forceinline__ void emit16(uint8_t *dst, uint32_t x, uint32_t W, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                       uint32_t color, uint32_t clr_a) {
    if (x >= W) return;
    if (FMT == FTL_MATTE8) {  // store, colour ignored (fig.rs:632-636; imgbuf.rs:59,93)
        uint8_t *d = dst + x;
        if (ALIGNED && x + 16 <= W) {
            *reinterpret_cast<uint4 *>(d) = make_uint4(a0, a1, a2, a3);
        } else {
#pragma unroll
            for (uint32_t i = 0; i < 16; i++) {
                uint32_t w = i < 4 ? a0 : (i < 8 ? a1 : (i < 12 ? a2 : a3));
                if (x + i < W) d[i] = (uint8_t)(w >> (8 * (i & 3)));
            }
        }
    } else if (FMT == FTL_RGBA8P) {  // fig.rs:641-642,662-663 via pix (pix_compat.cuh)
        uint32_t *d = reinterpret_cast<uint32_t *>(dst) + x;
#pragma unroll 1
        for (int j = 0; j < 4; j++) {  // not unrolled: the general blend is large and this body is instantiated at six call sites
            const uint32_t w = j == 0 ? a0 : (j == 1 ? a1 : (j == 2 ? a2 : a3));
            if (ALIGNED && x + 4 * j + 4 <= W) {
                uint4 *q4 = reinterpret_cast<uint4 *>(d) + j;
                if (w == 0xFFFFFFFFu && clr_a == 255) {  // four opaque pixels: no read
                    const uint32_t c = mul255_x4(color);
                    *q4 = make_uint4(c, c, c, c);
                } else {
                    uint4 t = *q4;
                    if (w == 0) {
                        mul255_rmw(q4, t);
                        continue;
                    } else {
                        t.x = blend_rgba(t.x, color, w & 0xFF, clr_a);
                        t.y = blend_rgba(t.y, color, (w >> 8) & 0xFF, clr_a);
                        t.z = blend_rgba(t.z, color, (w >> 16) & 0xFF, clr_a);
                        t.w = blend_rgba(t.w, color, w >> 24, clr_a);
                    }
                    *q4 = t;
                }
            } else {
#pragma unroll
                for (uint32_t i = 0; i < 4; i++)
                    if (x + 4 * j + i < W) d[4 * j + i] = blend_rgba(d[4 * j + i], color, (w >> (8 * i)) & 0xFF, clr_a);
            }
        }
    } else {  // Graya8p
        uint16_t *d = reinterpret_cast<uint16_t *>(dst) + x;
#pragma unroll
        for (uint32_t i = 0; i < 16; i++) {
            uint32_t w = i < 4 ? a0 : (i < 8 ? a1 : (i < 12 ? a2 : a3));
            if (x + i < W) {
                uint32_t al = (w >> (8 * (i & 3))) & 0xFF, p = d[i];
                uint32_t sa1 = 255u - pix::ch8_mul(al, clr_a);
                d[i] = (uint16_t)(pix::src_over_ch(p & 0xFF, color & 0xFF, al, sa1) | (pix::src_over_ch(p >> 8, (color >> 8) & 0xFF, al, sa1) << 8));
            }
        }
    }
}

// Pixels [16 * lo, 16 * hi) of a row (whole 16-pixel groups, 16-byte aligned) take one constant alpha
// `a`: consecutive lanes on consecutive 16 bytes.  Matte8 stores; Graya8p / Rgba8p blend SrcOver with
// the two cheap cases of pix (alpha 0: d * Ch8(255); opaque coverage of an opaque colour: no read).
template <int FMT>
__device__ __forceinline__ void fill_const(uint8_t *drow, uint32_t lo, uint32_t hi, uint32_t a, uint32_t color, uint32_t clr_a) {
    const uint32_t lane = threadIdx.x & 31;
    constexpr uint32_t U = FMT == FTL_MATTE8 ? 1u : (FMT == FTL_GRAYA8P ? 2u : 4u);  // uint4 per group
    uint4 *p = reinterpret_cast<uint4 *>(drow);
    uint32_t u = lo * U + lane;
    const uint32_t end = hi * U;
    if (FMT == FTL_MATTE8 || (a == 255u && clr_a == 255u)) {
        uint32_t w = a * 0x01010101u;
        if (FMT == FTL_RGBA8P) w = mul255_x4(color);
        if (FMT == FTL_GRAYA8P) w = mul255_x4((color & 0xFFFFu) * 0x00010001u);
        uint32_t q0 = w, q1 = w, q2 = w, q3 = w;
        asm volatile("" : "+r"(q0), "+r"(q1), "+r"(q2), "+r"(q3));  // four resident registers: no per-store moves
        const uint4 v = make_uint4(q0, q1, q2, q3);
#pragma unroll 1
        for (; u + 32 < end; u += 64) {
            p[u] = v;
            p[u + 32] = v;
        }
        if (u < end) p[u] = v;
    } else if (a == 0u) {
#pragma unroll 1
        for (; u + 96 < end; u += 128) {  // four loads in flight per lane
            const uint4 t0 = p[u], t1 = p[u + 32], t2 = p[u + 64], t3 = p[u + 96];
            mul255_rmw(p + u, t0);
            mul255_rmw(p + u + 32, t1);
            mul255_rmw(p + u + 64, t2);
            mul255_rmw(p + u + 96, t3);
        }
#pragma unroll 1
        for (; u < end; u += 32) mul255_rmw(p + u, p[u]);
    } else {
        const uint32_t sa1 = 255u - pix::ch8_mul(a, clr_a);
#pragma unroll 1
        for (; u < end; u += 32) {
            uint4 t = p[u];
#pragma unroll 1
            for (int k = 0; k < 4; k++) {  // rolled: the general blend is large
                const uint32_t wk = k == 0 ? t.x : (k == 1 ? t.y : (k == 2 ? t.z : t.w));
                uint32_t o = 0;
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const uint32_t sc = FMT == FTL_RGBA8P ? (color >> (8 * ch)) & 0xFF : (color >> (8 * (ch & 1))) & 0xFF;
                    o |= pix::src_over_ch((wk >> (8 * ch)) & 0xFF, sc, a, sa1) << (8 * ch);
                }
                if (k == 0) t.x = o;
                else if (k == 1) t.y = o;
                else if (k == 2) t.z = o;
                else t.w = o;
            }
            p[u] = t;
        }
    }
}

// One step of an inclusive add-scan over segments of WIDTH lanes: v += the value `d` lanes below,
// predicated by the shuffle's own in-range result (no lane compare).
template <int WIDTH>
__device__ __forceinline__ void scan_step(int32_t &v, int d) {
    asm volatile("{ .reg .pred p; .reg .s32 t; shfl.sync.up.b32 t|p, %0, %1, %2, 0xffffffff; @p add.s32 %0, %0, t; }"
                 : "+r"(v)
                 : "r"(d), "r"((32 - WIDTH) << 8));
}

// Resolve chunks [c_begin, c_end) of one row held in shared memory, by one
// warp.  Per 512-cell chunk each lane owns 16 consecutive cells: it reads them
// (4 LDS.128), zeroes them, scans them serially, one 5-step shuffle scan
// carries the lane totals across the warp, then the fill rule turns the 16
// sums into 16 alpha bytes which are stored (Matte8: one STG.128 per lane) or
// blended SrcOver into the raster row (Graya8p/Rgba8p).  A chunk whose mask
// word is zero holds no edge: its pixels take the constant alpha of the
// running sum without touching shared memory.
template <int FMT, bool EVEN_ODD, bool ALIGNED>
__device__ __forceinline__ void resolve_row(uint32_t row, uint32_t mask, uint8_t *dst, uint32_t W, uint32_t c_begin, uint32_t c_end,
                                            int32_t &carry_io, uint32_t color) {
    int32_t carry = carry_io;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sw = (lane >> 1) & 3u;
    const uint32_t clr_a = FMT == FTL_RGBA8P ? (color >> 24) : ((color >> 8) & 0xFF);
    const uint32_t n = c_end - c_begin;  // <= 32 chunks per warp
    // the warp's chunk masks: lane i holds (and clears) the mask of chunk c_begin + i
    uint32_t mym = 0;
    if (lane < n) {
        mym = slds(mask + 4u * (c_begin + lane));
        if (mym) ssts(mask + 4u * (c_begin + lane), 0u);
    }
    const uint32_t dense = __ballot_sync(0xFFFFFFFFu, mym != 0);
    uint32_t q = quad_alpha<EVEN_ODD>(0, 0, 0, 0, carry);  // alpha of an edge-free span at the current sum
    uint4 *out4 = reinterpret_cast<uint4 *>(dst) + lane;
    const uint32_t full_end = min(c_end, W / CHUNK);  // chunks below this index lie entirely inside the row
    // Walk the chunks that hold edges (bits of `dense`); the edge-free chunks between them are runs of
    // one constant alpha.
    uint32_t pend = dense, ch = c_begin;
    for (;;) {
        const uint32_t nxt = c_begin + (pend ? (uint32_t)__ffs((int)pend) - 1u : n);
        if (ALIGNED && ch < min(nxt, full_end)) {
            const uint32_t stop = min(nxt, full_end);
            fill_const<FMT>(dst, ch * 32u, stop * 32u, q & 0xFFu, color, clr_a);
            ch = stop;
        }
#pragma unroll 1
        for (; ch < nxt; ch++) emit16<FMT, ALIGNED>(dst, ch * CHUNK + lane * 16, W, q, q, q, q, color, clr_a);  // ragged last chunk
        if (pend == 0) break;
        pend &= pend - 1;
        const uint32_t cur = ch++;  // == nxt: the chunk to resolve
        const uint32_t x = cur * CHUNK + lane * 16;
        const bool full = cur < full_end;
        const uint32_t m = __shfl_sync(0xFFFFFFFFu, mym, cur - c_begin);
        if (__popc(m) <= 4) {
            // Sparse chunk: each touched 16-cell group is scanned by a half-warp (one cell per lane,
            // two groups per step); the other groups take the constant alpha of the sum reaching them.
            int32_t mybase = carry;  // sum reaching group `lane` of this chunk
            const uint32_t half = lane >> 4, l16 = lane & 15u;
            for (uint32_t mm = m; mm;) {
                const int32_t g = __ffs(mm) - 1;
                mm &= mm - 1;
                const int32_t g2 = mm ? __ffs(mm) - 1 : -1;
                if (g2 >= 0) mm &= mm - 1;
                const int32_t gg = half ? g2 : g;
                int32_t inc = 0;
                if (gg >= 0) {
                    const uint32_t p = row + 4u * (cur * CHUNK + cell_phys((uint32_t)gg * 16u + l16));
                    inc = (int32_t)slds(p);
                    ssts(p, 0u);
                }
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) scan_step<16>(inc, d);
                const int32_t t_lo = __shfl_sync(0xFFFFFFFFu, inc, 15), t_hi = __shfl_sync(0xFFFFFFFFu, inc, 31);
                if (gg >= 0) emit1<FMT>(dst, cur * CHUNK + (uint32_t)gg * 16 + l16, W, rule_alpha<EVEN_ODD>(carry + inc + (half ? t_lo : 0)), color, clr_a);
                if ((int32_t)lane > g) mybase += t_lo;
                if (g2 >= 0 && (int32_t)lane > g2) mybase += t_hi;
                carry += t_lo + t_hi;
            }
            if (!((m >> lane) & 1u)) {
                const uint32_t qq = quad_alpha<EVEN_ODD>(0, 0, 0, 0, mybase);
                if (FMT == FTL_MATTE8 && ALIGNED && full) out4[cur * 32] = make_uint4(qq, qq, qq, qq);
                else emit16<FMT, ALIGNED>(dst, x, W, qq, qq, qq, qq, color, clr_a);
            }
            q = quad_alpha<EVEN_ODD>(0, 0, 0, 0, carry);
            continue;
        }
        int4 v0 = make_int4(0, 0, 0, 0), v1 = v0, v2 = v0, v3 = v0;
        if ((m >> lane) & 1u) {
            const uint32_t base = row + 4u * (cur * CHUNK + lane * 16);
            const uint32_t p0 = base + ((0 ^ sw) << 4), p1 = base + ((1 ^ sw) << 4), p2 = base + ((2 ^ sw) << 4), p3 = base + ((3 ^ sw) << 4);
            v0 = slds4(p0); v1 = slds4(p1); v2 = slds4(p2); v3 = slds4(p3);
            ssts4_zero(p0); ssts4_zero(p1); ssts4_zero(p2); ssts4_zero(p3);
        }
        // lane-local inclusive prefix: 4 independent quad scans, then quad offsets
        v0.y += v0.x; v0.z += v0.y; v0.w += v0.z;
        v1.y += v1.x; v1.z += v1.y; v1.w += v1.z;
        v2.y += v2.x; v2.z += v2.y; v2.w += v2.z;
        v3.y += v3.x; v3.z += v3.y; v3.w += v3.z;
        const int32_t o1 = v0.w, o2 = o1 + v1.w, o3 = o2 + v2.w, tot = o3 + v3.w;
        int32_t inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) scan_step<32>(inc, d);
        const int32_t b0 = carry + inc - tot;
        carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
        const uint32_t a0 = quad_alpha<EVEN_ODD>(v0.x, v0.y, v0.z, v0.w, b0);
        const uint32_t a1 = quad_alpha<EVEN_ODD>(v1.x, v1.y, v1.z, v1.w, b0 + o1);
        const uint32_t a2 = quad_alpha<EVEN_ODD>(v2.x, v2.y, v2.z, v2.w, b0 + o2);
        const uint32_t a3 = quad_alpha<EVEN_ODD>(v3.x, v3.y, v3.z, v3.w, b0 + o3);
        if (FMT == FTL_MATTE8 && ALIGNED && full) out4[cur * 32] = make_uint4(a0, a1, a2, a3);
        else emit16<FMT, ALIGNED>(dst, x, W, a0, a1, a2, a3, color, clr_a);
        q = quad_alpha<EVEN_ODD>(0, 0, 0, 0, carry);
    }
    carry_io = carry;
}

// ---------------------------------------------------------------------------
// Analytic rows: tiles with at most 8 edges and no shared-memory scatter.
// ---------------------------------------------------------------------------
// The sum the reference accumulates at pixel x of a row is  sum_e ed_e * X_e(x)  (mod 2^16), where X_e
// is edge e's running coverage (Edge::scan_area, fig.rs:285-302: the prefix of the deltas it adds):
// 0 left of the edge's span, min(pixel_cov(min(first + k*step, ONE)), cov) on the span's k-th cell and
// cov right of it.  With at most 8 edges the lanes of a warp are (row, edge) pairs of 4 rows: when the
// spans of a row lie in distinct 16-pixel groups, a lane knows the constant sum left of its span from
// an 8-lane exchange (`base`), builds the 16 alpha bytes of its group(s) in registers and stores them,
// and the warp stores the constant spans between the edges cooperatively.  Rows where two spans share
// a group, or a span covers more than two groups, are returned in a mask and take the shared-memory
// path.  Nothing here touches shared memory.
__device__ __forceinline__ uint32_t rule_alpha_rt(int32_t sum, bool even_odd) {
    int32_t s = (int32_t)(int16_t)sum;
    int32_t c = (s & 0xFF) - (s & 0x100);
    c = c < 0 ? -c : c;
    s = even_odd ? c : s;
    return (uint32_t)min(max(s, 0), 255);
}

// Rows ry_base .. ry_base + n_rows - 1 (n_rows <= 4) of one tile; `st` is this lane's (row my_r, edge)
// state from edge_row_setup with win_lo = 0 (cov == 0: nothing on this row).  `dst` is the first row.
// Returns the rows (bit r) that were NOT drawn and need the shared-memory path.
template <int FMT>
__device__ __forceinline__ uint32_t analytic_rows(const EdgeRowState &st, uint32_t my_r, int32_t n_rows, int32_t W, uint8_t *dst, uint32_t pitch,
                                                  bool even_odd, uint32_t color) {
    const uint32_t clr_a = FMT == FTL_RGBA8P ? (color >> 24) : ((color >> 8) & 0xFF);
    const uint32_t ngroups = (uint32_t)W >> 4;
    const bool active = st.cov > 0 && (int32_t)my_r < n_rows;
    // extent of the span: cells [st.c, c_last]
    int32_t c_last = 0;
    bool conflict = false;
    if (active) {
        int32_t xc = st.xc, c = st.c;
        for (int j = 0;; j++) {
            if (pixel_cov(xc) >= st.cov || c >= W - 1) break;
            if (j >= 31) {
                conflict = true;
                break;
            }
            c++;
            xc += st.step;
            if (xc > FX_ONE) xc = FX_ONE;
        }
        c_last = c;
    }
    const uint32_t ga = active ? (uint32_t)st.c >> 4 : 0xFFFFu, gb = active ? (uint32_t)c_last >> 4 : 0u;
    if (active && gb - ga > 1u) conflict = true;
    const int32_t wgt = active ? st.ed * st.cov : 0;
    // 8-lane exchange: constant sum left of the span, the group where the next span starts, overlaps
    const uint32_t packed = ga | (gb << 16);
    int32_t base = 0;
    uint32_t next_ga = ngroups, min_ga = ga;
#pragma unroll
    for (int d = 1; d < 8; d++) {
        const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, packed, d);
        const int32_t ow = __shfl_xor_sync(0xFFFFFFFFu, wgt, d);
        const uint32_t oga = o & 0xFFFFu, ogb = o >> 16;
        if (ogb < ga) base += ow;
        if (oga <= gb && ogb >= ga) conflict = true;
        if (oga > gb) next_ga = min(next_ga, oga);
        min_ga = min(min_ga, oga);
    }
    const uint32_t conf_bal = __ballot_sync(0xFFFFFFFFu, conflict);
    uint32_t redo = 0;  // rows for the shared-memory path
#pragma unroll
    for (int r = 0; r < 4; r++)
        if (r < n_rows && ((conf_bal >> (8 * r)) & 0xFFu)) redo |= 1u << r;
    const bool row_ok = (int32_t)my_r < n_rows && !((redo >> my_r) & 1u);
    uint8_t *drow = dst + (size_t)my_r * pitch;
    const uint32_t after_a = rule_alpha_rt(base + wgt, even_odd), after = after_a * 0x01010101u;
    // ---- the groups holding the span ----
    if (active && row_ok) {
        const uint32_t before = rule_alpha_rt(base, even_odd) * 0x01010101u;
        // group ga: `before` left of the span start, `after` right of it; the span's cells are inserted
        const uint32_t p0 = (uint32_t)st.c & 15u;
        uint32_t a[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
            const uint32_t nb = p0 > 4 * j ? min(p0 - 4 * j, 4u) : 0u;  // bytes of word j left of the span
            const uint32_t m = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
            a[j] = (before & m) | (after & ~m);
        }
        int32_t xc = st.xc, c = st.c;
#pragma unroll 1
        for (uint32_t g = ga;; g++) {
            bool done;
            do {  // the span's cells inside group g
                int32_t xk = pixel_cov(xc);
                if (xk > st.cov) xk = st.cov;
                const uint32_t al = rule_alpha_rt(base + st.ed * xk, even_odd);
                const uint32_t sh = ((uint32_t)c & 3u) * 8u, m = 0xFFu << sh, v = al << sh, wj = ((uint32_t)c >> 2) & 3u;
#pragma unroll
                for (uint32_t j = 0; j < 4; j++)
                    if (wj == j) a[j] = (a[j] & ~m) | v;
                done = xk >= st.cov || c >= W - 1;
                c++;
                xc += st.step;
                if (xc > FX_ONE) xc = FX_ONE;
            } while (!done && ((uint32_t)c & 15u) != 0);
            emit16<FMT, true>(drow, g * 16, (uint32_t)W, a[0], a[1], a[2], a[3], color, clr_a);
            if (done) break;
            a[0] = a[1] = a[2] = a[3] = after;  // the span continues in the next group
        }
    }
    // ---- the constant spans: right of every span, and left of the first one ----
    const uint32_t owners = __ballot_sync(0xFFFFFFFFu, active && row_ok);
    const uint32_t my_span = (gb + 1u) | (next_ga << 16);
#pragma unroll 1
    for (uint32_t m = owners; m; m &= m - 1) {
        const uint32_t s = (uint32_t)__ffs((int)m) - 1u;
        const uint32_t sp = __shfl_sync(0xFFFFFFFFu, my_span, s), q = __shfl_sync(0xFFFFFFFFu, after_a, s);
        fill_const<FMT>(dst + (size_t)(s >> 3) * pitch, sp & 0xFFFFu, sp >> 16, q, color, clr_a);
    }
#pragma unroll 1
    for (int r = 0; r < n_rows; r++) {
        const uint32_t hi = min(__shfl_sync(0xFFFFFFFFu, min_ga, 8 * r), ngroups);
        if (!((redo >> r) & 1u)) fill_const<FMT>(dst + (size_t)r * pitch, 0u, hi, 0u, color, clr_a);
    }
    return redo;
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// The tile kernel.  Every WARP owns a private shared-memory row window
// (`win_chunks` chunks of 512 cells + their masks) and walks the rows of a
// (job, band) tile on its own: its lanes scatter the coverage of the edges
// crossing the row, then the row is resolved and written, window after
// window, with the running sum carried across windows.  Warps never wait for
// each other, and the small window keeps many warps resident per SM, which is
// what hides the latency of the serial scatter -> scan -> store chain.
template <int FMT, bool ALIGNED, bool GENERAL>
__global__ void __launch_bounds__(128, 5) raster_tiles(const EdgeRec *__restrict__ E, const JobDesc *__restrict__ jobs,
                                                       const JobState *__restrict__ JS, Params P, const uint32_t *__restrict__ tile_off,
                                                       const uint32_t *__restrict__ entries, const Counters *__restrict__ C) {
    if (C->overflow) return;
    extern __shared__ __align__(16) int32_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
    const uint32_t cells = smem_addr(smem) + 4u * warp * P.warp_words, mask = cells + 4u * P.win_rows * P.win_chunks * CHUNK;
    for (uint32_t i = lane; i < P.warp_words; i += 32) ssts(cells + 4u * i, 0u);
    __syncwarp();
    const int32_t W = (int32_t)P.W, win_cells = (int32_t)(P.win_chunks * CHUNK);
    const uint32_t bpp = P.bpp;
    const uint32_t n_warps = gridDim.x * warps_per_cta;
    // analytic rows need full 16-pixel groups on 16-byte boundaries (and group indices below 0xFFFF)
    constexpr bool ANALYTIC = ALIGNED && !GENERAL;
    const bool analytic_ok = ANALYTIC && (P.W & 15u) == 0 && (P.W >> 4) < 0xFFFFu;
    uint32_t j = (P.tile_begin + blockIdx.x * warps_per_cta + warp) / P.n_bands, j_next = 0;
    for (uint32_t tile = P.tile_begin + blockIdx.x * warps_per_cta + warp; tile < P.tile_end; tile += n_warps, j = j_next) {
        const uint32_t band = tile - j * P.n_bands;
        // the records the next tile of this warp starts with: in flight while this tile is drawn
        uint32_t vb_next = 0, ve_next = 0;
        if (tile + n_warps < P.tile_end) {
            j_next = (tile + n_warps) / P.n_bands;
            prefetch_l1(&jobs[j_next]);
            vb_next = JS[j_next].vtx_begin;
            ve_next = JS[j_next].vtx_end;
        }
        const JobState js = JS[j];
        int32_t row0 = (int32_t)P.row_begin + (int32_t)(band << P.log2R);
        int32_t row_hi = row0 + (int32_t)P.R;
        if (row_hi > (int32_t)P.row_end) row_hi = (int32_t)P.row_end;
        if (row0 < js.first_row) row0 = js.first_row;  // rows above the figure are untouched (fig.rs:497)
        if (row0 >= row_hi) continue;
        const unsigned long long raster = jobs[j].raster;
        const uint32_t rule = jobs[j].rule, color = jobs[j].color;
        const uint32_t n_slots = js.vtx_end - js.vtx_begin;
        // Edge list of the tile: the job's own edges (direct: at most DIRECT_MAX slots; the only case
        // when GENERAL is false), or the (tile, window) bins.  With a single list for all windows the
        // first 32 edges stay in registers for all rows of the tile.
        const bool direct = !GENERAL || n_slots <= DIRECT_MAX;
        const bool one_list = !GENERAL || direct || P.n_win == 1;
        uint32_t e0 = direct ? js.vtx_begin : tile_off[tile * P.n_win];
        uint32_t ne = direct ? n_slots : tile_off[tile * P.n_win + 1] - e0;
        // Row groups: with few edges the 32 lanes are (row, edge) pairs of 4 (or 2) consecutive rows, so
        // the per-(edge,row) set-up of several rows costs one pass; each row then scatters with its own lanes.
        const uint32_t gl = !one_list ? 5u : (ne <= 8 ? 3u : (ne <= 16 ? 4u : 5u));
        const uint32_t my_e = lane & ((1u << gl) - 1u), my_r = lane >> gl;
        // Narrow rasters (one window per row) with lanes = edges: the window holds `win_rows` rows, every
        // lane scatters all the rows of its edges in one pass, then the rows are resolved one by one.
        const bool multi = gl == 5u && one_list && P.win_rows > 1u;
        const int32_t rows_per_pass = multi ? (int32_t)P.win_rows : (int32_t)(32u >> gl);
        const uint32_t row_bytes = 4u * P.win_chunks * CHUNK, rmask_bytes = 4u * P.win_chunks;
        EdgeRec mine;
        mine.flags = 0;
        if (one_list && my_e < ne) mine = E[direct ? e0 + my_e : entries[e0 + my_e]];
        uint8_t *dst = reinterpret_cast<uint8_t *>(raster) + (size_t)(row0 - (int32_t)P.row_begin) * P.pitch;
        const uint32_t win_bytes = (uint32_t)win_cells * bpp;
        for (int32_t ry_base = row0; ry_base < row_hi; ry_base += rows_per_pass) {
            EdgeRowState st;
            st.cov = 0;
            {
                const int32_t my_ry = ry_base + (int32_t)my_r;
                if (!multi && (mine.flags & 1u) && my_ry < row_hi && my_ry >= mine.ry0 && my_ry <= mine.ry1) st = edge_row_setup(mine, my_ry, W, 0);
            }
            const int32_t rr_end = min(rows_per_pass, row_hi - ry_base);
            uint32_t redo = 0xFFFFFFFFu;
            if (ry_base + rows_per_pass >= row_hi && vb_next + lane < ve_next && lane < 8) prefetch_l1(&E[vb_next + lane]);  // last pass
            if (analytic_ok && gl == 3u) redo = analytic_rows<FMT>(st, my_r, rr_end, W, dst, P.pitch, rule == FTL_EVENODD, color);
            const int32_t ry_last = ry_base + rr_end - 1;
            for (int32_t rr = 0; rr < rr_end; rr++, dst += P.pitch) {
                if (!((redo >> rr) & 1u)) continue;
                const int32_t ry = ry_base + rr;
                const uint32_t rcells = multi ? cells + (uint32_t)rr * row_bytes : cells, rmask = multi ? mask + (uint32_t)rr * rmask_bytes : mask;
                int32_t carry = 0;
                uint32_t bin = tile * P.n_win;
                uint8_t *dwin = dst;
                for (int32_t win_lo = 0; win_lo < W; win_lo += win_cells, bin++, dwin += win_bytes) {
                    const int32_t win_hi = min(W, win_lo + win_cells);
                    // ---- (c) scatter: one lane per edge crossing this row (multi: all rows of the pass at once) ----
                    if (!multi || rr == 0) {
                        uint32_t first = 32;
                        if (multi) first = 0;
                        else if (one_list) {
                            if ((int32_t)my_r == rr) edge_row_scatter(st, win_lo, win_hi, cells, mask);
                        } else {  // wide raster with many edges: each window has its own bin
                            e0 = tile_off[bin];
                            ne = tile_off[bin + 1] - e0;
                            first = 0;
                        }
                        const int32_t ra = multi ? ry_base : ry, rb = multi ? ry_last : ry;
                        for (uint32_t i = lane + first; i < ne; i += 32) {
                            const EdgeRec e = (multi && i < 32) ? mine : E[direct ? e0 + i : entries[e0 + i]];
                            if (!(e.flags & 1u)) continue;
                            const int32_t r1 = min(e.ry1, rb);
                            for (int32_t r = max(e.ry0, ra); r <= r1; r++) {
                                EdgeRowState s2 = edge_row_setup(e, r, W, win_lo);
                                const uint32_t ro = (uint32_t)(r - ra);
                                edge_row_scatter(s2, win_lo, win_hi, cells + ro * row_bytes, mask + ro * rmask_bytes);
                            }
                        }
                        __syncwarp();
                    }
                    // ---- (d) resolve ----
                    const uint32_t nch = ((uint32_t)(win_hi - win_lo) + CHUNK - 1) / CHUNK;
                    if (rule == FTL_EVENODD) resolve_row<FMT, true, ALIGNED>(rcells, rmask, dwin, (uint32_t)(W - win_lo), 0, nch, carry, color);
                    else resolve_row<FMT, false, ALIGNED>(rcells, rmask, dwin, (uint32_t)(W - win_lo), 0, nch, carry, color);
                    __syncwarp();
                }
            }
        }
    }
}

// Kernel (d) alone, for the imgbuf.rs KATs: one warp per row of i16 cells.
__global__ void __launch_bounds__(32) accumulate_rows_kernel(const int16_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t n,
                                                             uint32_t chunks, int even_odd) {
    extern __shared__ __align__(16) int32_t area[];
    uint32_t *mask = reinterpret_cast<uint32_t *>(area + chunks * CHUNK);
    const int16_t *s = src + (size_t)blockIdx.x * n;
    for (uint32_t i = threadIdx.x; i < chunks * CHUNK; i += 32) area[cell_phys(i)] = i < n ? (int32_t)s[i] : 0;
    for (uint32_t i = threadIdx.x; i < chunks; i += 32) mask[i] = 0xFFFFFFFFu;
    __syncwarp();
    uint8_t *d = dst + (size_t)blockIdx.x * n;
    const bool al = (n & 15u) == 0;
    int32_t carry = 0;
    if (even_odd) {
        if (al) resolve_row<FTL_MATTE8, true, true>(smem_addr(area), smem_addr(mask), d, n, 0, chunks, carry, 0);
        else resolve_row<FTL_MATTE8, true, false>(smem_addr(area), smem_addr(mask), d, n, 0, chunks, carry, 0);
    } else {
        if (al) resolve_row<FTL_MATTE8, false, true>(smem_addr(area), smem_addr(mask), d, n, 0, chunks, carry, 0);
        else resolve_row<FTL_MATTE8, false, false>(smem_addr(area), smem_addr(mask), d, n, 0, chunks, carry, 0);
    }
}

// ---------------------------------------------------------------------------
// packed read-back: rasters are mostly long constant spans, and PCIe is ~100x
// slower than HBM, so device->host copies of large rasters travel as
//   code[b]   : the byte value of 32-byte block b if the block is uniform
//   bitmap[u] : bit i set = block 32u+i is literal (not uniform)
//   off[u]    : literal blocks before unit u (exclusive scan of the popcounts)
//   literals  : the literal blocks, 32 bytes each, in order
// and are expanded into the caller's buffer by host threads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_classify(const uint4 *__restrict__ src, size_t n_blocks, uint8_t *__restrict__ code,
                                                     uint32_t *__restrict__ bitmap, uint32_t *__restrict__ cnt) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // n_blocks is a multiple of 32: warps are full or empty
    if (b >= n_blocks) return;
    const uint4 lo = src[2 * b], hi = src[2 * b + 1];
    const uint32_t v = (lo.x & 0xFFu) * 0x01010101u;
    const bool uniform = lo.x == v && lo.y == v && lo.z == v && lo.w == v && hi.x == v && hi.y == v && hi.z == v && hi.w == v;
    code[b] = (uint8_t)(v & 0xFFu);
    const uint32_t lit = __ballot_sync(0xFFFFFFFFu, !uniform);
    if ((threadIdx.x & 31) == 0) {
        bitmap[b >> 5] = lit;
        cnt[b >> 5] = __popc(lit);
    }
}
__global__ void __launch_bounds__(256) pack_literals(const uint4 *__restrict__ src, size_t n_blocks, const uint32_t *__restrict__ bitmap,
                                                     const uint32_t *__restrict__ off, uint4 *__restrict__ lit) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t m = bitmap[b >> 5], lane = threadIdx.x & 31;
    if ((m >> lane) & 1u) {
        const size_t k = (size_t)off[b >> 5] + __popc(m & ((1u << lane) - 1u));
        lit[2 * k] = src[2 * b];
        lit[2 * k + 1] = src[2 * b + 1];
    }
}

// 64-bit FNV-1a per raster (parity checks of large batches): one CTA per
// raster hashes 256 interleaved lanes, then lane digests are folded in order.
__global__ void __launch_bounds__(256) fnv_rasters(const uint8_t *__restrict__ base, size_t raster_bytes, uint64_t *__restrict__ out) {
    __shared__ uint64_t part[256];
    const uint8_t *p = base + (size_t)blockIdx.x * raster_bytes;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = threadIdx.x; i < raster_bytes; i += 256) h = (h ^ p[i]) * 0x100000001b3ull;
    part[threadIdx.x] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t g = 0xcbf29ce484222325ull;
        for (int i = 0; i < 256; i++)
            for (int b = 0; b < 8; b++) g = (g ^ ((part[i] >> (8 * b)) & 0xFF)) * 0x100000001b3ull;
        out[blockIdx.x] = g;
    }
}

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

struct ProfSpan {
    cudaEvent_t a, b;
};
static std::vector<ProfSpan> g_spans;
static std::mutex g_spans_mu;
static double g_tile_ms = 0.0;
static uint64_t g_tile_launches = 0;

struct Engine::Impl {
    cudaStream_t st = nullptr;
    int n_sms = 148;
    size_t max_smem = 0;
    DevBuf ops, jobs, jstate, cnt, off, partials, vtx, edges, sub_last, tcount, toff, tpart, entries, counters, opw, wide, misc;
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
void Engine::set_profiling(bool on) { g_profiling.store(on); }
void Engine::tile_kernel_time(bool reset, double *ms, uint64_t *launches) {
    std::lock_guard<std::mutex> lock(g_spans_mu);
    for (ProfSpan &s : g_spans) {
        float t = 0.f;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
            g_tile_ms += t;
            g_tile_launches++;
        }
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    g_spans.clear();
    if (ms) *ms = g_tile_ms;
    if (launches) *launches = g_tile_launches;
    if (reset) {
        g_tile_ms = 0.0;
        g_tile_launches = 0;
    }
}

Engine::Engine(int device) : impl_(new Impl()), device_(device), stream_(nullptr) {}

Engine::~Engine() {
    if (impl_) {
        if (impl_->st) {
            cudaSetDevice(device_);
            cudaStreamSynchronize(impl_->st);
            Impl &m = *impl_;
            for (DevBuf *b : {&m.ops, &m.jobs, &m.jstate, &m.cnt, &m.off, &m.partials, &m.vtx, &m.edges, &m.sub_last, &m.tcount, &m.toff,
                              &m.tpart, &m.entries, &m.counters, &m.opw, &m.wide, &m.misc, &m.pack_fixed, &m.pack_cnt, &m.pack_lit})
                b->release();
            m.drop_graph();
            for (PinBuf *b : {&m.pin_ops, &m.pin_jobs, &m.pin_small, &m.pin_misc, &m.pin_ring, &m.pin_pack[0], &m.pin_pack[1], &m.pin_lit[0], &m.pin_lit[1]}) b->release();
            cudaStreamDestroy(impl_->st);
        }
        delete impl_;
    }
}

static int engine_init(Engine::Impl *m, int device, void **stream_out);
static int run_pipeline(Engine::Impl &m, bool exact);
static int resolve_pending(Engine::Impl &m);

typedef void (*TileKernel)(const EdgeRec *, const JobDesc *, const JobState *, Params, const uint32_t *, const uint32_t *, const Counters *);
static TileKernel tile_kernel(int fmt, bool aligned, bool general) {
    if (general) switch (fmt) {
        case FTL_MATTE8: return aligned ? raster_tiles<FTL_MATTE8, true, true> : raster_tiles<FTL_MATTE8, false, true>;
        case FTL_GRAYA8P: return aligned ? raster_tiles<FTL_GRAYA8P, true, true> : raster_tiles<FTL_GRAYA8P, false, true>;
        default: return aligned ? raster_tiles<FTL_RGBA8P, true, true> : raster_tiles<FTL_RGBA8P, false, true>;
        }
    switch (fmt) {
    case FTL_MATTE8: return aligned ? raster_tiles<FTL_MATTE8, true, false> : raster_tiles<FTL_MATTE8, false, false>;
    case FTL_GRAYA8P: return aligned ? raster_tiles<FTL_GRAYA8P, true, false> : raster_tiles<FTL_GRAYA8P, false, false>;
    default: return aligned ? raster_tiles<FTL_RGBA8P, true, false> : raster_tiles<FTL_RGBA8P, false, false>;
    }
}

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
                for (int a = 0; a < 4; a++) {
                    CK(cudaFuncSetAttribute(tile_kernel(f, (a & 1) != 0, (a & 2) != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)di.max_smem));
                    CK(cudaFuncSetAttribute(tile_kernel(f, (a & 1) != 0, (a & 2) != 0), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                }
            CK(cudaFuncSetAttribute(accumulate_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)di.max_smem));
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

static int choose_tiling(const Geometry &g, size_t max_smem, uint32_t jobs_per_launch, uint32_t warp_slots, bool all_direct, Params *P) {
    P->W = g.width; P->H = g.height; P->row_begin = g.row_begin; P->row_end = g.row_end;
    P->fmt = (uint32_t)g.format; P->bpp = g.bpp(); P->pitch = (uint32_t)g.pitch();
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

    // Band height: 8 rows per warp amortise the per-tile set-up when every tile scans its job's own few
    // edges; binned jobs do better with 4 (fewer edges per bin to test against each row); fewer rows per
    // band when one launch would otherwise leave most of the GPU's warp slots empty (a single raster, a
    // layer of a scene).
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
    return FTL_OK;
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

int Engine::upload(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops, bool layered) {
    ENSURE_INIT();
    Impl &m = *impl_;
    {
        int rc0 = resolve_pending(m);  // earlier replays refer to the job set that is about to be replaced
        if (rc0) return rc0;
    }
    m.have_jobs = false;  // stays false if anything below fails
    m.layered = layered;
    if (jobs.empty() || g.rows() == 0 || g.width == 0) {
        m.have_jobs = false;
        return FTL_OK;
    }
    if (n_ops >= 0x7FFFFFFFull || jobs.size() >= 0x7FFFFFFFull) {
        set_error("too many ops/jobs for one call");
        return FTL_ERR_INVALID;
    }
    int rc = validate_ops(ops, n_ops);
    if (rc) return rc;
    Params P{};
    // Line-only jobs of at most DIRECT_MAX ops cannot exceed DIRECT_MAX vertices: no binning at all.
    P.all_direct = 1;
    P.all_tiny = 1;
    for (const HostJob &h : jobs) {
        if (h.op_end - h.op_begin > 8u) P.all_tiny = 0;
        if (h.op_end - h.op_begin > DIRECT_MAX) P.all_direct = 0;
        for (uint32_t i = h.op_begin; i < h.op_end && P.all_direct; i++)
            if (ops[i].tag == FTL_OP_QUAD || ops[i].tag == FTL_OP_CUBIC) P.all_direct = 0;
        if (!P.all_direct) break;
    }
    rc = choose_tiling(g, m.max_smem, layered ? 1u : (uint32_t)jobs.size(), (uint32_t)m.n_sms * 16u, P.all_direct != 0, &P);
    if (rc) return rc;
    P.n_jobs = (uint32_t)jobs.size();
    P.n_ops = (uint32_t)n_ops;
    uint64_t nt = (uint64_t)P.n_jobs * P.n_bands;
    if (nt >= 0x7FFFFFFFull) {
        set_error("too many tiles for one call");
        return FTL_ERR_INVALID;
    }
    P.n_tiles = (uint32_t)nt;
    if (nt * P.n_win >= 0x7FFFFFFFull) {
        set_error("too many tiles for one call");
        return FTL_ERR_INVALID;
    }
    P.n_bins = (uint32_t)(nt * P.n_win);
    // stage + upload ops and job descriptors
    size_t ops_bytes = n_ops * sizeof(ftl_path_op), jobs_bytes = jobs.size() * sizeof(JobDesc);
    if ((rc = m.pin_ops.ensure(ops_bytes ? ops_bytes : 1))) return rc;
    if ((rc = m.pin_jobs.ensure(jobs_bytes))) return rc;
    // the previous call's async copies out of the pinned staging must be done before it is overwritten
    CK(cudaStreamSynchronize(m.st));
    if (ops_bytes) memcpy(m.pin_ops.p, ops, ops_bytes);
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
    if ((rc = m.ops.ensure(ops_bytes ? ops_bytes : 1, m.st))) return rc;
    if ((rc = m.jobs.ensure(jobs_bytes, m.st))) return rc;
    if (ops_bytes) CK(cudaMemcpyAsync(m.ops.p, m.pin_ops.p, ops_bytes, cudaMemcpyHostToDevice, m.st));
    CK(cudaMemcpyAsync(m.jobs.p, m.pin_jobs.p, jobs_bytes, cudaMemcpyHostToDevice, m.st));
    m.P = P;
    m.smem_bytes = (int)((size_t)P.warp_words * 4 * P.cta_warps);
    m.have_jobs = true;
    return FTL_OK;
}

int Engine::fill(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops) {
    int rc = upload(g, jobs, ops, n_ops);
    if (rc) return rc;
    return replay();
}

int Engine::fill_layers(const Geometry &g, const std::vector<HostJob> &jobs, const ftl_path_op *ops, size_t n_ops) {
    int rc = upload(g, jobs, ops, n_ops, true);
    if (rc) return rc;
    return replay();
}

// One pass of the device pipeline over the resident job set.
//   exact = true : sizes are read back after each scan (two host round trips) and the scratch
//                  buffers are grown to fit; used for the first call and after an overflow.
//   exact = false: buffers keep the capacity of earlier calls, every kernel takes its counts
//                  from device memory and nothing waits for the host; the stages before the tile
//                  kernel are replayed from a CUDA graph.
static int run_pipeline(Engine::Impl &m, bool exact) {
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
    if ((rc = m.tcount.ensure((size_t)P.n_bins * sizeof(uint32_t), st))) return rc;
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
    auto front = [&](bool sync_sizes) -> int {
        CK(cudaMemsetAsync(d_cnt, 0, sizeof(Counters), st));
        uint32_t nv_hint = cap_v;
        if (P.n_ops > 0) {
            flatten_ops<false, false><<<fb, 128, 0, st>>>(d_ops, d_jobs, P, nullptr, (SumHead *)m.cnt.p, nullptr, nullptr, nullptr, nullptr); LAUNCHED();
            int r2 = run_scan<SumHeadOp>(st, (const SumHead *)m.cnt.p, P.n_ops, (SumHead *)m.off.p, m.partials);
            if (r2) return r2;
            if (sync_sizes) {
                CK(cudaMemcpyAsync(m.pin_small.p, &((SumHead *)m.off.p)[P.n_ops].sum, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                uint32_t nv = *(uint32_t *)m.pin_small.p;
                size_t want = nv ? nv : 1;
                if ((r2 = m.vtx.ensure(want * sizeof(Vtx), st))) return r2;
                if ((r2 = m.edges.ensure(want * sizeof(EdgeRec), st))) return r2;
                if ((r2 = m.sub_last.ensure(want * sizeof(uint32_t), st))) return r2;
                cap_v = (uint32_t)std::min<size_t>({m.vtx.cap / sizeof(Vtx), m.edges.cap / sizeof(EdgeRec), m.sub_last.cap / sizeof(uint32_t), (size_t)0x7FFFFFFF});
                nv_hint = nv;
            }
            set_vertex_count<<<1, 1, 0, st>>>(d_cnt, (const SumHead *)m.off.p, P.n_ops, cap_v); LAUNCHED();
            flatten_ops<false, true><<<fb, 128, 0, st>>>(d_ops, d_jobs, P, nullptr, nullptr, (const SumHead *)m.off.p, (Vtx *)m.vtx.p, nullptr, d_cnt); LAUNCHED();
        }
        init_job_state<<<div_up(P.n_jobs, 256), 256, 0, st>>>(d_js, d_jobs, P.n_ops > 0 ? (const SumHead *)m.off.p : nullptr, P.n_jobs); LAUNCHED();
        const uint32_t vb = std::max<uint32_t>(1u, std::min<uint32_t>(div_up(nv_hint, 256), (uint32_t)m.n_sms * 8));
        vtx_topkey<<<vb, 256, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js); LAUNCHED();
        vtx_topvid<<<vb, 256, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js, (uint32_t *)m.sub_last.p); LAUNCHED();
        job_finalize<<<div_up(P.n_jobs, 128), 128, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js, (const uint32_t *)m.sub_last.p, P.n_jobs); LAUNCHED();
        edge_build<<<vb, 256, 0, st>>>((const Vtx *)m.vtx.p, d_cnt, d_js, (EdgeRec *)m.edges.p); LAUNCHED();
        if (P.all_direct) return FTL_OK;  // every tile scans its job's own edges: nothing to bin
        CK(cudaMemsetAsync(m.tcount.p, 0, (size_t)P.n_bins * sizeof(uint32_t), st));
        bin_edges<false><<<vb, 256, 0, st>>>((const EdgeRec *)m.edges.p, d_cnt, d_js, P, (uint32_t *)m.tcount.p, nullptr, nullptr); LAUNCHED();
        int r3 = run_scan<AddU32>(st, (const uint32_t *)m.tcount.p, P.n_bins, (uint32_t *)m.toff.p, m.tpart);
        if (r3) return r3;
        if (sync_sizes) {
            CK(cudaMemcpyAsync(m.pin_small.p, &((uint32_t *)m.toff.p)[P.n_bins], sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            uint32_t n_entries = *(uint32_t *)m.pin_small.p;
            if ((r3 = m.entries.ensure((size_t)(n_entries ? n_entries : 1) * sizeof(uint32_t), st))) return r3;
            cap_e = (uint32_t)std::min<size_t>(m.entries.cap / sizeof(uint32_t), (size_t)0x7FFFFFFF);
        }
        set_entry_count<<<1, 1, 0, st>>>(d_cnt, (const uint32_t *)m.toff.p, P.n_bins, cap_e); LAUNCHED();
        CK(cudaMemsetAsync(m.tcount.p, 0, (size_t)P.n_bins * sizeof(uint32_t), st));
        bin_edges<true><<<vb, 256, 0, st>>>((const EdgeRec *)m.edges.p, d_cnt, d_js, P, (uint32_t *)m.tcount.p, (const uint32_t *)m.toff.p,
                                            (uint32_t *)m.entries.p); LAUNCHED();
        CK(cudaGetLastError());
        return FTL_OK;
    };

    if (exact) {
        m.drop_graph();
        if ((rc = front(true))) return rc;
    } else if (!m.use_graph) {
        if ((rc = front(false))) return rc;
    } else {
        // the sequence depends only on these values: replay the captured graph while they are unchanged
        std::vector<uint64_t> key = {(uint64_t)(uintptr_t)m.ops.p, (uint64_t)(uintptr_t)m.jobs.p, (uint64_t)(uintptr_t)m.cnt.p, (uint64_t)(uintptr_t)m.off.p,
                                     (uint64_t)(uintptr_t)m.partials.p, (uint64_t)(uintptr_t)m.vtx.p, (uint64_t)(uintptr_t)m.edges.p,
                                     (uint64_t)(uintptr_t)m.sub_last.p, (uint64_t)(uintptr_t)m.tcount.p, (uint64_t)(uintptr_t)m.toff.p,
                                     (uint64_t)(uintptr_t)m.tpart.p, (uint64_t)(uintptr_t)m.entries.p, (uint64_t)(uintptr_t)m.counters.p,
                                     (uint64_t)(uintptr_t)m.jstate.p, cap_v, cap_e, P.W, P.H, P.row_begin, P.row_end, P.fmt, P.log2R, P.n_jobs, P.n_ops,
                                     P.n_tiles, P.win_chunks, P.n_bins, P.all_direct};
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
    TileKernel tk = tile_kernel((int)P.fmt, aligned, !P.all_direct);
    const int tile_threads = (int)P.cta_warps * 32;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tk, tile_threads, m.smem_bytes));
    if (occ < 1) occ = 1;
    // A launch whose tiles all take the analytic rows into Matte8 rasters is a stream of stores with little
    // else to hide, and the wider the rows the fewer resident warps it wants: 1024/2048-px rows are fastest
    // at 5 CTAs per SM, 4096-px rows at 3 (13 % faster than 5 at every batch size tried), 8192-px rows at 2
    // (tools/occ_probe.py).  Everything that reads or scatters wants all the warps it can get.
    if (P.fmt == FTL_MATTE8 && P.all_direct && P.all_tiny && aligned && (P.W & 15u) == 0) occ = std::min(occ, P.W >= 6144u ? 2 : (P.W >= 3072u ? 3 : 5));
    if (const char *ev = getenv("FTL_OCC")) occ = std::max(1, std::min(16, atoi(ev)));  // tuning knob
    // Independent rasters: one launch over all tiles.  Layers of one raster: one launch per job, in
    // order, each over that job's tiles (the stages before ran once for all layers).
    const uint32_t n_launches = m.layered ? P.n_jobs : 1u;
    for (uint32_t l = 0; l < n_launches; l++) {
        Params PL = P;
        PL.tile_begin = m.layered ? l * P.n_bands : 0u;
        PL.tile_end = m.layered ? (l + 1) * P.n_bands : P.n_tiles;
        uint32_t grid = std::min<uint32_t>(div_up(PL.tile_end - PL.tile_begin, P.cta_warps), (uint32_t)(m.n_sms * occ));
        ProfSpan span{};
        const bool prof = g_profiling.load();
        if (prof) {
            CK(cudaEventCreate(&span.a));
            CK(cudaEventCreate(&span.b));
            CK(cudaEventRecord(span.a, st));
        }
        tk<<<grid, tile_threads, m.smem_bytes, st>>>((const EdgeRec *)m.edges.p, d_jobs, d_js, PL, (const uint32_t *)m.toff.p,
                                                     (const uint32_t *)m.entries.p, d_cnt); LAUNCHED();
        if (prof) {
            CK(cudaEventRecord(span.b, st));
            std::lock_guard<std::mutex> lock(g_spans_mu);
            g_spans.push_back(span);
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
static int resolve_pending(Engine::Impl &m) {
    if (m.pending == 0) return FTL_OK;
    CK(cudaStreamSynchronize(m.st));
    uint32_t redo = 0;
    const Counters *ring = (const Counters *)m.pin_ring.p;
    for (uint32_t i = 0; i < m.pending; i++) redo += ring[i].overflow ? 1u : 0u;
    m.pending = 0;
    for (uint32_t i = 0; i < redo; i++) {
        int rc = run_pipeline(m, true);
        if (rc) return rc;
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
    if (!m.have_jobs) return FTL_OK;
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
    flatten_ops<false, false><<<fb, 128, 0, st>>>((const ftl_path_op *)m.ops.p, (const JobDesc *)m.jobs.p, P, nullptr, (SumHead *)m.cnt.p, nullptr, nullptr, nullptr, nullptr); LAUNCHED();
    if ((rc = run_scan<SumHeadOp>(st, (const SumHead *)m.cnt.p, P.n_ops, (SumHead *)m.off.p, m.partials))) return rc;
    SumHead tot;
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(&tot, &((SumHead *)m.off.p)[n_ops], sizeof(tot), cudaMemcpyDeviceToHost));
    uint32_t nv = tot.sum;
    if (nv == 0) return FTL_OK;
    if ((rc = m.vtx.ensure((size_t)nv * sizeof(Vtx), st))) return rc;
    flatten_ops<false, true><<<fb, 128, 0, st>>>((const ftl_path_op *)m.ops.p, (const JobDesc *)m.jobs.p, P, nullptr, nullptr, (const SumHead *)m.off.p, (Vtx *)m.vtx.p, nullptr, nullptr); LAUNCHED();
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
    flatten_ops<true, false><<<fb, 128, 0, st>>>((const ftl_path_op *)m.ops.p, (const JobDesc *)m.jobs.p, P, (const float *)m.opw.p, (SumHead *)m.cnt.p, nullptr, nullptr, nullptr, nullptr); LAUNCHED();
    if ((rc = run_scan<SumHeadOp>(st, (const SumHead *)m.cnt.p, P.n_ops, (SumHead *)m.off.p, m.partials))) return rc;
    CK(cudaStreamSynchronize(st));
    std::vector<SumHead> off(n_ops + 1);
    CK(cudaMemcpy(off.data(), m.off.p, (n_ops + 1) * sizeof(SumHead), cudaMemcpyDeviceToHost));
    uint32_t np = off[n_ops].sum;
    for (size_t i = 0; i < n_ops; i++) out->counts[i] = off[i + 1].sum - off[i].sum;
    if (np == 0) return FTL_OK;
    if ((rc = m.wide.ensure((size_t)np * 3 * sizeof(float), st))) return rc;
    flatten_ops<true, true><<<fb, 128, 0, st>>>((const ftl_path_op *)m.ops.p, (const JobDesc *)m.jobs.p, P, (const float *)m.opw.p, nullptr, (const SumHead *)m.off.p, nullptr, (float *)m.wide.p, nullptr); LAUNCHED();
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
    fnv_rasters<<<count, 256, 0, m.st>>>((const uint8_t *)rasters, raster_bytes, (uint64_t *)m.misc.p); LAUNCHED();
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, m.misc.p, (size_t)count * 8, cudaMemcpyDeviceToHost, m.st));
    CK(cudaStreamSynchronize(m.st));
    return FTL_OK;
}

int Engine::sync() {
    ENSURE_INIT();
    int rc = resolve_pending(*impl_);
    if (rc) return rc;
    CK(cudaStreamSynchronize(impl_->st));
    return FTL_OK;
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
    CK(cudaMemsetAsync(dptr, value, bytes, impl_->st));
    return FTL_OK;
}
int Engine::copy_in(void *dptr, const void *src, size_t bytes) {
    ENSURE_INIT();
    {
        int rc = resolve_pending(*impl_);
        if (rc) return rc;
    }
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
// moved in 256 MiB pieces, and expanded by host threads while the next piece is in flight.
int Engine::copy_out(void *dst, const void *dptr, size_t bytes) {
    ENSURE_INIT();
    Impl &m = *impl_;
    {
        int rc = resolve_pending(m);
        if (rc) return rc;
    }
    cudaStream_t st = m.st;
    static const bool raw_only = getenv("FTL_RAW_READ") && atoi(getenv("FTL_RAW_READ")) != 0;
    const size_t PACK_MIN = 4u << 20;
    if (raw_only || bytes < PACK_MIN || (bytes & 1023u) != 0 || ((uintptr_t)dptr & 15u) != 0) {
        CK(cudaMemcpyAsync(dst, dptr, bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return FTL_OK;
    }
    unsigned nt = std::min(8u, std::thread::hardware_concurrency());  // the expansion is memory-bound well before 8 threads
    if (const char *ev = getenv("FTL_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(ev));
    nt = std::max(1u, std::min(nt, 64u));
    const size_t PIECE = 256u << 20;
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
                CK(cudaStreamSynchronize(st));
            }
            return FTL_OK;
        };
        int r = stage();  // overlaps the expansion of the previous piece
        if (worker.joinable()) worker.join();
        if (r == -1) {  // plain copy of this piece
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
