// front_kernels.cuh — stages (a)+(b): counters, flatten, edge preparation, binning
// Included by engine.cu inside namespace ftl (one translation unit: the kernels share Params / EdgeRec / ...).
#pragma once

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

// the same guards as functors for scan_small (the scan's grand total is the count)
struct FinVertexCount {
    Counters *C;
    uint32_t cap_v;
    __device__ void operator()(const SumHead &total) const {
        uint32_t nv = total.sum;
        if (nv > cap_v) {
            C->overflow = 1;
            C->need_v = nv;
            nv = 0;
        }
        C->nv = nv;
    }
};
struct FinEntryCount {
    Counters *C;
    uint32_t cap_e;
    __device__ void operator()(uint32_t total) const {
        C->n_entries = total;
        if (total > cap_e) {
            C->overflow = 1;
            C->need_e = total;
        }
    }
};

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
    uint32_t *wop = nullptr;  // wide mode, device stroker: the op each point came from
    uint32_t opi = 0;
    uint32_t sub = 0, job = 0;
    int2 *slab = nullptr;   // counting pass of a curve: the first slab_cap kept points are parked here, so that the
    uint32_t slab_cap = 0;  // emitting pass copies them instead of subdividing the curve a second time
    size_t slab_stride = 0; // point k of op i lives at slabs[k * n_ops + i]: neighbouring lanes (ops) share sectors
    __device__ __forceinline__ void put(WPt q) {
        if (WIDE) {
            if (EMIT) {
                wout[3 * (size_t)n] = q.p.x;
                wout[3 * (size_t)n + 1] = q.p.y;
                wout[3 * (size_t)n + 2] = q.w;
                if (wop) wop[n] = opi;
            }
            n++;
        } else {
            int32_t fx = fx_from_f32(q.p.x), fy = fx_from_f32(q.p.y);
            if (force || fx != px || fy != py) {
                if (EMIT) vout[n] = {fx, fy, sub, job};
                else if (n < slab_cap) slab[(size_t)n * slab_stride] = make_int2(fx, fy);
                n++;
            }
            force = false;
            px = fx;
            py = fy;
        }
    }
};

// The pending right halves of the subdivision (depth <= 16).  Fill mode keeps them in SHARED memory, one column per
// thread ([level][word][thread]: conflict-free); a dynamically indexed local array goes through local memory, and its
// pushes and pops were most of this kernel's time (92 -> 30 us on the 262 k curves of a config-4 step).  Wide mode (only
// used for very large strokes) keeps the local array: its entries carry a width as well.
constexpr uint32_t FLAT_THREADS = 128;
constexpr uint32_t FLAT_STACK_WORDS = 6;  // three control points (cubic) per level
constexpr size_t FLAT_SMEM_BYTES = (size_t)MAX_DEPTH * FLAT_STACK_WORDS * FLAT_THREADS * sizeof(float) + (size_t)MAX_DEPTH * FLAT_THREADS;
template <bool WIDE>
struct FlatStack;
template <>
struct FlatStack<false> {
    float *pts;      // [MAX_DEPTH][FLAT_STACK_WORDS][FLAT_THREADS]
    uint8_t *depth;  // [MAX_DEPTH][FLAT_THREADS]
    __device__ __forceinline__ FlatStack(float *smem) : pts(smem), depth(reinterpret_cast<uint8_t *>(smem + MAX_DEPTH * FLAT_STACK_WORDS * FLAT_THREADS)) {}
    __device__ __forceinline__ void push(int sp, WPt b, WPt c, WPt d, int dep) {
        float *q = pts + (size_t)sp * FLAT_STACK_WORDS * FLAT_THREADS + threadIdx.x;
        q[0] = b.p.x; q[FLAT_THREADS] = b.p.y; q[2 * FLAT_THREADS] = c.p.x; q[3 * FLAT_THREADS] = c.p.y; q[4 * FLAT_THREADS] = d.p.x; q[5 * FLAT_THREADS] = d.p.y;
        depth[sp * FLAT_THREADS + threadIdx.x] = (uint8_t)dep;
    }
    __device__ __forceinline__ void pop(int sp, WPt &b, WPt &c, WPt &d, int &dep) const {
        const float *q = pts + (size_t)sp * FLAT_STACK_WORDS * FLAT_THREADS + threadIdx.x;
        b = {{q[0], q[FLAT_THREADS]}, 0.0f};
        c = {{q[2 * FLAT_THREADS], q[3 * FLAT_THREADS]}, 0.0f};
        d = {{q[4 * FLAT_THREADS], q[5 * FLAT_THREADS]}, 0.0f};
        dep = depth[sp * FLAT_THREADS + threadIdx.x];
    }
};
template <>
struct FlatStack<true> {
    WPt sb[MAX_DEPTH], sc[MAX_DEPTH], sd[MAX_DEPTH];
    uint8_t sdep[MAX_DEPTH];
    __device__ __forceinline__ FlatStack(float *) {}
    __device__ __forceinline__ void push(int sp, WPt b, WPt c, WPt d, int dep) { sb[sp] = b; sc[sp] = c; sd[sp] = d; sdep[sp] = (uint8_t)dep; }
    __device__ __forceinline__ void pop(int sp, WPt &b, WPt &c, WPt &d, int &dep) const { b = sb[sp]; c = sc[sp]; d = sd[sp]; dep = sdep[sp]; }
};

template <bool WIDE, bool EMIT>
__device__ void flatten_quad(WPt a, WPt b, WPt c, float tol_sq, OpSink<WIDE, EMIT> &sink, FlatStack<WIDE> &stk) {  // plotter.rs:248-265
    int sp = 0, depth = 0;
    for (;;) {
        WPt ab = wmid<WIDE>(a, b), bc = wmid<WIDE>(b, c), ab_bc = wmid<WIDE>(ab, bc), ac = wmid<WIDE>(a, c);
        if (pointy::distance_sq(ab_bc.p, ac.p) <= tol_sq || depth >= MAX_DEPTH) {
            sink.put(c);
            if (sp == 0) break;
            sp--;
            a = c;
            WPt unused;
            stk.pop(sp, b, c, unused, depth);
        } else {
            stk.push(sp, bc, c, c, depth + 1);
            sp++;
            b = ab; c = ab_bc; depth++;
        }
    }
}

template <bool WIDE, bool EMIT>
__device__ void flatten_cubic(WPt a, WPt b, WPt c, WPt d, float tol_sq, OpSink<WIDE, EMIT> &sink, FlatStack<WIDE> &stk) {  // plotter.rs:311-332
    int sp = 0, depth = 0;
    for (;;) {
        WPt ab = wmid<WIDE>(a, b), bc = wmid<WIDE>(b, c), cd = wmid<WIDE>(c, d);
        WPt ab_bc = wmid<WIDE>(ab, bc), bc_cd = wmid<WIDE>(bc, cd);
        WPt pe = wmid<WIDE>(ab_bc, bc_cd), ad = wmid<WIDE>(a, d);
        if (pointy::distance_sq(pe.p, ad.p) <= tol_sq || depth >= MAX_DEPTH) {
            sink.put(d);
            if (sp == 0) break;
            sp--;
            a = d;
            stk.pop(sp, b, c, d, depth);
        } else {
            stk.push(sp, bc_cd, cd, d, depth + 1);
            sp++;
            b = ab; c = ab_bc; d = pe; depth++;
        }
    }
}

// ---------------------------------------------------------------------------
// row-band culling (one raster split over several GPUs: SURVEY 8e-ii)
// ---------------------------------------------------------------------------
// A handle that owns rows [row_begin, row_end) of a raster needs only the sub-figures whose edges can
// reach those rows - plus every sub-figure that could hold the figure's top-left vertex, because the
// winding direction and the top row are properties of the WHOLE figure (fig.rs:402-411,493-497).
// Flattened points are midpoints of control points, so they stay inside the y range of the op's
// control points (the f32 midpoint and the f32 -> Fixed conversion are both monotone): the extents
// below are exact bounds, and dropping a sub-figure outside them changes no pixel and no (dir, top_row).
struct CullBufs {
    const uint32_t *head;  // per op: 1 + index of the op that starts its sub-figure (inclusive max-scan of the start marks)
    const int32_t *sub_lo, *sub_hi;  // per start op: Fixed y range of the sub-figure's control points
    const int32_t *job_bound;        // per job: [2j] min Fixed y over op end points (real vertices), [2j+1] min over all control points
};
struct MaxU32 {
    typedef uint32_t T;
    static __device__ __forceinline__ T identity() { return 0u; }
    static __device__ __forceinline__ T combine(T a, T b) { return a > b ? a : b; }
    static __device__ __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
};
__global__ void cull_init_jobs(int32_t *job_bound, uint32_t n_jobs) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 2 * n_jobs) job_bound[j] = INT32_MAX;
}
// y range (Fixed) of the control points of drawing op i
__device__ __forceinline__ void op_y_range(const ftl_path_op &op, const float *e, const PenInfo &pi, int32_t *lo, int32_t *hi, int32_t *y_end) {
    const int n_pts = op.tag == FTL_OP_QUAD ? 2 : (op.tag == FTL_OP_CUBIC ? 3 : 1);
    float ymin = 0.0f, ymax = 0.0f, ye = 0.0f;
    for (int k = 0; k < n_pts; k++) {
        const float y = pointy::transform(e, {op.v[2 * k], op.v[2 * k + 1]}).y;
        ymin = k == 0 ? y : fminf(ymin, y);
        ymax = k == 0 ? y : fmaxf(ymax, y);
        ye = y;
    }
    if (n_pts > 1) {  // a curve starts at the pen
        const float y = pointy::transform(e, pi.pen).y;
        ymin = fminf(ymin, y);
        ymax = fmaxf(ymax, y);
    }
    *lo = fx_from_f32(ymin);
    *hi = fx_from_f32(ymax);
    *y_end = fx_from_f32(ye);
}
__global__ void __launch_bounds__(256) cull_op_extents(const ftl_path_op *__restrict__ ops, const JobDesc *__restrict__ jobs, Params P,
                                                       uint32_t *__restrict__ headmark, int32_t *__restrict__ sub_lo, int32_t *__restrict__ sub_hi,
                                                       int32_t *__restrict__ job_bound) {
    const uint32_t lane = threadIdx.x & 31;
    // The two minima of a job are kept in registers across the thread's whole grid-stride loop and flushed when the job
    // changes and once at the end, there with one pair of atomics per warp: one raster with 10 M ops otherwise sends
    // 650 k same-address atomics (one pair per warp and iteration) through one L2 slice, and the loads queue up behind them.
    uint32_t cur_j = NONE32;
    int32_t cur_ye = INT32_MAX, cur_lo = INT32_MAX;
    for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < P.n_ops; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + lane;
        if (i >= P.n_ops) continue;
        const ftl_path_op op = ops[i];
        uint32_t mark = 0;
        if (op.tag >= FTL_OP_MOVE && op.tag <= FTL_OP_CUBIC) {
            const uint32_t j = job_of_op(jobs, P.n_jobs, i);
            const JobDesc &jd = jobs[j];
            const PenInfo pi = find_pen(ops, jd.op_begin, i);
            if (pi.starts_sub || op.tag == FTL_OP_MOVE) {
                mark = i + 1;
                sub_lo[i] = INT32_MAX;
                sub_hi[i] = INT32_MIN;
            }
            float e[6];
#pragma unroll
            for (int k = 0; k < 6; k++) e[k] = jd.e[k];
            int32_t lo, hi, ye;
            op_y_range(op, e, pi, &lo, &hi, &ye);
            if (j != cur_j) {
                if (cur_j != NONE32) {
                    atomicMin(&job_bound[2 * cur_j], cur_ye);
                    atomicMin(&job_bound[2 * cur_j + 1], cur_lo);
                }
                cur_j = j;
                cur_ye = INT32_MAX;
                cur_lo = INT32_MAX;
            }
            cur_ye = min(cur_ye, ye);
            cur_lo = min(cur_lo, lo);
        }
        headmark[i] = mark;
    }
    // every lane of the warp arrives here (the loop bound is warp-uniform)
    const uint32_t have = __ballot_sync(0xFFFFFFFFu, cur_j != NONE32);
    if (!have) return;
    const uint32_t j0 = __shfl_sync(0xFFFFFFFFu, cur_j, __ffs((int)have) - 1);
    if (__all_sync(0xFFFFFFFFu, cur_j == NONE32 || cur_j == j0)) {
        const int32_t ye_min = __reduce_min_sync(0xFFFFFFFFu, cur_ye), lo_min = __reduce_min_sync(0xFFFFFFFFu, cur_lo);
        if (lane == 0) {
            atomicMin(&job_bound[2 * j0], ye_min);
            atomicMin(&job_bound[2 * j0 + 1], lo_min);
        }
    } else if (cur_j != NONE32) {
        atomicMin(&job_bound[2 * cur_j], cur_ye);
        atomicMin(&job_bound[2 * cur_j + 1], cur_lo);
    }
}
__global__ void __launch_bounds__(256) cull_sub_extents(const ftl_path_op *__restrict__ ops, const JobDesc *__restrict__ jobs, Params P,
                                                        const uint32_t *__restrict__ head, int32_t *__restrict__ sub_lo, int32_t *__restrict__ sub_hi) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_ops; i += gridDim.x * blockDim.x) {
        const ftl_path_op op = ops[i];
        if (op.tag < FTL_OP_MOVE || op.tag > FTL_OP_CUBIC) continue;
        const uint32_t j = job_of_op(jobs, P.n_jobs, i);
        const JobDesc &jd = jobs[j];
        const PenInfo pi = find_pen(ops, jd.op_begin, i);
        float e[6];
#pragma unroll
        for (int k = 0; k < 6; k++) e[k] = jd.e[k];
        int32_t lo, hi, ye;
        op_y_range(op, e, pi, &lo, &hi, &ye);
        const uint32_t h = head[i + 1] - 1u;  // exclusive scan: entry i + 1 covers ops 0 .. i
        atomicMin(&sub_lo[h], lo);
        atomicMax(&sub_hi[h], hi);
    }
}
// Does drawing op i belong to a sub-figure this handle must flatten?
__device__ __forceinline__ bool cull_keep(const CullBufs &cb, const Params &P, uint32_t i, uint32_t j) {
    const uint32_t h = cb.head[i + 1] - 1u;
    const int32_t lo = cb.sub_lo[h], hi = cb.sub_hi[h];
    const int32_t y_vertex = cb.job_bound[2 * j], y_hull = cb.job_bound[2 * j + 1];
    if (lo <= y_vertex) return true;  // may hold the top-left vertex of the figure
    // geometry row r lands on raster row r - shift, shift = min(top_row, 0); top_row is between these two
    const int32_t s_lo = min(fx_to_i32(y_hull), 0), s_hi = min(fx_to_i32(y_vertex), 0);
    return fx_to_i32(hi) - s_lo >= (int32_t)P.row_begin && fx_to_i32(lo) - s_hi <= (int32_t)P.row_end - 1;
}

// One thread per PathOp.  Pass 1 (EMIT=false) counts the vertices the op
// contributes; after the scan, pass 2 (EMIT=true) repeats the identical
// subdivision and writes them at the scanned offset, so the output order is
// the reference's depth-first order.
template <bool WIDE, bool EMIT>
__global__ void __launch_bounds__(FLAT_THREADS) flatten_ops(const ftl_path_op *__restrict__ ops, const JobDesc *__restrict__ jobs, Params P,
                                                   const float *__restrict__ opw, SumHead *__restrict__ cnt,
                                                   const SumHead *__restrict__ off, Vtx *__restrict__ vout,
                                                   float *__restrict__ wout, const uint32_t *__restrict__ stop, CullBufs cull,
                                                   int2 *__restrict__ slabs = nullptr, uint32_t slab_pts = 0, uint32_t *__restrict__ wop = nullptr) {
    if (EMIT && stop && *stop) return;  // a capacity guard tripped (Counters::overflow / StrokeCounters::overflow): draw nothing
    extern __shared__ __align__(16) float flat_smem[];
    FlatStack<WIDE> stk(flat_smem);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_ops; i += gridDim.x * blockDim.x) {
        const ftl_path_op op = ops[i];
        OpSink<WIDE, EMIT> sink;
        uint32_t j = job_of_op(jobs, P.n_jobs, i);
        const JobDesc &jd = jobs[j];
        bool starts = false;
        if (op.tag >= FTL_OP_MOVE && op.tag <= FTL_OP_CUBIC && (WIDE || !cull.head || cull_keep(cull, P, i, j))) {
            PenInfo pi = find_pen(ops, jd.op_begin, i);
            starts = pi.starts_sub || op.tag == FTL_OP_MOVE;  // Move closes the current sub-figure (plotter.rs:210)
            float e[6];
#pragma unroll
            for (int k = 0; k < 6; k++) e[k] = jd.e[k];
            sink.force = starts;
            if (EMIT) {
                SumHead o = off[i];
                if (WIDE) {
                    sink.wout = wout + 3 * (size_t)o.sum;
                    if (wop) {
                        sink.wop = wop + o.sum;
                        sink.opi = i;
                    }
                } else {
                    sink.vout = vout + o.sum;
                    sink.sub = starts ? o.sum : o.head;
                    sink.job = j;
                }
            }
            const bool curve = op.tag == FTL_OP_QUAD || op.tag == FTL_OP_CUBIC;
            if (!WIDE && curve && slab_pts) {
                int2 *slab = slabs + i;
                if (!EMIT) {
                    sink.slab = slab;
                    sink.slab_cap = slab_pts;
                    sink.slab_stride = P.n_ops;
                } else {
                    const uint32_t cnt_i = off[i + 1].sum - off[i].sum;
                    if (cnt_i <= slab_pts) {  // the counting pass parked every point of this curve: copy, do not subdivide again
                        for (uint32_t k = 0; k < cnt_i; k++) {
                            const int2 q = slab[(size_t)k * P.n_ops];
                            sink.vout[k] = {q.x, q.y, sink.sub, sink.job};
                        }
                        continue;
                    }
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
                flatten_quad<WIDE, EMIT>(a, b, c, jd.tol_sq, sink, stk);
            } else {  // plotter.rs:286-305; float_lerp(a,b,t) = b + (a-b)*t (geom.rs:14-16)
                float w0 = WIDE ? w_now + (w_pen - w_now) * (1.0f / 3.0f) : 0.0f;
                float w1 = WIDE ? w_now + (w_pen - w_now) * (2.0f / 3.0f) : 0.0f;
                WPt b = {pointy::transform(e, {op.v[0], op.v[1]}), w0};
                WPt c = {pointy::transform(e, {op.v[2], op.v[3]}), w1};
                WPt d = {pointy::transform(e, {op.v[4], op.v[5]}), w_now};
                flatten_cubic<WIDE, EMIT>(a, b, c, d, jd.tol_sq, sink, stk);
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

__global__ void init_job_state(JobState *JS, const JobDesc *__restrict__ jobs, const SumHead *__restrict__ off, uint32_t n_jobs, Counters *C, uint32_t direct_max) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool big = j < n_jobs && off && off[jobs[j].op_end].sum - off[jobs[j].op_begin].sum > direct_max;
    const uint32_t n_big = __popc(__ballot_sync(0xFFFFFFFFu, big));
    if ((threadIdx.x & 31u) == 0 && n_big) atomicAdd(&C->n_big, n_big);
    if (j >= n_jobs) return;
    JobState s;
    s.top_key = ~0ull; s.top_vid = NONE32; s.dir = 0; s.top_row = 0; s.first_row = 0x7FFFFFFF; s.shift = 0;
    s.vtx_begin = off ? off[jobs[j].op_begin].sum : 0u;
    s.vtx_end = off ? off[jobs[j].op_end].sum : 0u;
    s.pad[0] = s.pad[1] = s.pad[2] = 0;
    JS[j] = s;
}

// Top-left vertex, pass 1: minimum (y,x) over the live vertices of each job (fig.rs:493-494).
// One atomic for the warp's minima when its lanes hold one job, else one per lane.
__device__ __forceinline__ void topkey_flush(JobState *JS, uint32_t cur_j, unsigned long long cur_key) {
    const uint32_t have = __ballot_sync(0xFFFFFFFFu, cur_j != NONE32);
    if (!have) return;
    const uint32_t j0 = __shfl_sync(0xFFFFFFFFu, cur_j, __ffs((int)have) - 1);
    if (__all_sync(0xFFFFFFFFu, cur_j == NONE32 || cur_j == j0)) {
        unsigned long long key = cur_key;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, key, d);
            key = o < key ? o : key;
        }
        if ((threadIdx.x & 31u) == 0 && key != ~0ull) atomicMin(&JS[j0].top_key, key);
    } else if (cur_j != NONE32 && cur_key != ~0ull) atomicMin(&JS[cur_j].top_key, cur_key);
}
__global__ void __launch_bounds__(256) vtx_topkey(const Vtx *__restrict__ V, const Counters *__restrict__ C, JobState *JS) {
    const uint32_t nv = C->nv;
    const uint32_t stride = gridDim.x * blockDim.x;
    // Whole warps iterate together (k0 is the warp's first vertex; vertices are stored job after job).  The minimum is kept
    // in registers across the grid-stride loop and flushed - one atomic per warp - only when a lane moves on to another
    // job, and at the end: a batch of small jobs flushes every iteration (as many atomics as warps x iterations), one
    // raster with 10 M vertices flushes once per warp instead of sending 320 k atomics to one address with the loads
    // queueing up behind them (241 -> 33 us).
    uint32_t cur_j = NONE32;
    unsigned long long cur_key = ~0ull;
    for (uint32_t k0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; k0 < nv; k0 += stride) {
        const uint32_t k = k0 + (threadIdx.x & 31u);
        unsigned long long key = ~0ull;
        uint32_t job = NONE32;
        if (k < nv) {
            const Vtx v = V[k];
            job = v.job;
            if (!(vtx_is_last(V, nv, k) && vtx_same(v, V[v.sub]))) key = vtx_key(v);  // else: popped by the close rule
        }
        if (__any_sync(0xFFFFFFFFu, job != NONE32 && cur_j != NONE32 && job != cur_j)) {
            topkey_flush(JS, cur_j, cur_key);
            cur_j = NONE32;
            cur_key = ~0ull;
        }
        if (job != NONE32) {
            cur_j = job;
            cur_key = key < cur_key ? key : cur_key;
        }
    }
    topkey_flush(JS, cur_j, cur_key);
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

// Band range (bands of 32 rows, raster_bins) of an edge inside this device's rows; returns false if none.
__device__ __forceinline__ bool edge_bands(const EdgeRec &e, const Params &P, uint32_t *b0, uint32_t *b1) {
    int32_t lo = e.ry0, hi = e.ry1;  // ry0 >= first_row >= 0 by construction
    if (lo < (int32_t)P.row_begin) lo = (int32_t)P.row_begin;
    if (hi > (int32_t)P.row_end - 1) hi = (int32_t)P.row_end - 1;
    if (lo > hi) return false;
    *b0 = (uint32_t)(lo - (int32_t)P.row_begin) >> 5;
    *b1 = (uint32_t)(hi - (int32_t)P.row_begin) >> 5;
    return true;
}

// Conservative range of column windows an edge can write to on the rows [ra, rb] of one band (both
// inside the edge's own rows).  The span of a row is linear in the row, so the extremes are at the
// two end rows, evaluated as the scatter does; the scatter loop can run at most |dx/dy| + 2 cells
// past the leftmost one.  Anything that would wrap 32-bit arithmetic falls back to "all windows".
__device__ __forceinline__ void edge_windows(const EdgeRec &e, int32_t ra, int32_t rb, const Params &P, uint32_t *w0, uint32_t *w1) {
    *w0 = 0;
    *w1 = P.b_nwin - 1;
    if (P.b_nwin == 1) return;
    // X at the bottom / top of the two end rows; in 64 bits only to detect what would wrap the scatter's 32-bit arithmetic
    const int64_t slope = e.inv_slope;
    const int64_t xa_bot = (int64_t)e.x_bot0 + (int64_t)(ra - e.ry0) * slope, xb_bot = (int64_t)e.x_bot0 + (int64_t)(rb - e.ry0) * slope;
    const int64_t xa_top = xa_bot - slope, xb_top = xb_bot - slope;
    if (xa_bot != (int32_t)xa_bot || xa_top != (int32_t)xa_top || xb_bot != (int32_t)xb_bot || xb_top != (int32_t)xb_top) return;
    // X is monotone in the row: the extremes are the top of the first row and the bottom of the last one
    const int32_t lo = slope >= 0 ? (int32_t)xa_top : (int32_t)xb_bot, hi = slope >= 0 ? (int32_t)xb_bot : (int32_t)xa_top;
    const int32_t run = (int32_t)((slope < 0 ? -slope : slope) >> 16);
    const int32_t wmax = (int32_t)P.W - 1;
    const int32_t lo_pix = min(max((lo >> 16) - 1, 0), wmax), hi_pix = min(max((hi >> 16) + run + 3, 0), wmax);
    const int sh = 31 - __clz((int)P.b_wc);  // the window width is a power of two
    *w0 = (uint32_t)lo_pix >> sh;
    *w1 = (uint32_t)hi_pix >> sh;
}

// Counting sort of edges by (job, band of 32 rows, column window): pass FILL=false counts, pass FILL=true
// writes edge ids at the scanned offsets.  Short edges are handled by their own thread; an edge
// crossing many bands is spread over the warp.  Jobs with at most Params::direct_max edge slots are not
// binned at all.
template <bool FILL>
__device__ __forceinline__ void bin_one(const EdgeRec &e, uint32_t k, uint32_t tile, uint32_t band, const Params &P, uint32_t *bin_count,
                                        const uint32_t *bin_off, uint32_t *entries) {
    const int32_t row0 = (int32_t)P.row_begin + (int32_t)(band << 5);
    const int32_t ra = max(e.ry0, row0), rb = min(min(e.ry1, row0 + 31), (int32_t)P.row_end - 1);
    uint32_t w0, w1;
    edge_windows(e, ra, rb, P, &w0, &w1);
    for (uint32_t w = w0; w <= w1; w++) {
        const uint32_t bin = tile * P.b_nwin + w;
        const uint32_t slot = atomicAdd(&bin_count[bin], 1u);
        if (FILL) entries[bin_off[bin] + slot] = k;
    }
}
// The bins of the warp's 32 edges (lane = edge k0 + lane; e.flags == 0: nothing).  Short edges are handled by their own
// lane; an edge crossing many bands is spread over the warp (its record travels by shuffle).
template <bool FILL>
__device__ __forceinline__ void bin_warp(EdgeRec e, uint32_t k0, const JobState *__restrict__ JS, const Params &P, uint32_t *bin_count,
                                         const uint32_t *bin_off, uint32_t *entries) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t b0 = 0, nb = 0, tbase = 0, b1;
    if ((e.flags & 1u) && edge_bands(e, P, &b0, &b1)) {
        const JobState &js = JS[e.job];
        if (js.vtx_end - js.vtx_begin > P.direct_max) {
            nb = b1 - b0 + 1;
            tbase = e.job * P.b_nbands;
        }
    }
    if (nb > 0 && nb <= 4)
        for (uint32_t b = b0; b < b0 + nb; b++) bin_one<FILL>(e, k0 + lane, tbase + b, b, P, bin_count, bin_off, entries);
    uint32_t tall = __ballot_sync(0xFFFFFFFFu, nb > 4);
    // FILL: the slot of an entry comes back from an atomic, and the store that needs it would stall the warp for an L2 round
    // trip per edge.  The stores of one tall edge are therefore issued after the atomics of the NEXT one (up to three windows
    // of the lane's band are kept pending; anything beyond that is stored at once).
    uint32_t p_off0 = 0, p_off1 = 0, p_off2 = 0, p_slot0 = 0, p_slot1 = 0, p_slot2 = 0, p_n = 0, p_k = 0;
    while (tall) {
        const int src = __ffs(tall) - 1;
        tall &= tall - 1;
        const uint32_t sb0 = __shfl_sync(0xFFFFFFFFu, b0, src), snb = __shfl_sync(0xFFFFFFFFu, nb, src), stb = __shfl_sync(0xFFFFFFFFu, tbase, src);
        EdgeRec es;
        es.x_bot0 = __shfl_sync(0xFFFFFFFFu, e.x_bot0, src);
        es.inv_slope = __shfl_sync(0xFFFFFFFFu, e.inv_slope, src);
        es.step_pix = 0;
        es.ry0 = __shfl_sync(0xFFFFFFFFu, e.ry0, src);
        es.ry1 = __shfl_sync(0xFFFFFFFFu, e.ry1, src);
        es.fr = 0; es.job = 0; es.flags = 1u;
        if (!FILL) {
            for (uint32_t b = lane; b < snb; b += 32) bin_one<false>(es, k0 + (uint32_t)src, stb + sb0 + b, sb0 + b, P, bin_count, bin_off, entries);
            continue;
        }
        uint32_t c_off0 = 0, c_off1 = 0, c_off2 = 0, c_slot0 = 0, c_slot1 = 0, c_slot2 = 0, c_n = 0;
        if (lane < snb) {  // the lane's first band of this edge: up to three windows pending
            const uint32_t band = sb0 + lane, tile = stb + band;
            const int32_t row0 = (int32_t)P.row_begin + (int32_t)(band << 5);
            const int32_t ra = max(es.ry0, row0), rb = min(min(es.ry1, row0 + 31), (int32_t)P.row_end - 1);
            uint32_t w0, w1;
            edge_windows(es, ra, rb, P, &w0, &w1);
            const uint32_t bin = tile * P.b_nwin + w0;
            c_n = min(w1 - w0 + 1u, 3u);
            c_off0 = bin_off[bin];
            c_slot0 = atomicAdd(&bin_count[bin], 1u);
            if (c_n > 1) {
                c_off1 = bin_off[bin + 1];
                c_slot1 = atomicAdd(&bin_count[bin + 1], 1u);
            }
            if (c_n > 2) {
                c_off2 = bin_off[bin + 2];
                c_slot2 = atomicAdd(&bin_count[bin + 2], 1u);
            }
            for (uint32_t w = w0 + 3; w <= w1; w++) {
                const uint32_t bw = tile * P.b_nwin + w;
                entries[bin_off[bw] + atomicAdd(&bin_count[bw], 1u)] = k0 + (uint32_t)src;
            }
        }
        for (uint32_t b = lane + 32; b < snb; b += 32) bin_one<true>(es, k0 + (uint32_t)src, stb + sb0 + b, sb0 + b, P, bin_count, bin_off, entries);  // taller than 1024 rows
        // the previous edge's stores: its atomics have had this iteration to complete
        if (p_n > 0) entries[p_off0 + p_slot0] = p_k;
        if (p_n > 1) entries[p_off1 + p_slot1] = p_k;
        if (p_n > 2) entries[p_off2 + p_slot2] = p_k;
        p_off0 = c_off0; p_off1 = c_off1; p_off2 = c_off2;
        p_slot0 = c_slot0; p_slot1 = c_slot1; p_slot2 = c_slot2;
        p_n = c_n;
        p_k = k0 + (uint32_t)src;
    }
    if (FILL) {
        if (p_n > 0) entries[p_off0 + p_slot0] = p_k;
        if (p_n > 1) entries[p_off1 + p_slot1] = p_k;
        if (p_n > 2) entries[p_off2 + p_slot2] = p_k;
    }
}

// Counting sort of edges by (job, band of 32 rows, column window): the COUNT pass rides in edge_build (the edge is in
// registers there), this FILL pass writes edge ids at the scanned offsets.  Jobs with at most Params::direct_max edge slots are
// not binned at all.
__global__ void __launch_bounds__(256) bin_fill(const EdgeRec *__restrict__ E, const Counters *__restrict__ C, const JobState *__restrict__ JS, Params P,
                                                uint32_t *__restrict__ bin_count, const uint32_t *__restrict__ bin_off, uint32_t *__restrict__ entries) {
    if (C->overflow) return;
    const uint32_t nv = C->nv;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t span = gridDim.x * blockDim.x;
    for (uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x - lane; k0 < nv; k0 += span) {
        EdgeRec e;
        e.flags = 0;
        if (k0 + lane < nv) e = E[k0 + lane];
        bin_warp<true>(e, k0, JS, P, bin_count, bin_off, entries);
    }
}

// One thread per vertex k: the ring segment (k, next_fwd(k)) becomes at most one edge, directed from its upper to its
// lower vertex.  This is the same set of edges the reference creates in update_edges/add_edge (fig.rs:576-600) when it
// visits both neighbours of every vertex.  COUNT: the edge's bins are counted here (first pass of the counting sort).
template <bool COUNT>
__global__ void __launch_bounds__(256) edge_build(const Vtx *__restrict__ V, const Counters *__restrict__ C, const JobState *__restrict__ JS,
                                                  EdgeRec *__restrict__ E, Params P, uint32_t *__restrict__ bin_count) {
    const uint32_t nv = C->nv;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t span = gridDim.x * blockDim.x;
    for (uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x - lane; k0 < nv; k0 += span) {
        const uint32_t k = k0 + lane;
        EdgeRec e;
        e.flags = 0;
        if (k < nv) {
            Vtx v = V[k];
            bool last = vtx_is_last(V, nv, k);
            bool pop = last && vtx_same(v, V[v.sub]);
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
        if (COUNT) bin_warp<false>(e, k0, JS, P, bin_count, nullptr, nullptr);
    }
}
