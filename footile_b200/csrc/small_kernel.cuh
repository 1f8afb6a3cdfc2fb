// small_kernel.cuh — one small fill in ONE launch: stages (a) to (d) inside a CTA.
// Included by engine.cu inside namespace ftl (one translation unit: the kernels share Params / EdgeRec / ...).
#pragma once

// ---------------------------------------------------------------------------
// The reference's own benchmark (benches/fishyb.rs:10-39) and its examples draw ONE small path into a
// small raster per call: a few path ops, tens of edges, 16x16 to 256x256 pixels.  The general pipeline
// spends such a call on ~15 dependent kernels that each wait on two or three global loads.  Here the
// whole fill is one launch and nothing but the pixels touches HBM:
//   - the path ops travel as kernel arguments (no staging copy, no H2D memcpy);
//   - curves are subdivided LEVEL-PARALLEL by one warp each: the lanes hold the nodes of one level of the
//     De Casteljau recursion (plotter.rs:248-332), a flat node emits its end point tagged with its position
//     in the recursion tree, the others split into two lanes of the next level; sorting the leaves by tag
//     restores the reference's depth-first point order, with exactly the reference's f32 arithmetic per node;
//   - point intake (Fixed conversion, de-duplication: fig.rs:428-461), the closing rule, the top-left vertex,
//     winding direction and Edge::new (fig.rs:179-210,373-411,493-497) run on shared-memory arrays with the
//     same device functions the general pipeline uses;
//   - every warp then owns bands of 32 rows and scatters / resolves them with the binned kernel's
//     lanes-are-rows code (bin_kernel.cuh).
// A CTA draws 256 rows; larger rasters take several CTAs which each repeat the (cheap) geometry stages.
// Anything that does not fit (too many points, a level wider than a warp) sets an overflow flag in mapped
// host memory and draws nothing: the host repeats that fill through the general pipeline, and a device-side
// poison flag keeps later small fills from overtaking it.
// ---------------------------------------------------------------------------
constexpr uint32_t SMALL_MAX_OPS = 112;   // 112 * 28 B of ops + the job record stay under the 4 KB kernel-argument limit
constexpr uint32_t SMALL_MAX_V = 1024;    // vertices (= edge slots) per fill
constexpr uint32_t SMALL_MAX_LEAVES = 64; // points per curve op
constexpr uint32_t SMALL_MAX_DIM = 1024;  // raster width / owned rows
constexpr uint32_t SMALL_WARPS = 8;

struct SmallArgs {
    JobDesc job;  // op_begin = 0, op_end = n_ops
    uint32_t n_ops, W, H, row_begin, row_end, pitch, bpp, seq;  // seq: written to the host's completion word by the last CTA
    ftl_path_op ops[SMALL_MAX_OPS];
};
static_assert(sizeof(SmallArgs) <= 3400, "SmallArgs must fit the kernel argument space with the pointers beside it");

struct SmallNode {  // one node of a subdivision level: four control points and its tag
    float v[8];
    uint32_t key;
};
struct SmallShared {
    Vtx V[SMALL_MAX_V];
    EdgeRec E[SMALL_MAX_V];
    int32_t pool[2 * SMALL_MAX_V];          // Fixed (x, y) of the kept points of every op, op after op (allocation order)
    uint32_t sub_last[SMALL_MAX_V];
    uint32_t op_off[SMALL_MAX_OPS + 1];     // pool offset of the op's points
    uint32_t op_cnt[SMALL_MAX_OPS + 1];     // kept points of the op
    uint32_t op_start[SMALL_MAX_OPS + 1];   // 1: the op starts a sub-figure
    uint32_t op_voff[SMALL_MAX_OPS + 1];    // first vertex of the op
    uint32_t op_sub[SMALL_MAX_OPS + 1];     // first vertex of the op's sub-figure
    float leaf_x[SMALL_WARPS][SMALL_MAX_LEAVES], leaf_y[SMALL_WARPS][SMALL_MAX_LEAVES];
    uint32_t leaf_key[SMALL_WARPS][SMALL_MAX_LEAVES];
    int32_t sorted[SMALL_WARPS][2 * SMALL_MAX_LEAVES];
    SmallNode nodes[SMALL_WARPS][32];
    int4 stage[2 * BIN_PACKED_MAX];         // the band's edges, prepared (two int4 each)
    uint32_t item_end[BIN_PACKED_MAX];      // inclusive prefix of the rows each staged edge has inside the band
    JobState js;
    unsigned long long top_key;
    uint32_t top_vid, pool_used, nv, n_popped, overflow, n_edges, n_staged, touched;
};

// One curve op, one warp: level-parallel subdivision.  Returns the number of leaves (points) written in
// depth-first order to out_xy (Fixed pairs), or 0xFFFFFFFF when a level or the leaf list overflows.
template <bool CUBIC>
__device__ __forceinline__ uint32_t small_flatten_curve(pointy::Pt a, pointy::Pt b, pointy::Pt c, pointy::Pt d, float tol_sq, float *lx, float *ly,
                                                        uint32_t *lkey, SmallNode *nodes, int32_t *out_xy) {
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    bool valid = lane == 0;
    uint32_t key = 0, n_leaf = 0;
    for (int depth = 0;; depth++) {
        pointy::Pt la, lb, lc, ld, ra, rb, rc, rd;  // children
        bool flat = false;
        if (CUBIC) {  // plotter.rs:311-332
            const pointy::Pt ab = pointy::midpoint(a, b), bc = pointy::midpoint(b, c), cd = pointy::midpoint(c, d);
            const pointy::Pt ab_bc = pointy::midpoint(ab, bc), bc_cd = pointy::midpoint(bc, cd);
            const pointy::Pt pe = pointy::midpoint(ab_bc, bc_cd), ad = pointy::midpoint(a, d);
            flat = pointy::distance_sq(pe, ad) <= tol_sq || depth >= MAX_DEPTH;
            la = a; lb = ab; lc = ab_bc; ld = pe;
            ra = pe; rb = bc_cd; rc = cd; rd = d;
        } else {  // plotter.rs:248-265; the quad's points are (a, b, c): d mirrors c
            const pointy::Pt ab = pointy::midpoint(a, b), bc = pointy::midpoint(b, c), ab_bc = pointy::midpoint(ab, bc), ac = pointy::midpoint(a, c);
            flat = pointy::distance_sq(ab_bc, ac) <= tol_sq || depth >= MAX_DEPTH;
            la = a; lb = ab; lc = ab_bc; ld = ab_bc;
            ra = ab_bc; rb = bc; rc = c; rd = c;
        }
        const uint32_t leafs = __ballot_sync(0xFFFFFFFFu, valid && flat), splits = __ballot_sync(0xFFFFFFFFu, valid && !flat);
        if (n_leaf + __popc(leafs) > SMALL_MAX_LEAVES || 2 * __popc(splits) > 32) return 0xFFFFFFFFu;
        if (valid && flat) {
            const uint32_t slot = n_leaf + __popc(leafs & lt);
            lx[slot] = CUBIC ? d.x : c.x;
            ly[slot] = CUBIC ? d.y : c.y;
            lkey[slot] = key;
        }
        n_leaf += __popc(leafs);
        if (splits == 0) break;
        // the r-th splitting lane hands its two children to lanes 2r and 2r + 1 of the next level through shared memory
        if (valid && !flat) {
            const uint32_t r = __popc(splits & lt);
            SmallNode &nl = nodes[2 * r], &nr = nodes[2 * r + 1];
            nl.v[0] = la.x; nl.v[1] = la.y; nl.v[2] = lb.x; nl.v[3] = lb.y; nl.v[4] = lc.x; nl.v[5] = lc.y; nl.v[6] = ld.x; nl.v[7] = ld.y;
            nr.v[0] = ra.x; nr.v[1] = ra.y; nr.v[2] = rb.x; nr.v[3] = rb.y; nr.v[4] = rc.x; nr.v[5] = rc.y; nr.v[6] = rd.x; nr.v[7] = rd.y;
            nl.key = key;
            nr.key = key | (1u << (15 - depth));
        }
        __syncwarp();
        valid = lane < 2u * __popc(splits);
        if (valid) {
            const SmallNode &n = nodes[lane];
            a = {n.v[0], n.v[1]}; b = {n.v[2], n.v[3]}; c = {n.v[4], n.v[5]}; d = {n.v[6], n.v[7]};
            key = n.key;
        }
        __syncwarp();
    }
    __syncwarp();
    // depth-first order = ascending tag: rank every leaf (n_leaf <= 64)
    for (uint32_t i = lane; i < n_leaf; i += 32) {
        const uint32_t k = lkey[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n_leaf; j++) rank += lkey[j] < k ? 1u : 0u;
        out_xy[2 * rank] = fx_from_f32(lx[i]);
        out_xy[2 * rank + 1] = fx_from_f32(ly[i]);
    }
    __syncwarp();
    return n_leaf;
}

template <int FMT, bool ALIGNED>
__global__ void __launch_bounds__(SMALL_WARPS * 32) small_fill(const __grid_constant__ SmallArgs A, JobState *__restrict__ js_out, Counters *__restrict__ cnt_out,
                                                              EdgeRec *__restrict__ edges_out, uint32_t *__restrict__ poison, uint32_t *__restrict__ host_flag,
                                                              uint32_t *__restrict__ done_count, uint32_t *__restrict__ host_done, long long *__restrict__ prof) {
    extern __shared__ __align__(16) uint8_t small_smem[];
#define FTL_STAMP(k) do { if (prof && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) prof[k] = clock64(); } while (0)
    FTL_STAMP(0);
    SmallShared &S = *reinterpret_cast<SmallShared *>(small_smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_ops = A.n_ops;
    // an earlier small fill may be waiting to be repeated by the host: then this one keeps the order and draws nothing.
    // The load is issued now and consumed after the flatten stage, off the critical path.
    const uint32_t poisoned = tid == 0 ? *reinterpret_cast<const volatile uint32_t *>(poison) : 0u;
    if (tid == 0) {
        S.pool_used = 0; S.n_popped = 0;
        S.overflow = 0;
        S.top_key = ~0ull; S.top_vid = NONE32;
        S.n_edges = 0; S.n_staged = 0; S.touched = 0;
    }
    for (uint32_t i = tid; i <= SMALL_MAX_OPS; i += blockDim.x) { S.op_cnt[i] = 0; S.op_start[i] = 0; S.op_off[i] = 0; }
    __syncthreads();
    float e[6];
#pragma unroll
    for (int k = 0; k < 6; k++) e[k] = A.job.e[k];

    // ---- (a) flatten: a warp per op (curves level-parallel, lines and moves by lane 0) ----
    for (uint32_t i = warp; i < n_ops; i += SMALL_WARPS) {
        const ftl_path_op &op = A.ops[i];
        if (op.tag < FTL_OP_MOVE || op.tag > FTL_OP_CUBIC) continue;
        const PenInfo pi = find_pen(A.ops, 0, i);
        const bool starts = pi.starts_sub || op.tag == FTL_OP_MOVE;  // Move closes the current sub-figure (plotter.rs:210)
        const pointy::Pt a = pointy::transform(e, pi.pen);
        int32_t *out = S.sorted[warp];
        uint32_t n;
        if (op.tag == FTL_OP_MOVE || op.tag == FTL_OP_LINE) {
            const pointy::Pt p = pointy::transform(e, {op.v[0], op.v[1]});
            if (lane == 0) { out[0] = fx_from_f32(p.x); out[1] = fx_from_f32(p.y); }
            n = 1;
            __syncwarp();
        } else if (op.tag == FTL_OP_QUAD) {
            const pointy::Pt b = pointy::transform(e, {op.v[0], op.v[1]}), c = pointy::transform(e, {op.v[2], op.v[3]});
            n = small_flatten_curve<false>(a, b, c, c, A.job.tol_sq, S.leaf_x[warp], S.leaf_y[warp], S.leaf_key[warp], S.nodes[warp], out);
        } else {
            const pointy::Pt b = pointy::transform(e, {op.v[0], op.v[1]}), c = pointy::transform(e, {op.v[2], op.v[3]}), d = pointy::transform(e, {op.v[4], op.v[5]});
            n = small_flatten_curve<true>(a, b, c, d, A.job.tol_sq, S.leaf_x[warp], S.leaf_y[warp], S.leaf_key[warp], S.nodes[warp], out);
        }
        if (n == 0xFFFFFFFFu) {
            if (lane == 0) S.overflow = 1;
            continue;
        }
        // point intake: a point equal to its predecessor is dropped unless it starts a sub-figure (fig.rs:428-440);
        // the predecessor of the op's first point is the pen position
        const int32_t pen_x = fx_from_f32(a.x), pen_y = fx_from_f32(a.y);
        uint32_t kept = 0, base = 0;
        for (uint32_t t0 = 0; t0 < n; t0 += 32) {
            const uint32_t t = t0 + lane;
            bool keep = false;
            int32_t x = 0, y = 0;
            if (t < n) {
                x = out[2 * t]; y = out[2 * t + 1];
                const int32_t qx = t ? out[2 * t - 2] : pen_x, qy = t ? out[2 * t - 1] : pen_y;
                keep = (t == 0 && starts) || x != qx || y != qy;
            }
            const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
            if (t0 == 0) {  // the op's pool space: at most n points
                if (lane == 0) base = atomicAdd(&S.pool_used, n);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base + n > SMALL_MAX_V) {
                    if (lane == 0) S.overflow = 1;
                    break;
                }
            }
            if (keep) {
                const uint32_t slot = base + kept + __popc(km & ((1u << lane) - 1u));
                S.pool[2 * slot] = x;
                S.pool[2 * slot + 1] = y;
            }
            kept += __popc(km);
        }
        if (lane == 0) { S.op_off[i] = base; S.op_cnt[i] = kept; S.op_start[i] = starts ? 1u : 0u; }
    }
    if (tid == 0 && poisoned) S.overflow = 1;
    __syncthreads();
    FTL_STAMP(1);  // flatten done
    bool draw = !S.overflow;
    if (draw) {
    // ---- vertex offsets and sub-figure heads over the ops: n_ops <= 112, so every op's thread just adds up the
    // counts before it (broadcast shared-memory reads, no shuffle chains, no extra barrier) ----
    if (tid <= n_ops) {
        uint32_t voff = 0, head = NONE32;
        for (uint32_t k = 0; k < tid; k++) {
            if (S.op_start[k]) head = voff;
            voff += S.op_cnt[k];
        }
        S.op_voff[tid] = voff;
        S.op_sub[tid] = head;  // first vertex of the latest sub-figure started before this op
        if (tid == n_ops) S.nv = voff;
    }
    __syncthreads();
    const uint32_t nv = S.nv;
    // ---- vertices in op order ----
    for (uint32_t i = warp; i < n_ops; i += SMALL_WARPS) {
        const uint32_t cnt = S.op_cnt[i], off = S.op_off[i], voff = S.op_voff[i];
        const uint32_t sub = S.op_start[i] ? voff : S.op_sub[i];
        for (uint32_t t = lane; t < cnt; t += 32) S.V[voff + t] = {S.pool[2 * (off + t)], S.pool[2 * (off + t) + 1], sub, 0u};
    }
    __syncthreads();
    FTL_STAMP(2);  // vertices in place
    // ---- (b) top-left vertex (fig.rs:493-494), closing rule (fig.rs:373-383) ----
    for (uint32_t k = tid; k < nv; k += blockDim.x) {
        const Vtx v = S.V[k];
        const bool last = vtx_is_last(S.V, nv, k);
        const bool pop = last && vtx_same(v, S.V[v.sub]);
        if (last) S.sub_last[v.sub] = pop ? (k > v.sub ? k - 1 : NONE32) : k;
        if (pop) atomicAdd(&S.n_popped, 1u);
        else atomicMin(&S.top_key, vtx_key(v));
    }
    __syncthreads();
    for (uint32_t k = tid; k < nv; k += blockDim.x) {
        const Vtx v = S.V[k];
        const bool pop = vtx_is_last(S.V, nv, k) && vtx_same(v, S.V[v.sub]);
        if (!pop && vtx_key(v) == S.top_key) atomicMin(&S.top_vid, k);
    }
    __syncthreads();
    if (tid == 0) {  // Fig::get_dir + top_row (fig.rs:402-411,495-497)
        JobState s;
        s.top_key = S.top_key; s.top_vid = S.top_vid; s.dir = 0; s.top_row = 0; s.first_row = 0x7FFFFFFF; s.shift = 0;
        s.vtx_begin = 0; s.vtx_end = nv; s.pad[0] = s.pad[1] = s.pad[2] = 0;
        const uint32_t k = S.top_vid;
        if (k != NONE32) {
            const Vtx v = S.V[k];
            const uint32_t f = vtx_next_fwd(S.V, nv, k, v, vtx_is_last(S.V, nv, k));
            const uint32_t r = k > v.sub ? k - 1 : S.sub_last[v.sub];
            const Vtx pf = S.V[f], pr = S.V[r];
            const fx_t ax = fx_sub(pr.x, v.x), ay = fx_sub(pr.y, v.y), bx = fx_sub(pf.x, v.x), by = fx_sub(pf.y, v.y);
            const bool widdershins = fx_mul(ax, by) > fx_mul(bx, ay);  // fig.rs:116-119
            const int32_t top = fx_to_i32(v.y);
            s.dir = widdershins ? 0 : 1;
            s.top_row = top;
            s.first_row = top > 0 ? top : 0;
            s.shift = top < 0 ? top : 0;
        }
        S.js = s;
        if (blockIdx.x == 0) {
            *js_out = s;
            cnt_out->nv = nv;
            cnt_out->n_popped = S.n_popped;
        }
    }
    __syncthreads();
    FTL_STAMP(3);  // top vertex, direction
    const JobState js = S.js;
    draw = js.top_vid != NONE32;  // no vertices: nothing is drawn (fig.rs:491)
    // ---- edges: one per ring segment (fig.rs:179-210,576-600) ----
    for (uint32_t k = tid; k < nv && draw; k += blockDim.x) {
        const Vtx v = S.V[k];
        const bool last = vtx_is_last(S.V, nv, k);
        const bool pop = last && vtx_same(v, S.V[v.sub]);
        EdgeRec ed;
        ed.flags = 0;
        if (!pop) {
            const uint32_t w = vtx_next_fwd(S.V, nv, k, v, last);
            if (w != k) {
                const Vtx q = S.V[w];
                if (q.y > v.y) ed = make_edge(v, q, 0u, 0u, js);
                else if (q.y < v.y) ed = make_edge(q, v, 0u, 1u, js);
            }
        }
        if (!ed.flags) { ed.x_bot0 = 0; ed.inv_slope = 0; ed.step_pix = 0; ed.ry0 = 0; ed.ry1 = -1; ed.fr = 0; ed.job = 0; }
        else atomicAdd(&S.n_edges, 1u);
        S.E[k] = ed;
        if (blockIdx.x == 0 && edges_out) edges_out[k] = ed;
    }
    __syncthreads();
    // the scatter below adds packed 32-bit words: exact while fewer than 128 edges can meet in a cell (bin_kernel.cuh)
    if (draw && S.n_edges >= BIN_PACKED_MAX) {
        if (tid == 0) S.overflow = 1;
        draw = false;
    }
    __syncthreads();

    FTL_STAMP(4);  // edges
    // ---- (c)+(d): this CTA draws one band of 32 rows; the tile spans the whole raster width ----
    const int32_t W = (int32_t)A.W;
    const uint32_t seg = A.W <= 256u ? 256u : 512u;                       // columns per resolve pass
    const uint32_t row_bytes = ((A.W + seg - 1) / seg) * seg * 2u + BIN_ROW_PAD;
    uint32_t cells = smem_addr(small_smem) + (uint32_t)((sizeof(SmallShared) + 15u) & ~15u);
    asm volatile("" : "+r"(cells));
    const int32_t row0 = (int32_t)A.row_begin + (int32_t)(blockIdx.x << BIN_LOG2R);
    const int32_t row = row0 + (int32_t)lane;
    const bool row_ok = draw && row >= js.first_row && row < (int32_t)A.row_end;  // rows above the figure are untouched (fig.rs:497)
    const uint32_t valid_mask = __ballot_sync(0xFFFFFFFFu, row_ok);
    if (valid_mask != 0) {
        const int32_t v_lo = __ffs((int)valid_mask) - 1, v_hi = 31 - __clz((int)valid_mask);
        for (uint32_t i = tid; i < BIN_ROWS * row_bytes / 16; i += blockDim.x) ssts4_bias(cells + 16u * i);
        // the edges crossing the band, prepared once (compacted in any order: sums do not depend on it)
        for (uint32_t k = tid; k < nv; k += blockDim.x) {
            const EdgeRec ed = S.E[k];
            const int32_t r0 = ed.ry0 - row0, r1 = ed.ry1 - row0;
            const int32_t nrows = min(r1, v_hi) - max(r0, v_lo) + 1;
            if (!(ed.flags & 1u) || nrows <= 0) continue;
            const uint32_t slot = atomicAdd(&S.n_staged, 1u);
            const fx_t fr0 = (fx_t)(ed.fr & 0xFFFFu), fr1 = (fx_t)(ed.fr >> 16);
            int4 a, b;
            a.x = (int32_t)((uint32_t)ed.x_bot0 + (uint32_t)(row0 - ed.ry0) * (uint32_t)ed.inv_slope);
            a.y = ed.inv_slope;
            a.z = ed.step_pix > 0 ? ed.step_pix : FX_ONE;
            a.w = (max(r0, -1) + 1) | (min(r1, (int32_t)BIN_ROWS) << 8);
            b.x = (ed.flags & 2u) ? -1 : 1;
            b.y = fx_mul(ed.inv_slope, FX_ONE - fr0);
            b.z = fx_mul(ed.inv_slope, (FX_ONE - fr1) & FX_MASK);
            b.w = (int32_t)((uint32_t)pixel_cov(fr0) | ((uint32_t)pixel_cov(fr1) << 9));
            S.stage[2 * slot] = a;
            S.stage[2 * slot + 1] = b;
            S.item_end[slot] = (uint32_t)nrows;
            atomicOr(&S.touched, ((2u << (nrows - 1)) - 1u) << max(r0, v_lo));
        }
        __syncthreads();
        const uint32_t n_staged = S.n_staged;
        if (warp == 0) {  // inclusive prefix of the item counts (fewer than 128 edges: four per lane)
            uint32_t c4[4], sum = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t i = lane * 4 + k;
                c4[k] = i < n_staged ? S.item_end[i] : 0u;
                sum += c4[k];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if ((int)lane >= d) inc += o;
            }
            uint32_t run = inc - sum;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t i = lane * 4 + k;
                run += c4[k];
                if (i < BIN_PACKED_MAX) S.item_end[i] = run;
            }
        }
        __syncthreads();
        FTL_STAMP(5);  // tile cleared, edges staged, prefix
        // (c) scatter: one (edge, row) item per thread and pass
        const uint32_t n_items = n_staged ? S.item_end[n_staged - 1] : 0u;
        for (uint32_t q = tid; q < n_items; q += blockDim.x) {
            uint32_t k = 0;  // first staged edge with item_end > q (entries beyond n_staged repeat the total)
#pragma unroll
            for (uint32_t h = 64; h > 0; h >>= 1)
                if (S.item_end[k + h - 1] <= q) k += h;
            const uint32_t before = k ? S.item_end[k - 1] : 0u;
            const int4 a = S.stage[2 * k], b = S.stage[2 * k + 1];
            const int32_t er0 = (a.w & 0xFF) - 1, er1 = a.w >> 8;
            bin_item_packed<false>(a.x, a.y, a.z, er0, er1, b.y, b.z, (uint32_t)b.w, b.x, max(er0, v_lo) + (int32_t)(q - before), W, 0, W, cells, row_bytes, 0u);
        }
        __syncthreads();
        FTL_STAMP(6);  // scatter
        // (d) resolve: the rows are dealt to the warps; a row wider than a pass carries its sum from pass to pass
        const uint32_t touched = S.touched & valid_mask, rule = A.job.rule, color = A.job.color;
        uint8_t *dst = reinterpret_cast<uint8_t *>(A.job.raster) + (size_t)(row0 - (int32_t)A.row_begin) * A.pitch;
        int32_t carry = 0;
        for (uint32_t x0 = 0; x0 < A.W; x0 += seg) {
            uint8_t *dwin = dst + (size_t)x0 * A.bpp;
            const uint32_t c0 = cells + 2u * x0, w_rel = A.W - x0;
            int32_t tot;
            if (seg == 256u) {
                if (rule == FTL_EVENODD) tot = bin_resolve<FMT, true, ALIGNED, 256, true>(c0, touched, carry, valid_mask, dwin, A.pitch, w_rel, color, row_bytes, 2 * warp, 2 * SMALL_WARPS);
                else tot = bin_resolve<FMT, false, ALIGNED, 256, true>(c0, touched, carry, valid_mask, dwin, A.pitch, w_rel, color, row_bytes, 2 * warp, 2 * SMALL_WARPS);
            } else {
                if (rule == FTL_EVENODD) tot = bin_resolve<FMT, true, ALIGNED, 512, true>(c0, touched, carry, valid_mask, dwin, A.pitch, w_rel, color, row_bytes, warp, SMALL_WARPS);
                else tot = bin_resolve<FMT, false, ALIGNED, 512, true>(c0, touched, carry, valid_mask, dwin, A.pitch, w_rel, color, row_bytes, warp, SMALL_WARPS);
            }
            carry += tot;
        }
    }
    }  // if (draw)
    FTL_STAMP(7);  // resolve
    // ---- completion: the last CTA tells the host (a word in mapped memory, so ftl_sync can watch it instead of the stream) ----
    __syncthreads();
    if (tid == 0) {
        __threadfence();  // this CTA's pixels are visible device-wide before it counts itself done
        if (atomicAdd(done_count, 1u) == gridDim.x - 1) {
            *done_count = 0u;
            if (S.overflow) {  // every CTA reaches the same verdict: the last one reports it
                *poison = 1u;
                cnt_out->overflow = 1u;
                *reinterpret_cast<volatile uint32_t *>(host_flag) = 1u;
                __threadfence_system();  // the flag reaches the host before the completion word
            }
            *reinterpret_cast<volatile uint32_t *>(host_done) = A.seq;
        }
    }
    FTL_STAMP(8);
#undef FTL_STAMP
}
