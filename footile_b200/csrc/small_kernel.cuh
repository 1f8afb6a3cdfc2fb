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
constexpr int SMALL_WC = 256;
constexpr uint32_t SMALL_WARPS = 8;

struct SmallArgs {
    JobDesc job;  // op_begin = 0, op_end = n_ops
    uint32_t n_ops, W, H, row_begin, row_end, pitch, bpp, pad;
    ftl_path_op ops[SMALL_MAX_OPS];
};
static_assert(sizeof(SmallArgs) <= 3400, "SmallArgs must fit the kernel argument space with the pointers beside it");

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
    JobState js;
    unsigned long long top_key;
    uint32_t top_vid, pool_used, nv, n_popped, overflow;
};

// One curve op, one warp: level-parallel subdivision.  Returns the number of leaves (points) written in
// depth-first order to out_xy (Fixed pairs), or 0xFFFFFFFF when a level or the leaf list overflows.
template <bool CUBIC>
__device__ __forceinline__ uint32_t small_flatten_curve(pointy::Pt a, pointy::Pt b, pointy::Pt c, pointy::Pt d, float tol_sq, float *lx, float *ly,
                                                        uint32_t *lkey, int32_t *out_xy) {
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
        // lane L of the next level is child (L & 1) of the (L >> 1)-th splitting lane
        const bool child = lane < 2u * __popc(splits);
        const uint32_t parent = child ? __fns(splits, 0, (lane >> 1) + 1) : 0u;
        const bool right = lane & 1u;
#define FTL_CHILD(f) (right ? __shfl_sync(0xFFFFFFFFu, r##f, parent) : __shfl_sync(0xFFFFFFFFu, l##f, parent))
        const float nax = FTL_CHILD(a.x), nay = FTL_CHILD(a.y), nbx = FTL_CHILD(b.x), nby = FTL_CHILD(b.y);
        const float ncx = FTL_CHILD(c.x), ncy = FTL_CHILD(c.y), ndx = FTL_CHILD(d.x), ndy = FTL_CHILD(d.y);
#undef FTL_CHILD
        const uint32_t pkey = __shfl_sync(0xFFFFFFFFu, key, parent);
        a = {nax, nay}; b = {nbx, nby}; c = {ncx, ncy}; d = {ndx, ndy};
        key = pkey | ((right ? 1u : 0u) << (15 - depth));
        valid = child;
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
                                                              EdgeRec *__restrict__ edges_out, uint32_t *__restrict__ poison, uint32_t *__restrict__ host_flag) {
    extern __shared__ __align__(16) uint8_t small_smem[];
    SmallShared &S = *reinterpret_cast<SmallShared *>(small_smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_ops = A.n_ops;
    if (tid == 0) {
        S.pool_used = 0; S.n_popped = 0;
        S.overflow = *poison;  // an earlier small fill is waiting to be repeated by the host: keep the order, draw nothing
        S.top_key = ~0ull; S.top_vid = NONE32;
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
            n = small_flatten_curve<false>(a, b, c, c, A.job.tol_sq, S.leaf_x[warp], S.leaf_y[warp], S.leaf_key[warp], out);
        } else {
            const pointy::Pt b = pointy::transform(e, {op.v[0], op.v[1]}), c = pointy::transform(e, {op.v[2], op.v[3]}), d = pointy::transform(e, {op.v[4], op.v[5]});
            n = small_flatten_curve<true>(a, b, c, d, A.job.tol_sq, S.leaf_x[warp], S.leaf_y[warp], S.leaf_key[warp], out);
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
    __syncthreads();
    if (S.overflow) {
        if (blockIdx.x == 0 && tid == 0) {
            *poison = 1u;
            *host_flag = 1u;
            cnt_out->overflow = 1u;
        }
        return;
    }
    // ---- vertex offsets and sub-figure heads over the ops (n_ops <= 112: one warp, four ops per lane) ----
    if (warp == 0) {
        uint32_t cnt[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = lane * 4 + k;
            cnt[k] = i < n_ops ? S.op_cnt[i] : 0u;
            sum += cnt[k];
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
            if (i <= n_ops && i <= SMALL_MAX_OPS) S.op_voff[i] = run;
            run += cnt[k];
        }
        if (lane == 31) S.nv = inc;
        __syncwarp();
        // head of the sub-figure of every op: the latest starting op at or before it
        uint32_t head = NONE32, hs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = lane * 4 + k;
            if (i < n_ops && S.op_start[i]) head = S.op_voff[i];
            hs[k] = head;
        }
        // carry the last head of the lanes below (a max-scan does it: heads grow with the op index, NONE32 = none)
        uint32_t hin = head == NONE32 ? 0u : head + 1u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, hin, d);
            if ((int)lane >= d) hin = max(hin, o);
        }
        const uint32_t below = __shfl_up_sync(0xFFFFFFFFu, hin, 1);
        const uint32_t carry_head = lane == 0 ? NONE32 : (below == 0 ? NONE32 : below - 1u);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = lane * 4 + k;
            if (i < n_ops) S.op_sub[i] = hs[k] != NONE32 ? hs[k] : carry_head;
        }
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
    const JobState js = S.js;
    if (js.top_vid == NONE32) return;  // no vertices: nothing is drawn (fig.rs:491)
    // ---- edges: one per ring segment (fig.rs:179-210,576-600) ----
    for (uint32_t k = tid; k < nv; k += blockDim.x) {
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
        S.E[k] = ed;
        if (blockIdx.x == 0 && edges_out) edges_out[k] = ed;
    }
    __syncthreads();

    // ---- (c)+(d): every warp draws bands of 32 rows with the lanes-are-rows scatter of bin_kernel.cuh ----
    typedef BinTile<SMALL_WC> T;
    uint32_t cells = smem_addr(small_smem) + (uint32_t)((sizeof(SmallShared) + 15u) & ~15u) + warp * T::BYTES;
    asm volatile("" : "+r"(cells));
    const uint32_t stage = cells + T::STAGE;
    for (uint32_t i = lane; i < T::CELL_BYTES / 16; i += 32) ssts4_bias(cells + 16u * i);
    __syncwarp();
    const int32_t W = (int32_t)A.W;
    const uint32_t n_bands = (A.row_end - A.row_begin + BIN_ROWS - 1) / BIN_ROWS, n_win = (A.W + SMALL_WC - 1) / SMALL_WC;
    const uint32_t rbase = cells + lane * T::ROW_BYTES;
    const uint32_t rule = A.job.rule, color = A.job.color;
    for (uint32_t band = blockIdx.x * SMALL_WARPS + warp; band < n_bands; band += gridDim.x * SMALL_WARPS) {
        const int32_t row0 = (int32_t)A.row_begin + (int32_t)(band << BIN_LOG2R);
        const int32_t row = row0 + (int32_t)lane;
        const bool row_ok = row >= js.first_row && row < (int32_t)A.row_end;
        const uint32_t valid_mask = __ballot_sync(0xFFFFFFFFu, row_ok);
        if (valid_mask == 0) continue;
        const bool band_full = valid_mask == 0xFFFFFFFFu;
        uint8_t *dst = reinterpret_cast<uint8_t *>(A.job.raster) + (size_t)(row0 - (int32_t)A.row_begin) * A.pitch;
        int32_t carry = 0;
        for (uint32_t w = 0; w < n_win; w++) {
            const int32_t win_lo = (int32_t)(w * (uint32_t)SMALL_WC), win_hi = min(W, win_lo + SMALL_WC);
            int32_t tot = 0;
            bool row_touched = false;
            for (uint32_t base = 0; base < nv; base += 32) {
                // stage the edges of this round that cross the band (compacted: the item loop visits only those)
                const EdgeRec ed = base + lane < nv ? S.E[base + lane] : S.E[0];
                const int32_t r0 = ed.ry0 - row0, r1 = ed.ry1 - row0;
                const bool hit = base + lane < nv && (ed.flags & 1u) && r0 < (int32_t)BIN_ROWS && r1 >= 0;
                const uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
                if (hits == 0) continue;
                if (hit) {
                    const uint32_t slot = __popc(hits & ((1u << lane) - 1u));
                    const fx_t fr0 = (fx_t)(ed.fr & 0xFFFFu), fr1 = (fx_t)(ed.fr >> 16);
                    const bool full = band_full && r0 < 0 && r1 >= (int32_t)BIN_ROWS;
                    int4 a, b;
                    a.x = (int32_t)((uint32_t)ed.x_bot0 + (uint32_t)(row0 - ed.ry0) * (uint32_t)ed.inv_slope);
                    a.y = ed.inv_slope;
                    a.z = ed.step_pix > 0 ? ed.step_pix : FX_ONE;
                    a.w = (max(r0, -1) + 1) | (min(r1, (int32_t)BIN_ROWS) << 8);
                    b.x = (ed.flags & 2u) ? -1 : 1;
                    b.y = fx_mul(ed.inv_slope, FX_ONE - fr0);
                    b.z = fx_mul(ed.inv_slope, (FX_ONE - fr1) & FX_MASK);
                    b.w = (int32_t)((uint32_t)pixel_cov(fr0) | ((uint32_t)pixel_cov(fr1) << 9) | (full ? 1u << 19 : 0u));
                    ssts4(stage + slot * 32u, a);
                    ssts4(stage + slot * 32u + 16u, b);
                }
                __syncwarp();
                row_touched = true;
                const uint32_t cnt = __popc(hits);
#pragma unroll 1
                for (uint32_t k = 0; k < cnt; k++) {
                    const int4 a = slds4(stage + k * 32u), b = slds4(stage + k * 32u + 16u);
                    if (b.w & (1 << 19))
                        bin_item_rows<true>(a.x, a.y, a.z, 0, 0, b.y, b.z, (uint32_t)b.w, b.x, (int32_t)lane, true, W, win_lo, win_hi, rbase, tot);
                    else
                        bin_item_rows<false>(a.x, a.y, a.z, (a.w & 0xFF) - 1, a.w >> 8, b.y, b.z, (uint32_t)b.w, b.x, (int32_t)lane, row_ok, W, win_lo, win_hi, rbase, tot);
                }
                __syncwarp();
            }
            const uint32_t touched = __ballot_sync(0xFFFFFFFFu, row_touched && row_ok);
            uint8_t *dwin = dst + (size_t)win_lo * A.bpp;
            if (rule == FTL_EVENODD) bin_resolve<FMT, true, ALIGNED, SMALL_WC>(cells, touched, carry, valid_mask, dwin, A.pitch, (uint32_t)(W - win_lo), color);
            else bin_resolve<FMT, false, ALIGNED, SMALL_WC>(cells, touched, carry, valid_mask, dwin, A.pitch, (uint32_t)(W - win_lo), color);
            __syncwarp();
            carry += tot;
        }
    }
}
