// bin_kernel.cuh — stages (c)+(d) for jobs with many edges: the binned tile kernel.
// Included by engine.cu inside namespace ftl (one translation unit: the kernels share Params / EdgeRec / ...).
#pragma once

// ---------------------------------------------------------------------------
// (c)+(d) binned tiles
// ---------------------------------------------------------------------------
// A tile is (job, band of 32 raster rows, window of WC columns).  One warp owns a tile and a private
// shared-memory image of it: 32 rows of WC wrapping i16 cells — the reference's own cell type
// (plotter.rs:45) — two cells per 32-bit word, the low one stored with a bias of 0x8000 (see below).
// The warp walks the tile's edge list (bin_edges) 32 entries per round and scatters each round in one
// of two ways:
//
//   T  LANES ARE ROWS (tall edges, dense tiles).  Every lane evaluates the SAME edge on ITS row of the
//      band in closed form (fig.rs:238-321,557-600) and adds the span's coverage deltas to
//      its own row with plain 16-bit ld/add/st: lanes never share a cell, so there are no atomics
//      (the round-1 scatter spent 2.9 wavefronts per shared atomic on bank conflicts), the edge record
//      is warp-uniform (staged in shared memory, read by broadcast), control flow is uniform (an edge
//      has nearly the same span length on neighbouring rows) and the row totals live in lane
//      registers.  An edge crossing the whole band takes a path with no start / end handling.
//   S  LANES ARE (EDGE, ROW) ITEMS (short edges, sparse tiles: curves flattened to segments a few rows
//      tall would leave most row-lanes idle).  A prefix sum over the rows each staged edge has inside
//      the band numbers the round's items; every lane takes an item, finds its edge by a 5-step
//      search of the prefix and evaluates that row.  Lanes can meet in a cell, so deltas are added with ONE
//      32-bit red.shared per cell to the word holding the cell pair: delta for the low cell,
//      delta << 16 for the high one.  A 32-bit add carries from the low cell into the high one when
//      the low half crosses 0 / 65536; the bias keeps the low half at 0x8000 + (sum so far), so no
//      add can carry while |sum of the deltas of one cell| < 32768.  A cell receives at most 256 per
//      edge of the tile, hence S is used only in tiles of fewer than 128 edges: exact, not "almost".
//
// The rows are then resolved WC/16 lanes per row (16 cells per lane, packed 16-bit prefix sums), the
// fill rule applied and the pixels stored / blended exactly as the direct kernel does (emit16).
//
// Wide rasters: the running sum of a row crosses windows.  Each (band, window) tile is claimed from
// a ticket counter in window-major order; a tile publishes row sums (carry-in + its own row totals,
// known right after the scatter) to its right neighbour through one 32-bit word per row holding
// {launch epoch, 16-bit sum}: the consumer spins on that word only, so no fence is needed.  A
// predecessor always has a smaller ticket, hence is running or done: no deadlock, whatever the
// residency.  Narrow rasters (few windows, many tiles) walk the windows of a band serially with
// the sums in registers instead.
constexpr uint32_t BIN_ROWS = 32;
constexpr uint32_t BIN_LOG2R = 5;
constexpr uint32_t BIN_ROW_PAD = 16;       // bytes between rows beyond WC * 2: neighbouring rows start 4 banks apart
constexpr uint32_t BIN_BIAS = 0x00008000u; // resting value of a cell-pair word
constexpr uint32_t BIN_PACKED_MAX = 128;   // tiles with fewer edges may add deltas as packed 32-bit words

template <int WC>
struct BinTile {
    static constexpr uint32_t ROW_BYTES = WC * 2 + BIN_ROW_PAD;
    static constexpr uint32_t CELL_BYTES = BIN_ROWS * ROW_BYTES;
    static constexpr uint32_t STAGE = CELL_BYTES;             // 32 staged edge records of 32 bytes
    static constexpr uint32_t TOT = STAGE + 32 * 32;          // 32 row totals (S rounds)
    static constexpr uint32_t PREFIX = TOT + 32 * 4;          // inclusive prefix of the items per staged edge (S rounds)
    static constexpr uint32_t FLAGS = PREFIX + 32 * 4;        // touched-row mask (S rounds)
    static constexpr uint32_t BYTES = FLAGS + 16;
};

__device__ __forceinline__ uint32_t slds16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void ssts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v)); }
__device__ __forceinline__ void ssts4(uint32_t a, int4 v) {
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ void ssts4_bias(uint32_t a) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(BIN_BIAS));
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v)); }

// One edge on one row, cells of [win_lo, win_hi) only: the span of the row and the coverage of its first cell.
//   x_bot     X at the bottom of the row (closed form of advance_edges, fig.rs:569-573)
//   dxs, dxe  inv_slope * (ONE - fract(y_upper)), inv_slope * ((ONE - fract(y_lower)) & MASK)   (fig.rs:244-249,262-278)
//   misc      pixel_cov(fract(y_upper)) | pixel_cov(fract(y_lower)) << 9 | negative sign << 18
// Returns false when the row receives nothing inside the window.  On success: cells [c, ...) receive
// X(k) - X(k-1) with X(k) = min(pixel_cov(min(xc + k*step, ONE)), cov) and X(-1) = prev (fig.rs:285-321).
// PLAIN: neither the first nor the last row of the edge (cov = 256).  VOTE: all lanes are here (lanes = rows).
template <bool PLAIN, bool VOTE>
__device__ __forceinline__ bool bin_span(fx_t x_bot, int32_t inv_slope, int32_t step, bool starting, bool ending, int32_t dxs, int32_t dxe, uint32_t misc,
                                         bool act, int32_t W, int32_t win_lo, int32_t win_hi, int32_t &c, int32_t &xc, int32_t &prev, int32_t &cov,
                                         bool &resumed) {
    fx_t x0, x1;
    if (PLAIN) {
        x0 = fx_sub(x_bot, inv_slope);
        x1 = x_bot;
        cov = 256;
    } else {
        x0 = fx_sub(x_bot, starting ? dxs : inv_slope);
        x1 = ending ? fx_sub(x_bot, dxe) : x_bot;
        cov = (ending ? (int32_t)((misc >> 9) & 0x1FFu) : 256) - (starting ? (int32_t)(misc & 0x1FFu) : 0);
        if (cov <= 0) act = false;
    }
    const fx_t min_x = fx_min(x0, x1), max_x = fx_max(x0, x1);
    const int32_t min_pix = fx_to_i32(min_x), max_pix = fx_to_i32(max_x);
    const int32_t c0 = min_pix > 0 ? min_pix : 0;
    c = c0 > win_lo ? c0 : win_lo;
    if (min_pix >= W || c >= win_hi) act = false;
    // first_cov / step_cov (fig.rs:305-321): full_cov = cov / 256 in Fixed = cov << 8, so fx_mul(a, cov << 8) = (a * cov) >> 8
    const int32_t rr = min_pix == max_pix ? (int32_t)(((uint32_t)(FX_ONE - fx_fract(fx_avg(max_x, min_x))) * (uint32_t)cov) >> 8)
                                          : (FX_ONE - fx_fract(min_x)) >> 1;
    const int32_t first = (int32_t)(((uint64_t)(uint32_t)rr * (uint64_t)(uint32_t)step) >> 16);  // step == ONE for a vertical edge: first = rr
    xc = first;
    prev = 0;
    // a span that starts left of the window (or of the raster): resume inside it (rare)
    const bool resume = act && c != min_pix;
    resumed = VOTE ? __any_sync(0xFFFFFFFFu, resume) : resume;
    if (resumed) {
        const int64_t xc64 = (int64_t)first + (int64_t)(c - min_pix) * (int64_t)step;
        xc = (int32_t)(xc64 < (int64_t)FX_ONE ? xc64 : (int64_t)FX_ONE);
        if (c > c0) {
            const int64_t xq = xc64 - (int64_t)step;
            const int32_t xk = pixel_cov((fx_t)(xq < (int64_t)FX_ONE ? xq : (int64_t)FX_ONE));
            prev = xk < cov ? xk : cov;
            if (prev >= cov) act = false;
        }
    }
    return act;
}

// T: the lane's row of one staged edge, plain 16-bit read-modify-write of the lane's own cells.
//   xb  X at the bottom of band row 0;  r0, r1  first / last row of the edge relative to the band;  ed  +1 / -1 (fig.rs:286)
template <bool FULL>
__device__ __forceinline__ void bin_item_rows(int32_t xb, int32_t inv_slope, int32_t step, int32_t r0, int32_t r1, int32_t dxs, int32_t dxe, uint32_t misc,
                                              int32_t ed, int32_t rel_row, bool row_ok, int32_t W, int32_t win_lo, int32_t win_hi, uint32_t rbase, int32_t &tot) {
    const fx_t x_bot = (fx_t)((uint32_t)xb + (uint32_t)rel_row * (uint32_t)inv_slope);
    int32_t c, xc, prev, cov;
    bool resumed;
    const bool act = bin_span<FULL, true>(x_bot, inv_slope, step, rel_row == r0, rel_row == r1, dxs, dxe, misc,
                                          FULL || (row_ok && rel_row >= r0 && rel_row <= r1), W, win_lo, win_hi, c, xc, prev, cov, resumed);
    uint32_t a = rbase + ((uint32_t)(c - win_lo) << 1);
    const uint32_t a_end = rbase + ((uint32_t)(win_hi - win_lo) << 1);
    if (step == FX_ONE && !resumed) {
        // |dx/dy| <= 1 (warp-uniform): the span covers at most two cells, X(0) = pixel_cov(first) and X(1) = cov - no loop
        if (act) {
            int32_t x0 = (xc + 128) >> 8;
            if (x0 > cov) x0 = cov;
            ssts16(a, slds16(a) + (uint32_t)(ed * x0));
            const bool two = x0 < cov && a + 2 < a_end;
            if (two) ssts16(a + 2, slds16(a + 2) + (uint32_t)(ed * (cov - x0)));
            tot += ed * (two ? cov : x0);
        }
        return;
    }
    if (act) {
        const int32_t prev0 = prev;
        int32_t xr = xc + 128;  // pixel_cov of a value in [0, ONE] is (x + 128) >> 8
        for (;;) {
            int32_t xk = xr >> 8;
            if (xk > cov) xk = cov;
            ssts16(a, slds16(a) + (uint32_t)(ed * (xk - prev)));
            prev = xk;
            a += 2;
            xr = min(xr + step, FX_ONE + 128);
            if (xk >= cov || a >= a_end) break;
        }
        tot += ed * (prev - prev0);
    }
}

// S: one (edge, row) item of a staged edge, packed 32-bit adds (see the file header for why they are exact here).
// TOT: also add the item's total to the row's word at totbase (tiles that hand row sums to a neighbour window).
template <bool TOT>
__device__ __forceinline__ void bin_item_packed(int32_t xb, int32_t inv_slope, int32_t step, int32_t r0, int32_t r1, int32_t dxs, int32_t dxe, uint32_t misc,
                                                int32_t ed, int32_t rel_row, int32_t W, int32_t win_lo, int32_t win_hi, uint32_t cells, uint32_t row_bytes,
                                                uint32_t totbase) {
    const fx_t x_bot = (fx_t)((uint32_t)xb + (uint32_t)rel_row * (uint32_t)inv_slope);
    int32_t c, xc, prev, cov;
    bool resumed;
    if (!bin_span<false, false>(x_bot, inv_slope, step, rel_row == r0, rel_row == r1, dxs, dxe, misc, true, W, win_lo, win_hi, c, xc, prev, cov, resumed)) return;
    const int32_t prev0 = prev;
    const uint32_t rbase = cells + (uint32_t)rel_row * row_bytes;
    int32_t rel = c - win_lo;
    const int32_t end_rel = win_hi - win_lo;
    int32_t xr = xc + 128;
    for (;;) {
        int32_t xk = xr >> 8;
        if (xk > cov) xk = cov;
        sred_add(rbase + (((uint32_t)rel >> 1) << 2), (ed * (xk - prev)) << ((rel & 1) << 4));
        prev = xk;
        rel++;
        xr = min(xr + step, FX_ONE + 128);
        if (xk >= cov || rel >= end_rel) break;
    }
    if (TOT) sred_add(totbase + 4u * (uint32_t)rel_row, ed * (prev - prev0));
}

// alpha bytes of four consecutive pixels from two words of packed wrapped-i16 sums (lo = pixel 2j, hi = pixel 2j+1)
// and the sums reaching them (imgbuf.rs:54-66,157-167; fig.rs:637-664)
template <bool EVEN_ODD>
__device__ __forceinline__ uint32_t packed_alpha(uint32_t pa, uint32_t pb, int32_t base_a, int32_t base_b) {
    const uint32_t ba = __byte_perm((uint32_t)base_a, 0u, 0x1010), bb = __byte_perm((uint32_t)base_b, 0u, 0x1010);
    if (!EVEN_ODD) {
        const uint32_t lo = __viaddmin_s16x2_relu(pa, ba, 0x00FF00FFu);  // clamp(i16(p + base), 0, 255) per halfword
        const uint32_t hi = __viaddmin_s16x2_relu(pb, bb, 0x00FF00FFu);
        return __byte_perm(lo, hi, 0x6420);
    } else {
        uint32_t lo = __viaddmin_s16x2(pa, ba, 0x7FFF7FFFu);  // wrapping i16 add of the base, per halfword
        uint32_t hi = __viaddmin_s16x2(pb, bb, 0x7FFF7FFFu);
        const uint32_t bl = (lo >> 8) & 0x00010001u, bh = (hi >> 8) & 0x00010001u;
        lo = ((lo & 0x00FF00FFu) ^ (bl * 0xFFu)) + bl;
        hi = ((hi & 0x00FF00FFu) ^ (bh * 0xFFu)) + bh;
        lo = __vimin_s16x2_relu(lo, 0x00FF00FFu);
        hi = __vimin_s16x2_relu(hi, 0x00FF00FFu);
        return __byte_perm(lo, hi, 0x6420);
    }
}

// SrcOver of one constant alpha over 16 bytes of pixels already in registers (the three cases of fill_const)
template <int FMT>
__device__ __forceinline__ void blend_const_u4(uint4 *p, uint4 t, uint32_t a, uint32_t color, uint32_t clr_a) {
    if (a == 255u && clr_a == 255u) {
        const uint32_t w = FMT == FTL_RGBA8P ? mul255_x4(color) : mul255_x4((color & 0xFFFFu) * 0x00010001u);
        *p = make_uint4(w, w, w, w);
    } else if (a == 0u) {
        mul255_rmw(p, t);
    } else {
        const uint32_t sa1 = 255u - pix::ch8_mul(a, clr_a);
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
        *p = t;
    }
}

// A tile without edges in a read-modify-write format: every row takes the constant alpha of its running sum.
// Lanes = 4 rows x 8 columns of 16 bytes, eight loads in flight per lane (the blend of a constant span is
// bound by read latency: alpha 0 changes almost no pixel, so nearly all of it is loads).
template <int FMT, bool EVEN_ODD>
__device__ __forceinline__ void bin_const_tile(int32_t carry, uint32_t valid_mask, uint8_t *dst_win, uint32_t pitch, uint32_t row_u4, uint32_t color,
                                               uint32_t clr_a) {
    const uint32_t lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
#pragma unroll 1
    for (uint32_t s = 0; s < BIN_ROWS; s += 4) {
        const uint32_t row = s + sub;
        const int32_t cr = __shfl_sync(0xFFFFFFFFu, carry, row);
        if (!((valid_mask >> s) & 0xFu)) continue;
        const bool ok = (valid_mask >> row) & 1u;
        const uint32_t a = rule_alpha<EVEN_ODD>(cr);
        uint4 *p = reinterpret_cast<uint4 *>(dst_win + (size_t)row * pitch);
#pragma unroll 1
        for (uint32_t u0 = 0; u0 < row_u4; u0 += 64) {
            uint4 t[8];
            const bool opaque = a == 255u && clr_a == 255u;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t u = u0 + l8 + 8u * i;
                t[i] = make_uint4(0u, 0u, 0u, 0u);
                if (ok && u < row_u4 && !opaque) t[i] = p[u];
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t u = u0 + l8 + 8u * i;
                if (ok && u < row_u4) blend_const_u4<FMT>(p + u, t[i], a, color, clr_a);
            }
        }
    }
}

// Resolve a tile: rows of WC cells, WC/16 lanes per row, 32/(WC/16) rows per step.
//   touched : rows that received deltas;   carry : this LANE's row (lane = row): running sum reaching the window
//   row_bytes : distance between the rows of the tile;   steps s_begin, s_begin + s_stride, ... (rows s .. s + RPS - 1 each)
//   TOTALS : return the sum of the row's cells (lane = row) for the rows of the steps taken, 0 elsewhere
template <int FMT, bool EVEN_ODD, bool ALIGNED, int WC, bool TOTALS = false>
__device__ __forceinline__ int32_t bin_resolve(uint32_t cells, uint32_t touched, int32_t carry, uint32_t valid_mask, uint8_t *dst_win, uint32_t pitch,
                                               uint32_t w_rel, uint32_t color, uint32_t row_bytes = BinTile<WC>::ROW_BYTES, uint32_t s_begin = 0,
                                               uint32_t s_stride = 32 / (WC / 16)) {
    constexpr uint32_t LPR = WC / 16, RPS = 32 / LPR;
    const uint32_t lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
    const uint32_t clr_a = FMT == FTL_RGBA8P ? (color >> 24) : ((color >> 8) & 0xFF);
    int32_t totals = 0;
    if (!TOTALS && FMT != FTL_MATTE8 && ALIGNED && touched == 0 && w_rel >= (uint32_t)WC && s_begin == 0 && s_stride == RPS) {
        constexpr uint32_t U = FMT == FTL_GRAYA8P ? 2u : 4u;  // 16-byte words per 16 pixels
        bin_const_tile<FMT, EVEN_ODD>(carry, valid_mask, dst_win, pitch, LPR * U, color, clr_a);
        return 0;
    }
#pragma unroll 1
    for (uint32_t s = s_begin; s < BIN_ROWS; s += s_stride) {
        const uint32_t row = s + sub;
        const int32_t cr = __shfl_sync(0xFFFFFFFFu, carry, row);
        const uint32_t step_rows = ((1u << RPS) - 1u) << s;
        if (!(valid_mask & step_rows)) continue;
        const bool ok = (valid_mask >> row) & 1u;
        uint8_t *drow = dst_win + (size_t)row * pitch;
        if (!(touched & step_rows)) {  // edge-free rows: one constant alpha
            const uint32_t q = rule_alpha<EVEN_ODD>(cr) * 0x01010101u;
            if (ok) emit16<FMT, ALIGNED>(drow, 16 * l, w_rel, q, q, q, q, color, clr_a);
            continue;
        }
        // every lane reads its 16 cells, touched row or not (an untouched row holds the resting value): a per-lane
        // branch here leaves the half-warps diverged for the rest of the step and every instruction issues twice
        int4 v0, v1;
        {
            const uint32_t a = cells + row * row_bytes + 32u * l;
            v0 = slds4(a);
            v1 = slds4(a + 16u);
            ssts4_bias(a);
            ssts4_bias(a + 16u);
        }
        // packed prefix inside each word: (lo, hi) -> (lo, lo + hi) mod 2^16.  The resting bias 0x8000 of the low cell is not
        // removed here: it rides into both halves of p_j (+0x8000 each) and into the running offsets (+0x8000 per word), and is
        // taken out again for free in the constants added to the bases below (kBiasEven / kBiasOdd).
        const uint32_t p0 = (uint32_t)v0.x * 0x00010001u, p1 = (uint32_t)v0.y * 0x00010001u, p2 = (uint32_t)v0.z * 0x00010001u,
                       p3 = (uint32_t)v0.w * 0x00010001u, p4 = (uint32_t)v1.x * 0x00010001u, p5 = (uint32_t)v1.y * 0x00010001u,
                       p6 = (uint32_t)v1.z * 0x00010001u, p7 = (uint32_t)v1.w * 0x00010001u;
        const int32_t o1 = (int32_t)(p0 >> 16), o2 = o1 + (int32_t)(p1 >> 16), o3 = o2 + (int32_t)(p2 >> 16), o4 = o3 + (int32_t)(p3 >> 16),
                      o5 = o4 + (int32_t)(p4 >> 16), o6 = o5 + (int32_t)(p5 >> 16), o7 = o6 + (int32_t)(p6 >> 16), tot = o7 + (int32_t)(p7 >> 16);
        int32_t inc = tot;
#pragma unroll
        for (int d = 1; d < (int)LPR; d <<= 1) scan_step<(int)LPR>(inc, d);
        const int32_t b = cr + inc - tot;
        if (TOTALS) {
#pragma unroll
            for (uint32_t q = 0; q < RPS; q++) {
                const int32_t tq = __shfl_sync(0xFFFFFFFFu, inc, q * LPR + LPR - 1);
                if (lane == s + q) totals = tq;
            }
        }
        // word j carries (j + 1) * 0x8000 of bias in both halves (its own + j words before it): mod 2^16 that is 0x8000 for
        // even j and 0 for odd j; tot carries 8 * 0x8000 = 0, so the scan and the row totals need no correction
        constexpr int32_t kBiasEven = -0x8000, kBiasOdd = 0;
        const uint32_t a0 = packed_alpha<EVEN_ODD>(p0, p1, b + kBiasEven, b + o1 + kBiasOdd);
        const uint32_t a1 = packed_alpha<EVEN_ODD>(p2, p3, b + o2 + kBiasEven, b + o3 + kBiasOdd);
        const uint32_t a2 = packed_alpha<EVEN_ODD>(p4, p5, b + o4 + kBiasEven, b + o5 + kBiasOdd);
        const uint32_t a3 = packed_alpha<EVEN_ODD>(p6, p7, b + o6 + kBiasEven, b + o7 + kBiasOdd);
        if (ok) emit16<FMT, ALIGNED>(drow, 16 * l, w_rel, a0, a1, a2, a3, color, clr_a);
    }
    return totals;
}

template <int FMT, bool ALIGNED, int WC>
__global__ void __launch_bounds__(32) raster_bins(const EdgeRec *__restrict__ E, const JobDesc *__restrict__ jobs, const JobState *__restrict__ JS, Params P,
                                                  const uint32_t *__restrict__ bin_off, const uint32_t *__restrict__ entries,
                                                  const Counters *__restrict__ C, uint32_t *__restrict__ ticket, uint32_t *__restrict__ look,
                                                  uint32_t epoch) {
    if (C->overflow || C->n_big == 0) return;
    typedef BinTile<WC> T;
    extern __shared__ __align__(16) uint8_t bin_smem[];
    const uint32_t lane = threadIdx.x;
    uint32_t cells = smem_addr(bin_smem);
    asm volatile("" : "+r"(cells));  // opaque: keep the shared-window base in a register instead of rebuilding it at every access
    const uint32_t stage = cells + T::STAGE, totbase = cells + T::TOT, prefix = cells + T::PREFIX, flags = cells + T::FLAGS;
    uint32_t sp = stage;
    for (uint32_t i = lane; i < T::CELL_BYTES / 16; i += 32) ssts4_bias(cells + 16u * i);
    ssts(totbase + 4u * lane, 0u);
    if (lane == 0) ssts(flags, 0u);
    __syncwarp();
    const int32_t W = (int32_t)P.W;
    const uint32_t tiles_per_win = (P.job_end - P.job_begin) * P.b_nbands;
    const uint32_t n_tasks = P.b_lookback ? tiles_per_win * P.b_nwin : tiles_per_win;
    const uint32_t rbase = cells + lane * T::ROW_BYTES;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= n_tasks) break;
        uint32_t w_begin = 0, w_end = P.b_nwin, jb = t;
        if (P.b_lookback) {  // window-major: the left neighbour of a tile always has a smaller ticket
            w_begin = t / tiles_per_win;
            jb = t - w_begin * tiles_per_win;
            w_end = w_begin + 1;
        }
        const uint32_t j = P.job_begin + jb / P.b_nbands, band = jb % P.b_nbands;
        const JobState js = JS[j];
        if (js.vtx_end - js.vtx_begin <= P.direct_max) continue;  // drawn by raster_tiles from the job's own edge range
        const int32_t row0 = (int32_t)P.row_begin + (int32_t)(band << BIN_LOG2R);
        const int32_t row = row0 + (int32_t)lane;
        const bool row_ok = row >= js.first_row && row < (int32_t)P.row_end;  // rows above the figure are untouched (fig.rs:497)
        const uint32_t valid_mask = __ballot_sync(0xFFFFFFFFu, row_ok);
        if (valid_mask == 0) continue;
        const bool band_full = valid_mask == 0xFFFFFFFFu;
        const int32_t v_lo = __ffs((int)valid_mask) - 1, v_hi = 31 - __clz((int)valid_mask);  // the drawn rows are one contiguous range
        const unsigned long long raster = jobs[j].raster;
        const uint32_t rule = jobs[j].rule, color = jobs[j].color;
        uint8_t *dst = reinterpret_cast<uint8_t *>(raster) + (size_t)(row0 - (int32_t)P.row_begin) * P.pitch;
        int32_t carry = 0;
        for (uint32_t w = w_begin; w < w_end; w++) {
            const uint32_t bin = (j * P.b_nbands + band) * P.b_nwin + w;
            const uint32_t e0 = bin_off[bin], ne = bin_off[bin + 1] - e0;
            const int32_t win_lo = (int32_t)(w * (uint32_t)WC), win_hi = min(W, win_lo + WC);
            int32_t tot = 0;
            bool row_touched = false, any_s = false;
            // ---- (c) scatter: the warp walks the bin, 32 records per round, the next round's gather in flight ----
            EdgeRec nxt;
            nxt.flags = 0;
            nxt.ry0 = 0;
            nxt.ry1 = -1;
            if (lane < ne) nxt = E[entries[e0 + lane]];
            for (uint32_t base = 0; base < ne; base += 32) {
                const EdgeRec e = nxt;
                const bool have = base + lane < ne;
                if (base + 32 + lane < ne) nxt = E[entries[e0 + base + 32 + lane]];
                // everything that does not depend on the row, once per (edge, band)
                const int32_t r0 = e.ry0 - row0, r1 = e.ry1 - row0;
                const int32_t nrows = have ? max(min(r1, v_hi) - max(r0, v_lo) + 1, 0) : 0;
                if (have) {
                    const fx_t fr0 = (fx_t)(e.fr & 0xFFFFu), fr1 = (fx_t)(e.fr >> 16);
                    const bool full = band_full && r0 < 0 && r1 >= (int32_t)BIN_ROWS;
                    int4 a, b;
                    a.x = (int32_t)((uint32_t)e.x_bot0 + (uint32_t)(row0 - e.ry0) * (uint32_t)e.inv_slope);
                    a.y = e.inv_slope;
                    a.z = e.step_pix > 0 ? e.step_pix : FX_ONE;
                    a.w = (max(r0, -1) + 1) | (min(r1, (int32_t)BIN_ROWS) << 8);  // rows relative to the band, clamped just outside it
                    b.x = (e.flags & 2u) ? -1 : 1;
                    b.y = fx_mul(e.inv_slope, FX_ONE - fr0);
                    b.z = fx_mul(e.inv_slope, (FX_ONE - fr1) & FX_MASK);
                    b.w = (int32_t)((uint32_t)pixel_cov(fr0) | ((uint32_t)pixel_cov(fr1) << 9) | (full ? 1u << 19 : 0u));
                    ssts4(stage + lane * 32u, a);
                    ssts4(stage + lane * 32u + 16u, b);
                }
                // items (rows inside the band) of this round: short edges go lanes = items, tall ones lanes = rows
                int32_t incl = nrows;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) scan_step<32>(incl, d);
                const int32_t n_items = __shfl_sync(0xFFFFFFFFu, incl, 31);
                const uint32_t cnt = min(32u, ne - base);
                const bool packed = ne < BIN_PACKED_MAX && (uint32_t)n_items < 12u * cnt;
                if (packed) ssts(prefix + 4u * lane, (uint32_t)incl);
                __syncwarp();
                if (packed) {
                    if (nrows > 0) sred_or(flags, ((2u << (nrows - 1)) - 1u) << max(r0, v_lo));
                    any_s = true;
#pragma unroll 1
                    for (int32_t q = (int32_t)lane; q < n_items; q += 32) {
                        uint32_t k = 0;  // the staged edge owning item q: first k with incl[k] > q
#pragma unroll
                        for (uint32_t h = 16; h > 0; h >>= 1)
                            if ((int32_t)slds(prefix + 4u * (k + h - 1)) <= q) k += h;
                        const int32_t before = k ? (int32_t)slds(prefix + 4u * (k - 1)) : 0;
                        const int4 a = slds4(stage + k * 32u), b = slds4(stage + k * 32u + 16u);
                        const int32_t er0 = (a.w & 0xFF) - 1, er1 = a.w >> 8;
                        bin_item_packed<true>(a.x, a.y, a.z, er0, er1, b.y, b.z, (uint32_t)b.w, b.x, max(er0, v_lo) + (q - before), W, win_lo, win_hi, cells, T::ROW_BYTES, totbase);
                    }
                } else {
                    row_touched = true;
#pragma unroll 1
                    for (uint32_t k = 0; k < cnt; k++, sp += 32u) {
                        const int4 a = slds4(sp), b = slds4(sp + 16u);
                        if (b.w & (1 << 19))
                            bin_item_rows<true>(a.x, a.y, a.z, 0, 0, b.y, b.z, (uint32_t)b.w, b.x, (int32_t)lane, true, W, win_lo, win_hi, rbase, tot);
                        else
                            bin_item_rows<false>(a.x, a.y, a.z, (a.w & 0xFF) - 1, a.w >> 8, b.y, b.z, (uint32_t)b.w, b.x, (int32_t)lane, row_ok, W, win_lo, win_hi, rbase, tot);
                    }
                    sp = stage;
                }
                __syncwarp();
            }
            uint32_t touched = __ballot_sync(0xFFFFFFFFu, row_touched && row_ok);
            if (any_s) {  // totals and touched rows of the lanes = edges rounds
                __syncwarp();
                tot += (int32_t)slds(totbase + 4u * lane);
                touched |= slds(flags);
                __syncwarp();
                ssts(totbase + 4u * lane, 0u);
                if (lane == 0) ssts(flags, 0u);
            }
            // ---- the sums reaching this window / leaving it ----
            if (P.b_lookback) {
                if (w > 0) {
                    const uint32_t *src = look + (size_t)(bin - 1) * 32u + lane;
                    uint32_t v = ld_relaxed(src);
                    while (!__all_sync(0xFFFFFFFFu, (v >> 16) == epoch)) {
                        __nanosleep(64);
                        v = ld_relaxed(src);
                    }
                    carry = (int32_t)(v & 0xFFFFu);
                }
                if (w + 1 < P.b_nwin) st_relaxed(look + (size_t)bin * 32u + lane, (epoch << 16) | ((uint32_t)(carry + tot) & 0xFFFFu));
            }
            // ---- (d) resolve ----
            uint8_t *dwin = dst + (size_t)win_lo * P.bpp;
            if (rule == FTL_EVENODD) bin_resolve<FMT, true, ALIGNED, WC>(cells, touched, carry, valid_mask, dwin, P.pitch, (uint32_t)(W - win_lo), color);
            else bin_resolve<FMT, false, ALIGNED, WC>(cells, touched, carry, valid_mask, dwin, P.pitch, (uint32_t)(W - win_lo), color);
            __syncwarp();
            carry += tot;
        }
    }
}

// Parity probe of stage (c) alone: the signed-area deltas one raster row receives from the edges of a job, before any
// prefix sum (what Edge::scan_area adds into the i16 area buffer, fig.rs:285-302).  One thread per edge, the same
// closed-form span as the scatter (bin_span), global atomics into an i32 row that the host truncates to the reference's i16.
__global__ void __launch_bounds__(256) area_row_probe(const EdgeRec *__restrict__ E, uint32_t e_begin, uint32_t e_end, int32_t row, int32_t W,
                                                      int32_t *__restrict__ area) {
    const uint32_t k = e_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= e_end) return;
    const EdgeRec e = E[k];
    if (!(e.flags & 1u) || row < e.ry0 || row > e.ry1) return;
    const fx_t fr0 = (fx_t)(e.fr & 0xFFFFu), fr1 = (fx_t)(e.fr >> 16);
    const int32_t step = e.step_pix > 0 ? e.step_pix : FX_ONE;
    const fx_t x_bot = (fx_t)((uint32_t)e.x_bot0 + (uint32_t)(row - e.ry0) * (uint32_t)e.inv_slope);
    const uint32_t misc = (uint32_t)pixel_cov(fr0) | ((uint32_t)pixel_cov(fr1) << 9);
    int32_t c, xc, prev, cov;
    bool resumed;
    if (!bin_span<false, false>(x_bot, e.inv_slope, step, row == e.ry0, row == e.ry1, fx_mul(e.inv_slope, FX_ONE - fr0),
                                fx_mul(e.inv_slope, (FX_ONE - fr1) & FX_MASK), misc, true, W, 0, W, c, xc, prev, cov, resumed))
        return;
    const int32_t ed = (e.flags & 2u) ? -1 : 1;
    int32_t xr = xc + 128;
    for (;;) {
        int32_t xk = xr >> 8;
        if (xk > cov) xk = cov;
        atomicAdd(&area[c], ed * (xk - prev));
        prev = xk;
        c++;
        xr = min(xr + step, FX_ONE + 128);
        if (xk >= cov || c >= W) break;
    }
}
