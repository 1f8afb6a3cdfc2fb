// tile_kernel.cuh — stages (c)+(d): the tile raster kernel (scatter rows, analytic rows, resolve, composite)
// Included by engine.cu inside namespace ftl (one translation unit: the kernels share Params / EdgeRec / ...).
#pragma once

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
#pragma unroll 1
                for (uint32_t i = 0; i < 4; i++)  // ragged end of a row: rolled, the general blend is large
                    if (x + 4 * j + i < W) d[4 * j + i] = blend_rgba(d[4 * j + i], color, (w >> (8 * i)) & 0xFF, clr_a);
            }
        }
    } else {  // Graya8p
        uint16_t *d = reinterpret_cast<uint16_t *>(dst) + x;
        if (ALIGNED && x + 16 <= W) {
            // two 16-byte words of 8 pixels each: opaque full coverage is written unread, alpha 0 goes through
            // the changed-byte test, anything else is blended in registers (one load and one store per 8 pixels)
            uint4 *q4 = reinterpret_cast<uint4 *>(d);
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint32_t wl = h == 0 ? a0 : a2, wh = h == 0 ? a1 : a3;  // alphas of pixels 8h .. 8h+7
                if ((wl & wh) == 0xFFFFFFFFu && clr_a == 255) {
                    const uint32_t c = mul255_x4((color & 0xFFFFu) * 0x00010001u);
                    q4[h] = make_uint4(c, c, c, c);
                    continue;
                }
                uint4 t = q4[h];
                if ((wl | wh) == 0) {
                    mul255_rmw(q4 + h, t);
                    continue;
                }
                uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {  // word k holds pixels 2k and 2k+1 of this half
                    const uint32_t aw = k < 2 ? wl : wh;
                    uint32_t o = 0;
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const uint32_t al = (aw >> (8 * (2 * (k & 1) + e))) & 0xFF, p = (w[k] >> (16 * e)) & 0xFFFFu;
                        const uint32_t sa1 = 255u - pix::ch8_mul(al, clr_a);
                        o |= (pix::src_over_ch(p & 0xFF, color & 0xFF, al, sa1) | (pix::src_over_ch(p >> 8, (color >> 8) & 0xFF, al, sa1) << 8)) << (16 * e);
                    }
                    w[k] = o;
                }
                q4[h] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            return;
        }
#pragma unroll 1
        for (uint32_t i = 0; i < 16; i++) {  // ragged end of a row: rolled, the general blend is large
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

// Alpha 0 over a long span, staged through the warp's shared-memory window.  The register version above has four
// 16-byte loads in flight per lane and pays one DRAM round trip per 2 KiB of the span (plus one per leftover 512 bytes):
// the Rgba8p / Graya8p composite of a scene that is mostly background is bound by exactly that latency.  Here every
// lane copies its 16-byte words with cp.async into the window (no registers held), two half-windows in flight, and tests
// them from shared memory: one round trip per half-window (4 KiB for the usual 8 KiB window), the next one already under
// way.  Every lane reads back only what it copied itself, so no warp synchronisation is needed.  The window must be
// re-zeroed before the scatter path uses it again (the caller keeps a `dirty` flag).
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ uint4 slds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void fill_zero_staged(uint8_t *drow, uint32_t u0, uint32_t end, uint32_t stage, uint32_t half_u4) {
    const uint32_t lane = threadIdx.x & 31;
    uint4 *p = reinterpret_cast<uint4 *>(drow);
    const uint32_t nb = (end - u0 + half_u4 - 1u) / half_u4;
    const uint32_t my = stage + lane * 16u;
    auto issue = [&](uint32_t b) {
        const uint32_t sbase = my + (b & 1u) * half_u4 * 16u;
        uint32_t u = u0 + b * half_u4 + lane;
        const uint32_t stop = min(end, u0 + (b + 1u) * half_u4);
#pragma unroll 2
        for (uint32_t k = 0; u < stop; u += 32, k++) cp_async16(sbase + k * 512u, p + u);
        cp_async_commit();
    };
    issue(0);
    if (nb > 1) issue(1);
#pragma unroll 1
    for (uint32_t b = 0; b < nb; b++) {
        if (b + 1 < nb) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        const uint32_t sbase = my + (b & 1u) * half_u4 * 16u;
        uint32_t u = u0 + b * half_u4 + lane;
        const uint32_t stop = min(end, u0 + (b + 1u) * half_u4);
#pragma unroll 2
        for (uint32_t k = 0; u < stop; u += 32, k++) mul255_rmw(p + u, slds_u4(sbase + k * 512u));
        if (b + 2 < nb) issue(b + 2);
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
                                                  bool even_odd, uint32_t color, uint32_t stage, uint32_t half_u4, bool &dirty) {
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
    // ---- the constant spans: left of the first span of every row, then right of every span ----
    // ONE loop and one call site for both kinds (the fill routines are large, and this kernel's hot path must stay inside
    // the instruction cache): iterations 0 .. n_rows - 1 are the leading spans, the rest walk the owners.
    const uint32_t owners = __ballot_sync(0xFFFFFFFFu, active && row_ok);
    if (FMT != FTL_RGBA8P) {  // Matte8 (plain stores) and Graya8p (8 KiB rows: one or two round trips per span anyway): two tight loops; the merged loop below costs them 7-9 %
        // The trailing alpha-0 span of a row and the leading span of the next row are one run of bytes (rows have no
        // padding): the rightmost span of row r runs on into row r + 1, whose leading span is then skipped.
        uint32_t lead_next = 0;
#pragma unroll
        for (int r = 1; r < 4; r++) {
            const uint32_t l = min(__shfl_sync(0xFFFFFFFFu, min_ga, 8 * r), ngroups);
            if ((int32_t)my_r + 1 == r && r < n_rows && !((redo >> r) & 1u)) lead_next = l;
        }
        const bool extend = active && row_ok && next_ga == ngroups && after_a == 0u && lead_next > 0u && ngroups + lead_next < 0x10000u;
        const uint32_t skip = __reduce_or_sync(0xFFFFFFFFu, extend ? 2u << my_r : 0u);
        const uint32_t my_span8 = (gb + 1u) | ((next_ga + (extend ? lead_next : 0u)) << 16);
#pragma unroll 1
        for (uint32_t mm = owners; mm; mm &= mm - 1) {
            const uint32_t s = (uint32_t)__ffs((int)mm) - 1u;
            const uint32_t sp = __shfl_sync(0xFFFFFFFFu, my_span8, s), q = __shfl_sync(0xFFFFFFFFu, after_a, s);
            fill_const<FMT>(dst + (size_t)(s >> 3) * pitch, sp & 0xFFFFu, sp >> 16, q, color, clr_a);
        }
#pragma unroll 1
        for (int r = 0; r < n_rows; r++) {
            const uint32_t hi = min(__shfl_sync(0xFFFFFFFFu, min_ga, 8 * r), ngroups);
            if (!(((redo | skip) >> r) & 1u)) fill_const<FMT>(dst + (size_t)r * pitch, 0u, hi, 0u, color, clr_a);
        }
        return redo;
    }
    const uint32_t my_span = (gb + 1u) | (next_ga << 16);
    constexpr uint32_t U = 4u;  // 16-byte words per 16-pixel group of Rgba8p
    uint32_t m = owners;
#pragma unroll 1
    for (int it = 0;; it++) {
        uint32_t row, lo, hi, q;
        if (it < n_rows) {
            row = (uint32_t)it;
            lo = 0u;
            hi = min(__shfl_sync(0xFFFFFFFFu, min_ga, 8 * it), ngroups);
            q = 0u;
            if ((redo >> it) & 1u) continue;
        } else {
            if (!m) break;
            const uint32_t s = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1;
            const uint32_t sp = __shfl_sync(0xFFFFFFFFu, my_span, s);
            q = __shfl_sync(0xFFFFFFFFu, after_a, s);
            row = s >> 3;
            lo = sp & 0xFFFFu;
            hi = sp >> 16;
        }
        uint8_t *d = dst + (size_t)row * pitch;
        if (q == 0u && half_u4 && hi > lo && (hi - lo) * U > 128u) {
            fill_zero_staged(d, lo * U, hi * U, stage, half_u4);
            dirty = true;
        } else fill_const<FMT>(d, lo, hi, q, color, clr_a);
    }
    return redo;
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// The direct tile kernel: jobs with at most Params::direct_max (8) edge slots (polygons, stars, simple layers of a scene).
// Every WARP owns a private shared-memory row window (`win_chunks` chunks of 512 cells + their masks) and
// walks the rows of a (job, band) tile on its own: its lanes scatter the coverage of the job's edges
// crossing the row, then the row is resolved and written, window after window, with the running sum
// carried across windows.  Warps never wait for each other, and the small window keeps many warps
// resident per SM, which is what hides the latency of the serial scatter -> scan -> store chain.
// Jobs with more edges are drawn by raster_bins (bin_kernel.cuh).
template <int FMT, bool ALIGNED>
__global__ void __launch_bounds__(128, 5) raster_tiles(const EdgeRec *__restrict__ E, const JobDesc *__restrict__ jobs,
                                                       const JobState *__restrict__ JS, Params P, const Counters *__restrict__ C) {
    if (C->overflow || C->n_big == P.n_jobs) return;  // every job goes to raster_bins: nothing to scan for
    extern __shared__ __align__(16) int32_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
    const uint32_t cells = smem_addr(smem) + 4u * warp * P.warp_words, mask = cells + 4u * P.win_rows * P.win_chunks * CHUNK;
    for (uint32_t i = lane; i < P.warp_words; i += 32) ssts(cells + 4u * i, 0u);
    __syncwarp();
    const int32_t W = (int32_t)P.W, win_cells = (int32_t)(P.win_chunks * CHUNK);
    const uint32_t bpp = P.bpp;
    const uint32_t n_warps = gridDim.x * warps_per_cta;
    // analytic rows need full 16-pixel groups on 16-byte boundaries (and group indices below 0xFFFF)
    const bool analytic_ok = ALIGNED && (P.W & 15u) == 0 && (P.W >> 4) < 0xFFFFu;
    // the analytic rows use the (otherwise idle) cell window as a staging area for long alpha-0 spans: two halves of a
    // power-of-two number of 16-byte words; `dirty`: the window no longer holds the zeros the scatter path relies on
    const uint32_t win_u4 = (P.win_rows * P.win_chunks * CHUNK) / 4u;
    const uint32_t half_u4 = FMT != FTL_RGBA8P || win_u4 < 256u ? 0u : (1u << (31 - __clz(win_u4))) / 2u;
    bool dirty = false;
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
        const uint32_t ne = js.vtx_end - js.vtx_begin, e0 = js.vtx_begin;
        if (ne > P.direct_max) continue;  // a binned job: raster_bins draws it
        const unsigned long long raster = jobs[j].raster;
        const uint32_t rule = jobs[j].rule, color = jobs[j].color;
        // Row groups: with few edges the 32 lanes are (row, edge) pairs of 4 (or 2) consecutive rows, so
        // the per-(edge,row) set-up of several rows costs one pass; each row then scatters with its own lanes.
        const uint32_t gl = ne <= 8 ? 3u : (ne <= 16 ? 4u : 5u);
        const uint32_t my_e = lane & ((1u << gl) - 1u), my_r = lane >> gl;
        // Narrow rasters (one window per row) with lanes = edges: the window holds `win_rows` rows, every
        // lane scatters all the rows of its edges in one pass, then the rows are resolved one by one.
        const bool multi = gl == 5u && P.win_rows > 1u;
        const int32_t rows_per_pass = multi ? (int32_t)P.win_rows : (int32_t)(32u >> gl);
        const uint32_t row_bytes = 4u * P.win_chunks * CHUNK, rmask_bytes = 4u * P.win_chunks;
        EdgeRec mine;
        mine.flags = 0;
        if (my_e < ne) mine = E[e0 + my_e];
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
            if (analytic_ok && gl == 3u) redo = analytic_rows<FMT>(st, my_r, rr_end, W, dst, P.pitch, rule == FTL_EVENODD, color, cells, half_u4, dirty);
            if (dirty && redo) {  // rows for the scatter path follow: give it back a window of zeros
                for (uint32_t i = lane; i < win_u4; i += 32) ssts4_zero(cells + 16u * i);
                __syncwarp();
                dirty = false;
            }
            const int32_t ry_last = ry_base + rr_end - 1;
            for (int32_t rr = 0; rr < rr_end; rr++, dst += P.pitch) {
                if (!((redo >> rr) & 1u)) continue;
                const int32_t ry = ry_base + rr;
                const uint32_t rcells = multi ? cells + (uint32_t)rr * row_bytes : cells, rmask = multi ? mask + (uint32_t)rr * rmask_bytes : mask;
                int32_t carry = 0;
                uint8_t *dwin = dst;
                for (int32_t win_lo = 0; win_lo < W; win_lo += win_cells, dwin += win_bytes) {
                    const int32_t win_hi = min(W, win_lo + win_cells);
                    // ---- (c) scatter: one lane per edge crossing this row (multi: all rows of the pass at once) ----
                    if (!multi || rr == 0) {
                        uint32_t first = 32;
                        if (multi) first = 0;
                        else if ((int32_t)my_r == rr) edge_row_scatter(st, win_lo, win_hi, cells, mask);
                        const int32_t ra = multi ? ry_base : ry, rb = multi ? ry_last : ry;
                        for (uint32_t i = lane + first; i < ne; i += 32) {
                            const EdgeRec e = (multi && i < 32) ? mine : E[e0 + i];
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
