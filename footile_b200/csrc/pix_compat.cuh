// pix_compat.cuh — the Ch8 channel arithmetic and SrcOver compositing the hot
// path takes from the `pix 0.14` crate (Cargo.toml:15; call sites
// fig.rs:641-642,662-663).  Source absent from the reference tree: RECALLED
// semantics, isolated here.  The only reference test at
// this boundary is fig_3x3 (fig.rs:702-721), the alpha=255-over-clear case.
#pragma once
#include "fixed.cuh"

namespace ftl {
namespace pix {

// Ch8 * Ch8: both operands widened to 12 bits by bit replication.
FTL_HD uint32_t ch8_mul(uint32_t a, uint32_t b) {
    uint32_t l = (a << 4) | (a >> 4);
    uint32_t r = (b << 4) | (b >> 4);
    return (l * r) >> 16;
}
// Ch8 + Ch8 saturates.
FTL_HD uint32_t ch8_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return s > 255u ? 255u : s;
}
// Ch8 / Ch8 (the Premultiplied -> Straight alpha decode of a conversion): (c << 8) / a, at most 255; 0 for a = 0.
FTL_HD uint32_t ch8_div(uint32_t c, uint32_t a) {
    if (a == 0u) return 0u;
    const uint32_t q = (c << 8) / a;
    return q > 255u ? 255u : q;
}
// One channel of dst.composite_channels_alpha(&src, SrcOver, &alpha):
//   d' = (s * alpha) + d * (255 - alpha * src_alpha)
FTL_HD uint32_t src_over_ch(uint32_t d, uint32_t s, uint32_t alpha, uint32_t sa1) {
    return ch8_add(ch8_mul(s, alpha), ch8_mul(d, sa1));
}

}  // namespace pix
}  // namespace ftl
