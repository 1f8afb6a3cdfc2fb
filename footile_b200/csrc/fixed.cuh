// fixed.cuh — 16.16 fixed point, bit-compatible with footile's `Fixed`
// (reference: src/fixed.rs:10-159).  Every i32 op wraps, as Rust release
// builds do; (i64) results are truncated to their low 32 bits.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FTL_HD __host__ __device__ __forceinline__
#else
#define FTL_HD inline
#endif

namespace ftl {

typedef int32_t fx_t;
constexpr fx_t FX_ONE = 1 << 16;
constexpr fx_t FX_HALF = 1 << 15;
constexpr fx_t FX_MASK = (1 << 16) - 1;

FTL_HD fx_t fx_add(fx_t a, fx_t b) { return (fx_t)((uint32_t)a + (uint32_t)b); }            // fixed.rs:24-30
FTL_HD fx_t fx_sub(fx_t a, fx_t b) { return (fx_t)((uint32_t)a - (uint32_t)b); }            // fixed.rs:32-38
FTL_HD fx_t fx_mul(fx_t a, fx_t b) { return (fx_t)(((int64_t)a * (int64_t)b) >> 16); }      // fixed.rs:40-47
FTL_HD fx_t fx_div(fx_t a, fx_t b) {                                                        // fixed.rs:49-56
#if defined(__CUDA_ARCH__)
    // The 48-by-32-bit truncating division through one double-precision division rounded towards zero: the numerator
    // (|a| << 16 < 2^47) and the divisor are exact doubles, the integer part of the true quotient (< 2^48) is itself a
    // double not above it in magnitude, so the quotient rounded towards zero lies between the two and truncates to exactly
    // that integer part.  About a third of the instructions of the 64-bit integer division it replaces.
    return (fx_t)__double2ll_rz(__ddiv_rz((double)a * 65536.0, (double)b));
#else
    return (fx_t)((int64_t)((uint64_t)(int64_t)a << 16) / (int64_t)b);
#endif
}
FTL_HD int32_t fx_to_i32(fx_t a) { return a >> 16; }                                        // fixed.rs:81-86
FTL_HD fx_t fx_abs(fx_t a) { return a < 0 ? (fx_t)(0u - (uint32_t)a) : a; }                 // fixed.rs:122-124
FTL_HD fx_t fx_floor(fx_t a) { return a & ~FX_MASK; }                                       // fixed.rs:127-129
FTL_HD fx_t fx_ceil(fx_t a) { return fx_floor(fx_sub(fx_add(a, FX_ONE), 1)); }              // fixed.rs:132-134
FTL_HD fx_t fx_fract(fx_t a) { return a & FX_MASK; }                                        // fixed.rs:151-153
FTL_HD fx_t fx_avg(fx_t a, fx_t b) { return fx_add(a, b) >> 1; }                            // fixed.rs:156-158
FTL_HD fx_t fx_min(fx_t a, fx_t b) { return a < b ? a : b; }
FTL_HD fx_t fx_max(fx_t a, fx_t b) { return a > b ? a : b; }

// Fixed::from(f32) = (f * 65536.0) as i32 — truncate toward zero, saturate,
// NaN -> 0 (fixed.rs:88-93).  On the device cvt.rzi.s32.f32 has exactly these
// semantics.
FTL_HD fx_t fx_from_f32(float f) {
    float v = f * 65536.0f;
#if defined(__CUDA_ARCH__)
    return __float2int_rz(v);
#else
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (fx_t)v;
#endif
}

// pixel_cov(fcov) = round(fcov * 256) as an integer 0..256 (fig.rs:677-682)
FTL_HD int32_t pixel_cov(fx_t fcov) { return (fx_t)(((uint32_t)fcov << 8) + (uint32_t)FX_HALF) >> 16; }

}  // namespace ftl
