// libm_compat.cuh — the three libm calls of the stroker (stroker.rs:305,354-355,389,407: `normalize()` = hypot,
// `angle_rel()` = atan2, `(th / 2).sin()`), restated so that the DEVICE returns the bits the host's glibc returns.
// Rust's f32::{hypot, atan2, sin} lower to the platform's hypotf / atan2f / sinf, glibc 2.39 here.
//
//   hypotf : glibc (>= 2.35) evaluates sqrt(x*x + y*y) in double and rounds once — every step is a correctly rounded
//            IEEE operation, so the device repeats it exactly.
//   atan2f : glibc's float atan2 / atan are the fdlibm float routines (e_atan2f.c, s_atanf.c): plain IEEE f32 adds,
//            multiplies and divisions, no FMA.  Restated below; -fmad=false / -ffp-contract=off keep the op order.
//   sinf   : glibc's sinf is a double-precision polynomial with a CPU-dependent (FMA) variant; the stroker only
//            COMPARES |sin(th / 2)| with two thresholds (stroker.rs:355-356), so the device evaluates sin in double and
//            reports "undecided" when the value is too close to a threshold for a < 1 ulp float sinf to be predicted;
//            the caller then takes the host stroker for that call (stroke_kernels.cuh).
//
// ftl_debug_libm_selftest (capi.cpp) compares hypotf_glibc / atan2f_glibc with the host's libm on random inputs: that is
// the pin of this restatement (tests/test_host.py::test_libm_restatement_matches_glibc), on this image's glibc.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "fixed.cuh"

namespace ftl {
namespace libm {

FTL_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

FTL_HD float hypotf_glibc(float x, float y) {
    const double dx = (double)x, dy = (double)y;
    return (float)sqrt(dx * dx + dy * dy);  // the products are exact in double; one rounding in the sum, one in sqrt, one to float
}

// fdlibm s_atanf.c
FTL_HD float atanf_glibc(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                          6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    const int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {  // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {  // |x| < 0.4375
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {  // |x| < 1.1875
            if (ix < 0x3f300000) {
                id = 0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                id = 1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else if (ix < 0x401c0000) {  // |x| < 2.4375
            id = 2;
            x = (x - 1.5f) / (1.0f + 1.5f * x);
        } else {
            id = 3;
            x = -1.0f / x;
        }
    }
    const float z = x * x, w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return hx < 0 ? -r : r;
}

// fdlibm e_atan2f.c
FTL_HD float atan2f_glibc(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff, hy = (int32_t)f2u(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_glibc(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) return m == 0 ? pi_o_4 + tiny : (m == 1 ? -pi_o_4 - tiny : (m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny));
        return m == 0 ? 0.0f : (m == 1 ? -0.0f : (m == 2 ? pi + tiny : -pi - tiny));
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_glibc(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

// Three-valued comparisons of |sinf(x)| (x = th / 2, |x| <= pi/2 + a few ulp) with the miter thresholds of
// stroker.rs:355-356: 1 = certainly true, 0 = certainly false, -1 = too close to call.
//   sm >= sm_min : glibc's sinf is within one ulp (6e-8) of the true value; the margin is 1.5e-7.
//   sm < 1.0     : decided by joins of almost collinear segments (every flattened curve has some), where sinf returns
//                  1.0f exactly when its double-precision result rounds up, i.e. reaches 1 - 2^-25.  Next to pi/2 that
//                  result is 1 - d*d/2 to within a few 1e-16 (the cosine polynomial's leading coefficients are 1 and
//                  -0.5 * (1 - 2.8e-9)), so the margin here is 4e-15: in practice never hit.
// ftl_debug_libm_selftest checks both predictions against the host's sinf, the second one over EVERY float within
// 2e-3 of +-pi/2.
FTL_HD int abs_sin_ge(float x, float sm_min) {
    const double s = fabs(sin((double)x)), t = (double)sm_min;
    if (s >= t + 1.5e-7) return 1;
    if (s <= t - 1.5e-7) return 0;
    return -1;
}
FTL_HD int abs_sin_lt_one(float x) {
    const double s = fabs(sin((double)x)), edge = 1.0 - 2.98023223876953125e-8;
    if (s <= edge - 4.0e-15) return 1;
    if (s >= edge + 4.0e-15) return 0;
    return -1;
}

}  // namespace libm
}  // namespace ftl
