// tools/ubench/cvt_exact.cuh (experiment, not used by the product) -- float <-> double conversions on the INTEGER pipe, bit-identical to the hardware conversions
// (cvt.f64.f32 and cvt.rn.f32.f64).
//
// Why: the recursive-Gaussian consumer (canvas_kernels.cu) converts every sample float -> double on load and double ->
// float on store.  On B200 those are F2F instructions, which issue at about one warp instruction per 10 cycles on a
// unit shared by the whole SM (tools/ubench/f2f_lat.cu): with c consumer warps resident on an SM the kernel runs at
// 20 c cycles per sample (measured 80 at 4 CTAs/SM, 160 at 8) although the dependent fp64 chain is only 32.  The
// integer ALUs are per sub-partition and idle in that kernel, so the common cases are done there:
//   f32 -> f64: zero and normal numbers by re-biasing the exponent and shifting the mantissa;
//   f64 -> f32: results that are normal floats (round to nearest even on the 29 dropped bits; a carry out of the
//               mantissa correctly bumps the exponent, up to infinity) and magnitudes below 2^-150 (signed zero).
// Everything else (float denormals in or out, infinities, NaN) takes the hardware instruction on a rare branch.
// tools/ubench/cvt_exact.cu checks both against the hardware over all 2^32 floats and 2^34 structured doubles.
#pragma once
#include <cuda_runtime.h>

namespace pb {

// Branch-free: the fast result is always computed and the hardware conversion is issued under a predicate that is
// false in the common case (a predicated-off F2F costs an issue slot, not the shared conversion unit).  Branches here
// would stop the compiler from overlapping the conversions of one sample with the fp64 chain of another.
__device__ __forceinline__ double f2d_int(float f) {
    const unsigned u = __float_as_uint(f);
    const unsigned a = u & 0x7fffffffu;
    const bool normal = a - 0x00800000u < 0x7f000000u;   // 0x00800000 <= a < 0x7f800000
    const unsigned hi = (u & 0x80000000u) | (normal ? (a >> 3) + 0x38000000u : 0u);   // zero: sign only
    double r = __hiloint2double((int)hi, (int)(u << 29));
    const unsigned slow = (!normal && a != 0u) ? 1u : 0u;   // denormal, infinity, NaN
#ifndef PB_CVT_NO_FALLBACK
    asm("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p cvt.f64.f32 %0, %1; }" : "+d"(r) : "f"(f), "r"(slow));
#endif
    return r;
}

__device__ __forceinline__ float d2f_int(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned ah = hi & 0x7fffffffu;
    const bool normal = ah - 0x38100000u < 0x0fe00000u;   // 2^-126 <= |v| < 2^128: normal float, or rounds up to infinity
    unsigned t = __funnelshift_l(lo, ah - 0x38000000u, 3);   // exponent re-biased, top 23 mantissa bits
    const unsigned rem = lo & 0x1fffffffu;                   // the 29 dropped bits; half = 0x10000000
    t += (rem + (t & 1u)) > 0x10000000u ? 1u : 0u;           // round to nearest, ties to even (a carry bumps the exponent)
    float r = __uint_as_float((hi & 0x80000000u) | (normal ? t : 0u));   // |v| < 2^-150 rounds to signed zero
    const unsigned slow = (!normal && ah >= 0x36900000u) ? 1u : 0u;      // float-denormal results, >= 2^128, infinity, NaN
#ifndef PB_CVT_NO_FALLBACK
    asm("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p cvt.rn.f32.f64 %0, %1; }" : "+f"(r) : "d"(v), "r"(slow));
#endif
    return r;
}

}  // namespace pb
