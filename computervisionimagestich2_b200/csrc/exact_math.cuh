// exact_math.cuh -- bit-reproducible scalar arithmetic shared by the sm_100a kernels.
//
// Every function here restates one small arithmetic helper of the reference so that the CUDA path produces the
// same bits as the CPU reference.  Rules (DESIGN.md "bit-exact policy"):
//   * the library is compiled with -fmad=false: no FMA contraction anywhere in parity-critical code;
//   * IEEE division / sqrt (nvcc defaults, no -use_fast_math);
//   * transcendentals (exp, pow, sin, cos, tan) are NEVER evaluated on the device: the host evaluates them with
//     glibc and uploads tables / per-item values;
//   * mixed float/double expressions keep the reference's promotion order, spelled out with explicit casts.
//
// The functions are __host__ __device__ so that tests/emul (a test-only g++ build of the same header) can check
// them against the compiled reference on a CPU-only box.  The product never runs them on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace pb {

constexpr double kPi = 3.141592653589793;      // vl/mathop.h:28
constexpr float kEpsF = 1.19209290E-07F;       // vl/mathop.h:37
constexpr double kEpsD = 2.220446049250313e-16; // vl/mathop.h:45

PB_HD float bits_to_float(int32_t i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    union { float f; int32_t i; } u; u.i = i; return u.f;
#endif
}
PB_HD int32_t float_to_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    union { float f; int32_t i; } u; u.f = f; return u.i;
#endif
}

// vl/mathop.h:109-115
PB_HD float mod_2pi_f(float x) {
    const float twopi = (float)(2 * kPi);
    while (x > twopi) x -= twopi;
    while (x < 0.0F) x += twopi;
    return x;
}

// vl/mathop.h:134-152 (long int in the reference; the values met on this path fit an int)
PB_HD int floor_f(float x) {
    int xi = (int)x;
    if (x >= 0 || (float)xi == x) return xi;
    return xi - 1;
}
PB_HD int floor_d(double x) {
    int xi = (int)x;
    if (x >= 0 || (double)xi == x) return xi;
    return xi - 1;
}

PB_HD double abs_d(double x) {
#if defined(__CUDA_ARCH__)
    return fabs(x);
#else
    return __builtin_fabs(x);
#endif
}
PB_HD float fabs_f(float x) {
#if defined(__CUDA_ARCH__)
    return fabsf(x);
#else
    return __builtin_fabsf(x);
#endif
}

// vl/mathop.h:407-425
PB_HD float fast_atan2_f(float y, float x) {
    float angle, r;
    const float c3 = 0.1821F;
    const float c1 = 0.9675F;
    float abs_y = fabs_f(y) + kEpsF;
    if (x >= 0) {
        r = (x - abs_y) / (x + abs_y);
        angle = (float)(kPi / 4);
    } else {
        r = (x + abs_y) / (abs_y - x);
        angle = (float)(3 * kPi / 4);
    }
    angle += (c3 * r * r - c1) * r;
    return (y < 0) ? -angle : angle;
}

// vl/mathop.h:479-500
PB_HD float fast_resqrt_f(float x) {
    float xhalf = (float)0.5 * x;
    int32_t i = float_to_bits(x);
    i = 0x5f3759df - (i >> 1);
    float u = bits_to_float(i);
    u = u * ((float)1.5 - xhalf * u * u);
    u = u * ((float)1.5 - xhalf * u * u);
    return u;
}

// vl/mathop.h:544-548 (the comparison is against the double literal 1e-8)
PB_HD float fast_sqrt_f(float x) { return ((double)x < 1e-8) ? 0.0f : x * fast_resqrt_f(x); }

// vl/sift.c:35-49; expn_tab[257] is built on the host with glibc exp (vl/sift.c:56-63) and uploaded.
PB_HD double fast_expn(const double* __restrict__ tab, double x) {
    if (x > 25.0) return 0.0;
    x *= 256 / 25.0;
    int i = floor_d(x);
    double r = x - i;
    double a = tab[i];
    double b = tab[i + 1];
    return a + r * (b - a);
}

}  // namespace pb
