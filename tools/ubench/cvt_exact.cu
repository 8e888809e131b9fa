// exhaustive / structured check of csrc/cvt_exact.cuh against the hardware conversions
#include "cvt_exact.cuh"
#include <cstdio>
using namespace pb;

__device__ unsigned long long g_bad[4];

__device__ __forceinline__ bool same_d(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }
__device__ __forceinline__ bool same_f(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }

__global__ void check_f2d() {   // all 2^32 float patterns
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
        const float f = __uint_as_float((unsigned)i);
        const double a = f2d_int(f), b = (double)f;
        if (!same_d(a, b) && !(b != b)) atomicAdd(&g_bad[0], 1ull);
    }
}
__global__ void check_d2f() {   // around every float: the value, both neighbours' midpoints, +-1 and +-2 double ulps of each
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned u = (unsigned)i;
        if ((u & 0x7f800000u) == 0x7f800000u) continue;
        const double d0 = (double)__uint_as_float(u);
        const double d1 = (double)__uint_as_float(u + 1);   // next float up in magnitude (may be inf)
        const double mid = 0.5 * d0 + 0.5 * d1;              // exact for finite d1
        const double base[2] = {d0, mid};
        for (int k = 0; k < 2; ++k) {
            const long long bits = __double_as_longlong(base[k]);
            for (int off = -2; off <= 2; ++off) {
                const double v = __longlong_as_double(bits + off);
                if (!same_f(d2f_int(v), (float)v)) atomicAdd(&g_bad[1], 1ull);
            }
        }
    }
}
__global__ void check_d2f_random(unsigned long long seed) {   // 2^32 pseudo-random 64-bit patterns + exponent sweep
    unsigned long long x = seed + (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    for (int it = 0; it < 4096; ++it) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const double v = __longlong_as_double((long long)x);
        if (!same_f(d2f_int(v), (float)v)) atomicAdd(&g_bad[2], 1ull);
        // same mantissa, exponent forced into the interesting band [2^-160, 2^-120] and near 2^128
        const unsigned long long m = x & 0x800fffffffffffffull;
        const unsigned e1 = 1023 - 160 + (unsigned)(it % 41), e2 = 1023 + 120 + (unsigned)(it % 10);
        const double v1 = __longlong_as_double((long long)(m | ((unsigned long long)e1 << 52)));
        const double v2 = __longlong_as_double((long long)(m | ((unsigned long long)e2 << 52)));
        if (!same_f(d2f_int(v1), (float)v1)) atomicAdd(&g_bad[3], 1ull);
        if (!same_f(d2f_int(v2), (float)v2)) atomicAdd(&g_bad[3], 1ull);
    }
}
int main() {
    unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_bad, z, sizeof z);
    check_f2d<<<148 * 8, 256>>>();
    check_d2f<<<148 * 8, 256>>>();
    check_d2f_random<<<1024, 1024>>>(12345);
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(z, g_bad, sizeof z);
    printf("f2d mismatches over all floats: %llu\nd2f mismatches around all floats (10 doubles each): %llu\n"
           "d2f mismatches over 2^32 random doubles: %llu\nd2f mismatches in the denormal / overflow exponent bands: %llu\n(%s)\n",
           z[0], z[1], z[2], z[3], cudaGetErrorString(cudaGetLastError()));
    return (z[0] | z[1] | z[2] | z[3]) ? 1 : 0;
}
