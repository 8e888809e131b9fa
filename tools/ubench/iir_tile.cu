// micro-benchmark: the consumer's tile loop of iir_pipe_kernel alone (no producer), 1..3 warps per CTA
#include <cstdio>
#include <cuda_runtime.h>
#include "../../computervisionimagestich2_b200/csrc/canvas_device.cuh"
namespace pb {
constexpr int kIirPitch = 33;
__device__ __forceinline__ double f2d_exact(float f) {
    const unsigned u = __float_as_uint(f);
    const unsigned a = u & 0x7fffffffu;
    const unsigned sign = u & 0x80000000u;
    unsigned hi = ((a >> 3) + 0x38000000u) | sign;
    if (a == 0u) hi = sign;
    double d = __hiloint2double((int)hi, (int)(a << 29));
    if (a != 0u && ((a >> 23) - 1u) >= 254u) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f));
    return d;
}
template <int MODE>
__device__ __forceinline__ void tile(float* t, double& v1, double& v2, double& v3, const IirCoef& c) {
    if (MODE == 0) {   // naive
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            double v0 = (double)t[e * kIirPitch];
            v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
            t[e * kIirPitch] = (float)v0;
            v3 = v2; v2 = v1; v1 = v0;
        }
    } else if (MODE == 1) {  // all loads + conversions first, all stores last
        double d[32]; float o[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) d[e] = f2d_exact(t[e * kIirPitch]);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            double v0 = d[e];
            v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
            d[e] = v0;
            v3 = v2; v2 = v1; v1 = v0;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = (float)d[e];
#pragma unroll
        for (int e = 0; e < 32; ++e) t[e * kIirPitch] = o[e];
    } else if (MODE == 3) {  // conversion of the next sample pinned one step ahead of its use (volatile cvt)
        double cur;
        { float f = t[0]; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(cur) : "f"(f)); }
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            double nxt = 0;
            if (e + 1 < 32) { float f = t[(e + 1) * kIirPitch]; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(nxt) : "f"(f)); }
            double v0 = cur;
            v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
            t[e * kIirPitch] = (float)v0;
            v3 = v2; v2 = v1; v1 = v0;
            cur = nxt;
        }
    } else if (MODE == 4) {  // same, two steps ahead
        double c0, c1;
        { float f = t[0]; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(c0) : "f"(f)); }
        { float f = t[kIirPitch]; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(c1) : "f"(f)); }
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            double nxt = 0;
            if (e + 2 < 32) { float f = t[(e + 2) * kIirPitch]; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(nxt) : "f"(f)); }
            double v0 = c0;
            v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
            t[e * kIirPitch] = (float)v0;
            v3 = v2; v2 = v1; v1 = v0;
            c0 = c1; c1 = nxt;
        }
    } else {   // no memory at all: registers only
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            double v0 = 3.0;
            v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
            v3 = v2; v2 = v1; v1 = v0;
        }
    }
}
template <int MODE>
__global__ void k(float* out, IirCoef c, int tiles, long long* cycles) {
    __shared__ float sm[3][32 * kIirPitch];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 32 * kIirPitch; i += 32) sm[warp][i] = (float)(i % 251);
    __syncwarp();
    double v1 = 1, v2 = 2, v3 = 3;
    long long t0 = clock64();
    for (int q = 0; q < tiles; ++q) tile<MODE>(&sm[warp][lane], v1, v2, v3, c);
    long long t1 = clock64();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    out[threadIdx.x] = (float)v1 + sm[warp][lane];
}
}
int main() {
    float* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    pb::IirCoef co; co.f1 = 0.5; co.f2 = -0.2; co.f3 = 0.05; co.sum = 0.3; co.sumsq = 0.09; co.bnd = 0.65;
    for (int i = 0; i < 9; ++i) co.M[i] = 0.1 * i;
    const int tiles = 2000;
    const char* names[] = {"naive (cvt in chain order)", "loads+int-cvt first, F2F+STS last", "registers only", "cvt pinned 1 step ahead", "cvt pinned 2 steps ahead"};
    for (int warps = 1; warps <= 1; warps += 2)
        for (int m = 0; m < 5; ++m) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) pb::k<0><<<1, 32 * warps>>>(d, co, tiles, c);
                if (m == 1) pb::k<1><<<1, 32 * warps>>>(d, co, tiles, c);
                if (m == 2) pb::k<5><<<1, 32 * warps>>>(d, co, tiles, c);
                if (m == 3) pb::k<3><<<1, 32 * warps>>>(d, co, tiles, c);
                if (m == 4) pb::k<4><<<1, 32 * warps>>>(d, co, tiles, c);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("warps %d  %-36s %.2f cycles per step\n", warps, names[m], (double)h / (tiles * 32.0));
        }
    return 0;
}
