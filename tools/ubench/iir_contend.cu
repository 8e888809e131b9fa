// micro-benchmark: does shared-memory traffic of OTHER warps slow the IIR consumer's tile loop?
// warp 0 = consumer (timed); warp 1 = optional traffic generator: per "tile" 32 cp.async.4 scatter writes + 32 LDS + 32 STG
#include <cstdio>
#include <cuda_runtime.h>
#include "../../computervisionimagestich2_b200/csrc/canvas_device.cuh"
namespace pb {
constexpr int P = 33;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(float* g, IirCoef c, int tiles, long long* cycles) {
    __shared__ float sm[4][32 * P];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 32 * P; i += 32) { sm[0][i] = (float)(i % 251); sm[1][i] = 1.f; sm[2][i] = 2.f; sm[3][i] = 3.f; }
    __syncthreads();
    if (warp == 0) {
        double v1 = 1, v2 = 2, v3 = 3;
        float* t = &sm[0][lane];
        long long t0 = clock64();
        for (int q = 0; q < tiles; ++q) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                double v0 = (double)t[e * P];
                v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
                t[e * P] = (float)v0;
                v3 = v2; v2 = v1; v1 = v0;
            }
        }
        long long t1 = clock64();
        if (lane == 0) *cycles = t1 - t0;
        g[lane] = (float)v1;
    } else if (MODE != 0) {
        // traffic generator, roughly paced like the loader + storer (one tile's worth per ~1200 cycles is the real rate;
        // here it simply runs flat out for the same number of tiles x 4)
        float acc = 0;
        for (int q = 0; q < tiles * (MODE == 3 ? 1 : 2); ++q) {
            float* t = sm[1 + (q % 3)];
            if (MODE == 1 || MODE == 3) {
#pragma unroll 8
                for (int r = 0; r < 32; ++r)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&t[lane * P + r])), "l"(g + 4096 + ((q * 32 + r) * 32 + lane) % 65536) : "memory");
                asm volatile("cp.async.wait_all;" ::: "memory");
            }
            if (MODE == 2 || MODE == 3) {
#pragma unroll 8
                for (int r = 0; r < 32; ++r) g[70000 + ((q * 32 + r) * 32 + lane) % 65536] = t[lane * P + r];
            }
            if (MODE == 3) __nanosleep(300);
        }
        g[64 + lane] = acc;
    }
}
}
int main() {
    float* d; long long* c; cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 8); cudaMemset(d, 0, 1 << 20);
    pb::IirCoef co; co.f1 = 0.5; co.f2 = -0.2; co.f3 = 0.05; co.sum = 0.3; co.sumsq = 0.09; co.bnd = 0.65;
    for (int i = 0; i < 9; ++i) co.M[i] = 0.1 * i;
    const int tiles = 2000;
    const char* names[] = {"consumer alone", "+ warp doing cp.async.4 scatter into smem", "+ warp doing LDS + STG", "+ warp doing both, paced"};
    for (int m = 0; m < 4; ++m) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) pb::k<0><<<1, 64>>>(d, co, tiles, c);
            if (m == 1) pb::k<1><<<1, 64>>>(d, co, tiles, c);
            if (m == 2) pb::k<2><<<1, 64>>>(d, co, tiles, c);
            if (m == 3) pb::k<3><<<1, 64>>>(d, co, tiles, c);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-46s %.2f cycles per step (%s)\n", names[m], (double)h / (tiles * 32.0), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
