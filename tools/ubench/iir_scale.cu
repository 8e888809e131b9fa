// micro-benchmark: how the recursive-Gaussian consumer loop scales with the number of consumer warps resident on ONE
// SM (1 CTA of 1..16 warps = 1..4 per sub-partition), for different float<->double conversion strategies.  Answers:
// which per-SM resource makes iir_pipe_kernel slow down from 60 to 160 cycles per sample between 1 and 8 CTAs per SM?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../computervisionimagestich2_b200/csrc/canvas_device.cuh"
#include "cvt_exact.cuh"
namespace pb {
constexpr int kIirPitch = 33;
// MODE 0: hardware conversions (F2F) as in the kernel; 1: integer-pipe conversions (cvt_exact.cuh); 2: no conversions
// at all (the tile is read / written as raw bits, fp64 chain kept); 3: registers only (fp64 chain alone)
template <int MODE>
__device__ __forceinline__ void tile(float* t, double& v1, double& v2, double& v3, const IirCoef& c) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        double v0;
        if (MODE == 0) v0 = (double)t[e * kIirPitch];
        else if (MODE == 1) v0 = f2d_int(t[e * kIirPitch]);
        else if (MODE == 2) v0 = __hiloint2double(0x40080000, __float_as_int(t[e * kIirPitch]));
        else v0 = 3.0;
        v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
        if (MODE == 0) t[e * kIirPitch] = (float)v0;
        else if (MODE == 1) t[e * kIirPitch] = d2f_int(v0);
        else if (MODE == 2) t[e * kIirPitch] = __int_as_float(__double2loint(v0));
        v3 = v2; v2 = v1; v1 = v0;
    }
}
template <int MODE>
__global__ void k(float* out, IirCoef c, int tiles, long long* cycles) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* my = sm + warp * 32 * kIirPitch;
    for (int i = lane; i < 32 * kIirPitch; i += 32) my[i] = (float)(i % 251);
    __syncthreads();
    double v1 = 1, v2 = 2, v3 = 3;
    long long t0 = clock64();
    for (int q = 0; q < tiles; ++q) tile<MODE>(my + lane, v1, v2, v3, c);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    out[threadIdx.x] = (float)v1 + my[lane];
}
}
int main() {
    float* d; long long* c; cudaMalloc(&d, 1 << 16); cudaMalloc(&c, 8);
    pb::IirCoef co; co.f1 = 0.5; co.f2 = -0.2; co.f3 = 0.05; co.sum = 0.3; co.sumsq = 0.09; co.bnd = 0.65;
    for (int i = 0; i < 9; ++i) co.M[i] = 0.1 * i;
    const int tiles = 1000;
    const char* names[] = {"F2F conversions", "integer conversions", "no conversions (LDS/STS + fp64 chain)", "fp64 chain only"};
    for (int m = 0; m < 4; ++m)
        for (int warps : {1, 2, 4, 8, 16}) {
            const size_t smem = (size_t)warps * 32 * pb::kIirPitch * 4;
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) { cudaFuncSetAttribute(pb::k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000); pb::k<0><<<1, 32 * warps, smem>>>(d, co, tiles, c); }
                if (m == 1) { cudaFuncSetAttribute(pb::k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000); pb::k<1><<<1, 32 * warps, smem>>>(d, co, tiles, c); }
                if (m == 2) { cudaFuncSetAttribute(pb::k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000); pb::k<2><<<1, 32 * warps, smem>>>(d, co, tiles, c); }
                if (m == 3) { cudaFuncSetAttribute(pb::k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000); pb::k<3><<<1, 32 * warps, smem>>>(d, co, tiles, c); }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%-40s warps/SM %2d  %.1f cycles per sample (%s)\n", names[m], warps, (double)h / (tiles * 32.0), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
