// micro-benchmark: the IIR recurrence step (3 DMUL + 3 DADD, chain = DMUL + 3 DADD) in registers, one or more warps
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, double f1, double f2, double f3, int iters, long long* cycles) {
    double v1 = 1.0 + threadIdx.x, v2 = 0.5, v3 = 0.25;
    double in = 3.0;
    float fin = 3.0f + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double v0;
            if (MODE == 0) v0 = in;
            if (MODE == 1) { v0 = (double)fin; fin += 1.0f; }
            if (MODE == 2) { v0 = in; }
            v0 = __dadd_rn(v0, __dmul_rn(v1, f1));
            v0 = __dadd_rn(v0, __dmul_rn(v2, f2));
            v0 = __dadd_rn(v0, __dmul_rn(v3, f3));
            if (MODE == 2) { fin = __fadd_rn(fin, (float)v0); }
            v3 = v2; v2 = v1; v1 = v0;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    out[threadIdx.x] = v1 + fin;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1 << 16); cudaMalloc(&c, 8);
    const int iters = 2048;
    const char* names[] = {"step (6 fp64)", "step + cvt.f64.f32 input", "step + cvt.f32.f64 output"};
    for (int warps = 1; warps <= 32; warps *= 2)
        for (int m = 0; m < 3; ++m) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) k<0><<<1, 32 * warps>>>(d, 0.5, 0.25, 0.125, iters, c);
                if (m == 1) k<1><<<1, 32 * warps>>>(d, 0.5, 0.25, 0.125, iters, c);
                if (m == 2) k<2><<<1, 32 * warps>>>(d, 0.5, 0.25, 0.125, iters, c);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("warps/CTA %2d  %-28s %.2f cycles per step\n", warps, names[m], (double)h / (iters * 16.0));
        }
    return 0;
}
