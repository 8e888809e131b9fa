// micro-benchmark: dependent-chain latency of FP64 / FP32 ops on one warp (tools only, not product code)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void chain(double* out, double a, double b, int iters, long long* cycles) {
    double x = a + threadIdx.x;
    float xf = (float)x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 0) x = __dadd_rn(x, b);
            if (MODE == 1) x = __dmul_rn(x, b);
            if (MODE == 2) x = __fma_rn(x, b, a);
            if (MODE == 3) xf = __fadd_rn(xf, (float)b);
            if (MODE == 4) { x = __dmul_rn(x, b); x = __dadd_rn(x, a); x = __dadd_rn(x, b); x = __dadd_rn(x, a); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    out[threadIdx.x] = x + xf;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    const int iters = 4096;
    const char* names[] = {"DADD", "DMUL", "DFMA", "FADD", "IIRSTEP(4 ops)"};
    for (int warps = 1; warps <= 8; warps *= 2)
    for (int m = 0; m < 5; ++m) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) chain<0><<<1, 32 * warps>>>(d, 1.0, 1.0000001, iters, c);
            if (m == 1) chain<1><<<1, 32 * warps>>>(d, 1.0, 1.0000001, iters, c);
            if (m == 2) chain<2><<<1, 32 * warps>>>(d, 1.0, 1.0000001, iters, c);
            if (m == 3) chain<3><<<1, 32 * warps>>>(d, 1.0, 1.0000001, iters, c);
            if (m == 4) chain<4><<<1, 32 * warps>>>(d, 1.0, 1.0000001, iters, c);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("warps/CTA %d  %-16s %.2f cycles per op-group\n", warps, names[m], (double)h / (iters * 16.0));
    }
    return 0;
}
