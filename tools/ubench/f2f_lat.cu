// micro-benchmark: latency / issue cost of float<->double conversions on one warp (tools only)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, float a, int iters, long long* cycles) {
    float x = a + threadIdx.x;
    float y0 = x + 1, y1 = x + 2, y2 = x + 3, y3 = x + 4;
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            if (MODE == 0) { double d = (double)x; d = __dadd_rn(d, 1.0); x = (float)d; }           // F2F + DADD + F2F dependent
            if (MODE == 1) { double d = (double)x; x = (float)d; }                                 // F2F + F2F dependent
            if (MODE == 2) { acc = __dadd_rn(acc, (double)y0); y0 += 1.f; }                          // dadd chain fed by independent cvt
            if (MODE == 3) { double d0 = (double)y0, d1 = (double)y1, d2 = (double)y2, d3 = (double)y3;   // 4 independent cvt pairs
                             y0 = (float)__dadd_rn(d0, 1.0); y1 = (float)__dadd_rn(d1, 1.0); y2 = (float)__dadd_rn(d2, 1.0); y3 = (float)__dadd_rn(d3, 1.0); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    out[threadIdx.x] = x + y0 + y1 + y2 + y3 + (float)acc;
}
int main() {
    float* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    const int iters = 4096;
    const char* names[] = {"cvt.f64.f32 -> dadd -> cvt.f32.f64 (dependent)", "cvt.f64.f32 -> cvt.f32.f64 (dependent)", "dadd chain + independent cvt.f64.f32", "4 independent (cvt,dadd,cvt) groups"};
    for (int m = 0; m < 4; ++m) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) k<0><<<1, 32>>>(d, 1.0f, iters, c);
            if (m == 1) k<1><<<1, 32>>>(d, 1.0f, iters, c);
            if (m == 2) k<2><<<1, 32>>>(d, 1.0f, iters, c);
            if (m == 3) k<3><<<1, 32>>>(d, 1.0f, iters, c);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-52s %.2f cycles per group\n", names[m], (double)h / (iters * 8.0));
    }
    return 0;
}
