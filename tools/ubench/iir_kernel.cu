// micro-benchmark: the real recursive-Gaussian kernel on synthetic planes, cycles per sample vs number of CTAs
#include "../../computervisionimagestich2_b200/csrc/canvas_kernels.cu"
#include "../../computervisionimagestich2_b200/csrc/host_numerics.h"
#include <cstdio>
#include <vector>
using namespace pb;
int main() {
    hostnum::VanVliet v = hostnum::vanvliet_coeffs(2.0f);
    IirCoef c;
    c.f1 = v.filter[1]; c.f2 = v.filter[2]; c.f3 = v.filter[3]; c.sumsq = v.filter[0]; c.sum = v.sum; c.bnd = v.bnd;
    for (int i = 0; i < 9; ++i) c.M[i] = v.M[i];
    const int w = 4096;
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int ctas : {1, 37, 148, 296, 444, 592, 1184}) {
        const int h = 32 * ctas;
        size_t n = (size_t)w * h;
        float *a, *b;
        cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4);
        std::vector<float> hsrc(n);
        for (size_t i = 0; i < n; ++i) hsrc[i] = (float)((i * 2654435761u) >> 24);
        cudaMemcpy(a, hsrc.data(), n * 4, cudaMemcpyHostToDevice);
        for (int pass = 0; pass < 2; ++pass) {   // x pass only (h lines of w samples); then y pass only via a transposed view
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0, st);
                if (pass == 0) iir_pipe_kernel<true, IirCoef><<<ctas, 128, 0, st>>>(a, b, w, (long)h, h, (long)w * h, 1L, c);
                else iir_pipe_kernel<false, IirCoef><<<ctas, 128, 0, st>>>(a, b, w, (long)h, h, (long)w * h, (long)h, c);  // plane viewed as [w rows][h cols]
                cudaEventRecord(e1, st);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            printf("%s pass  CTAs %5d  N %d  %.1f us  -> %.1f cycles per sample-step at %d MHz (err %s)\n", pass == 0 ? "x" : "y", ctas, w,
                   ms * 1e3, ms * 1e-3 * clk * 1e3 / (2.0 * w), clk / 1000, cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(a); cudaFree(b);
    }
    return 0;
}
