// micro-benchmark: issue rate of VABSDIFF4.U8.ACC (__vsadu4 + accumulate), the inner instruction of the u8 SAD
// pre-filter of the matcher, against FADD (the exact float-L1 matcher's instruction) and IADD3, on all SMs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t sad_acc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* in, int iters) {
    uint32_t q[8];
    for (int i = 0; i < 8; ++i) q[i] = in[(threadIdx.x + i * 32) & 1023];
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    float f0 = 0, f1 = 0, f2 = 0, f3 = 0;
    uint32_t r = in[threadIdx.x & 31];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) {
                a0 = sad_acc(q[j], r, a0); a1 = sad_acc(q[j], r + 1, a1);
                a2 = sad_acc(q[j], r + 2, a2); a3 = sad_acc(q[j], r + 3, a3);
            } else if (MODE == 1) {
                f0 = __fadd_rn(f0, __int_as_float(q[j])); f1 = __fadd_rn(f1, __int_as_float(q[j]));
                f2 = __fadd_rn(f2, __int_as_float(q[j])); f3 = __fadd_rn(f3, __int_as_float(q[j]));
            } else {
                a0 += q[j] ^ r; a1 += q[j] ^ a0; a2 += q[j] ^ a1; a3 += q[j] ^ a2;
            }
        }
        r += a0 & 1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __float_as_int(f0 + f1 + f2 + f3);
}

int main() {
    uint32_t *o, *in;
    cudaMalloc(&o, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 1, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    const char* names[] = {"VABSDIFF4.U8.ACC", "FADD", "LOP3+IADD"};
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int m = 0; m < 3; ++m)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (m == 0) k<0><<<148 * 8, 256>>>(o, in, iters);
            if (m == 1) k<1><<<148 * 8, 256>>>(o, in, iters);
            if (m == 2) k<2><<<148 * 8, 256>>>(o, in, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double n = 148.0 * 8 * 256 * iters * 32;   // thread-instructions of the measured kind
            if (rep) printf("%-18s %.3f ms  %.2f T thread-instr/s = %.1f lanes/clk/SM at %d MHz nominal\n", names[m], ms,
                            n / ms / 1e9, n / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
        }
    return 0;
}
