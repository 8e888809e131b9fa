// micro-benchmark: cost of CREDUX.MIN (warp min-reduction into a uniform register) interleaved with VABSDIFF4.U8.ACC,
// as in match_sad_sym_kernel: N VABSDIFF4 per CREDUX, all SMs, 4 CTAs x 128 threads per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t sad_acc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <int PER>
__global__ void __launch_bounds__(128) k(uint32_t* out, const uint32_t* in, int iters) {
    uint32_t q[8];
    for (int i = 0; i < 8; ++i) q[i] = in[(threadIdx.x + i * 32) & 1023];
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, r = in[threadIdx.x & 31], m = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a0 = sad_acc(q[j], r, a0); a1 = sad_acc(q[j], r + 1, a1);
            a2 = sad_acc(q[j], r + 2, a2); a3 = sad_acc(q[j], r + 3, a3);
            if (PER > 0 && (j % PER) == PER - 1) m += __reduce_min_sync(0xffffffffu, a0 + j);
        }
        r += m & 1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + m;
}
int main() {
    uint32_t *o, *in;
    cudaMalloc(&o, 148 * 8 * 128 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 1, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int m = 0; m < 4; ++m)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (m == 0) k<0><<<148 * 4, 128>>>(o, in, iters);
            if (m == 1) k<8><<<148 * 4, 128>>>(o, in, iters);   // 1 CREDUX per 32 VABSDIFF4
            if (m == 2) k<2><<<148 * 4, 128>>>(o, in, iters);   // 1 per 8
            if (m == 3) k<1><<<148 * 4, 128>>>(o, in, iters);   // 1 per 4
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const char* names[] = {"no CREDUX", "1 CREDUX / 32 VABSDIFF4", "1 CREDUX / 8 VABSDIFF4", "1 CREDUX / 4 VABSDIFF4"};
            if (rep) printf("%-26s %.3f ms\n", names[m], ms);
        }
    return 0;
}
