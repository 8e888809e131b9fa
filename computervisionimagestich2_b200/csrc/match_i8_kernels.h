// match_i8_kernels.h -- launch interface of the uint8 / tcgen05 matcher (match_i8_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace pb {

struct U8Top2 {   // per (query, database split): smallest / second smallest squared L2 distance, index of the smallest
    int d0, d1, i0;
};

// q = (uint8) min(512 x, 255) per component (VLFeat's descriptor quantisation) and |q|^2 per row
void launch_quantize_u8(const float* src, int n, unsigned char* dst, int* norm, cudaStream_t st);
void launch_norm_u8(const unsigned char* src, int n, int* norm, cudaStream_t st);

int match_u8_num_splits(int NA, int NB);
// A = database [NA][128] u8, B = queries [NB][128] u8, norms = |row|^2.  idx[b] = row of A or -1 (ratio rule
// 4 d0 < d1 on squared distances); d01 (optional) [NB][3] = d0, d1, index of the nearest row regardless of the rule.
// partial must hold nsplit * NB entries.
void launch_match_u8(const unsigned char* dA, const int* normA, int NA, const unsigned char* dB, const int* normB, int NB,
                     U8Top2* partial, int nsplit, int* idx, int* d01, cudaStream_t st);

}  // namespace pb
