// match_i8_kernels.h -- launch interface of the uint8 / tcgen05 matcher (match_i8_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace pb {

struct U8Top2 {   // per (query, database split): smallest / second smallest squared L2 distance, index of the smallest
    int d0, d1, i0;
};

// Device layout of a quantised table ("blocked256"): rows are grouped in blocks of 256; inside a block chunk c
// (16 bytes of K) of row r sits at c * 4096 + r * 16, so a block is one contiguous 32 KB UMMA operand tile.  A table
// of n rows occupies u8_blocked_bytes(n); rows past n must be zero.
inline size_t u8_blocked_bytes(int n) { return (size_t)((n + 255) / 256) * 32768; }
// q = (uint8) min(512 x, 255) per component (VLFeat's descriptor quantisation) into the blocked layout, |q|^2 per row
void launch_quantize_u8(const float* src, int n, unsigned char* dst_blocked, int* norm, cudaStream_t st);
void launch_norm_u8(const unsigned char* src_blocked, int n, int* norm, cudaStream_t st);
void launch_relayout_u8(const unsigned char* src, int n, unsigned char* dst, bool to_blocked, cudaStream_t st);

int match_u8_num_splits(int NA, int NB);
// A = database, B = queries, both blocked256 u8 tables; norms = |row|^2.  idx[b] = row of A or -1 (ratio rule
// 4 d0 < d1 on squared distances); d01 (optional) [NB][3] = d0, d1, index of the nearest row regardless of the rule.
// partial must hold nsplit * NB entries.
void launch_match_u8(const unsigned char* dA, const int* normA, int NA, const unsigned char* dB, const int* normB, int NB,
                     U8Top2* partial, int nsplit, int* idx, int* d01, cudaStream_t st);

}  // namespace pb
