// match_i8_kernels.h -- launch interface of the uint8 / tcgen05 matcher (match_i8_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace pb {

// Device layout of a quantised descriptor table ("blocked"): rows are grouped in blocks of 256; a block holds
// kU8Chunks 16-byte K-chunks per row, chunk c of row r at c * 4096 + r * 16, so that 128 consecutive rows of one chunk
// are one contiguous 2 KB run = a column of UMMA core matrices (8 rows x 16 B each).  Chunks 0..7 are the 128
// descriptor bytes; chunks 8.. hold the "norm extension" columns the matcher appends to DATABASE tables (see the .cu).
constexpr int kU8Chunks = 14;                   // 8 descriptor chunks + up to 3 extension K steps (2 chunks each)
constexpr int kU8BlockBytes = kU8Chunks * 4096;
inline size_t u8_table_bytes(int n) { return (size_t)((n + 255) / 256) * kU8BlockBytes; }

struct U8Table {
    unsigned char* blk = nullptr;   // blocked layout, u8_table_bytes(n) bytes, rows past n zero
    int* norm = nullptr;            // |row|^2
    int n = 0;
    int hmax = 0;                   // max over rows of floor(|row|^2 / 2)      (set by u8_table_prepare)
    int ext_steps = 0;              // extension K steps (of 32 columns) in use (set by u8_table_prepare)
};

struct U8Top3 {   // per (query, database split): the four largest unit maxima of the folded score, units of the first three
    int m1, u1, m2, u2, m3, u3, m4;
};

// q = (uint8) min(512 x, 255) per component (VLFeat's descriptor quantisation) into chunks 0..7 of the blocked layout
void launch_quantize_u8(const float* src, int n, unsigned char* dst_blocked, cudaStream_t st);
// row-major [n][128] u8 <-> chunks 0..7 of the blocked layout
void launch_relayout_u8(const unsigned char* src, int n, unsigned char* dst, bool to_blocked, cudaStream_t st);
// norms, and (for a table used as DATABASE) the extension columns; synchronises the stream once (reads the norm range)
void u8_table_prepare(U8Table& t, int* scratch2, cudaStream_t st);

int match_u8_num_splits(int NA, int NB);
// A = database, B = queries (both prepared).  idx[b] = row of A or -1 (ratio rule 4 d0 < d1 on squared distances);
// d01 (optional) [NB][3] = d0, d1, index of the nearest row regardless of the rule.  partial: nsplit * NB entries.
// mma_only: tensor-pipe peak measurement -- the same TMA + UTCIMMA stream, accumulators released unread, no results
void launch_match_u8(const U8Table& A, const U8Table& B, U8Top3* partial, int nsplit, int* idx, int* d01, cudaStream_t st,
                     bool mma_only = false);

}  // namespace pb
