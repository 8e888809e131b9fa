// match_kernels.cu -- exact float-L1 2-NN brute-force matcher and batched RANSAC hypothesis scoring (sm_100a).
//
// Matcher (replaces the kd-forest query loop of ImageProcess::getImgPair, ImageProcess.cpp:311-346): every thread
// owns one query descriptor (128 floats in registers); database rows stream through shared memory and are read
// with 128-bit broadcast loads; four rows are accumulated at a time so that four independent FADD chains hide the
// FP32 latency.  The accumulation over the 128 dimensions is sequential in dimension order with separate
// subtract / |.|-add (no FMA), i.e. exactly _vl_distance_l1_f (vl/mathop.c:307-318).  The database is split across
// blockIdx.y so small problems still fill the 148 SMs; a merge kernel combines the per-split top-2 and applies the
// ratio rule.
#include "match_kernels.h"
#include "match_device.cuh"
#include "ransac_device.cuh"
#include "common.h"
#include "ktimer.h"
#include <algorithm>

namespace pb {

constexpr int kQ = 128;   // queries per CTA (one per thread)
constexpr int kTA = 32;   // database rows per shared-memory tile

// LIST: scan only the queries of J.overflow[0 .. J.counters[1]) (the pre-filter's fallback); partial is then indexed
// by list position.  A short list (at most 1 / kListFine of the queries -- the usual case: a few dozen) would leave the
// machine empty with the database split sized for ALL queries, so it takes kListFine times as many splits; the partial
// buffer holds both layouts (kListFine x the splits, 1 / kListFine of the positions).  FINE selects which of the two
// layouts a launch serves; the other launch's CTAs leave at once.
constexpr int kListFine = 16;
__device__ __forceinline__ bool list_layout(const MatchJob& J, int count, int& nsplit, int& rps, int& pstride) {
    nsplit = J.nsplit; rps = J.rows_per_split; pstride = J.NB;
    if ((long)count * kListFine > J.NB) return false;
    const int fine = min(J.nsplit * kListFine, max(1, J.NA / (2 * kTA)));   // at least two tiles per split
    if (fine <= nsplit) return false;
    rps = ((J.NA + fine - 1) / fine + 3) & ~3;
    nsplit = (J.NA + rps - 1) / rps;
    pstride = J.NB / kListFine;
    return true;
}
template <bool LIST, bool FINE = false>
__global__ void __launch_bounds__(kQ) match_l1_kernel(const MatchJob* __restrict__ jobs) {
    __shared__ __align__(16) float tile[kTA][128];
    const MatchJob& J = jobs[blockIdx.z];
    const float* __restrict__ A = J.A;
    const float* __restrict__ B = J.B;
    const int NA = J.NA;
    const int NB = LIST ? J.counters[1] : J.NB;    // number of queries scanned
    int nsplit = J.nsplit, rows_per_split = J.rows_per_split, pstride = J.NB;
    if (LIST && list_layout(J, NB, nsplit, rows_per_split, pstride) != FINE) return;
    Top2* __restrict__ partial = J.partial;
    if (blockIdx.x * kQ >= NB || blockIdx.y >= nsplit) return;   // grid is sized for the largest job of the batch
    const int pos = blockIdx.x * kQ + threadIdx.x;
    const int b = LIST ? J.overflow[min(pos, NB - 1)] : pos;
    const int a_begin = blockIdx.y * rows_per_split;
    const int a_end = min(NA, a_begin + rows_per_split);
    float q[128];
    {
        const float4* src = reinterpret_cast<const float4*>(B + (size_t)min(b, J.NB - 1) * 128);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            float4 t = src[k];
            q[4 * k] = t.x; q[4 * k + 1] = t.y; q[4 * k + 2] = t.z; q[4 * k + 3] = t.w;
        }
    }
    Top2 best = top2_init();
    for (int a0 = a_begin; a0 < a_end; a0 += kTA) {
        __syncthreads();
        // cooperative, coalesced tile load: kTA rows x 32 float4
        for (int i = threadIdx.x; i < kTA * 32; i += kQ) {
            int r = i >> 5, c = i & 31;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a0 + r < a_end) v = reinterpret_cast<const float4*>(A + (size_t)(a0 + r) * 128)[c];
            reinterpret_cast<float4*>(&tile[r][0])[c] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int r = 0; r < kTA; r += 4) {
            if (a0 + r >= a_end) break;
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float4 t0 = reinterpret_cast<const float4*>(&tile[r + 0][0])[k];
                const float4 t1 = reinterpret_cast<const float4*>(&tile[r + 1][0])[k];
                const float4 t2 = reinterpret_cast<const float4*>(&tile[r + 2][0])[k];
                const float4 t3 = reinterpret_cast<const float4*>(&tile[r + 3][0])[k];
                acc0 += fabsf(q[4 * k] - t0.x); acc1 += fabsf(q[4 * k] - t1.x);
                acc2 += fabsf(q[4 * k] - t2.x); acc3 += fabsf(q[4 * k] - t3.x);
                acc0 += fabsf(q[4 * k + 1] - t0.y); acc1 += fabsf(q[4 * k + 1] - t1.y);
                acc2 += fabsf(q[4 * k + 1] - t2.y); acc3 += fabsf(q[4 * k + 1] - t3.y);
                acc0 += fabsf(q[4 * k + 2] - t0.z); acc1 += fabsf(q[4 * k + 2] - t1.z);
                acc2 += fabsf(q[4 * k + 2] - t2.z); acc3 += fabsf(q[4 * k + 2] - t3.z);
                acc0 += fabsf(q[4 * k + 3] - t0.w); acc1 += fabsf(q[4 * k + 3] - t1.w);
                acc2 += fabsf(q[4 * k + 3] - t2.w); acc3 += fabsf(q[4 * k + 3] - t3.w);
            }
            const int a = a0 + r;
            top2_push(best, acc0, a);
            if (a + 1 < a_end) top2_push(best, acc1, a + 1);
            if (a + 2 < a_end) top2_push(best, acc2, a + 2);
            if (a + 3 < a_end) top2_push(best, acc3, a + 3);
        }
    }
    if (pos < NB) partial[(size_t)blockIdx.y * pstride + pos] = best;
}

template <bool LIST>
__global__ void match_merge_kernel(const MatchJob* __restrict__ jobs) {
    const MatchJob& J = jobs[blockIdx.y];
    const Top2* __restrict__ partial = J.partial;
    const int NA = J.NA;
    const int NB = LIST ? J.counters[1] : J.NB;
    int nsplit = J.nsplit, rps = 0, pstride = J.NB;
    if (LIST) list_layout(J, NB, nsplit, rps, pstride);
    int* __restrict__ idx = J.idx;
    float* __restrict__ d01 = J.d01;
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= NB) return;
    const int b = LIST ? J.overflow[pos] : pos;
    Top2 t = partial[pos];
    for (int s = 1; s < nsplit; ++s) top2_merge(t, partial[(size_t)s * pstride + pos]);
    bool ok = NA >= 2 && t.i0 >= 0 && ratio_test(t.d0, t.d1);
    idx[b] = ok ? t.i0 : -1;
    if (!LIST && d01) { d01[2 * b] = t.d0; d01[2 * b + 1] = t.d1; }
}

// ---------------------------------------------------------------------------------------------------------
// Pre-filter (arithmetic and proof: match_device.cuh).  Tables are quantised once per image.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sad_quantize_kernel(const float* __restrict__ descr, int n,
                                                           unsigned* __restrict__ q8, int* __restrict__ qe,
                                                           int* __restrict__ emax, unsigned* __restrict__ g8,
                                                           unsigned short* __restrict__ w16) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    const float4 v = reinterpret_cast<const float4*>(descr + (size_t)row * 128)[lane];
    double e = 0.0;
    const unsigned b0 = sad_quantize(v.x, &e), b1 = sad_quantize(v.y, &e), b2 = sad_quantize(v.z, &e), b3 = sad_quantize(v.w, &e);
    q8[(size_t)row * 32 + lane] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);   // any order: the bound is rounded up
    const int er = sad_row_error(e);
    if (lane == 0) {
        qe[row] = er;
        if (emax) atomicMax(emax, er);
    }
    if (g8) {
        // grouped vector (match_device.cuh, sad_group_of_dim): lane = 4 consecutive bins of cell lane / 2; the 2 x 2 block
        // of cells differs in lane bits 2 (cx) and 8 (cy)
        unsigned s0 = b0, s1 = b1, s2 = b2, s3 = b3;
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 2); s3 += __shfl_xor_sync(0xffffffffu, s3, 2);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 8); s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 8); s3 += __shfl_xor_sync(0xffffffffu, s3, 8);
        if ((lane & 10) == 0) {
            const int word = (lane & 1) + 2 * (((lane >> 2) & 1) + 2 * (lane >> 4));
            g8[(size_t)row * 8 + word] = (s0 >> kGrpShift) | ((s1 >> kGrpShift) << 8) | ((s2 >> kGrpShift) << 16) |
                                         ((s3 >> kGrpShift) << 24);
        }
        if (lane == 0) w16[row] = (unsigned short)(er <= kGrpErrCap ? sad_w16(er) : 0u);
    }
}

void launch_sad_quantize(const float* descr, int n, unsigned* q8, int* qe, int* emax, cudaStream_t st, unsigned* g8,
                         unsigned short* w16) {
    if (emax) PB_CUDA(cudaMemsetAsync(emax, 0, sizeof(int), st));
    if (n <= 0) return;
    KScope ks("match.quantize", st, (double)n * (512 + 128 + 4 + 32 + 2));
    sad_quantize_kernel<<<div_up((long)n * 32, 128), 128, 0, st>>>(descr, n, q8, qe, emax, g8, w16);
    PB_KERNEL_CHECK();
}

constexpr int kST = 128;      // threads per CTA of the SAD kernels
constexpr int kSQ = 2;        // queries per thread (each database word read from shared memory serves both)
constexpr int kSRows = 64;    // database rows per shared-memory tile (8 KB of bytes + 256 B of error bounds)

__device__ __forceinline__ unsigned sad4_acc(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   // SASS: VABSDIFF4.U8.ACC
    return d;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}

// MODE 0: statistics pass over all queries (thread = kSQ queries; SadStat per query and database split).
// MODE 1: candidate pass over the surviving queries: rows with SAD - e(a) <= thr are appended to the query's list.
// One VABSDIFF4.U8.ACC per four dimensions: 32 instructions per (query, row) against 256 for the float scan; the
// accumulator starts at e(a), so SAD + e(a) costs nothing.
template <int MODE>
__global__ void __launch_bounds__(kST) match_sad_kernel(const MatchJob* __restrict__ jobs, const int* __restrict__ which) {
    __shared__ __align__(16) unsigned tile[2][kSRows][32];
    __shared__ int terr[2][kSRows];
    const MatchJob& J = jobs[which ? which[blockIdx.z] : blockIdx.z];
    const int NA = J.NA;
    const int nq = MODE == 0 ? J.NB : J.counters[0];
    const int rps = MODE == 0 ? J.sad_rows_per_split : J.cand_rows_per_split;
    if (blockIdx.x * (kST * kSQ) >= nq || blockIdx.y >= (MODE == 0 ? J.sad_nsplit : J.cand_nsplit)) return;
    const unsigned* __restrict__ A8 = J.A8;
    const int* __restrict__ Ae = J.Ae;
    const int a_begin = blockIdx.y * rps;
    const int a_end = min(NA, a_begin + rps);
    const int tid = threadIdx.x;

    unsigned q[kSQ][32];
    int slot[kSQ], thr[kSQ];
    bool live[kSQ];
#pragma unroll
    for (int j = 0; j < kSQ; ++j) {
        const int s = blockIdx.x * (kST * kSQ) + j * kST + tid;
        live[j] = s < nq;
        slot[j] = min(s, nq - 1);
        const int b = MODE == 0 ? slot[j] : J.surv[slot[j]];
        thr[j] = MODE == 0 ? 0 : J.thr[slot[j]];
        const uint4* src = reinterpret_cast<const uint4*>(J.B8 + (size_t)b * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 t = src[k];
            q[j][4 * k] = t.x; q[j][4 * k + 1] = t.y; q[j][4 * k + 2] = t.z; q[j][4 * k + 3] = t.w;
        }
    }
    SadStat st[kSQ];
#pragma unroll
    for (int j = 0; j < kSQ; ++j) st[j] = sadstat_init();

    const int ntiles = (a_end - a_begin + kSRows - 1) / kSRows;
    auto issue = [&](int t) {
        const int buf = t & 1, a0 = a_begin + t * kSRows, rows = min(kSRows, a_end - a0);
        for (int i = tid; i < rows * 8; i += kST)
            cp_async16(&tile[buf][i >> 3][(i & 7) * 4], A8 + (size_t)(a0 + (i >> 3)) * 32 + (i & 7) * 4);
        if (tid < rows) cp_async4(&terr[buf][tid], Ae + a0 + tid);
        asm volatile("cp.async.commit_group;");
    };
    if (ntiles > 0) issue(0);
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) {
            issue(t + 1);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const int buf = t & 1, a0 = a_begin + t * kSRows, rows = min(kSRows, a_end - a0);
#pragma unroll 1
        for (int r = 0; r < rows; r += 2) {
            const bool two = r + 1 < rows;     // uniform; row r + 1 of the tile holds stale bytes otherwise (ignored)
            const int e0 = terr[buf][r], e1 = two ? terr[buf][r + 1] : 0;
            unsigned acc[kSQ][2];
#pragma unroll
            for (int j = 0; j < kSQ; ++j) { acc[j][0] = (unsigned)e0; acc[j][1] = (unsigned)e1; }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 w0 = reinterpret_cast<const uint4*>(&tile[buf][r][0])[k];
                const uint4 w1 = reinterpret_cast<const uint4*>(&tile[buf][r + 1][0])[k];
#pragma unroll
                for (int j = 0; j < kSQ; ++j) {
                    acc[j][0] = sad4_acc(q[j][4 * k], w0.x, acc[j][0]); acc[j][1] = sad4_acc(q[j][4 * k], w1.x, acc[j][1]);
                    acc[j][0] = sad4_acc(q[j][4 * k + 1], w0.y, acc[j][0]); acc[j][1] = sad4_acc(q[j][4 * k + 1], w1.y, acc[j][1]);
                    acc[j][0] = sad4_acc(q[j][4 * k + 2], w0.z, acc[j][0]); acc[j][1] = sad4_acc(q[j][4 * k + 2], w1.z, acc[j][1]);
                    acc[j][0] = sad4_acc(q[j][4 * k + 3], w0.w, acc[j][0]); acc[j][1] = sad4_acc(q[j][4 * k + 3], w1.w, acc[j][1]);
                }
            }
#pragma unroll
            for (int j = 0; j < kSQ; ++j) {
                if (MODE == 0) {
                    sadstat_push(st[j], (int)acc[j][0], e0);
                    if (two) sadstat_push(st[j], (int)acc[j][1], e1);
                } else {
                    if (live[j] && (int)acc[j][0] - 2 * e0 <= thr[j]) {
                        const int p = atomicAdd(&J.cand_cnt[slot[j]], 1);
                        if (p < kMatchCand) J.cand[(size_t)slot[j] * kMatchCand + p] = a0 + r;
                    }
                    if (two && live[j] && (int)acc[j][1] - 2 * e1 <= thr[j]) {
                        const int p = atomicAdd(&J.cand_cnt[slot[j]], 1);
                        if (p < kMatchCand) J.cand[(size_t)slot[j] * kMatchCand + p] = a0 + r + 1;
                    }
                }
            }
        }
        __syncthreads();   // the buffer is refilled by the copy issued in the next iteration
    }
    if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < kSQ; ++j)
            if (live[j]) J.spartial[(size_t)blockIdx.y * J.NB + slot[j]] = st[j];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Both directions of an image pair from ONE pass over the SAD matrix.  getImgPair(X, Y) (database X, queries Y) and
// getImgPair(Y, X) need the same |X| x |Y| integer SADs; only the error terms differ.  Threads hold rows of Y in
// registers (kSQ each), rows of X stream through shared memory exactly as in match_sad_kernel:
//   forward  (job F: database X, queries Y): per held y, statistics over the streamed x   -- in-thread, sequential;
//   reverse  (job R: database Y, queries X): per streamed x, statistics over the CTA's y  -- warp reduction (CREDUX),
//            the four warps of the CTA combined through shared memory once per tile, one entry per (y block, x).
// The running bounds are kept as PACKED 16-bit pairs (Blackwell's VIADD.16x2 / VIMNMX.U16x2 / VIADDMNMX.U16x2): the two
// streamed rows of an iteration share one instruction, so the forward bookkeeping costs 2.5 ALU instructions per
// (x, y) instead of 4, and the whole pass ~37 ALU-pipe instructions per (x, y) for BOTH directions against 2 x 36.
// 16 bits suffice because SAD <= 128 * 255 = 32640 and every error bound of both tables is <= kSymErrCap (the launcher
// checks the tables' maxima; other pairs take the one-directional kernel).  Lower bounds are stored biased by +kSymErrCap.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSymErrCap = 255;
constexpr unsigned kSymDead = 1u << 20;   // added to the (32-bit) bounds of a held row that does not exist: never among the minima

__global__ void __launch_bounds__(kST) match_sad_sym_kernel(const MatchJob* __restrict__ jobs, const int2* __restrict__ pairs) {
    __shared__ __align__(16) unsigned tile[2][kSRows][32];
    __shared__ int terr[2][kSRows];
    __shared__ int rstat[kST / 32][kSRows][3];     // per warp: reverse statistics of the tile's rows
    const int2 fr = pairs[blockIdx.z];
    const MatchJob& F = jobs[fr.x];                // database X = F.A, queries Y = F.B
    const MatchJob& R = jobs[fr.y];                // database Y, queries X
    const int NX = F.NA, NY = F.NB;
    if (blockIdx.x * (kST * kSQ) >= NY || blockIdx.y >= F.sad_nsplit) return;
    const unsigned* __restrict__ X8 = F.A8;
    const int* __restrict__ Xe = F.Ae;
    const int x_begin = blockIdx.y * F.sad_rows_per_split;
    const int x_end = min(NX, x_begin + F.sad_rows_per_split);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    unsigned q[kSQ][32];
    int yrow[kSQ];
    bool live[kSQ];
    unsigned eyu[kSQ], eyl[kSQ];                   // added to a SAD: upper bound / biased lower bound of the reverse problem
#pragma unroll
    for (int j = 0; j < kSQ; ++j) {
        const int y = blockIdx.x * (kST * kSQ) + j * kST + tid;
        live[j] = y < NY;
        yrow[j] = min(y, NY - 1);
        const int ey = F.Be[yrow[j]];
        eyu[j] = live[j] ? (unsigned)ey : kSymDead;
        eyl[j] = live[j] ? (unsigned)(kSymErrCap - ey) : kSymDead;
        const uint4* src = reinterpret_cast<const uint4*>(F.B8 + (size_t)yrow[j] * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 t = src[k];
            q[j][4 * k] = t.x; q[j][4 * k + 1] = t.y; q[j][4 * k + 2] = t.z; q[j][4 * k + 3] = t.w;
        }
    }
    // forward statistics, two streamed rows per register: low half = even row of the iteration, high half = odd row
    unsigned flb[kSQ], fu0[kSQ], fu1[kSQ];
#pragma unroll
    for (int j = 0; j < kSQ; ++j) flb[j] = fu0[j] = fu1[j] = 0xffffffffu;

    const int ntiles = (x_end - x_begin + kSRows - 1) / kSRows;
    auto issue = [&](int t) {
        const int buf = t & 1, a0 = x_begin + t * kSRows, rows = min(kSRows, x_end - a0);
        for (int i = tid; i < rows * 8; i += kST)
            cp_async16(&tile[buf][i >> 3][(i & 7) * 4], X8 + (size_t)(a0 + (i >> 3)) * 32 + (i & 7) * 4);
        if (tid < rows) cp_async4(&terr[buf][tid], Xe + a0 + tid);
        asm volatile("cp.async.commit_group;");
    };
    if (ntiles > 0) issue(0);
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) {
            issue(t + 1);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const int buf = t & 1, a0 = x_begin + t * kSRows, rows = min(kSRows, x_end - a0);
#pragma unroll 1
        for (int r = 0; r < rows; r += 2) {
            const bool two = r + 1 < rows;
            const unsigned e0 = (unsigned)terr[buf][r], e1 = two ? (unsigned)terr[buf][r + 1] : 0u;
            // packed per-row constants of the forward problem; when the tile ends on an odd row the high halves are
            // forced to 0xffff below (larger than any real bound <= 32640 + 255), with nothing added to them
            const unsigned exu = e0 | (e1 << 16);
            const unsigned exl = (kSymErrCap - e0) | ((two ? (kSymErrCap - e1) : 0u) << 16);
            const unsigned deadhi = two ? 0u : 0xffff0000u;
            unsigned acc[kSQ][2];
#pragma unroll
            for (int j = 0; j < kSQ; ++j) acc[j][0] = acc[j][1] = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 w0 = reinterpret_cast<const uint4*>(&tile[buf][r][0])[k];
                const uint4 w1 = reinterpret_cast<const uint4*>(&tile[buf][r + 1][0])[k];
#pragma unroll
                for (int j = 0; j < kSQ; ++j) {
                    acc[j][0] = sad4_acc(q[j][4 * k], w0.x, acc[j][0]); acc[j][1] = sad4_acc(q[j][4 * k], w1.x, acc[j][1]);
                    acc[j][0] = sad4_acc(q[j][4 * k + 1], w0.y, acc[j][0]); acc[j][1] = sad4_acc(q[j][4 * k + 1], w1.y, acc[j][1]);
                    acc[j][0] = sad4_acc(q[j][4 * k + 2], w0.z, acc[j][0]); acc[j][1] = sad4_acc(q[j][4 * k + 2], w1.z, acc[j][1]);
                    acc[j][0] = sad4_acc(q[j][4 * k + 3], w0.w, acc[j][0]); acc[j][1] = sad4_acc(q[j][4 * k + 3], w1.w, acc[j][1]);
                }
            }
            // ---- forward: packed bounds of (row r, row r + 1) against the held y ----
#pragma unroll
            for (int j = 0; j < kSQ; ++j) {
                const unsigned pk = (acc[j][1] * 65536u + acc[j][0]) | deadhi;   // IMAD: FMA pipe
                const unsigned u = __vadd2(pk, exu);
                flb[j] = __viaddmin_u16x2(pk, exl, flb[j]);
                fu1[j] = __vminu2(fu1[j], __vmaxu2(fu0[j], u));
                fu0[j] = __vminu2(fu0[j], u);
            }
            // ---- reverse: bounds of the held y's against row r (and r + 1), reduced over the warp ----
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i == 1 && !two) break;
                unsigned lo = 0xffffffffu, hi = 0xffffffffu, lb = 0xffffffffu;   // two smallest upper bounds, smallest lower bound
#pragma unroll
                for (int j = 0; j < kSQ; ++j) {
                    const unsigned u = acc[j][i] + eyu[j];
                    hi = min(hi, max(lo, u));
                    lo = min(lo, u);
                    lb = min(lb, acc[j][i] + eyl[j]);
                }
                const unsigned m_lb = __reduce_min_sync(0xffffffffu, lb);
                const unsigned m0 = __reduce_min_sync(0xffffffffu, lo);
                const unsigned holders = __ballot_sync(0xffffffffu, lo == m0);
                const unsigned rest = __reduce_min_sync(0xffffffffu, lo == m0 ? hi : lo);
                const unsigned m1 = __popc(holders) >= 2 ? m0 : rest;
                if (lane == 0) {
                    rstat[warp][r + i][0] = (int)m_lb;
                    rstat[warp][r + i][1] = (int)m0;
                    rstat[warp][r + i][2] = (int)m1;
                }
            }
        }
        __syncthreads();   // the tile buffer is refilled by the copy issued in the next iteration; rstat is complete
        if (tid < rows) {  // combine the warps, one entry per (y block, x row): no other CTA writes it
            SadStat sst = sadstat_init();
#pragma unroll
            for (int w = 0; w < kST / 32; ++w) {
                SadStat o;
                o.lbmin = rstat[w][tid][0] - kSymErrCap; o.u0 = rstat[w][tid][1]; o.u1 = rstat[w][tid][2];
                sadstat_merge(sst, o);
            }
            R.spartial[(size_t)blockIdx.x * NX + a0 + tid] = sst;
        }
    }
#pragma unroll
    for (int j = 0; j < kSQ; ++j)
        if (live[j]) {
            SadStat a, b;
            a.lbmin = (int)(flb[j] & 0xffffu) - kSymErrCap; a.u0 = (int)(fu0[j] & 0xffffu); a.u1 = (int)(fu1[j] & 0xffffu);
            b.lbmin = (int)(flb[j] >> 16) - kSymErrCap; b.u0 = (int)(fu0[j] >> 16); b.u1 = (int)(fu1[j] >> 16);
            sadstat_merge(a, b);
            F.spartial[(size_t)blockIdx.y * NY + yrow[j]] = a;
        }
}

// ---------------------------------------------------------------------------------------------------------
// Grouped pass (match_device.cuh, second level of the pre-filter).
//   match_seed_kernel:      per query, exact SAD against a strided sample of the database -> u1 (a valid upper bound of the
//                           second-nearest distance) and the packed skip threshold c16 derived from tau(u1).
//   match_group_sym_kernel: both directions of an image pair.  Threads hold the 32-byte GROUP vectors of kGY rows of Y in
//                           registers; group vectors of X stream through shared memory.  Per (x, y): 8 VABSDIFF4 give S,
//                           and the pair is skipped when S + w16(x) + w16(y) >= max(c16(y), c16(x)) -- both directed
//                           problems then know SAD - e(row) >= tau(query).  Two streamed rows share every bookkeeping
//                           instruction (packed 16-bit halves).  The few pairs that fail the test get the exact 128-byte
//                           SAD, computed by the whole warp (lane = word), and enter the statistics of both problems:
//                           forward in the holder's registers, reverse through global atomics (lb: min; u0 / u1: the
//                           displaced-value rule keeps the two smallest of a multiset under any interleaving).
// ---------------------------------------------------------------------------------------------------------
// sampled database rows per query: NA / 48 within [512, 2048] -- the second-smallest of n samples sits at the 2 / n quantile,
// so a larger table needs a larger sample for the same tightness of the seed (72 k rows at 512 samples: 0.19 % of the pairs
// fail the grouped bound against 0.08 % for 25 k rows)
__host__ __device__ inline int seed_rows(int NA) { const int k = NA / 48; return k < 512 ? 512 : (k > 2048 ? 2048 : k); }
constexpr int kGY = 4;           // held rows of Y per thread
constexpr int kGRows = 128;      // streamed rows of X per tile (a barrier per tile: 64-row tiles spent 11 % of the samples at it)

__global__ void __launch_bounds__(128) match_seed_kernel(const MatchJob* __restrict__ jobs, const int* __restrict__ which) {
    __shared__ __align__(16) unsigned tile[32][32];
    __shared__ int terr[32];
    const MatchJob& J = jobs[which[blockIdx.z]];
    const int NA = J.NA, NB = J.NB, tid = threadIdx.x;
    if (blockIdx.x * 128 >= NB) return;
    const int b = min(blockIdx.x * 128 + tid, NB - 1);
    unsigned q[32];
    {
        const uint4* src = reinterpret_cast<const uint4*>(J.B8 + (size_t)b * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 t = src[k];
            q[4 * k] = t.x; q[4 * k + 1] = t.y; q[4 * k + 2] = t.z; q[4 * k + 3] = t.w;
        }
    }
    const int stride = max(1, NA / seed_rows(NA));
    const int nrows = (NA + stride - 1) / stride;
    SadStat st = sadstat_init();
    for (int r0 = 0; r0 < nrows; r0 += 32) {
        const int rows = min(32, nrows - r0);
        __syncthreads();
        for (int i = tid; i < rows * 8; i += 128)
            reinterpret_cast<uint4*>(&tile[i >> 3][0])[i & 7] =
                reinterpret_cast<const uint4*>(J.A8 + (size_t)(r0 + (i >> 3)) * stride * 32)[i & 7];
        if (tid < rows) terr[tid] = J.Ae[(size_t)(r0 + tid) * stride];
        __syncthreads();
#pragma unroll 1
        for (int r = 0; r < rows; ++r) {
            const int ea = terr[r];
            unsigned acc = (unsigned)ea;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 w = reinterpret_cast<const uint4*>(&tile[r][0])[k];
                acc = sad4_acc(q[4 * k], w.x, acc); acc = sad4_acc(q[4 * k + 1], w.y, acc);
                acc = sad4_acc(q[4 * k + 2], w.z, acc); acc = sad4_acc(q[4 * k + 3], w.w, acc);
            }
            sadstat_push(st, (int)acc, ea);
        }
    }
    if (blockIdx.x * 128 + tid < NB) {
        const int eb = J.Be[b];
        J.seed_u1[b] = st.u1;
        J.c16[b] = (unsigned short)sad_c16(sad_tau(st.u1, eb), eb);
    }
}

// two smallest of a multiset under concurrent insertion: the value displaced from (or refused by) slot 0 goes to slot 1
__device__ __forceinline__ void atomic_top2(int* u0, int* u1, int v) {
    const int old = atomicMin(u0, v);
    atomicMin(u1, old > v ? old : v);
}
// the same for keys (value << 32 | row); returns the second-smallest VALUE after the insertion
__device__ __forceinline__ int atomic_top2_key(unsigned long long* p0, unsigned long long* p1, unsigned long long v) {
    const unsigned long long old = atomicMin(p0, v);
    const unsigned long long d = old > v ? old : v;
    const unsigned long long old1 = atomicMin(p1, d);
    return (int)((old1 < d ? old1 : d) >> 32);
}

// Queue of the row pairs the grouped bound could not skip: entry = pair index << 52 | x << 26 | y.
struct GroupQueue {
    unsigned long long* items;
    unsigned long long* count;   // entries appended so far (may exceed cap: the excess is dropped and the batch is redone
                                 // without the grouped pass); 64 bits: a 24 x 8K batch has 10^12 row pairs
    unsigned long long cap;
};
constexpr int kGQLocal = 2048;   // entries a CTA collects in shared memory between flushes

__global__ void __launch_bounds__(128) match_group_sym_kernel(const MatchJob* __restrict__ jobs, const int2* __restrict__ pairs,
                                                             GroupQueue gq) {
    __shared__ __align__(16) unsigned tile[2][kGRows][8];
    __shared__ __align__(16) unsigned short xw[2][kGRows], xc[2][kGRows];
    __shared__ unsigned long long qbuf[kGQLocal];
    __shared__ int qn;
    __shared__ unsigned long long qbase;
    const int2 fr = pairs[blockIdx.z];
    const MatchJob& F = jobs[fr.x];                // database X = F.A, queries Y = F.B
    const MatchJob& R = jobs[fr.y];                // database Y, queries X
    const int NX = F.NA, NY = F.NB;
    if (blockIdx.x * (128 * kGY) >= NY || blockIdx.y >= F.grp_nsplit) return;
    const unsigned* __restrict__ Xg = F.Ag;
    const unsigned short* __restrict__ Xw = F.Aw;
    const unsigned short* __restrict__ Xc = R.c16;
    const int x_begin = blockIdx.y * F.grp_rows_per_split;
    const int x_end = min(NX, x_begin + F.grp_rows_per_split);
    const int tid = threadIdx.x;
    const unsigned long long ztag = (unsigned long long)blockIdx.z << 52;

    unsigned g[kGY][8];
    int yrow[kGY];
    bool live[kGY];
    unsigned wy[kGY], cy[kGY];                     // w16(y); c16(y) replicated in both halves
#pragma unroll
    for (int j = 0; j < kGY; ++j) {
        const int y = blockIdx.x * (128 * kGY) + j * 128 + tid;
        live[j] = y < NY;
        yrow[j] = min(y, NY - 1);
        const uint4* src = reinterpret_cast<const uint4*>(F.Bg + (size_t)yrow[j] * 8);
        const uint4 t0 = src[0], t1 = src[1];
        g[j][0] = t0.x; g[j][1] = t0.y; g[j][2] = t0.z; g[j][3] = t0.w;
        g[j][4] = t1.x; g[j][5] = t1.y; g[j][6] = t1.z; g[j][7] = t1.w;
        wy[j] = live[j] ? (unsigned)F.Bw[yrow[j]] : 0x7f00u;   // accumulator start: w16(y); a row that does not exist never fails
        cy[j] = live[j] ? (unsigned)F.c16[yrow[j]] * 0x10001u : 0u;
    }
    if (tid == 0) qn = 0;

    // appends the CTA's collected pairs to the global queue (called by all threads, between two barriers)
    // Called by ALL threads right after a barrier, i.e. with no append in flight.  qn is reset between the two barriers:
    // after the first one every thread has read it, after the second one threads start appending again.
    auto flush = [&]() {
        const int n = min(qn, kGQLocal);
        if (tid == 0) qbase = n > 0 ? atomicAdd(gq.count, (unsigned long long)n) : 0ull;
        __syncthreads();
        if (tid == 0) qn = 0;
        for (int i = tid; i < n; i += 128)
            if (qbase + i < gq.cap) gq.items[qbase + i] = qbuf[i];
        __syncthreads();
    };

    const int ntiles = (x_end - x_begin + kGRows - 1) / kGRows;
    auto issue = [&](int t) {   // tables are padded to a multiple of kGRows rows: whole tiles are always readable
        const int buf = t & 1, a0 = x_begin + t * kGRows;
#pragma unroll
        for (int i = tid; i < kGRows * 2; i += 128)
            cp_async16(&tile[buf][i >> 1][(i & 1) * 4], Xg + (size_t)(a0 + (i >> 1)) * 8 + (i & 1) * 4);
        if (tid < kGRows / 8) cp_async16(&xw[buf][tid * 8], Xw + a0 + tid * 8);
        else if (tid < kGRows / 4) cp_async16(&xc[buf][(tid - kGRows / 8) * 8], Xc + a0 + (tid - kGRows / 8) * 8);
        asm volatile("cp.async.commit_group;");
    };
    if (ntiles > 0) issue(0);
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) {
            issue(t + 1);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const int buf = t & 1, a0 = x_begin + t * kGRows, rows = min(kGRows, x_end - a0);
#pragma unroll 1
        for (int r = 0; r < rows; r += 2) {
            const bool two = r + 1 < rows;
            const unsigned deadhi = two ? 0u : 0xffff0000u;
            const unsigned xwp = *reinterpret_cast<const unsigned*>(&xw[buf][r]);
            const unsigned xcp = *reinterpret_cast<const unsigned*>(&xc[buf][r]);
            const uint4 a0w = reinterpret_cast<const uint4*>(&tile[buf][r][0])[0];
            const uint4 a1w = reinterpret_cast<const uint4*>(&tile[buf][r][0])[1];
            const uint4 b0w = reinterpret_cast<const uint4*>(&tile[buf][r + 1][0])[0];
            const uint4 b1w = reinterpret_cast<const uint4*>(&tile[buf][r + 1][0])[1];
            unsigned Z[kGY], P[kGY], bad = 0u;
#pragma unroll
            for (int j = 0; j < kGY; ++j) {
                unsigned s0 = wy[j], s1 = wy[j];   // the accumulators start at w16(y): S + w16(y) costs nothing
                s0 = sad4_acc(g[j][0], a0w.x, s0); s1 = sad4_acc(g[j][0], b0w.x, s1);
                s0 = sad4_acc(g[j][1], a0w.y, s0); s1 = sad4_acc(g[j][1], b0w.y, s1);
                s0 = sad4_acc(g[j][2], a0w.z, s0); s1 = sad4_acc(g[j][2], b0w.z, s1);
                s0 = sad4_acc(g[j][3], a0w.w, s0); s1 = sad4_acc(g[j][3], b0w.w, s1);
                s0 = sad4_acc(g[j][4], a1w.x, s0); s1 = sad4_acc(g[j][4], b1w.x, s1);
                s0 = sad4_acc(g[j][5], a1w.y, s0); s1 = sad4_acc(g[j][5], b1w.y, s1);
                s0 = sad4_acc(g[j][6], a1w.z, s0); s1 = sad4_acc(g[j][6], b1w.z, s1);
                s0 = sad4_acc(g[j][7], a1w.w, s0); s1 = sad4_acc(g[j][7], b1w.w, s1);
                const unsigned pk = s1 * 65536u + s0;                          // IMAD: FMA pipe
                Z[j] = __vadd2(pk, xwp) | deadhi;
                P[j] = __vmaxu2(cy[j], xcp);
                bad |= __vminu2(Z[j], P[j]) ^ P[j];                            // a half differs iff its Z < P
            }
            if (bad != 0u) {
                // ---- rare (~0.1 % of the pairs): the grouped bound cannot skip the pair -> queue it for the exact SAD ----
#pragma unroll
                for (int j = 0; j < kGY; ++j) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const unsigned z = i ? (Z[j] >> 16) : (Z[j] & 0xffffu), pth = i ? (P[j] >> 16) : (P[j] & 0xffffu);
                        if (live[j] && z < pth) {
                            const unsigned long long item = ztag | ((unsigned long long)(a0 + r + i) << 26) | (unsigned)yrow[j];
                            const int pos = atomicAdd(&qn, 1);
                            if (pos < kGQLocal) qbuf[pos] = item;
                            else {   // the local buffer is full (degenerate tables): straight to the global queue
                                const unsigned long long gp = atomicAdd(gq.count, 1ull);
                                if (gp < gq.cap) gq.items[gp] = item;
                            }
                        }
                    }
                }
            }
        }
        // barrier: the tile buffer is refilled by the copy issued in the next iteration.  The flush decision is taken BY the
        // barrier (a thread that reads qn on its own could see appends of threads already in the next tile)
        if (__syncthreads_or(qn >= kGQLocal / 2)) flush();
    }
    __syncthreads();
    flush();
}

// exact 128-byte SAD of the queued pairs, eight lanes per pair (lane = 16 bytes), statistics of both directed problems
// through atomics: the two smallest keys (SAD + e(a)) << 32 | a and the two smallest SAD - e(a) (displaced-value rule).
// A value that is not below the CURRENT second smallest can never enter the final pair (the slots only decrease), so it
// is dropped after a plain read: ~2 ln(n) of a query's n queued pairs reach an atomic.
__device__ __forceinline__ void group_stat_push(int* s6, int sad, int e, int row) {
    const int lb = sad - e;
    if (lb < __ldcg(s6 + 5)) atomic_top2(s6 + 4, s6 + 5, lb);
    const unsigned long long key = ((unsigned long long)(unsigned)(sad + e) << 32) | (unsigned)row;
    if (key < __ldcg(reinterpret_cast<const unsigned long long*>(s6) + 1))
        atomic_top2_key(reinterpret_cast<unsigned long long*>(s6), reinterpret_cast<unsigned long long*>(s6) + 1, key);
}
__global__ void __launch_bounds__(256) match_group_exact_kernel(const MatchJob* __restrict__ jobs, const int2* __restrict__ pairs,
                                                               GroupQueue gq) {
    const int sub = threadIdx.x & 7;
    const unsigned long long ngroups = (gridDim.x * blockDim.x) >> 3;
    const unsigned long long n = min(*gq.count, gq.cap);
    const unsigned long long nround = (n + ngroups - 1) / ngroups * ngroups;   // whole warps stay in the loop: the shuffles need all lanes
    for (unsigned long long it = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; it < nround; it += ngroups) {
        const bool ok = it < n;
        const unsigned long long item = gq.items[ok ? it : 0];
        const int2 fr = pairs[(int)(item >> 52)];
        const MatchJob& F = jobs[fr.x];
        const MatchJob& R = jobs[fr.y];
        const int x = (int)((item >> 26) & 0x3ffffffu), y = (int)(item & 0x3ffffffu);
        const uint4 a = reinterpret_cast<const uint4*>(F.A8 + (size_t)x * 32)[sub];
        const uint4 b = reinterpret_cast<const uint4*>(F.B8 + (size_t)y * 32)[sub];
        unsigned d = sad4_acc(a.x, b.x, 0u);
        d = sad4_acc(a.y, b.y, d); d = sad4_acc(a.z, b.z, d); d = sad4_acc(a.w, b.w, d);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if (!ok) continue;
        if (sub == 0) group_stat_push(F.stat6 + 6 * (size_t)y, (int)d, F.Ae[x], x);        // forward: database row x, query y
        else if (sub == 1) group_stat_push(R.stat6 + 6 * (size_t)x, (int)d, F.Be[y], y);   // reverse: database row y, query x
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long total = *gq.count;
        atomicAdd(&jobs[pairs[0].x].counters[2], (int)min(min(total, gq.cap), 0x7fffffffull));   // bookkeeping: exact SADs of the batch ...
        if (total > gq.cap) atomicAdd(&jobs[pairs[0].x].counters[3], 1 << 30);   // ... and the overflow flag (host redoes the batch)
    }
}

// decision of the grouped pass: exact statistics of the evaluated rows + the skip guarantee + the seed's bound
__global__ void match_group_decide_kernel(const MatchJob* __restrict__ jobs, const int* __restrict__ which) {
    const MatchJob& J = jobs[which[blockIdx.y]];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= J.NB) return;
    const int* s6 = J.stat6 + 6 * (size_t)b;
    const int i0 = s6[0], u0 = s6[1], u1x = s6[3], l0 = s6[4], l1 = s6[5];   // keys are (value << 32 | row): little-endian halves
    const int eb = J.Be[b];
    SadStat s;
    s.u0 = u0;
    s.u1 = min(u1x, J.seed_u1[b]);                  // each is the second smallest over a set of distinct rows
    const int tau = sad_tau(s.u1, eb);              // every skipped row has SAD - e(a) >= tau(u1 at that time) >= tau(final u1)
    s.lbmin = min(l0, tau);
    int thr = 0;
    J.idx[b] = -1;
    if (sad_certain_reject(s, eb, &thr)) return;
    if (u0 < (1 << 24) && sad_certain_accept(u0, J.Ae[i0], l0, l1, tau, eb)) {
        J.idx[b] = i0;
        atomicAdd(&J.counters[3], 1);
        return;
    }
    const int slot = atomicAdd(&J.counters[0], 1);
    J.surv[slot] = b;
    J.thr[slot] = thr;
    J.cand_cnt[slot] = 0;
}

// merges the splits' statistics, rejects what can be rejected with certainty, compacts the rest
__global__ void match_sad_decide_kernel(const MatchJob* __restrict__ jobs, const int* __restrict__ which) {
    const MatchJob& J = jobs[which ? which[blockIdx.y] : blockIdx.y];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= J.NB) return;
    SadStat s = J.spartial[b];
    for (int k = 1; k < J.sad_nsplit; ++k) sadstat_merge(s, J.spartial[(size_t)k * J.NB + b]);
    int thr = 0;
    J.idx[b] = -1;
    if (sad_certain_reject(s, J.Be[b], &thr)) return;
    const int slot = atomicAdd(&J.counters[0], 1);
    J.surv[slot] = b;
    J.thr[slot] = thr;
    J.cand_cnt[slot] = 0;
}

// one warp per surviving query: lane = candidate row, exact sequential float L1 (the reference's arithmetic), top-2
// over the lanes, ratio rule.  Lists that overflowed go to the full scan.
__global__ void __launch_bounds__(128) match_exact_kernel(const MatchJob* __restrict__ jobs) {
    const MatchJob& J = jobs[blockIdx.y];
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (slot >= J.counters[0]) return;
    const int n = J.cand_cnt[slot], b = J.surv[slot];
    if (n > kMatchCand) {
        if (lane == 0) J.overflow[atomicAdd(&J.counters[1], 1)] = b;
        return;
    }
    Top2 t = top2_init();
    if (lane < n) {
        const int a = J.cand[(size_t)slot * kMatchCand + lane];
        const float4* __restrict__ qa = reinterpret_cast<const float4*>(J.B + (size_t)b * 128);
        const float4* __restrict__ ra = reinterpret_cast<const float4*>(J.A + (size_t)a * 128);
        float acc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float4 x = qa[k], y = ra[k];
            acc += fabsf(x.x - y.x); acc += fabsf(x.y - y.y); acc += fabsf(x.z - y.z); acc += fabsf(x.w - y.w);
        }
        t.d0 = acc; t.i0 = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Top2 u;
        u.d0 = __shfl_xor_sync(0xffffffffu, t.d0, o);
        u.d1 = __shfl_xor_sync(0xffffffffu, t.d1, o);
        u.i0 = __shfl_xor_sync(0xffffffffu, t.i0, o);
        top2_merge(t, u);
    }
    if (lane == 0) J.idx[b] = (J.NA >= 2 && t.i0 >= 0 && ratio_test(t.d0, t.d1)) ? t.i0 : -1;
}

int match_num_splits(int NA, int NB) {
    int qblocks = div_up(NB, kQ);
    int want = div_up(148 * 3, qblocks);          // aim at ~3 CTAs per SM
    int maxs = div_up(NA, 2 * kTA);               // at least two tiles per split
    int s = want < maxs ? want : maxs;
    return s < 1 ? 1 : s;
}

MatchJob make_match_job(const float* dA, int NA, const float* dB, int NB, Top2* partial, int nsplit, int* idx,
                        float* d01) {
    MatchJob J{};
    J.A = dA; J.B = dB; J.NA = NA; J.NB = NB;
    J.rows_per_split = align_up(div_up(NA > 0 ? NA : 1, nsplit > 0 ? nsplit : 1), 4);
    J.nsplit = div_up(NA > 0 ? NA : 1, J.rows_per_split);
    J.partial = partial; J.idx = idx; J.d01 = d01;
    return J;
}

void launch_match_batch(const MatchJob* d_jobs, const MatchJob* h_jobs, int njobs, cudaStream_t st) {
    if (njobs <= 0) return;
    int gx = 1, gy = 1;
    double work = 0;
    for (int i = 0; i < njobs; ++i) {
        gx = std::max(gx, div_up(h_jobs[i].NB, kQ));
        gy = std::max(gy, h_jobs[i].nsplit);
        work += 256.0 * h_jobs[i].NA * h_jobs[i].NB;
    }
    {
        KScope ks("match.l1", st, work);
        match_l1_kernel<false><<<dim3(gx, gy, njobs), kQ, 0, st>>>(d_jobs);
        PB_KERNEL_CHECK();
    }
    KScope ks2("match.merge", st, 0);
    match_merge_kernel<false><<<dim3(div_up(gx * kQ, 128), njobs), 128, 0, st>>>(d_jobs);
    PB_KERNEL_CHECK();
}

int match_sad_num_splits(int NA, int NB, int njobs) {
    const int qblocks = div_up(NB, kST * kSQ) * (njobs > 0 ? njobs : 1);
    const int want = div_up(148 * 6, qblocks);            // ~6 CTAs per SM over the whole batch
    const int maxs = div_up(NA, 2 * kSRows);              // at least two tiles per split
    const int s = want < maxs ? want : maxs;
    return s < 1 ? 1 : s;
}

size_t match_prefilter_ints(int NB) { return (size_t)align_up(NB, 4) * (4 + kMatchCand); }

void match_prefilter_attach(MatchJob& J, const unsigned* A8, const int* Ae, const unsigned* B8, const int* Be,
                            int sad_nsplit, SadStat* spartial, int* scratch, int* counters) {
    J.A8 = A8; J.Ae = Ae; J.B8 = B8; J.Be = Be;
    J.sad_rows_per_split = align_up(div_up(J.NA > 0 ? J.NA : 1, sad_nsplit > 0 ? sad_nsplit : 1), 2);
    J.sad_nsplit = div_up(J.NA > 0 ? J.NA : 1, J.sad_rows_per_split);
    const int cns = std::max(J.sad_nsplit, std::min(8, div_up(J.NA > 0 ? J.NA : 1, 4 * kSRows)));
    J.cand_rows_per_split = align_up(div_up(J.NA > 0 ? J.NA : 1, cns), 2);
    J.cand_nsplit = div_up(J.NA > 0 ? J.NA : 1, J.cand_rows_per_split);
    J.spartial = spartial;
    const size_t n = (size_t)align_up(J.NB, 4);
    J.counters = counters;
    J.surv = scratch;
    J.thr = J.surv + n;
    J.cand_cnt = J.thr + n;
    J.overflow = J.cand_cnt + n;
    J.cand = J.overflow + n;
}

int match_prefilter_pair(MatchJob& F, MatchJob& R) {
    // R's statistics arrive as one entry per (block of held rows of Y, row of X): its "splits" are the y blocks
    R.sad_nsplit = div_up(F.NB, kST * kSQ);
    return R.sad_nsplit;
}
int match_sym_err_cap() { return kSymErrCap; }
int match_sym_yblocks(int NY) { return div_up(NY, kST * kSQ); }

size_t match_group_ints(int NB) { return (size_t)align_up(NB, 4) * 7 + match_group_pad_rows(NB) / 2 + 16; }
void match_group_attach(MatchJob& J, const unsigned* Ag, const unsigned short* Aw, const unsigned* Bg, const unsigned short* Bw,
                        int nsplit, int* scratch) {
    J.Ag = Ag; J.Aw = Aw; J.Bg = Bg; J.Bw = Bw;
    J.grp_rows_per_split = align_up(div_up(J.NA > 0 ? J.NA : 1, nsplit > 0 ? nsplit : 1), kGRows);
    J.grp_nsplit = div_up(J.NA > 0 ? J.NA : 1, J.grp_rows_per_split);
    const size_t n4 = (size_t)align_up(J.NB, 4);
    J.stat6 = scratch;                              // 8-byte aligned keys: the scratch of a job starts 16-byte aligned
    J.seed_u1 = scratch + 6 * n4;
    J.c16 = reinterpret_cast<unsigned short*>(scratch + 7 * n4);
}
int match_group_yblocks(int NY) { return div_up(NY, 128 * kGY); }
int match_group_num_splits(int NA, int yblocks_total) {
    // ~8 waves of the 6 CTAs an SM holds (80 registers): with one split the 8 x 4K batch is 1400 CTAs = 1.6 waves, and the
    // half-empty second wave costs ~20 % of the pass
    const int want = div_up(148 * 6 * 8, yblocks_total > 0 ? yblocks_total : 1);
    const int maxs = div_up(NA, 4 * kGRows);                                   // at least four tiles per split
    const int s = want < maxs ? want : maxs;
    return s < 1 ? 1 : s;
}
int match_group_err_cap() { return kGrpErrCap; }

void launch_match_batch_prefilter(const MatchJob* d_jobs, const MatchJob* h_jobs, int njobs, const int2* d_pairs,
                                  const int2* h_pairs, int npairs, const int* d_singles, const int* h_singles, int nsingles,
                                  cudaStream_t st, bool grouped, unsigned long long* gq_items, unsigned long long* gq_count,
                                  size_t gq_cap) {
    if (njobs <= 0) return;
    int nbmax = 1, fy = 1, cy = 1;
    for (int i = 0; i < njobs; ++i) {
        nbmax = std::max(nbmax, h_jobs[i].NB);
        fy = std::max(fy, h_jobs[i].nsplit);
        cy = std::max(cy, h_jobs[i].cand_nsplit);
    }
    const int sx = div_up(nbmax, kST * kSQ);
    const int* d_pairjobs = reinterpret_cast<const int*>(d_pairs);   // the 2 * npairs job indices of the pair list
    if (npairs > 0 && grouped) {
        int px = 1, py = 1, nb = 1;
        double pairs = 0, seed = 0;
        for (int i = 0; i < npairs; ++i) {
            const MatchJob& F = h_jobs[h_pairs[i].x];
            px = std::max(px, match_group_yblocks(F.NB));
            py = std::max(py, F.grp_nsplit);
            nb = std::max(nb, std::max(F.NA, F.NB));
            pairs += (double)F.NA * F.NB;
            seed += 32.0 * std::min(F.NA, 2 * seed_rows(F.NA)) * F.NB + 32.0 * std::min(F.NB, 2 * seed_rows(F.NB)) * F.NA;
        }
        {
            KScope ks("match.seed", st, seed);
            match_seed_kernel<<<dim3(div_up(nb, 128), 1, 2 * npairs), 128, 0, st>>>(d_jobs, d_pairjobs);
            PB_KERNEL_CHECK();
        }
        {
            KScope ks("match.group_sym", st, 8.0 * pairs);   // VABSDIFF4 thread-instructions of the grouped bound
            match_group_sym_kernel<<<dim3(px, py, npairs), 128, 0, st>>>(d_jobs, d_pairs, GroupQueue{gq_items, gq_count, (unsigned long long)gq_cap});
            PB_KERNEL_CHECK();
        }
        {
            KScope ks("match.group_exact", st, 0);
            match_group_exact_kernel<<<148 * 8, 256, 0, st>>>(d_jobs, d_pairs, GroupQueue{gq_items, gq_count, (unsigned long long)gq_cap});
            PB_KERNEL_CHECK();
        }
        KScope ks("match.decide", st, 0);
        match_group_decide_kernel<<<dim3(div_up(nb, 128), 2 * npairs), 128, 0, st>>>(d_jobs, d_pairjobs);
        PB_KERNEL_CHECK();
    } else if (npairs > 0) {   // both directions of an image pair from one pass
        int px = 1, py = 1;
        double pairs = 0;
        for (int i = 0; i < npairs; ++i) {
            const MatchJob& F = h_jobs[h_pairs[i].x];
            px = std::max(px, div_up(F.NB, kST * kSQ));
            py = std::max(py, F.sad_nsplit);
            pairs += (double)F.NA * F.NB;
        }
        {
            KScope ks("match.sad_sym", st, 32.0 * pairs);   // VABSDIFF4 thread-instructions (they serve two directed problems)
            match_sad_sym_kernel<<<dim3(px, py, npairs), kST, 0, st>>>(d_jobs, d_pairs);
            PB_KERNEL_CHECK();
        }
        KScope ks("match.decide", st, 0);
        match_sad_decide_kernel<<<dim3(div_up(nbmax, 128), 2 * npairs), 128, 0, st>>>(d_jobs, d_pairjobs);
        PB_KERNEL_CHECK();
    }
    if (nsingles > 0) {
        int qx = 1, qy = 1;
        double pairs = 0;
        for (int i = 0; i < nsingles; ++i) {
            const MatchJob& J = h_jobs[h_singles[i]];
            qx = std::max(qx, div_up(J.NB, kST * kSQ));
            qy = std::max(qy, J.sad_nsplit);
            pairs += (double)J.NA * J.NB;
        }
        {
            KScope ks("match.sad", st, 32.0 * pairs);   // VABSDIFF4 thread-instructions
            match_sad_kernel<0><<<dim3(qx, qy, nsingles), kST, 0, st>>>(d_jobs, d_singles);
            PB_KERNEL_CHECK();
        }
        KScope ks("match.decide", st, 0);
        match_sad_decide_kernel<<<dim3(div_up(nbmax, 128), nsingles), 128, 0, st>>>(d_jobs, d_singles);
        PB_KERNEL_CHECK();
    }
    {   // grids are sized for "every query survives"; CTAs beyond the survivor count leave at once
        KScope ks("match.cand", st, 0);
        match_sad_kernel<1><<<dim3(sx, cy, njobs), kST, 0, st>>>(d_jobs, nullptr);
        PB_KERNEL_CHECK();
    }
    {
        KScope ks("match.exact", st, 0);
        match_exact_kernel<<<dim3(div_up((long)nbmax * 32, 128), njobs), 128, 0, st>>>(d_jobs);
        PB_KERNEL_CHECK();
    }
    {   // full float scan of the queries whose candidate list overflowed: short lists (the usual case) with fine splits
        KScope ks("match.l1", st, 0);
        match_l1_kernel<true, true><<<dim3(div_up(div_up(nbmax, kListFine), kQ), fy * kListFine, njobs), kQ, 0, st>>>(d_jobs);
        PB_KERNEL_CHECK();
        match_l1_kernel<true, false><<<dim3(div_up(nbmax, kQ), fy, njobs), kQ, 0, st>>>(d_jobs);
        PB_KERNEL_CHECK();
    }
    KScope ks2("match.merge", st, 0);
    match_merge_kernel<true><<<dim3(div_up(nbmax, 128), njobs), 128, 0, st>>>(d_jobs);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// RANSAC: one warp per (problem, hypothesis).  Every lane solves the two 4x4 systems redundantly (the hypothesis
// stays in registers), lanes then stride over the pairs and vote with __ballot_sync; the inlier set is kept as a
// bit mask so the host can apply the reference's "first strictly larger set wins" rule and refit.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ransac_score_kernel(const KeyPair* __restrict__ pairs,
                                                           const int* __restrict__ pair_off,
                                                           const int* __restrict__ samples, int nproblems,
                                                           int iters, int* __restrict__ counts, unsigned* __restrict__ masks,
                                                           int words_stride, double* __restrict__ hyp) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int p = warp / iters, k = warp - p * iters;
    if (p >= nproblems) return;
    const int off = pair_off[p], n = pair_off[p + 1] - off;
    const KeyPair* pp = pairs + off;
    const int* s = samples + ((size_t)p * iters + k) * 4;
    float sx[4], sy[4], dx[4], dy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const KeyPair& c = pp[s[i]];
        sx[i] = c.src.x; sy[i] = c.src.y; dx[i] = c.dst.x; dy[i] = c.dst.y;
    }
    double H[8];
    fit4(sx, sy, dx, dy, H);
    if (hyp && lane == 0)
        for (int i = 0; i < 8; ++i) hyp[((size_t)p * iters + k) * 8 + i] = H[i];
    unsigned* m = masks + ((size_t)p * iters + k) * words_stride;
    int cnt = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        bool in = false;
        if (i < n) in = is_inlier(H, pp[i].src.x, pp[i].src.y, pp[i].dst.x, pp[i].dst.y);
        unsigned bal = __ballot_sync(0xffffffffu, in);
        cnt += __popc(bal);
        if (lane == 0) m[base >> 5] = bal;
    }
    if (lane == 0) counts[(size_t)p * iters + k] = cnt;
}

void launch_ransac_score(const KeyPair* pairs, const int* pair_off, int nproblems, const int* samples, int iters,
                         int* counts, unsigned* masks, int words_stride, double* hyp, cudaStream_t st) {
    if (nproblems <= 0) return;
    KScope ks("ransac.score", st, 0);
    long warps = (long)nproblems * iters;
    // 4 warps per CTA; iters (72) is a multiple of 4 so a CTA never straddles two problems' tail
    long threads = warps * 32;
    ransac_score_kernel<<<div_up(threads, 128), 128, 0, st>>>(pairs, pair_off, samples, nproblems, iters, counts,
                                                              masks, words_stride, hyp);
    PB_KERNEL_CHECK();
}

}  // namespace pb
