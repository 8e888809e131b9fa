// match_i8_kernels.cu -- north-star stage (3): brute-force 2-NN on uint8-quantised descriptors as a dense integer
// contraction on the 5th-generation tensor cores (tcgen05.mma.kind::i8, accumulators in TMEM), with the top-2
// selection fused into the epilogue.
//
// This matcher has NO counterpart in the reference, which matches float descriptors under L1 (ImageProcess.cpp:
// 273-351); the exact matcher of the stitching pipeline is match_kernels.cu.  Here descriptors are
// quantised with VLFeat's own convention q = (uint8) min(512 x, 255) (the one its CLI / MATLAB drivers use) and
// compared under squared L2:  d2(q, a) = |q|^2 + |a|^2 - 2 q.a,  ratio rule 4 d0 < d1  (<=> sqrt(d0)/sqrt(d1) < 0.5).
// Everything is integer arithmetic (u8 x u8 -> s32), so distances and indices are exact; the oracle is
// oracle/match_u8_oracle.py.
//
// Kernel shape (one CTA = 128 queries x one slice of the database, 640 threads):
//   warp 0      producer: one lane streams 256-row database tiles HBM -> shared memory with ONE 32 KB TMA bulk copy
//               (cp.async.bulk) per tile, 4 stages, completion by mbarrier transaction bytes.  The quantised tables
//               are kept in HBM in the UMMA operand layout already ("blocked256": per block of 256 rows, chunk c
//               (16 bytes of K) of row r at c * 4096 + r * 16), so a tile is one contiguous 32 KB run and lands as
//               the canonical K-major no-swizzle layout (8-row x 16-byte core matrices) without any address math.
//               (A first version gathered row-major tables with 8 TMA tensor boxes of {16 B, 256 rows} per tile:
//               2048 16-byte requests per tile kept the TMA unit, not the tensor core, busy.)
//   warp 1      lane 0 issues 4 x tcgen05.mma (M128 x N256 x K32) per tile into one of two 256-column TMEM
//               accumulators; tcgen05.commit releases the shared-memory stage and publishes the accumulator
//   warps 4-19  epilogue: thread = (query row = TMEM lane, quarter of the tile's columns); tcgen05.ld 32 columns at a
//               time, two loads in flight; key = (|a|^2 - 2 q.a) * 256 + column with ONE integer multiply-add per
//               element and a branch-free 3-input min tree per 64-row unit (see the comment in the epilogue);
//               the second-nearest row inside the winning unit is recomputed by match_u8_finish_kernel
// Queries sit on the MMA's M side so that a thread owns a query and scans database columns: the top-2 needs no
// cross-thread reduction.  Ties between equal distances cannot change an accepted match (a tie of the two best
// fails the ratio rule), so the column packed in the key is only a payload.
#include "match_i8_kernels.h"
#include "common.h"
#include "ktimer.h"
#include <algorithm>
#include <climits>

namespace pb {

namespace {

constexpr int kMQ = 128;       // queries per MMA (UMMA M)
constexpr int kQB = 2;         // query blocks per CTA: every database tile in shared memory feeds kQB MMAs
constexpr int kND = 128;       // database rows per MMA (UMMA N)
constexpr int kStages = 6;     // shared-memory stages of database tiles
constexpr int kUnit = 64;      // database rows per epilogue unit (a quarter of a tile)
constexpr int kRowBytes = 128; // one descriptor
// shared-memory layout of a tile of R rows: chunk c (16 bytes of K) of row r at c * (R * 16) + r * 16, i.e. for every
// chunk the rows are contiguous: 8 rows x 16 B = one UMMA core matrix, 8-row groups 128 B apart (SBO), chunks R*16 B
// apart (LBO).  That is exactly what one TMA box {16 bytes, R rows} writes.
constexpr int kSbo = 128;
constexpr int kPadNorm = 0x7fffff;           // |a|^2 of a padding row: its key exceeds every real key

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy: `bytes` contiguous bytes HBM -> shared memory, completion counted on the mbarrier
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes, contiguous (128 B);
// LBO = distance between the two 16-byte K chunks of one K=32 step, SBO = distance between 8-row groups.
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;   // descriptor version for sm_100
    return d;          // base_offset = 0, lbo_mode = 0, layout_type = 0 (no swizzle)
}
// instruction descriptor: D = s32, A = B = unsigned 8-bit, both K-major, N = 256, M = 128
constexpr unsigned kIdesc = (2u << 4) | (0u << 7) | (0u << 10) | ((unsigned)(kND >> 3) << 17) | ((unsigned)(kMQ >> 4) << 24);

__device__ __forceinline__ void umma_i8(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

struct __align__(16) SmemLayout {
    unsigned char q[kQB][kMQ * kRowBytes];            // 2 x 16 KB: the CTA's queries (operand A), two blocks of 128
    unsigned char db[kStages][kND * kRowBytes];       // 6 x 16 KB: database tiles (operand B)
    int cstw[16][2][kUnit];                           // per epilogue warp, per accumulator buffer: |a|^2 * 256 + column
    unsigned long long full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], qfull;
    unsigned tmem_base;
    int mrg[kQB * kMQ][2][3];                         // merge of the two column halves of the epilogue
};

}  // namespace

__global__ void __launch_bounds__(640, 1)
match_u8_kernel(const unsigned char* __restrict__ Ablk, const int* __restrict__ normA, int NA,
                const unsigned char* __restrict__ Bblk, const int* __restrict__ normB, int NB, int rows_per_split,
                U8Top2* __restrict__ partial) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SmemLayout& S = *reinterpret_cast<SmemLayout*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (kQB * kMQ);
    const int a_begin = blockIdx.y * rows_per_split;
    const int a_end = min(NA, a_begin + rows_per_split);
    const int ntiles = (a_end - a_begin + kND - 1) / kND;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&S.tmem_full[b], 1); mbar_init(&S.tmem_empty[b], 16); }
        mbar_init(&S.qfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // one warp allocates all 512 TMEM columns (two 256-column accumulators) and later frees them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = S.tmem_base;

    if (warp == 0) {
        // ------------------------------------------------ producer ------------------------------------------------
        if (lane == 0) {
            // the CTA's queries (operand A of the MMA): 8 boxes of {16 B, 128 rows}
            // the CTA's 256 queries = one block of the blocked256 table: query block qb = rows qb*128 .. +127 of it
            mbar_expect_tx(&S.qfull, kQB * kMQ * kRowBytes);
            const unsigned char* qsrc = Bblk + (size_t)(q0 / 256) * 32768;
#pragma unroll
            for (int qb = 0; qb < kQB; ++qb)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    tma_bulk_load(S.q[qb] + c * (kMQ * 16), qsrc + c * 4096 + qb * (kMQ * 16), kMQ * 16, &S.qfull);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages;
                if (t >= kStages) mbar_wait(&S.empty[s], (unsigned)(((t / kStages) - 1) & 1));
                // tile t = rows a_begin + t*128 .. +127 = half (t & 1) of block (a_begin / 256 + t / 2): 8 runs of 2 KB
                mbar_expect_tx(&S.full[s], kND * kRowBytes);
                const unsigned char* src = Ablk + (size_t)(a_begin / 256 + (t >> 1)) * 32768 + (size_t)(t & 1) * (kND * 16);
#pragma unroll
                for (int c = 0; c < 8; ++c) tma_bulk_load(S.db[s] + c * (kND * 16), src + c * 4096, kND * 16, &S.full[s]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        if (lane == 0) {
            const unsigned qa = smem_u32(S.q[0]);
            mbar_wait(&S.qfull, 0);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages, b = t & 1;
                mbar_wait(&S.full[s], (unsigned)((t / kStages) & 1));
                if (t >= 2) mbar_wait(&S.tmem_empty[b], (unsigned)(((t >> 1) - 1) & 1));
                tc_fence_after();
                const unsigned ba = smem_u32(S.db[s]);
#pragma unroll
                for (int qb = 0; qb < kQB; ++qb)
#pragma unroll
                    for (int j = 0; j < 4; ++j)   // K = 128 = 4 steps of 32 bytes = chunks 2j, 2j+1
                        umma_i8(tmem + (unsigned)((b * kQB + qb) * kND),
                                umma_desc(qa + qb * (kMQ * kRowBytes) + j * 2 * (kMQ * 16), kMQ * 16, kSbo),
                                umma_desc(ba + j * 2 * (kND * 16), kND * 16, kSbo), j > 0 ? 1u : 0u);
                umma_commit(&S.empty[s]);        // the stage may be refilled once these MMAs have read it
                umma_commit(&S.tmem_full[b]);    // the accumulator is complete
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue ------------------------------------------------
        // 16 warps: warp w reads TMEM lanes 32 * (w % 4) .. +31 (its hardware lane quarter); warps 4-7 scan columns
        // 0..63 of every tile, warps 8-11 columns 64..127, and so on.  A (tile, quarter) of 64 database rows is a UNIT.
        // Per unit and query the thread computes only the unit MINIMUM of the packed keys (64 IMAD + a 3-input min
        // tree, no branches) and keeps, over its units, the best key, the unit it came from, and the second smallest unit
        // minimum.  The overall second-nearest row is either another unit's minimum (tracked here) or the second
        // smallest row INSIDE the best unit, which match_u8_finish_kernel recomputes exactly for that one unit.
        const int ew = warp & 3;
        const int part = ((warp - 4) >> 2) & 1;        // which 64-column half of the tile
        const int qb = (warp - 4) >> 3;                // which query block (accumulator) of the CTA
        const int qrow = qb * kMQ + ew * 32 + lane;    // query row within the CTA; TMEM lane = ew * 32 + lane
        int* cw = S.cstw[warp - 4][0];                 // this warp's private constants, double-buffered per tile
        int m1 = INT_MAX, s2 = INT_MAX, bestunit = 0;
        // |a|^2 of the 2 columns this lane prepares for the warp (columns lane*2, lane*2+1 of the warp's quarter)
        int2 nrm_next;
        {
            const int r0 = a_begin + part * kUnit + lane * 2;
            nrm_next.x = r0 + 0 < a_end ? normA[r0 + 0] : kPadNorm;
            nrm_next.y = r0 + 1 < a_end ? normA[r0 + 1] : kPadNorm;
        }
        for (int t = 0; t < ntiles; ++t) {
            const int b = t & 1;
            int* cst = cw + b * kUnit;
            {   // constants of this unit: |a|^2 * 256 + column-in-tile (padding rows: a key above every real key);
                // the norms of the next tile are fetched now so that their latency hides behind this tile's scan
                const int c0 = part * kUnit + lane * 2;
                *reinterpret_cast<int2*>(&cst[lane * 2]) = make_int2(nrm_next.x * 256 + c0, nrm_next.y * 256 + c0 + 1);
                const int r0 = a_begin + (t + 1) * kND + c0;
                const bool more = t + 1 < ntiles;
                nrm_next.x = more && r0 + 0 < a_end ? normA[r0 + 0] : kPadNorm;
                nrm_next.y = more && r0 + 1 < a_end ? normA[r0 + 1] : kPadNorm;
            }
            __syncwarp();
            mbar_wait(&S.tmem_full[b], (unsigned)((t >> 1) & 1));
            tc_fence_after();
            const unsigned taddr = tmem + ((unsigned)(ew * 32) << 16) + (unsigned)((b * kQB + qb) * kND + part * kUnit);
            int va[32], vb[32];
#define PB_LDTM(v, col)                                                                                                      \
    asm volatile(                                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "    \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                            \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),       \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),           \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),          \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                        \
        : "r"(taddr + (unsigned)(col)))
// unit minimum of key = cst - 512 * dot over 32 columns: 32 IMAD + 16 VIMNMX3
#define PB_SCAN(v, col)                                                                                                      \
    _Pragma("unroll") for (int g = 0; g < 32; g += 4) {                                                                      \
        const int4 cc = *reinterpret_cast<const int4*>(&cst[(col) + g]);                                                     \
        const int k0 = v[g] * -512 + cc.x, k1 = v[g + 1] * -512 + cc.y, k2 = v[g + 2] * -512 + cc.z,                       \
                  k3 = v[g + 3] * -512 + cc.w;                                                                               \
        umin = __vimin3_s32(umin, k0, k1);                                                                                   \
        umin = __vimin3_s32(umin, k2, k3);                                                                                   \
    }
            int umin = INT_MAX;
            PB_LDTM(va, 0);
            PB_LDTM(vb, 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.tmem_empty[b]);   // the accumulator quarter is in registers: release it early
            PB_SCAN(va, 0);
            PB_SCAN(vb, 32);
#undef PB_LDTM
#undef PB_SCAN
            s2 = min(s2, max(umin, m1));
            if (umin < m1) { m1 = umin; bestunit = t * 2 + part; }
        }
        // merge the two column halves of every query row and write (best distance, its row, 2nd smallest unit minimum)
        if (part != 0) { S.mrg[qrow][1][0] = m1; S.mrg[qrow][1][1] = s2; S.mrg[qrow][1][2] = bestunit; }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (part == 0) {
            int best = m1, bu = bestunit, second = s2;
            {
                const int o1 = S.mrg[qrow][1][0], o2 = S.mrg[qrow][1][1], ou = S.mrg[qrow][1][2];
                second = min(second, o2);
                if (o1 != INT_MAX && (best == INT_MAX || (o1 >> 8) < (best >> 8))) { second = min(second, best); best = o1; bu = ou; }
                else second = min(second, o1);
            }
            const int q = q0 + qrow;
            if (q < NB) {
                const int nq = normB[q];
                U8Top2 r;
                r.d0 = best == INT_MAX ? INT_MAX : (best >> 8) + nq;
                r.d1 = second == INT_MAX ? INT_MAX : (second >> 8) + nq;
                r.i0 = best == INT_MAX ? -1 : a_begin + (bu >> 1) * kND + (best & 255);
                partial[(size_t)blockIdx.y * NB + q] = r;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// Merge of the database splits + the exact second-nearest distance.  The main kernel reports, per (query, split), the
// nearest row, its distance, and the second smallest UNIT minimum (a unit = 64 consecutive database rows).  The true
// second-nearest distance is the smaller of that value (merged over splits, with the nearest rows of the losing splits)
// and the second smallest distance INSIDE the winning unit, which one warp recomputes here from the tables
// (64 rows x 128 bytes, dp4a).  Ratio rule sqrt(d0) / sqrt(d1) < 0.5  <=>  4 d0 < d1, exact in integers.
__global__ void __launch_bounds__(128) match_u8_finish_kernel(const U8Top2* __restrict__ partial, int nsplit,
                                                              const unsigned char* __restrict__ Ablk, const int* __restrict__ normA,
                                                              int NA, const unsigned char* __restrict__ Bblk,
                                                              const int* __restrict__ normB, int NB, int* __restrict__ idx,
                                                              int* __restrict__ d01) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= NB) return;
    int d0 = INT_MAX, d1 = INT_MAX, i0 = -1;
    for (int s = 0; s < nsplit; ++s) {
        const U8Top2 p = partial[(size_t)s * NB + q];
        if (p.i0 < 0) continue;
        if (p.d0 < d0) { d1 = min(min(d0, d1), p.d1); d0 = p.d0; i0 = p.i0; }
        else d1 = min(d1, min(p.d0, p.d1));
    }
    if (i0 >= 0) {
        // the query (128 bytes) in registers: 8 chunks of 16 bytes
        uint4 qv[8];
        const unsigned char* qp = Bblk + (size_t)(q >> 8) * 32768 + (size_t)(q & 255) * 16;
#pragma unroll
        for (int c = 0; c < 8; ++c) qv[c] = *reinterpret_cast<const uint4*>(qp + c * 4096);
        const int nq = normB[q];
        const int u0 = (i0 / kUnit) * kUnit;   // first row of the winning unit (units are kUnit-row aligned)
        int local = INT_MAX;
#pragma unroll 1
        for (int k = 0; k < kUnit / 32; ++k) {
            const int row = u0 + k * 32 + lane;
            if (row < NA && row != i0) {
                const unsigned char* ap = Ablk + (size_t)(row >> 8) * 32768 + (size_t)(row & 255) * 16;
                unsigned dot = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 av = *reinterpret_cast<const uint4*>(ap + c * 4096);
                    dot = __dp4a(av.x, qv[c].x, dot); dot = __dp4a(av.y, qv[c].y, dot);
                    dot = __dp4a(av.z, qv[c].z, dot); dot = __dp4a(av.w, qv[c].w, dot);
                }
                local = min(local, nq + normA[row] - 2 * (int)dot);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local = min(local, __shfl_xor_sync(0xffffffffu, local, o));
        d1 = min(d1, local);
    }
    if (lane == 0) {
        const bool ok = NA >= 2 && i0 >= 0 && d1 != INT_MAX && 4ll * d0 < (long long)d1;
        idx[q] = ok ? i0 : -1;
        if (d01) { d01[3 * q] = d0; d01[3 * q + 1] = d1; d01[3 * q + 2] = i0; }
    }
}

// VLFeat's uint8 descriptor convention: q = (uint8) min(512 x, 255); also |q|^2
__global__ void quantize_u8_kernel(const float* __restrict__ src, int n, unsigned char* __restrict__ dst, int* __restrict__ norm) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const float4 v = reinterpret_cast<const float4*>(src + (size_t)row * 128)[lane];
    const float f[4] = {v.x, v.y, v.z, v.w};
    unsigned packed = 0;
    int nrm = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float x = 512.0f * f[k];
        x = (x < 255.0f) ? x : 255.0f;
        const unsigned q = (unsigned)(unsigned char)x;
        packed |= q << (8 * k);
        nrm += (int)(q * q);
    }
    // blocked256 layout: block = row / 256, chunk = lane / 4 (16 bytes of K), 4 bytes at (lane % 4) * 4 inside the chunk
    *reinterpret_cast<unsigned*>(dst + (size_t)(row >> 8) * 32768 + (lane >> 2) * 4096 + (row & 255) * 16 + (lane & 3) * 4) = packed;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) norm[row] = nrm;
}
__global__ void norm_u8_kernel(const unsigned char* __restrict__ src, int n, int* __restrict__ norm) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const unsigned p = *reinterpret_cast<const unsigned*>(src + (size_t)(row >> 8) * 32768 + (lane >> 2) * 4096 + (row & 255) * 16 + (lane & 3) * 4);
    int nrm = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int q = (p >> (8 * k)) & 255; nrm += q * q; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) norm[row] = nrm;
}

// row-major [n][128] u8 <-> blocked256 (rows padded to a multiple of 256 must have been zeroed by the caller)
__global__ void relayout_u8_kernel(const unsigned char* __restrict__ src, int n, unsigned char* __restrict__ dst, int to_blocked) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte chunk per thread
    if (i >= (long)n * 8) return;
    const int row = (int)(i >> 3), c = (int)(i & 7);
    const size_t rm = (size_t)row * 128 + c * 16, bl = (size_t)(row >> 8) * 32768 + c * 4096 + (row & 255) * 16;
    if (to_blocked) *reinterpret_cast<uint4*>(dst + bl) = *reinterpret_cast<const uint4*>(src + rm);
    else *reinterpret_cast<uint4*>(dst + rm) = *reinterpret_cast<const uint4*>(src + bl);
}
void launch_relayout_u8(const unsigned char* src, int n, unsigned char* dst, bool to_blocked, cudaStream_t st) {
    if (n <= 0) return;
    relayout_u8_kernel<<<div_up((long)n * 8, 256), 256, 0, st>>>(src, n, dst, to_blocked ? 1 : 0);
    PB_KERNEL_CHECK();
}

void launch_quantize_u8(const float* src, int n, unsigned char* dst, int* norm, cudaStream_t st) {
    if (n <= 0) return;
    KScope ks("match_u8.quantize", st, 644.0 * n);
    quantize_u8_kernel<<<div_up(n, 8), 256, 0, st>>>(src, n, dst, norm);
    PB_KERNEL_CHECK();
}
void launch_norm_u8(const unsigned char* src, int n, int* norm, cudaStream_t st) {
    if (n <= 0) return;
    norm_u8_kernel<<<div_up(n, 8), 256, 0, st>>>(src, n, norm);
    PB_KERNEL_CHECK();
}

int match_u8_num_splits(int NA, int NB) {
    const int qtiles = div_up(NB, kQB * kMQ);
    int want = div_up(148, qtiles);                    // one CTA per SM (each CTA owns all 512 TMEM columns)
    const int maxs = std::max(1, div_up(NA, 1024));    // at least 1024 database rows per split
    want = std::min(want, maxs);
    return std::max(1, want);
}

void launch_match_u8(const unsigned char* dA, const int* normA, int NA, const unsigned char* dB, const int* normB, int NB,
                     U8Top2* partial, int nsplit, int* idx, int* d01, cudaStream_t st) {
    if (NB <= 0) return;
    if (NA <= 0) {
        PB_CUDA(cudaMemsetAsync(idx, 0xff, sizeof(int) * NB, st));
        return;
    }
    static bool attr_set = false;
    const int smem = (int)sizeof(SmemLayout) + 1024;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(match_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    int rps = align_up(div_up(NA, nsplit), 256);   // splits start on block boundaries of the blocked256 layout
    nsplit = div_up(NA, rps);
    {
        KScope ks("match_u8.mma", st, 2.0 * 128.0 * (double)NA * (double)NB);
        match_u8_kernel<<<dim3(div_up(NB, kQB * kMQ), nsplit), 640, smem, st>>>(dA, normA, NA, dB, normB, NB, rps, partial);
        PB_KERNEL_CHECK();
    }
    KScope ks2("match_u8.finish", st, 0);
    match_u8_finish_kernel<<<div_up(NB, 4), 128, 0, st>>>(partial, nsplit, dA, normA, NA, dB, normB, NB, idx, d01);
    PB_KERNEL_CHECK();
}

}  // namespace pb
