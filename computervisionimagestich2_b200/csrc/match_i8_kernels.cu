// match_i8_kernels.cu -- north-star stage (3): brute-force 2-NN on uint8-quantised descriptors as a dense integer
// contraction on the 5th-generation tensor cores (tcgen05.mma.kind::i8, accumulators in TMEM), with the nearest /
// second-nearest selection fused behind it.
//
// This matcher has NO counterpart in the reference, which matches float descriptors under L1 (ImageProcess.cpp:
// 273-351); the exact matcher of the stitching pipeline is match_kernels.cu.  Here descriptors are
// quantised with VLFeat's own convention q = (uint8) min(512 x, 255) (the one its CLI / MATLAB drivers use) and
// compared under squared L2:  d2(q, a) = |q|^2 + |a|^2 - 2 q.a,  ratio rule 4 d0 < d1  (<=> sqrt(d0)/sqrt(d1) < 0.5).
// Everything is integer arithmetic (u8 x u8 -> s32), so distances and indices are exact; the oracle is
// oracle/match_u8_oracle.py.
//
// Folding the norm into the GEMM.  For one query, ranking database rows by d2 is ranking by s = 2 q.a - |a|^2.  Write
// |a|^2 = 2 h + p (p = parity).  Database rows carry 32 E extra "extension" columns that encode e = hmax - h >= 0 in
// base 255 (32 E - 1 digits weighted 255, one weighted 1) and the queries carry the matching constants (255 ... 255, 1),
// so the tensor core delivers m = q.a + hmax - h directly and s = 2 (m - hmax) - p.  m orders rows exactly like s except
// for ties in m (where the parity decides), so the epilogue needs NO per-element arithmetic at all: it only takes maxima.
//
// Main kernel (one CTA = 256 queries x one slice of the database, 640 threads, all 512 TMEM columns):
//   warp 0      one lane streams 128-row database tiles HBM -> shared memory with TMA bulk copies (cp.async.bulk,
//               mbarrier transaction bytes, 4 stages).  Tables live in HBM in the UMMA operand layout already (see
//               match_i8_kernels.h), so a tile is (8 + 2 E) contiguous 2 KB runs and lands as the canonical K-major
//               no-swizzle layout (8-row x 16-byte core matrices) without any address arithmetic.
//   warp 1      one lane issues tcgen05.mma.cta_group::1.kind::i8 M128 x N128 x K32, (4 + E) per query block and tile
//               (descriptors prebuilt, K loop unrolled: the single issuing thread must not be the bottleneck);
//               two query blocks share every database tile; four 128-column accumulators (2 blocks x double buffer).
//   warps 4-19  epilogue: thread = (query row = TMEM lane, 64 of the tile's 128 columns); tcgen05.ld.32x32b.x32, the
//               accumulator is released as soon as it is in registers; per 32-row UNIT a 3-input max tree (half an
//               instruction per element); per query the four largest unit maxima and the units of the first three.
// match_u8_finish_kernel (one warp per query) merges the database splits, recomputes the exact distances of the rows
// of the best two units with dp4a (nearest row, second nearest inside its unit, nearest of the runner-up unit) and
// applies the ratio rule (the exactness argument is in the kernel); only a three-way tie of unit maxima falls back to
// an exact scan of the whole table for that query.
#include "match_i8_kernels.h"
#include "common.h"
#include "ktimer.h"
#include <algorithm>
#include <climits>

namespace pb {

namespace {

constexpr int kMQ = 128;       // queries per MMA (UMMA M)
constexpr int kQB = 2;         // query blocks per CTA: every database tile in shared memory feeds kQB MMA groups
constexpr int kND = 128;       // database rows per MMA (UMMA N)
constexpr int kBuf = 512 / (kQB * kND);   // TMEM accumulator buffers per query block (all 512 columns in use)
constexpr int kStages = 4;     // shared-memory stages of database tiles
constexpr int kUnit = 32;      // database rows per epilogue unit (one tcgen05.ld)
constexpr int kRun = kND * 16; // one chunk of one tile: kND rows x 16 bytes, contiguous in HBM and in shared memory
constexpr int kQRun = kMQ * 16; // one chunk of one query block
constexpr int kStageBytes = kU8Chunks * kRun;
constexpr int kQBytes = kU8Chunks * kQRun;
constexpr int kSbo = 128;      // distance between 8-row core-matrix groups inside a run

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// instruction descriptor: D = s32, A = B = unsigned 8-bit, both K-major, N = kND, M = 128
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

// the four largest values seen (m1 >= m2 >= m3 >= m4, -1 = none) and the units of the first three
struct Top3 {
    int m1, u1, m2, u2, m3, u3, m4;
};
__device__ __forceinline__ Top3 top3_init() { return Top3{-1, -1, -1, -1, -1, -1, -1}; }
__device__ __forceinline__ void top3_push(Top3& t, int v, int u) {
    if (v <= t.m4) return;   // the common case: one compare
    if (v > t.m1) { t.m4 = t.m3; t.m3 = t.m2; t.u3 = t.u2; t.m2 = t.m1; t.u2 = t.u1; t.m1 = v; t.u1 = u; }
    else if (v > t.m2) { t.m4 = t.m3; t.m3 = t.m2; t.u3 = t.u2; t.m2 = v; t.u2 = u; }
    else if (v > t.m3) { t.m4 = t.m3; t.m3 = v; t.u3 = u; }
    else t.m4 = v;
}

struct __align__(16) SmemLayout {
    unsigned char q[kQB][kQBytes];            // the CTA's queries (operand A): two blocks of 128 rows, incl. constants
    unsigned char db[kStages][kStageBytes];   // database tiles (operand B)
    unsigned long long full[kStages], empty[kStages], tmem_full[kBuf], tmem_empty[kBuf], qfull;
    unsigned tmem_base;
    int mrg[kQB * kMQ][7];                    // hand-over of the second column half of every query row
};

}  // namespace

namespace {
// The issue loop of the MMA lane.  One thread issues every tcgen05.mma of the CTA, so its instruction count per tile
// bounds the tensor-core rate: descriptors are built once (the K-step / stage offsets only touch the 14-bit start
// address field, i.e. an add on the low word) and the K loop is unrolled at compile time.
template <int KSTEPS>
__device__ __forceinline__ void mma_loop(SmemLayout& S, unsigned tmem, int ntiles) {
    unsigned long long da[kQB][KSTEPS], db0[KSTEPS];
#pragma unroll
    for (int j = 0; j < KSTEPS; ++j) {
#pragma unroll
        for (int qb = 0; qb < kQB; ++qb) da[qb][j] = umma_desc(smem_u32(S.q[qb]) + j * 2 * kQRun, kQRun, kSbo);
        db0[j] = umma_desc(smem_u32(S.db[0]) + j * 2 * kRun, kRun, kSbo);
    }
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages, b = t % kBuf;
        mbar_wait(&S.full[s], (unsigned)((t / kStages) & 1));
        if (t >= kBuf) mbar_wait(&S.tmem_empty[b], (unsigned)(((t / kBuf) - 1) & 1));
        tc_fence_after();
        const unsigned long long soff = (unsigned long long)((s * kStageBytes) >> 4);   // stage offset in the address field
#pragma unroll
        for (int qb = 0; qb < kQB; ++qb) {
            const unsigned d = tmem + (unsigned)((b * kQB + qb) * kND);
#pragma unroll
            for (int j = 0; j < KSTEPS; ++j) umma_i8(d, da[qb][j], db0[j] + soff, j > 0 ? 1u : 0u);
        }
        umma_commit(&S.empty[s]);        // the stage may be refilled once these MMAs have read it
        umma_commit(&S.tmem_full[b]);    // the accumulators are complete
    }
}
}  // namespace

__global__ void __launch_bounds__(640, 1)
match_u8_kernel(const unsigned char* __restrict__ Ablk, int NA, const unsigned char* __restrict__ Bblk, int NB,
                int rows_per_split, int ext_steps, U8Top3* __restrict__ partial, int mma_only) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SmemLayout& S = *reinterpret_cast<SmemLayout*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (kQB * kMQ);
    const int a_begin = blockIdx.y * rows_per_split;
    const int a_end = min(NA, a_begin + rows_per_split);
    const int ntiles = (a_end - a_begin + kND - 1) / kND;
    const int nch = 8 + 2 * ext_steps;      // 16-byte K chunks in use
    const int ksteps = 4 + ext_steps;       // MMA K steps of 32 bytes

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        for (int b = 0; b < kBuf; ++b) { mbar_init(&S.tmem_full[b], 1); mbar_init(&S.tmem_empty[b], 16); }
        mbar_init(&S.qfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // one warp allocates all 512 TMEM columns (four 128-column accumulators) and later frees them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the queries' extension columns are constants: 255 everywhere, 1 in the very last column
    for (int i = threadIdx.x; i < kQB * 2 * ext_steps * kMQ; i += blockDim.x) {
        const int qb = i / (2 * ext_steps * kMQ), rem = i - qb * (2 * ext_steps * kMQ);
        const int c = rem / kMQ, r = rem - c * kMQ;
        uint4 v = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (c == 2 * ext_steps - 1) v.w = 0x01ffffffu;
        *reinterpret_cast<uint4*>(S.q[qb] + (8 + c) * kQRun + r * 16) = v;
    }
    fence_proxy_async();   // generic-proxy writes above -> visible to the async proxy the MMA reads through
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = S.tmem_base;

    if (warp == 0) {
        // ------------------------------------------------ producer ------------------------------------------------
        if (lane == 0) {
            // the CTA's 256 queries = one block of the table: query block qb = rows qb*128 .. +127 of it
            mbar_expect_tx(&S.qfull, kQB * 8 * kQRun);
            const unsigned char* qsrc = Bblk + (size_t)(q0 / 256) * kU8BlockBytes;
#pragma unroll
            for (int qb = 0; qb < kQB; ++qb)
#pragma unroll
                for (int c = 0; c < 8; ++c) tma_bulk_load(S.q[qb] + c * kQRun, qsrc + c * 4096 + qb * kQRun, kQRun, &S.qfull);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages;
                if (t >= kStages) mbar_wait(&S.empty[s], (unsigned)(((t / kStages) - 1) & 1));
                // tile t = rows a_begin + t*kND .. : part (t % (256/kND)) of block (a_begin / 256 + t / (256/kND))
                mbar_expect_tx(&S.full[s], nch * kRun);
                const unsigned char* src = Ablk + (size_t)(a_begin / 256 + t / (256 / kND)) * kU8BlockBytes + (size_t)(t % (256 / kND)) * kRun;
                for (int c = 0; c < nch; ++c) tma_bulk_load(S.db[s] + c * kRun, src + c * 4096, kRun, &S.full[s]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        if (lane == 0) {
            mbar_wait(&S.qfull, 0);
            switch (ext_steps) {
            case 1: mma_loop<5>(S, tmem, ntiles); break;
            case 2: mma_loop<6>(S, tmem, ntiles); break;
            default: mma_loop<7>(S, tmem, ntiles); break;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue ------------------------------------------------
        const int ew = warp & 3;                       // TMEM lane quarter this warp may read
        const int part = ((warp - 4) >> 2) & 1;        // which half of the tile's columns (kND / 2 = 2 units)
        const int qb = (warp - 4) >> 3;                // which query block (accumulator) of the CTA
        const int qrow = qb * kMQ + ew * 32 + lane;    // query row within the CTA; TMEM lane = ew * 32 + lane
        Top3 T = top3_init();
        for (int t = 0; t < ntiles; ++t) {
            const int b = t % kBuf;
            mbar_wait(&S.tmem_full[b], (unsigned)((t / kBuf) & 1));
            tc_fence_after();
            if (mma_only) {   // tensor-pipe peak measurement: the accumulators are released unread (no LDTM, no max tree)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.tmem_empty[b]);
                continue;
            }
            const unsigned taddr = tmem + ((unsigned)(ew * 32) << 16) + (unsigned)((b * kQB + qb) * kND + part * (kND / 2));
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
// maximum of 32 accumulators: 16 three-input max instructions
#define PB_UMAX(v, out)                                                                                                      \
    {                                                                                                                        \
        int mx = __vimax3_s32(v[0], v[1], v[2]);                                                                             \
        _Pragma("unroll") for (int g = 3; g + 1 < 32; g += 2) mx = __vimax3_s32(mx, v[g], v[g + 1]);                         \
        out = max(mx, v[31]);                                                                                                \
    }
            PB_LDTM(va, 0);
            PB_LDTM(vb, 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.tmem_empty[b]);   // the accumulator part is in registers: release it
            int ua, ub;
            PB_UMAX(va, ua);
            PB_UMAX(vb, ub);
#undef PB_LDTM
#undef PB_UMAX
            const int unit0 = (a_begin + t * kND + part * (kND / 2)) / kUnit;
            top3_push(T, ua, unit0);
            top3_push(T, ub, unit0 + 1);
        }
        // hand the second column half of every query row to the first, merge, write
        if (part != 0) { int* m = S.mrg[qrow]; m[0] = T.m1; m[1] = T.u1; m[2] = T.m2; m[3] = T.u2; m[4] = T.m3; m[5] = T.u3; m[6] = T.m4; }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (part == 0) {
            const int* m = S.mrg[qrow];
            top3_push(T, m[0], m[1]);
            top3_push(T, m[2], m[3]);
            top3_push(T, m[4], m[5]);
            top3_push(T, m[6], -1);
            const int q = q0 + qrow;
            if (q < NB) partial[(size_t)blockIdx.y * NB + q] = U8Top3{T.m1, T.u1, T.m2, T.u2, T.m3, T.u3, T.m4};
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

namespace {
// exact squared distance between query (8 x uint4 in registers) and table row `row`
__device__ __forceinline__ int exact_d2(const uint4 (&qv)[8], int nq, const unsigned char* __restrict__ Ablk,
                                        const int* __restrict__ normA, int row) {
    const unsigned char* ap = Ablk + (size_t)(row >> 8) * kU8BlockBytes + (size_t)(row & 255) * 16;
    unsigned dot = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 av = *reinterpret_cast<const uint4*>(ap + c * 4096);
        dot = __dp4a(av.x, qv[c].x, dot); dot = __dp4a(av.y, qv[c].y, dot);
        dot = __dp4a(av.z, qv[c].z, dot); dot = __dp4a(av.w, qv[c].w, dot);
    }
    return nq + normA[row] - 2 * (int)dot;
}
__device__ __forceinline__ long long warp_min_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}
constexpr long long kNoKey = LLONG_MAX;
}  // namespace

__global__ void __launch_bounds__(128) match_u8_finish_kernel(const U8Top3* __restrict__ partial, int nsplit,
                                                              const unsigned char* __restrict__ Ablk, const int* __restrict__ normA,
                                                              int NA, const unsigned char* __restrict__ Bblk,
                                                              const int* __restrict__ normB, int NB, int* __restrict__ idx,
                                                              int* __restrict__ d01) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= NB) return;
    Top3 T = top3_init();
    for (int s = 0; s < nsplit; ++s) {
        const U8Top3 p = partial[(size_t)s * NB + q];
        top3_push(T, p.m1, p.u1);
        top3_push(T, p.m2, p.u2);
        top3_push(T, p.m3, p.u3);
        top3_push(T, p.m4, -1);
    }
    uint4 qv[8];
    const unsigned char* qp = Bblk + (size_t)(q >> 8) * kU8BlockBytes + (size_t)(q & 255) * 16;
#pragma unroll
    for (int c = 0; c < 8; ++c) qv[c] = *reinterpret_cast<const uint4*>(qp + c * 4096);
    const int nq = normB[q];
    // keys = (distance << 32) | row: the minimum is the nearest row, lowest index among equals.
    // Units 1 and 2 (and unit 3 when it ties with unit 2) are rescanned exactly; every row outside them has m <= m4.
    // If m4 < m2, such a row has s <= 2 (m2 - hmax) - 2, below the best row of each of the two leading units, so
    // both the nearest and the second nearest row are among the rescanned ones.  m4 == m2 (a three-way tie) is the
    // only ambiguous case: exact scan of the whole table for this query.
    long long best = kNoKey, second = kNoKey;
    const bool ambiguous = T.m4 >= 0 && T.m4 == T.m2;
    long long ka = kNoKey, kb = kNoKey, kc = kNoKey;   // this lane's candidates, sorted ka <= kb (<= kc dropped)
    auto add = [&](long long k) {
        if (k < ka) { kb = ka; ka = k; }
        else if (k < kb) kb = k;
    };
    (void)kc;
    if (!ambiguous) {
        if (T.u1 >= 0) {
            const int row = T.u1 * kUnit + lane;
            if (row < NA) add(((long long)exact_d2(qv, nq, Ablk, normA, row) << 32) | (unsigned)row);
        }
        if (T.m2 >= 0 && T.u2 >= 0) {
            const int row = T.u2 * kUnit + lane;
            if (row < NA) add(((long long)exact_d2(qv, nq, Ablk, normA, row) << 32) | (unsigned)row);
        }
        if (T.m3 >= 0 && T.m3 == T.m2 && T.u3 >= 0) {
            const int row = T.u3 * kUnit + lane;
            if (row < NA) add(((long long)exact_d2(qv, nq, Ablk, normA, row) << 32) | (unsigned)row);
        }
    } else {
        for (int row = lane; row < NA; row += 32) add(((long long)exact_d2(qv, nq, Ablk, normA, row) << 32) | (unsigned)row);
    }
    best = warp_min_ll(ka);
    second = warp_min_ll(ka == best ? kb : ka);
    if (lane == 0) {
        const int d0 = best == kNoKey ? INT_MAX : (int)(best >> 32), i0 = best == kNoKey ? -1 : (int)(best & 0xffffffffll);
        const int d1 = second == kNoKey ? INT_MAX : (int)(second >> 32);
        // ratio rule sqrt(d0) / sqrt(d1) < 0.5  <=>  4 d0 < d1 (exact in integers)
        const bool ok = NA >= 2 && i0 >= 0 && d1 != INT_MAX && 4ll * d0 < (long long)d1;
        idx[q] = ok ? i0 : -1;
        if (d01) { d01[3 * q] = d0; d01[3 * q + 1] = d1; d01[3 * q + 2] = i0; }
    }
}

// VLFeat's uint8 descriptor convention: q = (uint8) min(512 x, 255)
__global__ void quantize_u8_kernel(const float* __restrict__ src, int n, unsigned char* __restrict__ dst) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const float4 v = reinterpret_cast<const float4*>(src + (size_t)row * 128)[lane];
    const float f[4] = {v.x, v.y, v.z, v.w};
    unsigned packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float x = 512.0f * f[k];
        x = (x < 255.0f) ? x : 255.0f;
        packed |= (unsigned)(unsigned char)x << (8 * k);
    }
    // blocked layout: block = row / 256, chunk = lane / 4 (16 bytes of K), 4 bytes at (lane % 4) * 4 inside the chunk
    *reinterpret_cast<unsigned*>(dst + (size_t)(row >> 8) * kU8BlockBytes + (lane >> 2) * 4096 + (row & 255) * 16 + (lane & 3) * 4) = packed;
}
// |row|^2 and the range of h = floor(|row|^2 / 2) over the table (range[0] = min, range[1] = max)
__global__ void norm_u8_kernel(const unsigned char* __restrict__ src, int n, int* __restrict__ norm, int* __restrict__ range) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const unsigned p = *reinterpret_cast<const unsigned*>(src + (size_t)(row >> 8) * kU8BlockBytes + (lane >> 2) * 4096 + (row & 255) * 16 + (lane & 3) * 4);
    int nrm = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int q = (p >> (8 * k)) & 255; nrm += q * q; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) {
        norm[row] = nrm;
        atomicMin(&range[0], nrm >> 1);
        atomicMax(&range[1], nrm >> 1);
    }
}
// extension columns of a database table: e = hmax - floor(|row|^2 / 2) in base 255, 32 * steps - 1 digits (each <= 255,
// weighted 255 by the query constants) and the remainder e mod 255 in the last column (weighted 1)
__global__ void extend_u8_kernel(unsigned char* __restrict__ blk, const int* __restrict__ norm, int n, int hmax, int steps) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int e = hmax - (norm[row] >> 1);
    int S = e / 255;
    const int r = e - S * 255;
    unsigned char* base = blk + (size_t)(row >> 8) * kU8BlockBytes + (size_t)(row & 255) * 16;
    const int ncol = 32 * steps;
    for (int c = 0; c < 2 * steps; ++c) {
        unsigned w[4] = {0, 0, 0, 0};
        for (int k = 0; k < 16; ++k) {
            const int col = c * 16 + k;
            int v;
            if (col == ncol - 1) v = r;
            else { v = S < 255 ? S : 255; S -= v; }
            w[k >> 2] |= (unsigned)v << (8 * (k & 3));
        }
        *reinterpret_cast<uint4*>(base + (size_t)(8 + c) * 4096) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// row-major [n][128] u8 <-> chunks 0..7 of the blocked layout (rows past n must have been zeroed by the caller)
__global__ void relayout_u8_kernel(const unsigned char* __restrict__ src, int n, unsigned char* __restrict__ dst, int to_blocked) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte chunk per thread
    if (i >= (long)n * 8) return;
    const int row = (int)(i >> 3), c = (int)(i & 7);
    const size_t rm = (size_t)row * 128 + c * 16, bl = (size_t)(row >> 8) * kU8BlockBytes + c * 4096 + (row & 255) * 16;
    if (to_blocked) *reinterpret_cast<uint4*>(dst + bl) = *reinterpret_cast<const uint4*>(src + rm);
    else *reinterpret_cast<uint4*>(dst + rm) = *reinterpret_cast<const uint4*>(src + bl);
}
void launch_relayout_u8(const unsigned char* src, int n, unsigned char* dst, bool to_blocked, cudaStream_t st) {
    if (n <= 0) return;
    relayout_u8_kernel<<<div_up((long)n * 8, 256), 256, 0, st>>>(src, n, dst, to_blocked ? 1 : 0);
    PB_KERNEL_CHECK();
}

void launch_quantize_u8(const float* src, int n, unsigned char* dst, cudaStream_t st) {
    if (n <= 0) return;
    KScope ks("match_u8.quantize", st, 640.0 * n);
    quantize_u8_kernel<<<div_up(n, 8), 256, 0, st>>>(src, n, dst);
    PB_KERNEL_CHECK();
}

void u8_table_prepare(U8Table& t, int* scratch2, cudaStream_t st) {
    t.hmax = 0;
    t.ext_steps = 1;
    if (t.n <= 0) return;
    const int init[2] = {INT_MAX, 0};
    PB_CUDA(cudaMemcpyAsync(scratch2, init, sizeof init, cudaMemcpyHostToDevice, st));
    norm_u8_kernel<<<div_up(t.n, 8), 256, 0, st>>>(t.blk, t.n, t.norm, scratch2);
    PB_KERNEL_CHECK();
    int range[2];
    PB_CUDA(cudaMemcpyAsync(range, scratch2, sizeof range, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    t.hmax = range[1];
    const long span = (long)range[1] - range[0];
    int steps = 1;
    while (steps < 3 && span > 255L * 255L * (32 * steps - 1) + 254) ++steps;
    t.ext_steps = steps;
    extend_u8_kernel<<<div_up(t.n, 128), 128, 0, st>>>(t.blk, t.norm, t.n, t.hmax, steps);
    PB_KERNEL_CHECK();
}

int match_u8_num_splits(int NA, int NB) {
    const int qtiles = div_up(NB, kQB * kMQ);
    int want = div_up(148, qtiles);                    // one CTA per SM (each CTA owns all 512 TMEM columns)
    const int maxs = std::max(1, div_up(NA, 1024));    // at least 1024 database rows per split
    want = std::min(want, maxs);
    return std::max(1, want);
}

void launch_match_u8(const U8Table& A, const U8Table& B, U8Top3* partial, int nsplit, int* idx, int* d01, cudaStream_t st,
                     bool mma_only) {
    const int NA = A.n, NB = B.n;
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
    int rps = align_up(div_up(NA, nsplit), 256);   // splits start on block boundaries of the blocked layout
    nsplit = div_up(NA, rps);
    {
        KScope ks(mma_only ? "match_u8.mma_only" : "match_u8.mma", st, 2.0 * 128.0 * (double)NA * (double)NB);
        match_u8_kernel<<<dim3(div_up(NB, kQB * kMQ), nsplit), 640, smem, st>>>(A.blk, NA, B.blk, NB, rps, A.ext_steps, partial,
                                                                             mma_only ? 1 : 0);
        PB_KERNEL_CHECK();
    }
    if (mma_only) return;   // peak measurement: the same UTCIMMA stream with the epilogue reduced to releasing the accumulators
    KScope ks2("match_u8.finish", st, 0);
    match_u8_finish_kernel<<<div_up(NB, 4), 128, 0, st>>>(partial, nsplit, A.blk, A.norm, NA, B.blk, B.norm, NB, idx, d01);
    PB_KERNEL_CHECK();
}

}  // namespace pb
