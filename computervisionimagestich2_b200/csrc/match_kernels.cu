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

__global__ void __launch_bounds__(kQ) match_l1_kernel(const MatchJob* __restrict__ jobs) {
    __shared__ __align__(16) float tile[kTA][128];
    const MatchJob J = jobs[blockIdx.z];
    const float* __restrict__ A = J.A;
    const float* __restrict__ B = J.B;
    const int NA = J.NA, NB = J.NB, rows_per_split = J.rows_per_split;
    Top2* __restrict__ partial = J.partial;
    if (blockIdx.x * kQ >= NB || blockIdx.y >= J.nsplit) return;   // grid is sized for the largest job of the batch
    const int b = blockIdx.x * kQ + threadIdx.x;
    const int a_begin = blockIdx.y * rows_per_split;
    const int a_end = min(NA, a_begin + rows_per_split);
    float q[128];
    {
        const float4* src = reinterpret_cast<const float4*>(B + (size_t)min(b, NB - 1) * 128);
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
    if (b < NB) partial[(size_t)blockIdx.y * NB + b] = best;
}

__global__ void match_merge_kernel(const MatchJob* __restrict__ jobs) {
    const MatchJob J = jobs[blockIdx.y];
    const Top2* __restrict__ partial = J.partial;
    const int nsplit = J.nsplit, NA = J.NA, NB = J.NB;
    int* __restrict__ idx = J.idx;
    float* __restrict__ d01 = J.d01;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB) return;
    Top2 t = partial[b];
    for (int s = 1; s < nsplit; ++s) top2_merge(t, partial[(size_t)s * NB + b]);
    bool ok = NA >= 2 && t.i0 >= 0 && ratio_test(t.d0, t.d1);
    idx[b] = ok ? t.i0 : -1;
    if (d01) { d01[2 * b] = t.d0; d01[2 * b + 1] = t.d1; }
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
    MatchJob J;
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
        match_l1_kernel<<<dim3(gx, gy, njobs), kQ, 0, st>>>(d_jobs);
        PB_KERNEL_CHECK();
    }
    KScope ks2("match.merge", st, 0);
    match_merge_kernel<<<dim3(div_up(gx * kQ, 128), njobs), 128, 0, st>>>(d_jobs);
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
