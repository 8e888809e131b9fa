// match_kernels.h -- launch interface of the matcher and RANSAC kernels (match_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "types.h"
#include "match_device.cuh"

namespace pb {

// number of database splits (grid.y) used for an NA x NB problem; partial must hold nsplit * NB Top2 entries
int match_num_splits(int NA, int NB);

// One directed matching problem (ImageProcess::getImgPair): A = database, B = queries, both [n][128] f32 in HBM.
// idx[b] = index into A or -1; d01 (optional) receives (d0, d1) per query; partial holds nsplit * NB Top2 entries.
// Both tables must hold at least 2 / 1 rows (the host short-cuts degenerate problems).
struct MatchJob {
    const float* A;
    const float* B;
    int NA, NB, rows_per_split, nsplit;
    Top2* partial;
    int* idx;
    float* d01;
    // ---- rigorous uint8 pre-filter (match_device.cuh); unused (null) in the full-scan mode ----
    const unsigned* A8;     // [NA][32] words = 128 quantised bytes per row
    const int* Ae;          // [NA] quantisation error bound of each row, SAD units
    const unsigned* B8;
    const int* Be;
    int sad_rows_per_split, sad_nsplit;
    int cand_rows_per_split, cand_nsplit;   // the candidate pass runs over few queries: it splits the database finer
    SadStat* spartial;      // [sad_nsplit][NB]
    int* surv;              // [NB] query rows the SAD pass could not reject (compacted, any order)
    int* thr;               // [NB] per survivor slot: candidate threshold on SAD - e(a)
    int* cand_cnt;          // [NB] per survivor slot
    int* cand;              // [NB][kMatchCand] candidate rows of A
    int* overflow;          // [NB] query rows with more than kMatchCand candidates: full exact scan
    int* counters;          // [0] survivors, [1] overflow queries, [2] exact SADs of the grouped pass, [3] certain accepts
    // ---- grouped pass (second level of the pre-filter, match_device.cuh); null unless the job takes it ----
    const unsigned* Ag;     // [NA padded to whole tiles][8] words = 32 group bytes per row
    const unsigned short* Aw;   // [NA padded] w16 of each row
    const unsigned* Bg;
    const unsigned short* Bw;
    int* seed_u1;           // [NB] upper bound of the second-nearest SAD + e from the sampled rows
    unsigned short* c16;    // [NB padded to whole tiles] packed skip threshold of each query
    int* stat6;             // [NB][6] over the exactly evaluated rows (atomics; initialised to 0x7f bytes): two smallest keys
                            // (SAD + e(a)) << 32 | a as 64-bit values, then the two smallest SAD - e(a)
    int grp_rows_per_split, grp_nsplit;
};
constexpr int kMatchCand = 32;
MatchJob make_match_job(const float* dA, int NA, const float* dB, int NB, Top2* partial, int nsplit, int* idx, float* d01);
// Exact L1 2-NN + ratio rule for a batch of problems by a full float scan, one launch (grid.z = job); d_jobs = device
// copy of h_jobs.
void launch_match_batch(const MatchJob* d_jobs, const MatchJob* h_jobs, int njobs, cudaStream_t st);
// The same result through the pre-filter: SAD pass over the quantised tables, decision, candidate pass for the
// surviving queries, exact float distances of the candidates, full scan of the queries whose candidate list overflowed.
// The jobs' pre-filter fields must be set (match_prefilter_attach); counters must be zero.
// pairs: (F, R) job indices with R = the reverse problem of F (F.A == R.B, F.B == R.A): one SAD pass serves both
// (match_sad_sym_kernel); singles: the remaining job indices.  Every job appears in exactly one of the two lists.
// grouped: the pairs take the grouped pass (match_group_attach on both jobs of every pair) instead of the full SAD pass;
// gq_*: the queue of row pairs the grouped bound cannot skip.  If it overflows, counters[3] of the first pair's forward job
// has bit 30 set and the caller must redo the batch without the grouped pass.
void launch_match_batch_prefilter(const MatchJob* d_jobs, const MatchJob* h_jobs, int njobs, const int2* d_pairs,
                                  const int2* h_pairs, int npairs, const int* d_singles, const int* h_singles, int nsingles,
                                  cudaStream_t st, bool grouped = false, unsigned long long* gq_items = nullptr,
                                  unsigned long long* gq_count = nullptr /* zeroed */, size_t gq_cap = 0);
// grouped pass: ints of scratch one job needs, the attachment of that scratch (which the caller fills with 0x7f bytes
// before the launch), the number of database splits for a batch with `yblocks_total` blocks of held rows
size_t match_group_ints(int NB);
void match_group_attach(MatchJob& J, const unsigned* Ag, const unsigned short* Aw, const unsigned* Bg, const unsigned short* Bw,
                        int nsplit, int* scratch);
int match_group_yblocks(int NY);
int match_group_num_splits(int NA, int yblocks_total);
int match_group_err_cap();
inline size_t match_group_pad_rows(int n) { return (size_t)(n + 127) / 128 * 128 + 128; }   // rows readable by whole-tile copies
// makes R the reverse problem of F for the symmetric pass; returns the number of SadStat rows (of R.NB entries) R needs
int match_prefilter_pair(MatchJob& F, MatchJob& R);
int match_sym_yblocks(int NY);
int match_sym_err_cap();   // the symmetric pass needs every row error bound of both tables <= this
int match_sad_num_splits(int NA, int NB, int njobs);
// ints of scratch the pre-filter of one job needs (after the SadStat partials), and the attachment of that scratch
size_t match_prefilter_ints(int NB);
void match_prefilter_attach(MatchJob& J, const unsigned* A8, const int* Ae, const unsigned* B8, const int* Be,
                            int sad_nsplit, SadStat* spartial, int* scratch, int* counters /* 4 ints, zeroed */);
// float table [n][128] -> quantised words [n][32] + per-row error bound
// emax (optional, device): receives the largest error bound of the table
// g8 / w16 (optional): group vectors [n][8] words and w16 [n] for the grouped pass
void launch_sad_quantize(const float* descr, int n, unsigned* q8, int* qe, int* emax, cudaStream_t st, unsigned* g8 = nullptr,
                         unsigned short* w16 = nullptr);

// Scores `iters` hypotheses for each of nproblems pair lists.  pairs: concatenated lists, pair_off [nproblems+1];
// samples [nproblems][iters][4] indices into each list; counts [nproblems][iters]; masks [nproblems][iters][words_stride]
// inlier bit masks; hyp (optional) [nproblems][iters][8] the fitted coefficients.
void launch_ransac_score(const KeyPair* pairs, const int* pair_off, int nproblems, const int* samples, int iters,
                         int* counts, unsigned* masks, int words_stride, double* hyp, cudaStream_t st);

}  // namespace pb
