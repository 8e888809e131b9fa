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
};
MatchJob make_match_job(const float* dA, int NA, const float* dB, int NB, Top2* partial, int nsplit, int* idx, float* d01);
// Exact L1 2-NN + ratio rule for a batch of problems in one launch (grid.z = job); d_jobs = device copy of h_jobs.
void launch_match_batch(const MatchJob* d_jobs, const MatchJob* h_jobs, int njobs, cudaStream_t st);

// Scores `iters` hypotheses for each of nproblems pair lists.  pairs: concatenated lists, pair_off [nproblems+1];
// samples [nproblems][iters][4] indices into each list; counts [nproblems][iters]; masks [nproblems][iters][words_stride]
// inlier bit masks; hyp (optional) [nproblems][iters][8] the fitted coefficients.
void launch_ransac_score(const KeyPair* pairs, const int* pair_off, int nproblems, const int* samples, int iters,
                         int* counts, unsigned* masks, int words_stride, double* hyp, cudaStream_t st);

}  // namespace pb
