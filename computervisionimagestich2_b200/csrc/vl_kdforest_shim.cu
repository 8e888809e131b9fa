// vl_kdforest_shim.cu -- the vl_kdforest_* C surface (include/vl_b200/kdtree.h) as an exact brute-force k-NN on the GPU.
//
// What it replaces: the 1-tree, exact-search VLFeat kd-forest ImageProcess::getImgPair builds per image pair
// (ImageProcess.cpp:280-331; vl/kdtree.c:773-847).  Distances use VLFeat's arithmetic: a float accumulator over the
// dimensions in order (vl/mathop.c:296-318), so `distance` is bit-identical to what the tree reports.
//
// Kernel: grid = (query, database split); a CTA keeps the query in shared memory, every thread scans rows
// tid, tid + 128, ... of its split with a register-resident sorted list of the K best (distance, index) pairs
// (ties: smaller index first), the lists are merged pairwise through shared memory, and a second tiny kernel merges
// the splits.  The pipeline proper does not go through this file (it matches whole tables in one launch,
// match_kernels.cu); this is the compatibility surface for callers that keep VLFeat's per-query API.
#include "../../include/vl_b200/kdtree.h"
#include "common.h"
#include "exact_math.cuh"
#include <cstdio>
#include <cmath>
#include <vector>
#include <limits>

using namespace pb;

namespace {

constexpr int kMaxK = VL_B200_KDFOREST_MAX_NEIGHBORS;
constexpr int kThreads = 128;

struct Cand {
    float d;
    int i;
};
__device__ __forceinline__ bool before(float d, int i, const Cand& c) { return d < c.d || (d == c.d && (unsigned)i < (unsigned)c.i); }

// sorted insertion into a list of K (K <= kMaxK) held in registers: fully unrolled, no dynamic indexing
template <int K>
__device__ __forceinline__ void insert(Cand (&best)[K], float d, int i) {
    if (!before(d, i, best[K - 1])) return;
    best[K - 1] = Cand{d, i};
#pragma unroll
    for (int k = K - 1; k > 0; --k) {
        if (before(best[k].d, best[k].i, best[k - 1])) {
            const Cand t = best[k - 1];
            best[k - 1] = best[k];
            best[k] = t;
        }
    }
}

template <bool kL1>
__device__ __forceinline__ float distance(const float* __restrict__ q, const float* __restrict__ row, int dim) {
    float acc = 0.0f;
    for (int k = 0; k < dim; ++k) {
        const float d = q[k] - row[k];
        acc += kL1 ? fabs_f(d) : d * d;   // -fmad=false: the product is rounded before the add, as in vl/mathop.c:301
    }
    return acc;
}

// partial: [nq][nsplit][K]
template <int K, bool kL1>
__global__ void __launch_bounds__(kThreads) knn_scan_kernel(const float* __restrict__ data, int n, int dim,
                                                           const float* __restrict__ queries, int rows_per_split,
                                                           Cand* __restrict__ partial) {
    extern __shared__ __align__(8) unsigned char smem[];
    Cand* lists = reinterpret_cast<Cand*>(smem);                 // [kThreads][K]
    float* q = reinterpret_cast<float*>(lists + kThreads * K);   // [dim]
    const int qi = blockIdx.x, split = blockIdx.y;
    for (int k = threadIdx.x; k < dim; k += kThreads) q[k] = queries[(size_t)qi * dim + k];
    __syncthreads();
    Cand best[K];
#pragma unroll
    for (int k = 0; k < K; ++k) best[k] = Cand{INFINITY, -1};
    const int r0 = split * rows_per_split, r1 = min(n, r0 + rows_per_split);
    for (int r = r0 + threadIdx.x; r < r1; r += kThreads) insert<K>(best, distance<kL1>(q, data + (size_t)r * dim, dim), r);
#pragma unroll
    for (int k = 0; k < K; ++k) lists[threadIdx.x * K + k] = best[k];
    __syncthreads();
    for (int s = kThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const Cand c = lists[(threadIdx.x + s) * K + k];
                if (c.i >= 0) insert<K>(best, c.d, c.i);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) lists[threadIdx.x * K + k] = best[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) partial[((size_t)qi * gridDim.y + split) * K + k] = best[k];
    }
}

template <int K>
__global__ void knn_merge_kernel(const Cand* __restrict__ partial, int nq, int nsplit, float* __restrict__ dist,
                                 int* __restrict__ idx) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    Cand best[K];
#pragma unroll
    for (int k = 0; k < K; ++k) best[k] = Cand{INFINITY, -1};
    for (int s = 0; s < nsplit; ++s)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const Cand c = partial[((size_t)qi * nsplit + s) * K + k];
            if (c.i >= 0) insert<K>(best, c.d, c.i);
        }
#pragma unroll
    for (int k = 0; k < K; ++k) { dist[(size_t)qi * K + k] = best[k].d; idx[(size_t)qi * K + k] = best[k].i; }
}

int g_device = -1;

}  // namespace

struct _VlKDForest {
    int device = 0;
    cudaStream_t st = nullptr;
    vl_type data_type = VL_TYPE_FLOAT;
    int dim = 0;
    vl_size num_trees = 1;
    VlVectorComparisonType metric = VlDistanceL1;
    vl_size max_cmp = 0;
    VlKDTreeThresholdingMethod thresholding = VL_KDTREE_MEDIAN;
    int n = 0;
    DevBuf<float> data, queries, dist;
    DevBuf<int> idx;
    DevBuf<Cand> partial;
    std::vector<float> h_dist;
    std::vector<int> h_idx;
    int nsearchers = 0;
};
struct _VlKDForestSearcher {
    VlKDForest* forest;
};

namespace {

template <int K>
void launch_k(VlKDForest* f, int nq, int nsplit, int rps) {
    const size_t smem = (size_t)f->dim * sizeof(float) + (size_t)kThreads * K * sizeof(Cand);
    dim3 grid(nq, nsplit);
    if (f->metric == VlDistanceL1)
        knn_scan_kernel<K, true><<<grid, kThreads, smem, f->st>>>(f->data.p, f->n, f->dim, f->queries.p, rps, f->partial.p);
    else
        knn_scan_kernel<K, false><<<grid, kThreads, smem, f->st>>>(f->data.p, f->n, f->dim, f->queries.p, rps, f->partial.p);
    PB_KERNEL_CHECK();
    knn_merge_kernel<K><<<div_up(nq, 128), 128, 0, f->st>>>(f->partial.p, nq, nsplit, f->dist.p, f->idx.p);
    PB_KERNEL_CHECK();
}

// nq queries (host pointer) -> f->h_dist / f->h_idx as [nq][Kpad], Kpad = the instantiated list length >= K
int run_queries(VlKDForest* f, const float* queries, int nq, int K, int* Kpad_out) {
    PB_CUDA(cudaSetDevice(f->device));
    const int Kpad = K <= 1 ? 1 : K <= 2 ? 2 : K <= 4 ? 4 : 8;
    *Kpad_out = Kpad;
    // enough CTAs to fill the machine for a single query, no more splits than 128-row slices
    int nsplit = div_up(f->n, kThreads);
    const int want = nq >= 296 ? 1 : div_up(296, nq);
    if (nsplit > want) nsplit = want;
    if (nsplit < 1) nsplit = 1;
    const int rps = div_up(div_up(f->n, nsplit), kThreads) * kThreads;
    nsplit = div_up(f->n, rps);
    f->queries.ensure((size_t)nq * f->dim);
    f->partial.ensure((size_t)nq * nsplit * Kpad);
    f->dist.ensure((size_t)nq * Kpad);
    f->idx.ensure((size_t)nq * Kpad);
    PB_CUDA(cudaMemcpyAsync(f->queries.p, queries, (size_t)nq * f->dim * sizeof(float), cudaMemcpyHostToDevice, f->st));
    switch (Kpad) {
        case 1: launch_k<1>(f, nq, nsplit, rps); break;
        case 2: launch_k<2>(f, nq, nsplit, rps); break;
        case 4: launch_k<4>(f, nq, nsplit, rps); break;
        default: launch_k<8>(f, nq, nsplit, rps); break;
    }
    f->h_dist.resize((size_t)nq * Kpad);
    f->h_idx.resize((size_t)nq * Kpad);
    PB_CUDA(cudaMemcpyAsync(f->h_dist.data(), f->dist.p, f->h_dist.size() * sizeof(float), cudaMemcpyDeviceToHost, f->st));
    PB_CUDA(cudaMemcpyAsync(f->h_idx.data(), f->idx.p, f->h_idx.size() * sizeof(int), cudaMemcpyDeviceToHost, f->st));
    PB_CUDA(cudaStreamSynchronize(f->st));
    return 0;
}

bool usable(const VlKDForest* f, vl_size numNeighbors, const char* what) {
    if (!f || f->n <= 0) { fprintf(stderr, "vl_b200 kdforest: %s before vl_kdforest_build\n", what); return false; }
    if (numNeighbors < 1 || numNeighbors > (vl_size)kMaxK) {
        fprintf(stderr, "vl_b200 kdforest: %s with numNeighbors = %llu (supported: 1..%d)\n", what, numNeighbors, kMaxK);
        return false;
    }
    return true;
}

}  // namespace

extern "C" {

void vl_b200_kdforest_set_device(int device) { g_device = device; }

VlKDForest* vl_kdforest_new(vl_type dataType, vl_size dimension, vl_size numTrees, VlVectorComparisonType normType) {
    if (dataType != VL_TYPE_FLOAT) { fprintf(stderr, "vl_b200 kdforest: only VL_TYPE_FLOAT data is supported\n"); return nullptr; }
    if (normType != VlDistanceL1 && normType != VlDistanceL2) {
        fprintf(stderr, "vl_b200 kdforest: only VlDistanceL1 and VlDistanceL2 are supported\n");
        return nullptr;
    }
    if (dimension < 1 || dimension > 8192) { fprintf(stderr, "vl_b200 kdforest: dimension out of range\n"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        fprintf(stderr, "vl_b200 kdforest: no CUDA device (there is no CPU fallback)\n");
        return nullptr;
    }
    try {
        VlKDForest* f = new VlKDForest();
        f->device = g_device >= 0 ? g_device : 0;
        f->dim = (int)dimension;
        f->num_trees = numTrees;
        f->metric = normType;
        PB_CUDA(cudaSetDevice(f->device));
        PB_CUDA(cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking));
        return f;
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_b200 kdforest: %s\n", e.what());
        return nullptr;
    }
}

void vl_kdforest_delete(VlKDForest* self) {
    if (!self) return;
    cudaSetDevice(self->device);
    if (self->st) cudaStreamDestroy(self->st);
    delete self;
}

VlKDForestSearcher* vl_kdforest_new_searcher(VlKDForest* kdforest) {
    if (!kdforest) return nullptr;
    VlKDForestSearcher* s = new VlKDForestSearcher();
    s->forest = kdforest;
    kdforest->nsearchers++;
    return s;
}
void vl_kdforestsearcher_delete(VlKDForestSearcher* searcher) {
    if (!searcher) return;
    if (searcher->forest) searcher->forest->nsearchers--;
    delete searcher;
}

void vl_kdforest_build(VlKDForest* self, vl_size numData, void const* data) {
    if (!self || !data || numData < 1) return;
    try {
        PB_CUDA(cudaSetDevice(self->device));
        self->n = (int)numData;
        self->data.ensure((size_t)numData * self->dim);
        PB_CUDA(cudaMemcpyAsync(self->data.p, data, (size_t)numData * self->dim * sizeof(float), cudaMemcpyHostToDevice, self->st));
        PB_CUDA(cudaStreamSynchronize(self->st));   // the caller may free or change `data` only after delete in VLFeat; here right away
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_b200 kdforest: build failed: %s\n", e.what());
        self->n = 0;
    }
}

vl_size vl_kdforest_query(VlKDForest* self, VlKDForestNeighbor* neighbors, vl_size numNeighbors, void const* query) {
    if (!usable(self, numNeighbors, "vl_kdforest_query")) return 0;
    try {
        int Kpad = 0;
        run_queries(self, (const float*)query, 1, (int)numNeighbors, &Kpad);
        for (vl_size k = 0; k < numNeighbors; ++k) {
            const int i = self->h_idx[k];
            neighbors[k].index = i >= 0 ? (vl_uindex)i : (vl_uindex)-1;
            neighbors[k].distance = i >= 0 ? (double)self->h_dist[k] : std::numeric_limits<double>::quiet_NaN();
        }
        return (vl_size)self->n;
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_b200 kdforest: query failed: %s\n", e.what());
        return 0;
    }
}

vl_size vl_kdforestsearcher_query(VlKDForestSearcher* self, VlKDForestNeighbor* neighbors, vl_size numNeighbors,
                                  void const* query) {
    if (!self) return 0;
    return vl_kdforest_query(self->forest, neighbors, numNeighbors, query);
}

vl_size vl_kdforest_query_with_array(VlKDForest* self, vl_uint32* index, vl_size numNeighbors, vl_size numQueries,
                                     void* distance, void const* queries) {
    if (!usable(self, numNeighbors, "vl_kdforest_query_with_array") || numQueries < 1) return 0;
    try {
        int Kpad = 0;
        // chunked so that the partial lists stay small for very large query sets
        const vl_size chunk = 1 << 16;
        for (vl_size q0 = 0; q0 < numQueries; q0 += chunk) {
            const int nq = (int)((numQueries - q0) < chunk ? (numQueries - q0) : chunk);
            run_queries(self, (const float*)queries + q0 * self->dim, nq, (int)numNeighbors, &Kpad);
            for (int q = 0; q < nq; ++q)
                for (vl_size k = 0; k < numNeighbors; ++k) {
                    const int i = self->h_idx[(size_t)q * Kpad + k];
                    index[(q0 + q) * numNeighbors + k] = (vl_uint32)i;   // (vl_uint32)-1 when fewer than K data points
                    if (distance)
                        ((float*)distance)[(q0 + q) * numNeighbors + k] =
                            i >= 0 ? self->h_dist[(size_t)q * Kpad + k] : std::numeric_limits<float>::quiet_NaN();
                }
        }
        return (vl_size)self->n * numQueries;
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_b200 kdforest: query failed: %s\n", e.what());
        return 0;
    }
}

vl_size vl_kdforest_get_num_trees(VlKDForest const* self) { return self->num_trees; }
vl_size vl_kdforest_get_data_dimension(VlKDForest const* self) { return (vl_size)self->dim; }
vl_type vl_kdforest_get_data_type(VlKDForest const* self) { return self->data_type; }
void vl_kdforest_set_max_num_comparisons(VlKDForest* self, vl_size n) { self->max_cmp = n; }
vl_size vl_kdforest_get_max_num_comparisons(VlKDForest* self) { return self->max_cmp; }
void vl_kdforest_set_thresholding_method(VlKDForest* self, VlKDTreeThresholdingMethod method) { self->thresholding = method; }
VlKDTreeThresholdingMethod vl_kdforest_get_thresholding_method(VlKDForest const* self) { return self->thresholding; }
VlKDForest* vl_kdforest_searcher_get_forest(VlKDForestSearcher const* self) { return self->forest; }

}  // extern "C"
