// sift_kernels.h -- launch interface of the sm_100a SIFT kernels (sift_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "sift_device.cuh"

namespace pb {

constexpr int kMaxBlurW = 32;  // half-width limit of the in-parameter tap table (sigma <= 8)

struct BlurTaps {
    int W;                         // half width = max(ceil(4 sigma), 1)   (vl/sift.c:128)
    float c[2 * kMaxBlurW + 1];    // normalised taps, computed on the host (vl/sift.c:132-140)
};

struct Cand { int x, y, s; };

struct KeyIn {       // what the orientation / descriptor kernels need of a VlSiftKeypoint
    float x, y, sigma;
    short is, oct;   // level index; slot of the keypoint's octave in the OctaveSet
};

struct DescJob {     // one (keypoint, angle) pair; sin/cos evaluated on the host
    int key;         // index into the KeyIn array
    int pad;
    double angle, st0, ct0;
};

constexpr int kMaxOctaveSet = 8;
struct OctaveSet {   // the octaves of one image, so that one launch serves the keypoints of all of them
    OctaveView ov[kMaxOctaveSet];
    double xper[kMaxOctaveSet];   // 2^o
};

void launch_u8_to_f32(const unsigned char* src, int src_pitch, float* dst, int w, int h, int pitch, cudaStream_t st);
void launch_copy_f32(const float* src, int src_pitch, float* dst, int w, int h, int pitch, cudaStream_t st);
// dst row i = src row rows[i] (rows of 128 floats; rows: device array)
void launch_gather_rows128(const float* src, const int* rows, int n, float* dst, cudaStream_t st);

// dst = gaussian(src) (vertical pass into tmp, horizontal pass into dst; src may alias dst).
// If ds != nullptr the horizontal pass also writes dst sub-sampled by 2 into ds (w/2 x h/2, vl/sift.c:179-194).
void launch_blur(const float* src, float* tmp, float* dst, int w, int h, int pitch, const BlurTaps& taps, float* ds,
                 int ds_pitch, cudaStream_t st);
void launch_downsample2(const float* src, int w, int h, int pitch, float* dst, int dst_pitch, cudaStream_t st);

void launch_dog(const OctaveView& ov, float* dog, cudaStream_t st);  // materialise DoG (shim mirror / tests only)

// Extrema of DoG levels 1..nlevels-3 appended (unordered) to cand[0..cap); *count is incremented past cap on overflow.
void launch_detect(const OctaveView& ov, const SiftConsts& sc, Cand* cand, int* count, int cap, cudaStream_t st);
// out[i] = refine(cand[i]) for i < min(*count, cap)
void launch_refine(const OctaveView& ov, const SiftConsts& sc, const Cand* cand, const int* count, int cap,
                   RefinedKey* out, double xper, cudaStream_t st);
void launch_gradient(const OctaveView& ov, const SiftConsts& sc, float* grad, cudaStream_t st);
void launch_orient(const OctaveSet& os, const SiftConsts& sc, const double* expn_tab, const KeyIn* keys, int nkeys,
                   const int* order, int* nangles, double* angles, cudaStream_t st);
// order (both launchers, optional device array): slot -> item, so that the largest windows are issued first
// patch_bytes: algorithmic gradient-map bytes of the jobs, sum over jobs of (2W+1)^2 * 8 (reporting only)
void launch_descr(const OctaveSet& os, const SiftConsts& sc, const double* expn_tab, const KeyIn* keys,
                  const DescJob* jobs, int njobs, const int* order, float* descr, int* written, double patch_bytes,
                  cudaStream_t st);

}  // namespace pb
