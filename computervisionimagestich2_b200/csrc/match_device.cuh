// match_device.cuh -- arithmetic of the exact 2-NN matcher (shared by the CUDA kernel and the test emulator).
//
// The reference matches with a 1-tree VLFeat kd-forest configured for the L1 distance and exact search
// (ImageProcess.cpp:280, 327-331).  Its pruning bound never triggers on SIFT descriptors, so the
// result equals a brute-force scan with the same distance arithmetic:
//   _vl_distance_l1_f (vl/mathop.c:307-318): float accumulator, dimensions 0..127 in order, acc += max(d, -d).
// Ratio rule (ImageProcess.cpp:329-331): float ratio = (double)d0 / (double)d1; keep iff ratio < 0.5.
#pragma once
#include "exact_math.cuh"

namespace pb {

PB_HD float l1_distance_128(const float* __restrict__ q, const float* __restrict__ a) {
    float acc = 0.0f;
    for (int k = 0; k < 128; ++k) {
        float d = q[k] - a[k];
        acc += fabs_f(d);
    }
    return acc;
}

struct Top2 {
    float d0, d1;  // smallest, second smallest
    int i0;        // index of the smallest
};
PB_HD Top2 top2_init() { return Top2{3.0e38f, 3.0e38f, -1}; }
PB_HD void top2_push(Top2& t, float d, int i) {
    if (d < t.d1) {
        if (d < t.d0) { t.d1 = t.d0; t.d0 = d; t.i0 = i; }
        else t.d1 = d;
    }
}
PB_HD void top2_merge(Top2& t, const Top2& o) {
    if (o.i0 < 0) return;
    top2_push(t, o.d0, o.i0);
    top2_push(t, o.d1, -2);  // index of a second-best is never used
}
PB_HD bool ratio_test(float d0, float d1) {
    float ratio = (float)((double)d0 / (double)d1);
    return (double)ratio < 0.5;
}

}  // namespace pb
