// match_device.cuh -- arithmetic of the exact 2-NN matcher (shared by the CUDA kernel and the test emulator).
//
// The reference matches with a 1-tree VLFeat kd-forest configured for the L1 distance and exact search
// (ImageProcess.cpp:280, 327-331).  Its pruning bound never triggers on SIFT descriptors, so the
// result equals a brute-force scan with the same distance arithmetic:
//   _vl_distance_l1_f (vl/mathop.c:307-318): float accumulator, dimensions 0..127 in order, acc += max(d, -d).
// Ratio rule (ImageProcess.cpp:329-331): float ratio = (double)d0 / (double)d1; keep iff ratio < 0.5.
#pragma once
#include "exact_math.cuh"
#include <math.h>

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


// ---------------------------------------------------------------------------------------------------------
// Rigorous uint8 pre-filter of the exact matcher (match_kernels.cu: match_sad_kernel).
//
// The reference's distance is d(a, b) = the float sum above.  Scale by S = 512 (exact in binary floating point):
// S d(a, b) = fl-sum_i |S a_i - S b_i|.  Every table row x is quantised once to q_i = clamp(floor(S x_i + 1/2), 0, 255) with
// its own error e(x) >= sum_i |S x_i - q_i| (computed in double, rounded UP to an integer; clamped elements simply
// contribute their large error, so no assumption on the value range is needed).  With SAD(a, b) = sum_i |qa_i - qb_i|
// (an exact integer), the triangle inequality gives for the REAL sum D = sum_i |S a_i - S b_i|
//        SAD - e(a) - e(b)  <=  D  <=  SAD + e(a) + e(b),
// and the 128-term float accumulation of non-negative addends differs from D by at most a factor (1 +- g),
// g = 129 * 2^-24 / (1 - 129 * 2^-24) < 2^-16 =: kSadGamma.  So for the float distances the reference computes:
//        (SAD - e(a) - e(b)) (1 - 2^-16)  <=  S d(a, b)  <=  (SAD + e(a) + e(b)) (1 + 2^-16).            (*)
// Per query b the SAD pass keeps  lbmin = min_a (SAD - e(a))  and  u0 <= u1 = the two smallest (SAD + e(a)).
//   * d0 (nearest) >= L := max(lbmin - e(b), 0)(1 - 2^-16) / S, d1 (second nearest) <= U := (u1 + e(b))(1 + 2^-16) / S
//     (two rows lie below U).  If 2 L >= U then d0 / d1 >= 0.5 in real arithmetic; the reference's double quotient
//     and its float rounding are monotonic and 0.5 is representable, so `ratio < 0.5` is false (ImageProcess.cpp:
//     329-331; 0/0 = NaN is also "false"): the query is CERTAINLY rejected and needs no exact arithmetic at all.
//   * otherwise the exact pass needs d0, its row, and d1, i.e. every row with d(a, b) <= d1 <= U.  By (*) such a row
//     has (SAD - e(a) - e(b))(1 - 2^-16) <= S U, i.e.  SAD - e(a) <= thr := floor(S U / (1 - 2^-16)) + e(b) + 1.
//     The rows passing that test are the query's candidates; the exact float top-2 over them is the global top-2
//     (a tie for the minimum gives d0 == d1 and is rejected, so the row order among candidates is irrelevant).
// Nothing here is probabilistic: the miss count is 0 by construction, whatever the input values.
// ---------------------------------------------------------------------------------------------------------
constexpr float kSadScale = 512.0f;
constexpr double kSadGamma = 1.0 / 65536.0;
constexpr int kSadErrUnbounded = 1 << 28;   // row whose quantisation error is not finite / too large: always a candidate

// one element: returns q in 0..255 and adds |S x - q| to *err
PB_HD unsigned sad_quantize(float x, double* err) {
    const float s = x * kSadScale;
    float q = floorf(s + 0.5f);     // any rounding is admissible: the error actually made is what is accounted
    if (!(q >= 0.0f)) q = 0.0f;      // negative or NaN
    if (q > 255.0f) q = 255.0f;
    double e = (double)s - (double)q;
    if (e < 0) e = -e;
    if (!(e == e)) e = 1e300;        // NaN element: unbounded
    *err += e;
    return (unsigned)q;
}
PB_HD int sad_row_error(double err) {
    const double e = err * (1.0 + 1e-6) + 1.0;
    return e < (double)(kSadErrUnbounded - 1) ? (int)e + 1 : kSadErrUnbounded;
}

struct SadStat { int lbmin, u0, u1; };   // min (SAD - e(a)); two smallest (SAD + e(a)) as a multiset
PB_HD SadStat sadstat_init() { return SadStat{0x7fffffff, 0x7fffffff, 0x7fffffff}; }
PB_HD void sadstat_push(SadStat& s, int u, int ea) {
    const int lb = u - 2 * ea;
    if (lb < s.lbmin) s.lbmin = lb;
    if (u < s.u1) {
        if (u < s.u0) { s.u1 = s.u0; s.u0 = u; }
        else s.u1 = u;
    }
}
PB_HD void sadstat_merge(SadStat& s, const SadStat& o) {
    if (o.lbmin < s.lbmin) s.lbmin = o.lbmin;
    if (o.u0 < s.u1) { if (o.u0 < s.u0) { s.u1 = s.u0; s.u0 = o.u0; } else s.u1 = o.u0; }
    if (o.u1 < s.u1) { if (o.u1 < s.u0) { s.u1 = s.u0; s.u0 = o.u1; } else s.u1 = o.u1; }
}
// true: the reference certainly rejects this query.  false: *thr = candidate threshold on (SAD - e(a)).
PB_HD bool sad_certain_reject(const SadStat& s, int eb, int* thr) {
    const double ub_d1 = ((double)s.u1 + (double)eb) * (1.0 + kSadGamma);
    double lb_d0 = (double)s.lbmin - (double)eb;
    if (lb_d0 < 0) lb_d0 = 0;
    lb_d0 *= (1.0 - kSadGamma);
    if (2.0 * lb_d0 >= ub_d1) return true;
    const double t = ub_d1 / (1.0 - kSadGamma) + (double)eb + 1.0;
    *thr = t < 1.0e9 ? (int)t : 1000000000;
    return false;
}

}  // namespace pb
