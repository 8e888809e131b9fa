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

// ---------------------------------------------------------------------------------------------------------
// Second level of the pre-filter: a GROUPED lower bound of the SAD, 8 instructions per (row, row) instead of 32
// (match_kernels.cu: match_group_sym_kernel).
//
// A SIFT descriptor is 4 x 4 spatial cells x 8 orientation bins, dimension = bin + 8 (cx + 4 cy) (vl/sift.c:1268-1438).
// Group g = (bin, cx / 2, cy / 2) collects the SAME orientation bin of a 2 x 2 block of cells: 32 groups of 4
// dimensions.  With A_g = sum of the row's quantised bytes over group g (<= 1020) and its top 8 bits f_g = A_g >> 2:
//        SAD(a, b) = sum_i |qa_i - qb_i|  >=  sum_g |A_g(a) - A_g(b)|            (triangle inequality inside each group)
//                                         >=  sum_g (4 |f_g(a) - f_g(b)| - 3)     (A_g = 4 f_g + r, 0 <= r <= 3)
//                                         =   4 S(a, b) - 96,   S = SAD of the two 32-byte group vectors.          (**)
// On SIFT tables 4 S - 96 is ~0.83 of the SAD (orientation-selective, spatially smooth), which is far tighter than the
// certain-reject test of a query needs: once an upper bound u1 of the second-nearest distance is known (from ANY two
// rows), the query is certainly rejected as soon as every row has SAD - e(a) >= tau(u1, e(b)) below, i.e. about HALF of
// u1.  A row whose grouped bound already proves SAD - e(a) >= tau is SKIPPED (no exact SAD); all other rows get the exact
// SAD and enter the statistics as before.  For the decision,
//        min_a (SAD - e(a))  >=  min(lbmin over the exactly evaluated rows, tau),      u1 = second smallest over ANY rows,
// so sad_certain_reject stays rigorous; queries it does not reject go to the candidate pass (exact SAD of every row)
// exactly as before.  tau is monotone in u1, so a tau computed from the FINAL u1 is <= every tau used while skipping.
// ---------------------------------------------------------------------------------------------------------
constexpr int kGrpBytes = 32;          // groups per row
constexpr int kGrpShift = 2;           // f_g = A_g >> 2
constexpr int kGrpSlack = 3 * 32;      // (**)
constexpr int kGrpErrCap = 255;        // rows with a larger quantisation error bound do not take the grouped pass
constexpr int kGrpBias = 128;          // 2 * (kGrpErrCap + 1) / 4: keeps the packed 16-bit comparison non-negative
PB_HD int sad_group_of_dim(int d) {
    const int bin = d & 7, cx = (d >> 3) & 3, cy = d >> 5;
    return bin + 8 * ((cx >> 1) + 2 * (cy >> 1));
}
// q: the row's 128 quantised bytes -> g: 32 group bytes
PB_HD void sad_group_bytes(const unsigned char* q, unsigned char* g) {
    int sum[kGrpBytes];
    for (int i = 0; i < kGrpBytes; ++i) sum[i] = 0;
    for (int d = 0; d < 128; ++d) sum[sad_group_of_dim(d)] += q[d];
    for (int i = 0; i < kGrpBytes; ++i) g[i] = (unsigned char)(sum[i] >> kGrpShift);
}
// smallest lbmin with which sad_certain_reject(SadStat{lbmin, *, u1}, eb) returns true (plus one unit of margin; the
// decision itself is always taken by sad_certain_reject, so this value only steers how many rows are skipped)
PB_HD int sad_tau(int u1, int eb) {
    if (u1 >= (1 << 24)) return 1 << 24;
    const double half = ((double)u1 + (double)eb) * (1.0 + kSadGamma) / (2.0 * (1.0 - kSadGamma));
    return (int)half + eb + 2;
}
// Skip test of the pair (x, y) in quarter units: with ex4 = ceil(e(x) / 4), ey4 likewise,
//        S - ex4 - ey4 >= c4   ==>   4 S - 96 - e(x) >= tau        for  c4 = ceil((tau + 96 - e(y)) / 4)   [query y, row x]
// (4 ex4 >= e(x), 4 ey4 >= e(y)).  Stored biased for the packed unsigned 16-bit compare:
//        (S + w16(x) + w16(y))  >=  max(c16(y), c16(x)),   w16 = 64 - e4,   c16 = min(c4 + 128, 65535).
// CERTAIN ACCEPT.  Row r0 was evaluated exactly (u0 = SAD + e(r0)); every other row has SAD - e(a) >= `others` (the
// smaller of tau -- the skip guarantee -- and the smallest exact lower bound among the other evaluated rows: l0 <= l1 are
// the two smallest SAD - e(a) over the evaluated rows as a multiset, so without r0's own entry the minimum is l1 if r0
// holds l0, else l0).  By (*):  S d(r0) <= D0 := (u0 + e(b))(1 + 2^-16)  and  S d(r) >= L1 := (others - e(b))(1 - 2^-16)
// for r != r0.  If 2 D0 (1 + 2^-10) < L1 then r0 is the unique nearest row and d0 / d1 < 0.5 / (1 + 2^-10); the
// reference's double quotient and its float rounding (relative 2^-24) stay below 0.5: the query is CERTAINLY accepted with
// row r0 -- no exact float arithmetic is needed.
PB_HD bool sad_certain_accept(int u0, int e_r0, int l0, int l1, int tau, int eb) {
    if (u0 >= (1 << 24)) return false;                 // no row was evaluated
    int others = (u0 - 2 * e_r0 == l0) ? l1 : l0;
    if (tau < others) others = tau;
    double L1 = (double)others - (double)eb;
    if (!(L1 > 0.0)) return false;
    L1 *= (1.0 - kSadGamma);
    const double D0 = ((double)u0 + (double)eb) * (1.0 + kSadGamma);
    return 2.0 * D0 * (1.0 + 1.0 / 1024.0) < L1;
}
PB_HD unsigned sad_w16(int e) { return (unsigned)(64 - ((e + 3) >> 2)); }               // e <= kGrpErrCap
PB_HD unsigned sad_c16(int tau, int e) {
    const long c4 = ((long)tau + kGrpSlack - e + 3) >> 2;                               // tau > e: the numerator is positive
    const long c = c4 + kGrpBias;
    return (unsigned)(c < 0 ? 0 : (c > 65535 ? 65535 : c));
}

}  // namespace pb
