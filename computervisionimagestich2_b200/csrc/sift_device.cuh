// sift_device.cuh -- per-item bodies of the SIFT kernels (detector, refinement, gradient, orientation, descriptor).
//
// Each body is the work of ONE logical item (a pixel, a candidate extremum, a keypoint, a descriptor job) written
// as a __host__ __device__ function, so that the very same arithmetic can be exercised on a CPU-only box by the
// test-only emulator in tests/emul.  The CUDA kernels in sift_kernels.cu map items onto threads and call these.
//
// Reference semantics restated here (file:line relative to the reference tree):
//   blur taps + column convolution ... vl/sift.c:116-159, vl/imopv.c:118-201
//   DoG + 26-neighbour extrema ....... vl/sift.c:521-603
//   3-D quadratic refinement ......... vl/sift.c:610-772
//   gradient (mod, angle) map ........ vl/sift.c:792-876
//   orientation histogram ............ vl/sift.c:904-1037
//   4x4x8 descriptor ................. vl/sift.c:1268-1438
#pragma once
#include "exact_math.cuh"

namespace pb {

// One octave of the Gaussian scale space resident in HBM.  Rows are `pitch` floats apart (pitch % 32 == 0 so
// that every row starts on a 128-byte line); level l (= s - s_min) starts at gss + l * pitch * h.
// grad holds interleaved (modulus, angle) pairs for the detection levels s = s_min+1 .. s_max-2:
// grad[((l * h + y) * pitch + x) * 2 + {0,1}].
struct OctaveView {
    int w, h, pitch;
    int nlevels;  // s_max - s_min + 1
    const float* gss;
    const float* grad;
};

struct SiftConsts {
    int s_min, s_max, S;
    double peak_thresh, edge_thresh, norm_thresh, magnif, window_size;
};

// ---------------------------------------------------------------------------------------------------------
// separable blur, one output sample (vl/imopv.c:137-198 with VL_PAD_BY_CONTINUITY): taps are applied in
// increasing source index, the j-th source sample (p = centre - W + j, clamped) meets filt[2W - j].
// ---------------------------------------------------------------------------------------------------------
PB_HD float blur_sample(const float* __restrict__ line, int stride, int n, int centre, const float* __restrict__ filt,
                        int W) {
    float acc = 0.0f;
    for (int j = 0; j <= 2 * W; ++j) {
        int p = centre - W + j;
        p = p < 0 ? 0 : (p > n - 1 ? n - 1 : p);
        acc = acc + line[(long)p * stride] * filt[2 * W - j];
    }
    return acc;
}

// DoG value on the fly: dog[l](x,y) = gss[l+1](x,y) - gss[l](x,y)   (vl/sift.c:521-530)
PB_HD float dog_at(const OctaveView& ov, int x, int y, int l) {
    const long ls = (long)ov.pitch * ov.h;
    const long o = (long)y * ov.pitch + x;
    return ov.gss[(l + 1) * ls + o] - ov.gss[l * ls + o];
}

// 26-neighbour strict extremum test at interior pixel (x, y) of DoG level l (1 <= l <= nlevels-3).
// vl/sift.c:536-577 (CHECK_NEIGHBORS); tp = peak threshold.
PB_HD bool is_extremum(const OctaveView& ov, int x, int y, int l, double tp) {
    const float v = dog_at(ov, x, y, l);
    bool gt = ((double)v >= 0.8 * tp), lt = ((double)v <= -0.8 * tp);
    if (!gt && !lt) return false;
    for (int dl = -1; dl <= 1 && (gt || lt); ++dl)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (dl == 0 && dy == 0 && dx == 0) continue;
                const float n = dog_at(ov, x + dx, y + dy, l + dl);
                gt = gt && (v > n);
                lt = lt && (v < n);
            }
    return gt || lt;
}

struct RefinedKey {
    int ix0, iy0, is0;  // the candidate as detected (raster order key)
    int good;
    int ix, iy, is;     // after the <=5 moves
    float x, y, s;      // k->x, k->y, k->s   (sigma is finished on the host: it needs pow())
    double sn;          // s + b[2] in double, input of sigma0 * pow(2, sn / S) * xper
};

// vl/sift.c:610-772.  (x, y, s) is the detected extremum, s in reference units (s_min+1 .. s_max-2).
PB_HD RefinedKey refine_key(const OctaveView& ov, const SiftConsts& sc, int x, int y, int s, double xper) {
    RefinedKey out;
    out.ix0 = x; out.iy0 = y; out.is0 = s;
    const int w = ov.w, h = ov.h;
    const int l = s - sc.s_min;
    double Dx = 0, Dy = 0, Ds = 0, Dxx = 0, Dyy = 0, Dss = 0, Dxy = 0, Dxs = 0, Dys = 0;
    double A[9], b[3];
    b[0] = b[1] = b[2] = 0;
    int dx = 0, dy = 0;
#define PB_AT(ddx, ddy, dds) dog_at(ov, x + (ddx), y + (ddy), l + (dds))
#define PB_A(i, j) (A[(i) + (j)*3])
    for (int iter = 0; iter < 5; ++iter) {
        x += dx;
        y += dy;
        // float differences first (the reference subtracts/adds floats, then widens)
        Dx = 0.5 * (double)(PB_AT(+1, 0, 0) - PB_AT(-1, 0, 0));
        Dy = 0.5 * (double)(PB_AT(0, +1, 0) - PB_AT(0, -1, 0));
        Ds = 0.5 * (double)(PB_AT(0, 0, +1) - PB_AT(0, 0, -1));
        const double c2 = 2.0 * (double)PB_AT(0, 0, 0);
        Dxx = ((double)(PB_AT(+1, 0, 0) + PB_AT(-1, 0, 0)) - c2);
        Dyy = ((double)(PB_AT(0, +1, 0) + PB_AT(0, -1, 0)) - c2);
        Dss = ((double)(PB_AT(0, 0, +1) + PB_AT(0, 0, -1)) - c2);
        Dxy = 0.25 * (double)(PB_AT(+1, +1, 0) + PB_AT(-1, -1, 0) - PB_AT(-1, +1, 0) - PB_AT(+1, -1, 0));
        Dxs = 0.25 * (double)(PB_AT(+1, 0, +1) + PB_AT(-1, 0, -1) - PB_AT(-1, 0, +1) - PB_AT(+1, 0, -1));
        Dys = 0.25 * (double)(PB_AT(0, +1, +1) + PB_AT(0, -1, -1) - PB_AT(0, -1, +1) - PB_AT(0, +1, -1));

        PB_A(0, 0) = Dxx; PB_A(1, 1) = Dyy; PB_A(2, 2) = Dss;
        PB_A(0, 1) = PB_A(1, 0) = Dxy;
        PB_A(0, 2) = PB_A(2, 0) = Dxs;
        PB_A(1, 2) = PB_A(2, 1) = Dys;
        b[0] = -Dx; b[1] = -Dy; b[2] = -Ds;

        // Gauss elimination with maximal-magnitude pivot (vl/sift.c:661-712)
        bool singular = false;
        for (int j = 0; j < 3; ++j) {
            double maxa = 0, maxabsa = 0;
            int maxi = -1;
            for (int i = j; i < 3; ++i) {
                double a = PB_A(i, j);
                double absa = abs_d(a);
                if (absa > maxabsa) { maxa = a; maxabsa = absa; maxi = i; }
            }
            if (maxabsa < (double)1e-10f) {
                b[0] = 0; b[1] = 0; b[2] = 0;
                singular = true;
                break;
            }
            const int i = maxi;
            for (int jj = j; jj < 3; ++jj) {
                double tmp = PB_A(i, jj); PB_A(i, jj) = PB_A(j, jj); PB_A(j, jj) = tmp;
                PB_A(j, jj) /= maxa;
            }
            double tmp = b[j]; b[j] = b[i]; b[i] = tmp;
            b[j] /= maxa;
            for (int ii = j + 1; ii < 3; ++ii) {
                double xx = PB_A(ii, j);
                for (int jj = j; jj < 3; ++jj) PB_A(ii, jj) -= xx * PB_A(j, jj);
                b[ii] -= xx * b[j];
            }
        }
        (void)singular;
        // backward substitution (runs after a singular break as well, on b = 0: harmless, as in the reference)
        for (int i = 2; i > 0; --i) {
            double xx = b[i];
            for (int ii = i - 1; ii >= 0; --ii) b[ii] -= xx * PB_A(ii, i);
        }
        dx = ((b[0] > 0.6 && x < w - 2) ? 1 : 0) + ((b[0] < -0.6 && x > 1) ? -1 : 0);
        dy = ((b[1] > 0.6 && y < h - 2) ? 1 : 0) + ((b[1] < -0.6 && y > 1) ? -1 : 0);
        if (dx == 0 && dy == 0) break;
    }
    {
        const double te = sc.edge_thresh, tp = sc.peak_thresh;
        double val = (double)PB_AT(0, 0, 0) + 0.5 * (Dx * b[0] + Dy * b[1] + Ds * b[2]);
        double score = (Dxx + Dyy) * (Dxx + Dyy) / (Dxx * Dyy - Dxy * Dxy);
        double xn = x + b[0];
        double yn = y + b[1];
        double sn = s + b[2];
        bool good = abs_d(val) > tp && score < (te + 1) * (te + 1) / te && score >= 0 && abs_d(b[0]) < 1.5 &&
                    abs_d(b[1]) < 1.5 && abs_d(b[2]) < 1.5 && xn >= 0 && xn <= w - 1 && yn >= 0 && yn <= h - 1 &&
                    sn >= sc.s_min && sn <= sc.s_max;
        out.good = good ? 1 : 0;
        out.ix = x; out.iy = y; out.is = s;
        out.s = (float)sn;
        out.x = (float)(xn * xper);
        out.y = (float)(yn * xper);
        out.sn = sn;
    }
#undef PB_AT
#undef PB_A
    return out;
}

// vl/sift.c:792-876: gradient modulus and angle of GSS level l at (x, y); one-sided differences on the border.
PB_HD void gradient_at(const float* __restrict__ lev, int w, int h, int pitch, int x, int y, float* mod, float* ang) {
    const float* p = lev + (long)y * pitch + x;
    float gx, gy;
    if (x == 0) gx = p[1] - p[0];
    else if (x == w - 1) gx = p[0] - p[-1];
    else gx = 0.5f * (p[1] - p[-1]);
    if (y == 0) gy = p[pitch] - p[0];
    else if (y == h - 1) gy = p[0] - p[-pitch];
    else gy = 0.5f * (p[pitch] - p[-pitch]);
    *mod = fast_sqrt_f(gx * gx + gy * gy);
    *ang = mod_2pi_f((float)((double)fast_atan2_f(gy, gx) + 2 * kPi));
}

// ---------------------------------------------------------------------------------------------------------
// Descriptor, sample-parallel formulation (what the CUDA kernel runs).
//
// The reference scans the (2W+1)^2 window in raster order and scatters every sample into 2x2 spatial cells x 2
// orientation bins.  Floating-point addition order matters only WITHIN a bin, and the 8 bins one sample touches are
// all different.  The kernel therefore splits the work in two:
//   phase A (parallel over samples)  descriptor_sample(): everything that depends on one sample alone -- the bin
//            coordinates and the partial weight products, in the reference's evaluation order;
//   phase B (parallel over bins)     every bin owner adds, in raster order of the samples, the contributions that
//            land in its bin.  Each bin receives exactly the reference's addends in the reference's order.
// Samples that cannot reach the 4x4 grid (|nx| or |ny| beyond 2.5 cells) are skipped: descriptor_rows() /
// descriptor_row_range() give a conservative superset of the contributing samples and descriptor_sample() applies
// the reference's exact bin test to every sample of that superset.
// ---------------------------------------------------------------------------------------------------------
struct DescFrame {
    double x, y, SBP, st0, ct0, angle0, wden;
    int xi, yi, W, dx0, dx1, dy0, dy1;
    const float* pt;   // gradient plane of the keypoint's level
    int pitch;
    int valid;
};

PB_HD DescFrame descriptor_frame(const OctaveView& ov, const SiftConsts& sc, int o_cur, int ko, int kis, float kx,
                                 float ky, float ksigma, double xper, double angle0, double st0, double ct0) {
    enum { NBP = 4 };
    DescFrame F;
    const int w = ov.w, h = ov.h;
    F.x = (double)kx / xper;
    F.y = (double)ky / xper;
    const double sigma = (double)ksigma / xper;
    F.xi = (int)(F.x + 0.5);
    F.yi = (int)(F.y + 0.5);
    F.SBP = sc.magnif * sigma + kEpsD;
    F.W = (int)floor(1.4142135623730951 * F.SBP * (NBP + 1) / 2.0 + 0.5);
    F.st0 = st0; F.ct0 = ct0; F.angle0 = angle0;
    const float wsigma = (float)sc.window_size;
    F.wden = 2.0 * wsigma * wsigma;
    F.valid = !(ko != o_cur || F.xi < 0 || F.xi >= w || F.yi < 0 || F.yi >= h - 1 || kis < sc.s_min + 1 ||
                kis > sc.s_max - 2);
    F.pt = ov.grad + 2 * ((long)(kis - sc.s_min - 1) * ov.h * ov.pitch);
    F.pitch = ov.pitch;
    F.dy0 = (-F.W > 1 - F.yi) ? -F.W : 1 - F.yi;
    F.dy1 = (F.W < h - F.yi - 2) ? F.W : h - F.yi - 2;
    F.dx0 = (-F.W > 1 - F.xi) ? -F.W : 1 - F.xi;
    F.dx1 = (F.W < w - F.xi - 2) ? F.W : w - F.xi - 2;
    return F;
}

// Rows (dyi) that can hold contributing samples: |dy| <= 2.5 SBP (|cos| + |sin|) + slack, inside the window.
PB_HD void descriptor_rows(const DescFrame& F, int* ry0, int* ry1) {
    const double act = abs_d(F.ct0), ast = abs_d(F.st0);
    const double e = 2.5 * F.SBP * (act + ast) + 1.5;
    const double offy = (double)F.yi - F.y;
    int a = (int)floor(-e - offy), b = (int)floor(e - offy) + 1;
    if (a < F.dy0) a = F.dy0;
    if (b > F.dy1) b = F.dy1;
    *ry0 = a;
    *ry1 = b;
}

// Columns [x0, x1] (dxi) of row dyi that can hold contributing samples (empty when x0 > x1): both rotated-strip
// constraints |ct0 dx + st0 dy| <= 2.5 SBP and |-st0 dx + ct0 dy| <= 2.5 SBP, each widened by 1.5 px.
PB_HD void descriptor_row_range(const DescFrame& F, int dyi, int* px0, int* px1) {
    const double act = abs_d(F.ct0), ast = abs_d(F.st0);
    const double offx = (double)F.xi - F.x, offy = (double)F.yi - F.y;   // sample offset = d?i + off?
    const double half = 2.5 * F.SBP;
    const double dyr = (double)dyi + offy;
    double lo = (double)F.dx0, hi = (double)F.dx1;
    if (act > 1e-3) {
        const double c = -F.st0 * dyr;
        double a = (c - half) / F.ct0, b = (c + half) / F.ct0;
        if (a > b) { double t = a; a = b; b = t; }
        a = a - offx - 1.5; b = b - offx + 1.5;
        if (a > lo) lo = a;
        if (b < hi) hi = b;
    }
    if (ast > 1e-3) {
        const double c = F.ct0 * dyr;
        double a = (c - half) / F.st0, b = (c + half) / F.st0;
        if (a > b) { double t = a; a = b; b = t; }
        a = a - offx - 1.5; b = b - offx + 1.5;
        if (a > lo) lo = a;
        if (b < hi) hi = b;
    }
    if (lo > hi) { *px0 = 0; *px1 = -1; return; }
    int x0 = (int)floor(lo), x1 = (int)floor(hi) + 1;
    if (x0 < F.dx0) x0 = F.dx0;
    if (x1 > F.dx1) x1 = F.dx1;
    *px0 = x0;
    *px1 = x1;
}

// Everything one sample contributes (vl/sift.c:1351-1413).  A sample reaches cells (binx + {0,1}, biny + {0,1})
// that lie inside the 4x4 grid and orientation bins (bint + {0,1}) % 8 with
//   weight = (((win * mod) * |1 - dbx - rbinx|) * |1 - dby - rbiny|) * |1 - dbt - rbint|
// wxy[dbx][dby] holds the first three factors, at[dbt] the last.
struct DescSample {
    int active;            // 0: the sample reaches no cell of the grid
    int binx, biny, bint;  // binx, biny in -3..1, bint in 0..8
    float wxy[2][2];
    float at[2];
};

PB_HD DescSample descriptor_sample(const DescFrame& F, const double* __restrict__ expn_tab, int dxi, int dyi, float mod,
                                   float angle) {
    enum { NBO = 8 };
    DescSample S;
    S.active = 0;
    const float dy = (float)((double)(F.yi + dyi) - F.y);
    const float dx = (float)((double)(F.xi + dxi) - F.x);
    const float nx = (float)((F.ct0 * (double)dx + F.st0 * (double)dy) / F.SBP);
    const int binx = floor_f((float)((double)nx - 0.5));
    if (binx < -3 || binx > 1) return S;
    const float ny = (float)((-F.st0 * (double)dx + F.ct0 * (double)dy) / F.SBP);
    const int biny = floor_f((float)((double)ny - 0.5));
    if (biny < -3 || biny > 1) return S;
    const float theta = mod_2pi_f((float)((double)angle - F.angle0));
    const float nt = (float)((double)((float)NBO * theta) / (2 * kPi));
    const float win = (float)fast_expn(expn_tab, (double)(nx * nx + ny * ny) / F.wden);
    const int bint = floor_f(nt);
    const float rbinx = (float)((double)nx - ((double)binx + 0.5));
    const float rbiny = (float)((double)ny - ((double)biny + 0.5));
    const float rbint = nt - (float)bint;
    const float base = win * mod;
    const float bx0 = base * fabs_f((float)(1 - 0) - rbinx), bx1 = base * fabs_f((float)(1 - 1) - rbinx);
    const float ay0 = fabs_f((float)(1 - 0) - rbiny), ay1 = fabs_f((float)(1 - 1) - rbiny);
    S.wxy[0][0] = bx0 * ay0; S.wxy[0][1] = bx0 * ay1;
    S.wxy[1][0] = bx1 * ay0; S.wxy[1][1] = bx1 * ay1;
    S.at[0] = fabs_f((float)(1 - 0) - rbint);
    S.at[1] = fabs_f((float)(1 - 1) - rbint);
    S.binx = binx; S.biny = biny; S.bint = bint;
    S.active = 1;
    return S;
}

// normalise -> clamp 0.2 -> normalise over the 128 bins in index order (vl/sift.c:1048-1063, 1415-1436);
// hist: bin b at hist[b * hstride]; writes out[0..127]
PB_HD void descriptor_finish(const SiftConsts& sc, float* hist, int hstride, float* __restrict__ out) {
    const int n = 128;
    float norm = 0.0f;
    for (int i = 0; i < n; ++i) norm += hist[i * hstride] * hist[i * hstride];
    norm = fast_sqrt_f(norm) + kEpsF;
    for (int i = 0; i < n; ++i) hist[i * hstride] /= norm;
    if (sc.norm_thresh != 0 && (double)norm < sc.norm_thresh) {
        for (int i = 0; i < n; ++i) out[i] = 0;
        return;
    }
    for (int i = 0; i < n; ++i)
        if ((double)hist[i * hstride] > 0.2) hist[i * hstride] = (float)0.2;
    norm = 0.0f;
    for (int i = 0; i < n; ++i) norm += hist[i * hstride] * hist[i * hstride];
    norm = fast_sqrt_f(norm) + kEpsF;
    for (int i = 0; i < n; ++i) out[i] = hist[i * hstride] / norm;
}

}  // namespace pb
