// host_numerics.h -- the double-precision / libm pieces of the path that stay on the HOST (plain C++, no CUDA).
//
// The reference evaluates these with glibc (exp, pow, tan, sqrt) or with tiny dense linear algebra from CImg;
// they are O(1)..O(n_pairs) work per image pair and their results parameterise the kernels (filter taps,
// lookup tables, recursive-filter coefficients, resampling tables) or finish a RANSAC fit.  They are compiled with
// -ffp-contract=off semantics (nvcc host pass: -Xcompiler -ffp-contract=off).
//
// Restated from (reference file:line):
//   Gaussian taps ................. vl/sift.c:125-141
//   fast_expn table ............... vl/sift.c:56-63
//   (the 4x4 LU solve is in ransac_device.cuh: it also runs on the device)
//   SVD / pseudo-inverse (n x 4) .. CImg.h:25755-25890 (SVD), 25293-25302 (get_pseudoinvert), 12244-12262 (operator*)
//   van Vliet coefficients ........ CImg.h:35045-35065
//   moving-average / linear resampling tables .. CImg.h:29539-29560, 29618-29700
//   cylinder table ................ Projection.cpp:27-37
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <stdexcept>

namespace pb {
namespace hostnum {

// ---- SIFT ---------------------------------------------------------------------------------------------------
inline int gaussian_taps(double sigma, float* c, int max_half_width) {
    double wd = std::ceil(4.0 * sigma);
    long W = (long)(wd > 1 ? wd : 1);
    if (W > max_half_width) throw std::runtime_error("gaussian half-width too large for the tap table");
    float acc = 0;
    for (long j = 0; j < 2 * W + 1; ++j) {
        float d = ((float)(j - W)) / ((float)sigma);
        c[j] = (float)std::exp(-0.5 * (d * d));
        acc += c[j];
    }
    for (long j = 0; j < 2 * W + 1; ++j) c[j] /= acc;
    return (int)W;
}
inline void expn_table(double tab[257]) {
    for (int k = 0; k < 257; ++k) tab[k] = std::exp(-(double)k * (25.0 / 256));
}

// ---- SVD of an (H rows x W cols) matrix, CImg convention M(col,row) = m[row*W + col] -----------------------------
struct Mat {
    int w = 0, h = 0;  // width = columns, height = rows
    std::vector<double> d;
    Mat() {}
    Mat(int w_, int h_, double v = 0) : w(w_), h(h_), d((size_t)w_ * h_, v) {}
    double& operator()(int x, int y) { return d[(size_t)y * w + x]; }
    double operator()(int x, int y) const { return d[(size_t)y * w + x]; }
};

inline double cimg_hypot(double x, double y) {  // CImg.h:5845-5850
    double nx = std::fabs(x), ny = std::fabs(y), t;
    if (nx < ny) { t = nx; nx = ny; } else t = ny;
    if (nx > 0) { t /= nx; return nx * std::sqrt(1 + t * t); }
    return 0;
}

// CImg::_quicksort (CImg.h:25676-25740), decreasing order with permutation tracking.
inline void cimg_quicksort_dec(std::vector<double>& v, std::vector<int>& perm, long indm, long indM) {
    if (indm < indM) {
        const long mid = (indm + indM) / 2;
        if (v[indm] < v[mid]) { std::swap(v[indm], v[mid]); std::swap(perm[indm], perm[mid]); }
        if (v[mid] < v[indM]) { std::swap(v[indM], v[mid]); std::swap(perm[indM], perm[mid]); }
        if (v[indm] < v[mid]) { std::swap(v[indm], v[mid]); std::swap(perm[indm], perm[mid]); }
        if (indM - indm >= 3) {
            const double pivot = v[mid];
            long i = indm, j = indM;
            do {
                while (v[i] > pivot) ++i;
                while (v[j] < pivot) --j;
                if (i <= j) {
                    std::swap(perm[i], perm[j]);
                    std::swap(v[i], v[j]);
                    ++i; --j;
                }
            } while (i <= j);
            if (indm < j) cimg_quicksort_dec(v, perm, indm, j);
            if (i < indM) cimg_quicksort_dec(v, perm, i, indM);
        }
    }
}

// CImg<T>::SVD (CImg.h:25755-25890) with sorting=true, max_iteration=40, lambda=0.  A is (w x h).
inline void cimg_svd(const Mat& A, Mat& U, std::vector<double>& S, Mat& V) {
    const int width = A.w, height = A.h;
    U = A;
    S.assign(width, 0.0);
    V = Mat(width, width);
    std::vector<double> rv1(width, 0.0);
    double anorm = 0, c, f, g = 0, h, s, scale = 0;
    int l = 0, nm = 0;
    for (int i = 0; i < width; ++i) {
        l = i + 1; rv1[i] = scale * g; g = s = scale = 0;
        if (i < height) {
            for (int k = i; k < height; ++k) scale += std::fabs(U(i, k));
            if (scale) {
                for (int k = i; k < height; ++k) { U(i, k) /= scale; s += U(i, k) * U(i, k); }
                f = U(i, i); g = ((f >= 0 ? -1 : 1) * std::sqrt(s)); h = f * g - s; U(i, i) = f - g;
                for (int j = l; j < width; ++j) {
                    s = 0;
                    for (int k = i; k < height; ++k) s += U(i, k) * U(j, k);
                    f = s / h;
                    for (int k = i; k < height; ++k) U(j, k) += f * U(i, k);
                }
                for (int k = i; k < height; ++k) U(i, k) *= scale;
            }
        }
        S[i] = scale * g;
        g = s = scale = 0;
        if (i < height && i != width - 1) {
            for (int k = l; k < width; ++k) scale += std::fabs(U(k, i));
            if (scale) {
                for (int k = l; k < width; ++k) { U(k, i) /= scale; s += U(k, i) * U(k, i); }
                f = U(l, i); g = ((f >= 0 ? -1 : 1) * std::sqrt(s)); h = f * g - s; U(l, i) = f - g;
                for (int k = l; k < width; ++k) rv1[k] = U(k, i) / h;
                for (int j = l; j < height; ++j) {
                    s = 0;
                    for (int k = l; k < width; ++k) s += U(k, j) * U(k, i);
                    for (int k = l; k < width; ++k) U(k, j) += s * rv1[k];
                }
                for (int k = l; k < width; ++k) U(k, i) *= scale;
            }
        }
        anorm = (double)std::max((float)anorm, (float)(std::fabs(S[i]) + std::fabs(rv1[i])));
    }
    for (int i = width - 1; i >= 0; --i) {
        if (i < width - 1) {
            if (g) {
                for (int j = l; j < width; ++j) V(i, j) = (U(j, i) / U(l, i)) / g;
                for (int j = l; j < width; ++j) {
                    s = 0;
                    for (int k = l; k < width; ++k) s += U(k, i) * V(j, k);
                    for (int k = l; k < width; ++k) V(j, k) += s * V(i, k);
                }
            }
            for (int j = l; j < width; ++j) V(j, i) = V(i, j) = 0.0;
        }
        V(i, i) = 1.0; g = rv1[i]; l = i;
    }
    for (int i = std::min(width, height) - 1; i >= 0; --i) {
        l = i + 1; g = S[i];
        for (int j = l; j < width; ++j) U(j, i) = 0;
        if (g) {
            g = 1 / g;
            for (int j = l; j < width; ++j) {
                s = 0;
                for (int k = l; k < height; ++k) s += U(i, k) * U(j, k);
                f = (s / U(i, i)) * g;
                for (int k = i; k < height; ++k) U(j, k) += f * U(i, k);
            }
            for (int j = i; j < height; ++j) U(i, j) *= g;
        } else
            for (int j = i; j < height; ++j) U(i, j) = 0;
        ++U(i, i);
    }
    for (int k = width - 1; k >= 0; --k) {
        for (unsigned int its = 0; its < 40; ++its) {
            bool flag = true;
            for (l = k; l >= 1; --l) {
                nm = l - 1;
                if ((std::fabs(rv1[l]) + anorm) == anorm) { flag = false; break; }
                if ((std::fabs(S[nm]) + anorm) == anorm) break;
            }
            if (flag) {
                c = 0; s = 1;
                for (int i = l; i <= k; ++i) {
                    f = s * rv1[i]; rv1[i] = c * rv1[i];
                    if ((std::fabs(f) + anorm) == anorm) break;
                    g = S[i]; h = cimg_hypot(f, g); S[i] = h; h = 1 / h; c = g * h; s = -f * h;
                    for (int j = 0; j < height; ++j) {
                        const double y = U(nm, j), z = U(i, j);
                        U(nm, j) = y * c + z * s; U(i, j) = z * c - y * s;
                    }
                }
            }
            const double z = S[k];
            if (l == k) {
                if (z < 0) { S[k] = -z; for (int j = 0; j < width; ++j) V(k, j) = -V(k, j); }
                break;
            }
            nm = k - 1;
            double x = S[l], y = S[nm];
            g = rv1[nm]; h = rv1[k];
            f = ((y - z) * (y + z) + (g - h) * (g + h)) / std::max(1e-25, 2 * h * y);
            g = cimg_hypot(f, 1.0);
            f = ((x - z) * (x + z) + h * ((y / (f + (f >= 0 ? g : -g))) - h)) / std::max(1e-25, x);
            c = s = 1;
            for (int j = l; j <= nm; ++j) {
                const int i = j + 1;
                g = rv1[i]; h = s * g; g = c * g;
                double y2 = S[i];
                double z2 = cimg_hypot(f, h);
                rv1[j] = z2; c = f / std::max(1e-25, z2); s = h / std::max(1e-25, z2);
                f = x * c + g * s; g = g * c - x * s; h = y2 * s; y2 *= c;
                for (int jj = 0; jj < width; ++jj) {
                    const double xx = V(j, jj), zz = V(i, jj);
                    V(j, jj) = xx * c + zz * s; V(i, jj) = zz * c - xx * s;
                }
                z2 = cimg_hypot(f, h); S[j] = z2;
                if (z2) { z2 = 1 / std::max(1e-25, z2); c = f * z2; s = h * z2; }
                f = c * g + s * y2; x = c * y2 - s * g;
                for (int jj = 0; jj < height; ++jj) {
                    const double yy = U(j, jj);
                    z2 = U(i, jj);
                    U(j, jj) = yy * c + z2 * s; U(i, jj) = z2 * c - yy * s;
                }
            }
            rv1[l] = 0; rv1[k] = f; S[k] = x;
        }
    }
    // sorting (decreasing singular values), CImg.h:25874-25886
    std::vector<int> perm(width);
    for (int i = 0; i < width; ++i) perm[i] = i;
    cimg_quicksort_dec(S, perm, 0, width - 1);
    std::vector<double> tmp(width);
    for (int k = 0; k < height; ++k) {
        for (int y = 0; y < width; ++y) tmp[y] = U(perm[y], k);
        for (int y = 0; y < width; ++y) U(y, k) = tmp[y];
    }
    for (int k = 0; k < width; ++k) {
        for (int y = 0; y < width; ++y) tmp[y] = V(perm[y], k);
        for (int y = 0; y < width; ++y) V(y, k) = tmp[y];
    }
}

// P = pinv(A), A (4 cols x n rows) -> P (n cols x 4 rows).  CImg.h:25293-25302.
inline Mat pinv(const Mat& A) {
    Mat U, V;
    std::vector<double> S;
    cimg_svd(A, U, S, V);
    const int W = A.w, H = A.h;
    double smax = S[0];
    for (int i = 1; i < W; ++i) smax = std::max(smax, S[i]);
    const double tolerance = (double)(1.11e-16f * (float)std::max(W, H)) * smax;  // float literal x unsigned -> float
    for (int xx = 0; xx < W; ++xx) {
        const double s = S[xx], invs = s > tolerance ? 1 / s : 0;
        for (int y = 0; y < W; ++y) V(xx, y) *= invs;
    }
    // P = V * U^T : P(i,j) = sum_k V(k,j) * Ut(i,k), Ut(i,k) = U(k,i);  P is (H cols x W rows)
    Mat P(H, W);
    for (int j = 0; j < W; ++j)
        for (int i = 0; i < H; ++i) {
            double value = 0;
            for (int k = 0; k < W; ++k) value += V(k, j) * U(k, i);
            P(i, j) = value;
        }
    return P;
}
// x = P * b (operator*, CImg.h:12244-12262)
inline void pinv_apply(const Mat& P, const std::vector<double>& b, double* x) {
    const int H = P.w, W = P.h;
    for (int j = 0; j < W; ++j) {
        double value = 0;
        for (int k = 0; k < H; ++k) value += P(k, j) * b[k];
        x[j] = value;
    }
}
// x = pinv(A) * b.  The reference solves its two right-hand sides (x' and y', ImageProcess.cpp:500-529) with two
// get_solve calls on the same A; the pseudo-inverse depends on A alone, so callers with both sides use pinv once.
inline void pinv_solve(const Mat& A, const std::vector<double>& b, double* x) { pinv_apply(pinv(A), b, x); }

// ---- CImg van Vliet recursive Gaussian, order 0 (CImg.h:35045-35065) ---------------------------------------------
struct VanVliet {
    double filter[4];  // B, -b1, -b2, -b3
    double M[9];       // Triggs matrix
    double sum;        // filter[0]^2
    double bnd;        // 1 - a1 - a2 - a3
};
inline VanVliet vanvliet_coeffs(float sigma) {
    VanVliet v;
    const float nsigma = sigma;
    const double nnsigma = nsigma < 0.5f ? 0.5f : nsigma, m0 = 1.16680, m1 = 1.10783, m2 = 1.40586, m1sq = m1 * m1,
                 m2sq = m2 * m2,
                 q = (nnsigma < 3.556 ? -0.2568 + 0.5784 * nnsigma + 0.0561 * nnsigma * nnsigma
                                      : 2.5091 + 0.9804 * (nnsigma - 3.556)),
                 qsq = q * q, scale = (m0 + q) * (m1sq + m2sq + 2 * m1 * q + qsq),
                 b1 = -q * (2 * m0 * m1 + m1sq + m2sq + (2 * m0 + 4 * m1) * q + 3 * qsq) / scale,
                 b2 = qsq * (m0 + 2 * m1 + 3 * q) / scale, b3 = -qsq * q / scale, B = (m0 * (m1sq + m2sq)) / scale;
    v.filter[0] = B; v.filter[1] = -b1; v.filter[2] = -b2; v.filter[3] = -b3;
    const double sumsq = v.filter[0], a1 = v.filter[1], a2 = v.filter[2], a3 = v.filter[3];
    v.sum = sumsq * sumsq;
    const double scaleM = 1.0 / ((1.0 + a1 - a2 + a3) * (1.0 - a1 - a2 - a3) * (1.0 + a2 + (a1 - a3) * a3));
    v.M[0] = scaleM * (-a3 * a1 + 1.0 - a3 * a3 - a2);
    v.M[1] = scaleM * (a3 + a1) * (a2 + a3 * a1);
    v.M[2] = scaleM * a3 * (a1 + a3 * a2);
    v.M[3] = scaleM * (a1 + a3 * a2);
    v.M[4] = -scaleM * (a2 - 1.0) * (a2 + a3 * a1);
    v.M[5] = -scaleM * a3 * (a3 * a1 + a3 * a3 + a2 - 1.0);
    v.M[6] = scaleM * (a3 * a1 + a2 + a1 * a1 - a2 * a2);
    v.M[7] = scaleM * (a1 * a2 + a3 * a2 * a2 - a1 * a3 * a3 - a3 * a3 * a3 - a3 * a2 + a3);
    v.M[8] = scaleM * a3 * (a1 + a3 * a2);
    v.bnd = 1.0 - a1 - a2 - a3;
    return v;
}

// ---- CImg Deriche recursive filter, order 0 (CImg.h:34809-34823, 34843-34844) --------------------------------------
// All float, written in the reference's expression order; std::exp on a float is expf, evaluated at run time by the
// same libm as the reference (the volatile keeps the compiler from folding it with a differently rounded constant).
struct Deriche {
    float a0, a1, a2, a3, b1, b2, coefp, coefn;
};
inline Deriche deriche_coeffs(float sigma) {
    volatile float vs = sigma;
    const float nsigma = vs;
    const float nnsigma = nsigma < 0.1f ? 0.1f : nsigma, alpha = 1.695f / nnsigma, ema = (float)std::exp(-alpha),
                ema2 = (float)std::exp(-2 * alpha), b1 = -2 * ema, b2 = ema2;
    const float k = (1 - ema) * (1 - ema) / (1 + 2 * alpha * ema - ema2);
    Deriche d;
    d.a0 = k;
    d.a1 = k * (alpha - 1) * ema;
    d.a2 = k * (alpha + 1) * ema;
    d.a3 = -k * ema2;
    d.b1 = b1;
    d.b2 = b2;
    d.coefp = (d.a0 + d.a1) / (1 + b1 + b2);
    d.coefn = (d.a2 + d.a3) / (1 + b1 + b2);
    return d;
}

// ---- CImg resize tables ------------------------------------------------------------------------------------------
// moving average n -> m (m < n), CImg.h:29543-29555: out[t] = (sum_i in[src_i] * wgt_i) / n, terms in this order.
struct MovAvgTable {
    std::vector<int> start;   // [m+1] prefix offsets into src/wgt
    std::vector<int> src;
    std::vector<float> wgt;
};
inline MovAvgTable movavg_table(unsigned n, unsigned m) {
    MovAvgTable T;
    T.start.push_back(0);
    for (unsigned a = n * m, b = n, c = m, s = 0, t = 0; a;) {
        const unsigned d = std::min(b, c);
        a -= d; b -= d; c -= d;
        T.src.push_back((int)s);
        T.wgt.push_back((float)d);
        if (!b) { ++t; b = n; T.start.push_back((int)T.src.size()); }
        if (!c) { ++s; c = m; }
    }
    if (T.start.size() != (size_t)m + 1) throw std::runtime_error("moving-average table: 32-bit overflow (canvas too large)");
    return T;
}
// linear interpolation n -> m (m > n, n > 1), CImg.h:29628-29640: out[x] = (1-alpha[x])*in[pos[x]] + alpha[x]*in[min(pos+1,n-1)]
struct LinearTable {
    std::vector<int> pos;
    std::vector<double> alpha;
};
inline LinearTable linear_table(unsigned n, unsigned m) {
    LinearTable T;
    T.pos.resize(m);
    T.alpha.resize(m);
    const double fx = (m > 1 ? (n - 1.0) / (m - 1) : 0);
    double curr = 0, old = 0;
    unsigned p = 0;
    for (unsigned x = 0; x < m; ++x) {
        T.alpha[x] = curr - (unsigned int)curr;
        T.pos[x] = (int)p;
        old = curr;
        curr = std::min(n - 1.0, curr + fx);
        p += (unsigned int)curr - (unsigned int)old;
    }
    return T;
}

// ---- cylinder table (Projection.cpp:27-37): k[i] for i along the short side ---------------------------------------
inline void cylinder_table(int short_side, std::vector<float>& k) {
    const int width = short_side;
    const float tanVal = (float)std::tan(15 * 3.14159265358979323846 / 180.0);
    float r = (float)((width / 2.0) / tanVal);
    k.resize(width);
    for (int i = 0; i < width; ++i) {
        float dst_x = (float)(i - width / 2);
        k[i] = (float)(r / std::sqrt(std::pow((double)r, 2) + std::pow((double)dst_x, 2)));
    }
}

}  // namespace hostnum
}  // namespace pb
