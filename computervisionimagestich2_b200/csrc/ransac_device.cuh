// ransac_device.cuh -- arithmetic of the RANSAC hypothesis fit and scoring (ImageProcess.cpp:395-497).
//
//   getHomographyMat (ImageProcess.cpp:439-462): two 4x4 systems A h = b with rows [x, y, x*y, 1] solved by CImg's
//   LU with implicit row scaling (CImg.h:25351-25356 -> _LU 25911-25950 -> _solve 25400-25421), in double.
//   getInlinerIndex (ImageProcess.cpp:473-497): warp (double polynomial rounded to float), float distance, < 4.0.
#pragma once
#include "exact_math.cuh"
#include "canvas_device.cuh"

namespace pb {

// CImg LU solve of an N x N system; A row-major A[row*N + col] (CImg A(col,row)); b is overwritten by the solution.
template <int N>
PB_HD void lu_solve(const double* Ain, double* b) {
    double lu[N * N];
    for (int i = 0; i < N * N; ++i) lu[i] = Ain[i];
#define PB_LU(col, row) lu[(row)*N + (col)]
    double vv[N];
    double indx[N];
    int imax = 0;
    bool singular = false;
    for (int i = 0; i < N; ++i) {
        double vmax = 0;
        for (int j = 0; j < N; ++j) {
            const double tmp = abs_d(PB_LU(j, i));
            if (tmp > vmax) vmax = tmp;
        }
        if (vmax == 0) { singular = true; break; }
        vv[i] = 1 / vmax;
    }
    if (singular) {
        for (int i = 0; i < N; ++i) indx[i] = 0;
        for (int i = 0; i < N * N; ++i) lu[i] = 0;
    } else {
        for (int j = 0; j < N; ++j) {
            for (int i = 0; i < j; ++i) {
                double sum = PB_LU(j, i);
                for (int k = 0; k < i; ++k) sum -= PB_LU(k, i) * PB_LU(j, k);
                PB_LU(j, i) = sum;
            }
            double vmax = 0;
            for (int i = j; i < N; ++i) {
                double sum = PB_LU(j, i);
                for (int k = 0; k < j; ++k) sum -= PB_LU(k, i) * PB_LU(j, k);
                PB_LU(j, i) = sum;
                const double tmp = vv[i] * abs_d(sum);
                if (tmp >= vmax) { vmax = tmp; imax = i; }
            }
            if (j != imax) {
                for (int k = 0; k < N; ++k) {
                    double t = PB_LU(k, imax); PB_LU(k, imax) = PB_LU(k, j); PB_LU(k, j) = t;
                }
                vv[imax] = vv[j];
            }
            indx[j] = (double)imax;
            if (PB_LU(j, j) == 0) PB_LU(j, j) = 1e-20;
            const double tmp = 1 / PB_LU(j, j);
            for (int i = j + 1; i < N; ++i) PB_LU(j, i) = PB_LU(j, i) * tmp;
        }
    }
    int ii = -1;
    for (int i = 0; i < N; ++i) {
        const int ip = (int)indx[i];
        double sum = b[ip];
        b[ip] = b[i];
        if (ii >= 0) { for (int j = ii; j <= i - 1; ++j) sum -= PB_LU(j, i) * b[j]; }
        else if (sum != 0) ii = i;
        b[i] = sum;
    }
    for (int i = N - 1; i >= 0; --i) {
        double sum = b[i];
        for (int j = i + 1; j < N; ++j) sum -= PB_LU(j, i) * b[j];
        b[i] = sum / PB_LU(i, i);
    }
#undef PB_LU
}

// ImageProcess.cpp:439-462 on 4 pairs (sx, sy) -> (dx, dy); H8 in Homography constructor order.
PB_HD void fit4(const float* sx, const float* sy, const float* dx, const float* dy, double* H8) {
    double A[16], b[4];
    for (int i = 0; i < 4; ++i) {
        A[i * 4 + 0] = (double)sx[i];
        A[i * 4 + 1] = (double)sy[i];
        A[i * 4 + 2] = (double)sx[i] * (double)sy[i];
        A[i * 4 + 3] = 1.0;
        b[i] = (double)dx[i];
    }
    lu_solve<4>(A, b);
    for (int i = 0; i < 4; ++i) H8[i] = b[i];
    for (int i = 0; i < 4; ++i) b[i] = (double)dy[i];
    lu_solve<4>(A, b);
    for (int i = 0; i < 4; ++i) H8[4 + i] = b[i];
}

// ImageProcess.cpp:478-493
PB_HD bool is_inlier(const double* H8, float sx, float sy, float dx, float dy) {
    float x = warp_x(H8, sx, sy);
    float y = warp_y(H8, sx, sy);
#if defined(__CUDA_ARCH__)
    float distance = sqrtf((x - dx) * (x - dx) + (y - dy) * (y - dy));
#else
    float distance = __builtin_sqrtf((x - dx) * (x - dx) + (y - dy) * (y - dy));
#endif
    return (double)distance < 4.0;
}

}  // namespace pb
