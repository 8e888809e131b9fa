// canvas_device.cuh -- per-pixel / per-line bodies of the image-space kernels: cylindrical projection, grayscale,
// homography warp + shift, the multiband-blend primitives (recursive Gaussian, 2:1 reduce, linear expand, blend,
// collapse) and the equalisation / luminance-mix tail.  __host__ __device__ so tests/emul can run them on a CPU box.
//
// Reference (file:line):
//   Projection::bilinearInterpolation / imageProjection ... Projection.cpp:3-18, 20-73
//   toGrayScale ........................................... ImageProcess.cpp:27-40
//   getX/YAfterWarping, warpingImageByHomography .......... ImageProcess.cpp:465-471, 596-606
//   movingImageByOffset ................................... ImageProcess.cpp:608-620
//   blendTwoImages ........................................ ImageProcess.cpp:648-773
//   CImg vanvliet order 0 ................................. CImg.h:34905-34932
//   CImg moving-average / linear resize ................... CImg.h:29543-29560, 29641-29652
//   equalization mode 1, final Y mix ...................... equalization.cpp:74-131, ImageProcess.cpp:237-268
#pragma once
#include "exact_math.cuh"

namespace pb {

typedef unsigned char u8;

PB_HD float floor_float(float x) {
#if defined(__CUDA_ARCH__)
    return floorf(x);
#else
    return __builtin_floorf(x);
#endif
}
PB_HD float ceil_float(float x) {
#if defined(__CUDA_ARCH__)
    return ceilf(x);
#else
    return __builtin_ceilf(x);
#endif
}

// Projection.cpp:3-18 on one channel plane (w x h, row stride w).
PB_HD u8 bilinear_u8(const u8* __restrict__ plane, int w, int h, float x, float y) {
    int x_floor = (int)floor_float(x), y_floor = (int)floor_float(y);
    int x_ceil = ceil_float(x) >= (float)(w - 1) ? (w - 1) : (int)ceil_float(x);
    int y_ceil = ceil_float(y) >= (float)(h - 1) ? (h - 1) : (int)ceil_float(y);
    float a = x - (float)x_floor, b = y - (float)y_floor;
    float ld = (float)plane[(long)y_floor * w + x_floor];
    float lt = (float)plane[(long)y_ceil * w + x_floor];
    float rd = (float)plane[(long)y_floor * w + x_ceil];
    float rt = (float)plane[(long)y_ceil * w + x_ceil];
    float v = (1 - a) * (1 - b) * ld + a * (1 - b) * rd + a * b * rt + (1 - a) * b * lt;
    return (u8)v;
}

// Projection.cpp:20-73: source sample position of destination pixel (x, y); ktab[i] = r / sqrt(r^2 + (i - short/2)^2)
// along the short side (host table, hostnum::cylinder_table).  Returns false when the pixel stays black.
PB_HD bool project_source(int x, int y, int W, int H, const float* __restrict__ ktab, float* sx, float* sy) {
    const bool flag = W > H;
    const int width = flag ? H : W;
    const int height = flag ? W : H;
    if (flag) {
        float dst_x = (float)(y - width / 2);
        float dst_y = (float)(x - height / 2);
        float k = ktab[y];
        float src_x = dst_x / k;
        float src_y = dst_y / k;
        float u = src_x + (float)(width / 2), v = src_y + (float)(height / 2);
        if (u >= 0 && u < (float)H && v >= 0 && v < (float)W) { *sx = v; *sy = u; return true; }
        return false;
    } else {
        float dst_x = (float)(x - width / 2);
        float dst_y = (float)(y - height / 2);
        float k = ktab[x];
        float src_x = dst_x / k;
        float src_y = dst_y / k;
        float u = src_x + (float)(width / 2), v = src_y + (float)(height / 2);
        if (u >= 0 && u < (float)W && v >= 0 && v < (float)H) { *sx = u; *sy = v; return true; }
        return false;
    }
}

// ImageProcess.cpp:27-40
PB_HD u8 gray_u8(u8 r, u8 g, u8 b) {
    double v = 0.299 * (double)(float)r + 0.587 * (double)(float)g + 0.114 * (double)(float)b;
    return (u8)v;
}

// ImageProcess.cpp:465-471: H8 = {a,b,c,d, e,f,g,h}: x' = a x + b y + c x y + d, y' = e x + f y + g x y + h
PB_HD float warp_x(const double* H8, float x, float y) {
    return (float)(H8[0] * (double)x + H8[1] * (double)y + H8[2] * (double)x * (double)y + H8[3]);
}
PB_HD float warp_y(const double* H8, float x, float y) {
    return (float)(H8[4] * (double)x + H8[5] * (double)y + H8[6] * (double)x * (double)y + H8[7]);
}

// ImageProcess.cpp:596-606 (truncating sampler): source pixel index or -1
PB_HD long warp_source(const double* H8, int x, int y, float offx, float offy, int sw, int sh) {
    float fx = (float)x + offx, fy = (float)y + offy;
    int nx = (int)warp_x(H8, fx, fy);
    int ny = (int)warp_y(H8, fx, fy);
    if (nx >= 0 && nx < sw && ny >= 0 && ny < sh) return (long)ny * sw + nx;
    return -1;
}

// --- CImg vanvliet order 0 with Neumann boundary on one line (CImg.h:34905-34932), in place -----------------------
struct IirCoef {
    double f1, f2, f3;  // filter[1..3]
    double sumsq, sum;  // filter[0], filter[0]^2
    double M[9];
    double bnd;         // 1 - a1 - a2 - a3
};
PB_HD void iir_line(float* data, int N, long off, const IirCoef& c) {
    double val[4] = {0, 0, 0, 0};
    const double iplus = (double)data[(long)(N - 1) * off];
    // forward
    for (int k = 1; k < 4; ++k) val[k] = (double)data[0] / c.sumsq;
    float* p = data;
    for (int n = 0; n < N; ++n) {
        val[0] = (double)(*p);
        val[0] += val[1] * c.f1;
        val[0] += val[2] * c.f2;
        val[0] += val[3] * c.f3;
        *p = (float)val[0];
        p += off;
        val[3] = val[2]; val[2] = val[1]; val[1] = val[0];
    }
    p -= off;
    // backward with Triggs boundary
    {
        const double uplus = iplus / c.bnd, vplus = uplus / c.bnd;
        const double unp = val[1] - uplus, unp1 = val[2] - uplus, unp2 = val[3] - uplus;
        val[0] = (c.M[0] * unp + c.M[1] * unp1 + c.M[2] * unp2 + vplus) * c.sum;
        val[1] = (c.M[3] * unp + c.M[4] * unp1 + c.M[5] * unp2 + vplus) * c.sum;
        val[2] = (c.M[6] * unp + c.M[7] * unp1 + c.M[8] * unp2 + vplus) * c.sum;
        *p = (float)val[0];
        p -= off;
        val[3] = val[2]; val[2] = val[1]; val[1] = val[0];
    }
    for (int n = 1; n < N; ++n) {
        val[0] = (double)(*p);
        val[0] *= c.sum;
        val[0] += val[1] * c.f1;
        val[0] += val[2] * c.f2;
        val[0] += val[3] * c.f3;
        *p = (float)val[0];
        p -= off;
        val[3] = val[2]; val[2] = val[1]; val[1] = val[0];
    }
}

// --- CImg deriche order 0 with Neumann boundary on one line (CImg.h:34777-34808), src -> dst -----------------------
// The filter the src/ex6 variant's pyramid uses: get_blur(2) has is_gaussian = false (CImg.h:35139-35142,
// src/ex6/ImageProcess.cpp:702-705).  All float; every product is rounded before it is added (no contraction):
//   causal      Y[m]  = ((a0 x[m] + a1 x[m-1]) - b1 Y[m-1]) - b2 Y[m-2]
//   anticausal  yc[n] = ((a2 x[n+1] + a3 x[n+2]) - b1 yc[n+1]) - b2 yc[n+2];   out[n] = Y[n] + yc[n]
struct DericheCoef {
    float a0, a1, a2, a3, b1, b2, coefp, coefn;
};
struct DericheFwd {   // causal state
    float xp, yp, yb;
    PB_HD void init(float x0, const DericheCoef& c) { xp = x0; yb = yp = c.coefp * xp; }
    PB_HD float step(float xc, const DericheCoef& c) {
        const float yc = c.a0 * xc + c.a1 * xp - c.b1 * yp - c.b2 * yb;
        xp = xc; yb = yp; yp = yc;
        return yc;
    }
};
struct DericheBwd {   // anticausal state
    float xn, xa, yn, ya;
    PB_HD void init(float xlast, const DericheCoef& c) { xn = xa = xlast; yn = ya = c.coefn * xn; }
    PB_HD float step(float xc, const DericheCoef& c) {
        const float yc = c.a2 * xn + c.a3 * xa - c.b1 * yn - c.b2 * ya;
        xa = xn; xn = xc; ya = yn; yn = yc;
        return yc;
    }
};
PB_HD void deriche_line(const float* src, float* dst, int N, long off, const DericheCoef& c) {
    DericheFwd f;
    f.init(src[0], c);
    for (int m = 0; m < N; ++m) dst[(long)m * off] = f.step(src[(long)m * off], c);
    DericheBwd b;
    b.init(src[(long)(N - 1) * off], c);
    for (int n = N - 1; n >= 0; --n) {
        const float yc = b.step(src[(long)n * off], c);
        dst[(long)n * off] = dst[(long)n * off] + yc;
    }
}

// --- CImg moving-average resize along one axis (CImg.h:29543-29555): out = (sum_i in[src_i] * wgt_i) / n --------
PB_HD float movavg_sample(const float* __restrict__ line, long stride, const int* __restrict__ start,
                          const int* __restrict__ src, const float* __restrict__ wgt, int t, float n_as_float) {
    float acc = 0.0f;
    for (int i = start[t]; i < start[t + 1]; ++i) acc += line[(long)src[i] * stride] * wgt[i];
    return acc / n_as_float;
}

// --- CImg linear resize along one axis (CImg.h:29641-29652) ------------------------------------------------------
PB_HD float linear_sample(const float* __restrict__ line, long stride, int n, int pos, double alpha) {
    const float v1 = line[(long)pos * stride];
    const float v2 = pos < n - 1 ? line[(long)(pos + 1) * stride] : v1;
    return (float)((1 - alpha) * (double)v1 + alpha * (double)v2);
}

// --- blendTwoImages arithmetic -----------------------------------------------------------------------------------
PB_HD float blend_px(float a, float b, float m) {  // ImageProcess.cpp:748-751
    return (float)((double)(a * m) + (double)b * (1.0 - (double)m));
}
PB_HD float collapse_px(float lap, float expanded) {  // ImageProcess.cpp:766-769
    float e = lap + expanded;
    if (e > 255) e = 255;
    else if (e < 0) e = 0;
    return e;
}

// --- equalisation (equalization.cpp:74-100) and final mix (ImageProcess.cpp:240-268) ------------------------------
PB_HD float clamp256(float v) { return v > 0 ? (v < 256 ? v : 255.0f) : 0.0f; }
PB_HD void rgb_to_ycbcr_f(u8 r8, u8 g8, u8 b8, float* Y, float* Cb, float* Cr) {
    const double r = (double)(float)r8, g = (double)(float)g8, b = (double)(float)b8;
    float y = (float)(0.299 * r + 0.857 * g + 0.114 * b);
    float cb = (float)(128.0 - 0.168736 * r - 0.331264 * g + 0.5 * b);
    float cr = (float)(128.0 + 0.5 * r - 0.418688 * g - 0.081312 * b);
    *Y = clamp256(y); *Cb = clamp256(cb); *Cr = clamp256(cr);
}
PB_HD void ycbcr_to_rgb_u8(float Y, float Cb, float Cr, u8* r, u8* g, u8* b) {
    float R = (float)((double)Y + 1.402 * ((double)Cr - 128.0));
    float G = (float)((double)Y - 0.34414 * ((double)Cb - 128.0) - 0.71414 * ((double)Cr - 128.0));
    float B = (float)((double)Y + 1.772 * ((double)Cb - 128.0));
    *r = (u8)clamp256(R); *g = (u8)clamp256(G); *b = (u8)clamp256(B);
}
// One pixel of the whole tail: rgb (blended panorama) -> final rgb, given the equalisation LUT of the Y channel.
// The luminance mix is Y * num / den + Y_eq / den: 19/20 in the root variant (ImageProcess.cpp:261), 5/6 in src/ex6
// (src/ex6/ImageProcess.cpp:270).
PB_HD void equalize_mix_px(u8 r, u8 g, u8 b, const int* __restrict__ lut, u8* orr, u8* og, u8* ob, double num = 19.0,
                           double den = 20.0) {
    float Y, Cb, Cr;
    rgb_to_ycbcr_f(r, g, b, &Y, &Cb, &Cr);
    // equalization.cpp: the YCbCr image is stored as unsigned char (truncation), Y is mapped, then back to RGB
    u8 y8 = (u8)Y, cb8 = (u8)Cb, cr8 = (u8)Cr;
    u8 yeq = (u8)lut[y8];
    u8 tr, tg, tb;
    ycbcr_to_rgb_u8((float)yeq, (float)cb8, (float)cr8, &tr, &tg, &tb);
    float Yt, Cbt, Crt;
    rgb_to_ycbcr_f(tr, tg, tb, &Yt, &Cbt, &Crt);
    float Ym = (float)((double)Y * num / den + (double)Yt / den);
    ycbcr_to_rgb_u8(Ym, Cb, Cr, orr, og, ob);
}
PB_HD u8 luma_bin(u8 r, u8 g, u8 b) {
    float Y, Cb, Cr;
    rgb_to_ycbcr_f(r, g, b, &Y, &Cb, &Cr);
    return (u8)Y;
}

// --- Reinhard l-alpha-beta colour transfer (the reference's class transfer, transfer.cpp:125-225) -----------------------
// Per-pixel halves, in the reference's promotion order (double constants x float operands, results narrowed to float
// where the reference assigns to a float).  logf / pow are library functions: glibc's on the host, CUDA's on the device --
// each within an ulp of the true value, not necessarily of each other, which is why this stage's parity bar on the GPU
// is "within 1 LSB" (BASELINE.json north_star) rather than bit-exact; on the host the bodies reproduce the reference.
PB_HD void rgb_to_lab(float R, float G, float B, float* L, float* a, float* b) {
    float l = (float)(0.3811 * (double)R + 0.5783 * (double)G + 0.0402 * (double)B);
    float m = (float)(0.1967 * (double)R + 0.7244 * (double)G + 0.0782 * (double)B);
    float s = (float)(0.0241 * (double)R + 0.1288 * (double)G + 0.8444 * (double)B);
    if (l == 0) l = 1;
    if (m == 0) m = 1;
    if (s == 0) s = 1;
    // `log(l)` on a float with `using namespace std` is std::log(float) = logf; `log(10)` is the double overload
    // (transfer.cpp:187-189).  Host: glibc's logf.  Device: the correctly rounded float log via the double routine
    // (CUDA's logf is a 1-ulp approximation of its own; glibc's is within 0.82 ulp -- neither is the other).
    const double ln10 = log(10.0);
#ifdef __CUDA_ARCH__
    l = (float)((double)(float)log((double)l) / ln10);
    m = (float)((double)(float)log((double)m) / ln10);
    s = (float)((double)(float)log((double)s) / ln10);
#else
    l = (float)((double)logf(l) / ln10);
    m = (float)((double)logf(m) / ln10);
    s = (float)((double)logf(s) / ln10);
#endif
    const float paraA = (float)(1.0 / sqrt(3.0)), paraB = (float)(1.0 / sqrt(6.0)), paraC = (float)(1.0 / sqrt(2.0));
    *L = paraA * ((l + m) + s);
    *a = (float)((double)(paraB * l + paraB * m) - (2.0 * (double)paraB) * (double)s);
    *b = paraC * l - paraC * m;
}
PB_HD void lab_to_rgb(float L, float a, float b, float* R, float* G, float* B) {
    const float paraA = (float)(sqrt(3.0) / 3.0), paraB = (float)(sqrt(6.0) / 6.0), paraC = (float)(sqrt(2.0) / 2.0);
    float l = (paraA * L + paraB * a) + paraC * b;
    float m = (paraA * L + paraB * a) - paraC * b;
    float s = (float)((double)(paraA * L) - (2.0 * (double)paraB) * (double)a);
    l = (float)pow(10.0, (double)l);
    m = (float)pow(10.0, (double)m);
    s = (float)pow(10.0, (double)s);
    float r = (float)((4.4679 * (double)l - 3.5873 * (double)m) + 0.1193 * (double)s);
    float g = (float)(((-1.2186) * (double)l + 2.3809 * (double)m) - 0.1624 * (double)s);
    float bb = (float)((0.0497 * (double)l - 0.2439 * (double)m) + 1.2045 * (double)s);
    *R = r > 0.0f ? (r < 255.0f ? r : 255.0f) : 0.0f;
    *G = g > 0.0f ? (g < 255.0f ? g : 255.0f) : 0.0f;
    *B = bb > 0.0f ? (bb < 255.0f ? bb : 255.0f) : 0.0f;
}
// statistics matching (transfer.cpp:166-172): ((v - meanSrc) * sdTemplate) / sdSrc + meanTemplate, all float
PB_HD float lab_match(float v, float mean_src, float sd_src, float mean_tem, float sd_tem) {
    return ((v - mean_src) * sd_tem) / sd_src + mean_tem;
}

}  // namespace pb
