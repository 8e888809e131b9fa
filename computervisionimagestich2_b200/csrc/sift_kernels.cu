// sift_kernels.cu -- sm_100a kernels for the Gaussian scale space, DoG extrema, refinement, gradient map,
// orientation histograms and 128-d descriptors.  Arithmetic lives in sift_device.cuh (bit-exact bodies); this file
// maps items to threads, stages tiles in shared memory and vectorises HBM access.
//
// Compiled with -fmad=false (no FMA contraction): the blur is a chain of separate FMUL + FADD per tap, exactly
// the reference's `acc += v * c` (vl/imopv.c:163-187).
#include "sift_kernels.h"
#include "common.h"
#include "ktimer.h"

namespace pb {

// ---------------------------------------------------------------------------------------------------------
// format conversion
// ---------------------------------------------------------------------------------------------------------
__global__ void u8_to_f32_kernel(const unsigned char* __restrict__ src, int src_pitch, float* __restrict__ dst, int w,
                                 int h, int pitch) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < w && y < h) dst[(long)y * pitch + x] = (float)src[(long)y * src_pitch + x];
}
void launch_u8_to_f32(const unsigned char* src, int src_pitch, float* dst, int w, int h, int pitch, cudaStream_t st) {
    KScope ks("sift.u8_to_f32", st, 5.0 * w * h);
    dim3 b(128, 2), g(div_up(w, 128), div_up(h, 2));
    u8_to_f32_kernel<<<g, b, 0, st>>>(src, src_pitch, dst, w, h, pitch);
    PB_KERNEL_CHECK();
}
void launch_copy_f32(const float* src, int src_pitch, float* dst, int w, int h, int pitch, cudaStream_t st) {
    PB_CUDA(cudaMemcpy2DAsync(dst, (size_t)pitch * 4, src, (size_t)src_pitch * 4, (size_t)w * 4, h,
                              cudaMemcpyDeviceToDevice, st));
}

// rows[i] of src -> row i of dst (128 floats each): the sorted, de-duplicated feature table assembled on the device
__global__ void gather_rows128_kernel(const float* __restrict__ src, const int* __restrict__ rows, int n,
                                      float* __restrict__ dst) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    reinterpret_cast<float4*>(dst + (size_t)i * 128)[lane] = reinterpret_cast<const float4*>(src + (size_t)rows[i] * 128)[lane];
}
void launch_gather_rows128(const float* src, const int* rows, int n, float* dst, cudaStream_t st) {
    if (n <= 0) return;
    KScope ks("sift.table_gather", st, 1028.0 * n);
    gather_rows128_kernel<<<div_up(n, 8), 256, 0, st>>>(src, rows, n, dst);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// separable Gaussian blur
// ---------------------------------------------------------------------------------------------------------
// Vertical pass: each thread owns one column x and R consecutive rows.  It loads the R + 2W source samples it
// needs once into registers (coalesced across the warp) and reuses every sample for up to R outputs.
template <int W, int R>
__global__ void __launch_bounds__(256) blur_v_kernel(const float* __restrict__ src, float* __restrict__ dst, int w,
                                                     int h, int pitch, const BlurTaps taps) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y0 = (blockIdx.y * 8 + threadIdx.y) * R;
    if (x >= w || y0 >= h) return;
    float v[R + 2 * W];
#pragma unroll
    for (int i = 0; i < R + 2 * W; ++i) {
        int yy = y0 - W + i;
        yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
        v[i] = src[(long)yy * pitch + x];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j <= 2 * W; ++j) acc = acc + v[r + j] * taps.c[2 * W - j];
        if (y0 + r < h) dst[(long)(y0 + r) * pitch + x] = acc;
    }
}

// Horizontal pass: a CTA stages TY rows x (TX + 2 WP) columns in shared memory (edge columns replicated), then
// every thread produces 4 adjacent outputs from 2WP+4 samples fetched with 128-bit shared loads and writes them
// with one 128-bit store.  DS: also emit the 2:1 sub-sampled image for the next octave (vl/sift.c:179-194).
template <int W, bool DS>
__global__ void __launch_bounds__(256) blur_h_kernel(const float* __restrict__ src, float* __restrict__ dst, int w,
                                                     int h, int pitch, const BlurTaps taps, float* __restrict__ ds,
                                                     int ds_pitch, int w2, int h2) {
    constexpr int WP = (W + 3) / 4 * 4;
    constexpr int TX = 512, TY = 8;
    constexpr int L = TX + 2 * WP;
    __shared__ __align__(128) float tile[TY][L];
    __shared__ unsigned long long bar;
    const int tile_x0 = blockIdx.x * TX;
    const int tile_y0 = blockIdx.y * TY;
    const int tid = threadIdx.y * 128 + threadIdx.x;
    const int rows = min(TY, h - tile_y0);
    // Interior tiles (no column clamping needed): every tile row is one contiguous, 16-byte aligned run of L floats in
    // HBM, fetched by ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP) whose completion is counted in bytes on an mbarrier
    // -- no per-element LDG / STS, no clamp arithmetic.  Tiles touching the left or right image border replicate the edge
    // column (vl_imconvcol_vf pads by continuity, vl/imopv.c:137-198) with the element-wise fill.
    const bool interior = tile_x0 - WP >= 0 && tile_x0 - WP + L <= pitch && tile_x0 + TX + WP <= w;
    if (interior) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)(rows * L * 4)) : "memory");
            for (int r = 0; r < rows; ++r)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 (unsigned)__cvta_generic_to_shared(&tile[r][0])),
                             "l"(src + (long)(tile_y0 + r) * pitch + (tile_x0 - WP)), "r"((unsigned)(L * 4)), "r"(b)
                             : "memory");
        }
        unsigned ok;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok)
                         : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0u)
                         : "memory");
        } while (!ok);
    } else {
        for (int r = 0; r < rows; ++r) {
            const float* row = src + (long)(tile_y0 + r) * pitch;
            for (int c = tid; c < L; c += 256) {
                int gx = tile_x0 - WP + c;
                gx = gx < 0 ? 0 : (gx > w - 1 ? w - 1 : gx);
                tile[r][c] = row[gx];
            }
        }
        __syncthreads();
    }
    const int xl = threadIdx.x * 4;
    const int x0 = tile_x0 + xl;
    if (x0 >= w) return;
    constexpr int NV = (2 * WP + 4) / 4;
    constexpr int D = WP - W;
    for (int r = threadIdx.y; r < TY; r += 2) {
        const int y = tile_y0 + r;
        if (y >= h) break;
        float in[NV * 4];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            float4 t = *reinterpret_cast<const float4*>(&tile[r][xl + 4 * q]);
            in[4 * q + 0] = t.x; in[4 * q + 1] = t.y; in[4 * q + 2] = t.z; in[4 * q + 3] = t.w;
        }
        float acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float a = 0.0f;
#pragma unroll
            for (int j = 0; j <= 2 * W; ++j) a = a + in[D + i + j] * taps.c[2 * W - j];
            acc[i] = a;
        }
        *reinterpret_cast<float4*>(dst + (long)y * pitch + x0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (DS) {
            if ((y & 1) == 0 && (y >> 1) < h2) {
                float* drow = ds + (long)(y >> 1) * ds_pitch;
                const int xd = x0 >> 1;
                if (xd + 1 < w2) *reinterpret_cast<float2*>(drow + xd) = make_float2(acc[0], acc[2]);
                else if (xd < w2) drow[xd] = acc[0];
            }
        }
    }
}

// Any half-width: one thread per output sample, taps from the parameter table.
__global__ void blur_generic_kernel(const float* __restrict__ src, float* __restrict__ dst, int w, int h, int pitch,
                                    const BlurTaps taps, int vertical) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float r = vertical ? blur_sample(src + x, pitch, h, y, taps.c, taps.W)
                       : blur_sample(src + (long)y * pitch, 1, w, x, taps.c, taps.W);
    dst[(long)y * pitch + x] = r;
}

__global__ void downsample2_kernel(const float* __restrict__ src, int pitch, float* __restrict__ dst, int dst_pitch,
                                   int w2, int h2) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < w2 && y < h2) dst[(long)y * dst_pitch + x] = src[(long)(2 * y) * pitch + 2 * x];
}
void launch_downsample2(const float* src, int w, int h, int pitch, float* dst, int dst_pitch, cudaStream_t st) {
    int w2 = w / 2, h2 = h / 2;
    if (w2 <= 0 || h2 <= 0) return;
    dim3 b(128, 2), g(div_up(w2, 128), div_up(h2, 2));
    downsample2_kernel<<<g, b, 0, st>>>(src, pitch, dst, dst_pitch, w2, h2);
    PB_KERNEL_CHECK();
}

template <int W>
static void launch_blur_w(const float* src, float* tmp, float* dst, int w, int h, int pitch, const BlurTaps& taps,
                          float* ds, int ds_pitch, cudaStream_t st) {
    constexpr int R = 8;
    dim3 bv(32, 8), gv(div_up(w, 32), div_up(h, 8 * R));
    {
        KScope ks("sift.blur_v", st, 4.0 * w * h);
        blur_v_kernel<W, R><<<gv, bv, 0, st>>>(src, tmp, w, h, pitch, taps);
        PB_KERNEL_CHECK();
    }
    KScope ks("sift.blur_h", st, 4.0 * w * h + (ds ? 1.0 * w * h : 0.0));
    dim3 bh(128, 2), gh(div_up(w, 512), div_up(h, 8));
    if (ds)
        blur_h_kernel<W, true><<<gh, bh, 0, st>>>(tmp, dst, w, h, pitch, taps, ds, ds_pitch, w / 2, h / 2);
    else
        blur_h_kernel<W, false><<<gh, bh, 0, st>>>(tmp, dst, w, h, pitch, taps, nullptr, 0, 0, 0);
    PB_KERNEL_CHECK();
}

void launch_blur(const float* src, float* tmp, float* dst, int w, int h, int pitch, const BlurTaps& taps, float* ds,
                 int ds_pitch, cudaStream_t st) {
    switch (taps.W) {
    case 7: launch_blur_w<7>(src, tmp, dst, w, h, pitch, taps, ds, ds_pitch, st); return;
    case 10: launch_blur_w<10>(src, tmp, dst, w, h, pitch, taps, ds, ds_pitch, st); return;
    case 13: launch_blur_w<13>(src, tmp, dst, w, h, pitch, taps, ds, ds_pitch, st); return;
    case 19: launch_blur_w<19>(src, tmp, dst, w, h, pitch, taps, ds, ds_pitch, st); return;
    default: break;
    }
    KScope ks("sift.blur_generic", st, 8.0 * w * h);
    dim3 b(128, 2), g(div_up(w, 128), div_up(h, 2));
    blur_generic_kernel<<<g, b, 0, st>>>(src, tmp, w, h, pitch, taps, 1);
    PB_KERNEL_CHECK();
    blur_generic_kernel<<<g, b, 0, st>>>(tmp, dst, w, h, pitch, taps, 0);
    PB_KERNEL_CHECK();
    if (ds) launch_downsample2(dst, w, h, pitch, ds, ds_pitch, st);
}

// ---------------------------------------------------------------------------------------------------------
// DoG (only materialised for the vl_sift shim mirror and the parity tests; the detector forms it on the fly)
// ---------------------------------------------------------------------------------------------------------
__global__ void dog_kernel(OctaveView ov, float* __restrict__ dog) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int l = blockIdx.z;
    if (x < ov.w && y < ov.h) dog[((long)l * ov.h + y) * ov.pitch + x] = dog_at(ov, x, y, l);
}
void launch_dog(const OctaveView& ov, float* dog, cudaStream_t st) {
    dim3 b(128, 2), g(div_up(ov.w, 128), div_up(ov.h, 2), ov.nlevels - 1);
    dog_kernel<<<g, b, 0, st>>>(ov, dog);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// extrema detection: a CTA stages a (TX+2) x (TY+2) window of all DoG levels in shared memory (DoG is formed
// from the GSS levels while loading, so it never travels through HBM) and tests the interior pixels of the
// detection levels against their 26 neighbours.
// ---------------------------------------------------------------------------------------------------------
template <int NL>  // number of GSS levels (5 for S = 2)
__global__ void __launch_bounds__(256) detect_kernel(OctaveView ov, SiftConsts sc, Cand* __restrict__ cand,
                                                     int* __restrict__ count, int cap) {
    constexpr int TX = 64, TY = 16;
    constexpr int ND = NL - 1;
    __shared__ float d[ND][TY + 2][TX + 2];
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
    const long ls = (long)ov.pitch * ov.h;
    for (int i = threadIdx.x; i < (TY + 2) * (TX + 2); i += 256) {
        int ry = i / (TX + 2), rx = i - ry * (TX + 2);
        int gx = tx0 - 1 + rx, gy = ty0 - 1 + ry;
        if (gx >= 0 && gx < ov.w && gy >= 0 && gy < ov.h) {
            const float* p = ov.gss + (long)gy * ov.pitch + gx;
            float prev = p[0];
#pragma unroll
            for (int l = 0; l < ND; ++l) {
                float cur = p[(l + 1) * ls];
                d[l][ry][rx] = cur - prev;
                prev = cur;
            }
        } else {
#pragma unroll
            for (int l = 0; l < ND; ++l) d[l][ry][rx] = 0.0f;
        }
    }
    __syncthreads();
    const double tp = sc.peak_thresh;
    for (int i = threadIdx.x; i < TX * TY; i += 256) {
        int ry = i / TX, rx = i - ry * TX;
        int gx = tx0 + rx, gy = ty0 + ry;
        if (gx < 1 || gx > ov.w - 2 || gy < 1 || gy > ov.h - 2) continue;
#pragma unroll
        for (int l = 1; l <= ND - 2; ++l) {
            const float v = d[l][ry + 1][rx + 1];
            bool gt = ((double)v >= 0.8 * tp), lt = ((double)v <= -0.8 * tp);
#pragma unroll
            for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
                for (int dy = 0; dy <= 2; ++dy)
#pragma unroll
                    for (int dx = 0; dx <= 2; ++dx) {
                        if (dl == 0 && dy == 1 && dx == 1) continue;
                        const float n = d[l + dl][ry + dy][rx + dx];
                        gt = gt && (v > n);
                        lt = lt && (v < n);
                    }
            if (gt || lt) {
                int idx = atomicAdd(count, 1);
                if (idx < cap) cand[idx] = Cand{gx, gy, l + sc.s_min};
            }
        }
    }
}

__global__ void detect_generic_kernel(OctaveView ov, SiftConsts sc, Cand* __restrict__ cand, int* __restrict__ count,
                                      int cap) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < 1 || x > ov.w - 2 || y < 1 || y > ov.h - 2) return;
    for (int l = 1; l <= ov.nlevels - 3; ++l)
        if (is_extremum(ov, x, y, l, sc.peak_thresh)) {
            int idx = atomicAdd(count, 1);
            if (idx < cap) cand[idx] = Cand{x, y, l + sc.s_min};
        }
}

void launch_detect(const OctaveView& ov, const SiftConsts& sc, Cand* cand, int* count, int cap, cudaStream_t st) {
    if (ov.w < 3 || ov.h < 3) return;
    KScope ks("sift.detect", st, 16.0 * ov.w * ov.h);
    if (ov.nlevels == 5) {
        dim3 g(div_up(ov.w, 64), div_up(ov.h, 16));
        detect_kernel<5><<<g, 256, 0, st>>>(ov, sc, cand, count, cap);
    } else {
        dim3 b(128, 2), g(div_up(ov.w, 128), div_up(ov.h, 2));
        detect_generic_kernel<<<g, b, 0, st>>>(ov, sc, cand, count, cap);
    }
    PB_KERNEL_CHECK();
}

__global__ void refine_kernel(OctaveView ov, SiftConsts sc, const Cand* __restrict__ cand,
                              const int* __restrict__ count, int cap, RefinedKey* __restrict__ out, double xper) {
    int n = *count;
    n = n < cap ? n : cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Cand c = cand[i];
        out[i] = refine_key(ov, sc, c.x, c.y, c.s, xper);
    }
}
void launch_refine(const OctaveView& ov, const SiftConsts& sc, const Cand* cand, const int* count, int cap,
                   RefinedKey* out, double xper, cudaStream_t st) {
    KScope ks("sift.refine", st, 0);
    int blocks = div_up(cap, 128);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    refine_kernel<<<blocks, 128, 0, st>>>(ov, sc, cand, count, cap, out, xper);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// gradient map of the detection levels (modulus, angle) interleaved
// ---------------------------------------------------------------------------------------------------------
__global__ void gradient_kernel(OctaveView ov, int first_level, float* __restrict__ grad) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int l = blockIdx.z;
    if (x >= ov.w || y >= ov.h) return;
    const float* lev = ov.gss + (long)(first_level + l) * ov.pitch * ov.h;
    float m, a;
    gradient_at(lev, ov.w, ov.h, ov.pitch, x, y, &m, &a);
    reinterpret_cast<float2*>(grad)[((long)l * ov.h + y) * ov.pitch + x] = make_float2(m, a);
}
void launch_gradient(const OctaveView& ov, const SiftConsts& sc, float* grad, cudaStream_t st) {
    int nl = (sc.s_max - 2) - (sc.s_min + 1) + 1;
    if (nl <= 0) return;
    KScope ks("sift.gradient", st, 12.0 * nl * ov.w * ov.h);
    dim3 b(128, 2), g(div_up(ov.w, 128), div_up(ov.h, 2), nl);
    gradient_kernel<<<g, b, 0, st>>>(ov, 1, grad);  // grad level l <-> s = s_min+1+l <-> GSS level index 1+l
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// orientation + descriptor: ONE WARP per keypoint / per (keypoint, angle), items of all octaves of an image in one
// launch.  Both kernels run in two phases per chunk of 32 patch samples (raster order):
//   phase A  the 32 lanes evaluate 32 samples in parallel (the expensive double-precision part) and park the
//            per-sample contributions in shared memory;
//   phase B  the lanes switch roles and become BIN OWNERS: each owner adds, in sample order, the contributions that
//            land in its bins.  Every histogram bin therefore receives the reference's addends in the reference's
//            order (bit-exact), yet no lane ever walks the whole patch serially.
// ---------------------------------------------------------------------------------------------------------
// order (optional): slot -> keypoint index, largest window first.  The window radius grows with sigma (13^2 .. 49^2
// samples and more), one warp works on one keypoint, and the kernel ends when its slowest warp does: issued in array
// order the big ones land anywhere and the tail of the launch runs at a fraction of the machine.
#ifndef PB_ORIENT_MINB
#define PB_ORIENT_MINB 1
#endif
__global__ void __launch_bounds__(128, PB_ORIENT_MINB) orient_kernel(OctaveSet os, SiftConsts sc, const double* __restrict__ expn_tab,
                                                     const KeyIn* __restrict__ keys, int nkeys,
                                                     const int* __restrict__ order,
                                                     int* __restrict__ nangles, double* __restrict__ angles) {
    enum { nbins = 36 };
    __shared__ double tab[257];
    __shared__ double hsm[4][nbins];
    __shared__ double sv0[4][32], sv1[4][32];
    __shared__ unsigned binmask[4][nbins];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) tab[i] = expn_tab[i];
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + wid;
    if (slot >= nkeys) return;
    const int ki = order ? order[slot] : slot;
    const KeyIn k = keys[ki];
    const OctaveView ov = os.ov[k.oct];
    const double xper = os.xper[k.oct];
    const int w = ov.w, h = ov.h;
    const double x = (double)k.x / xper, y = (double)k.y / xper, sigma = (double)k.sigma / xper;
    const int xi = (int)(x + 0.5), yi = (int)(y + 0.5), si = k.is;
    const double sigmaw = 1.5 * sigma;
    const double Wd = floor(3.0 * sigmaw);
    const int W = (int)(Wd > 1 ? Wd : 1);
    if (xi < 0 || xi > w - 1 || yi < 0 || yi > h - 1 || si < sc.s_min + 1 || si > sc.s_max - 2) {
        if (lane == 0) { nangles[ki] = 0; for (int j = 0; j < 4; ++j) angles[ki * 4 + j] = 0; }
        return;
    }
    const float2* pt = reinterpret_cast<const float2*>(ov.grad) + (long)(si - sc.s_min - 1) * ov.h * ov.pitch;
    const int ys0 = (-W > -yi) ? -W : -yi, ys1 = (W < h - 1 - yi) ? W : h - 1 - yi;
    const int xs0 = (-W > -xi) ? -W : -xi, xs1 = (W < w - 1 - xi) ? W : w - 1 - xi;
    const int nxw = xs1 - xs0 + 1, total = nxw * (ys1 - ys0 + 1);
    const double r2max = W * W + 0.6, den = 2 * sigmaw * sigmaw;
    double h0 = 0.0, h1 = 0.0;  // bins lane and lane + 32
    unsigned* bm = binmask[wid];
    for (int base = 0; base < total; base += 32) {
        // ---- phase A: one sample per lane ----
        const int i = base + lane;
        int b0 = 99;  // 99 = no contribution
        if (i < total) {
            const int ry = i / nxw, ys = ys0 + ry, xs = xs0 + (i - ry * nxw);
            const double dx = (double)(xi + xs) - x, dy = (double)(yi + ys) - y;
            const double r2 = dx * dx + dy * dy;
            if (!(r2 >= r2max)) {
                const float2 g = pt[(long)(yi + ys) * ov.pitch + (xi + xs)];
                const double wgt = fast_expn(tab, r2 / den);
                const double mod = g.x, ang = g.y;
                const double fbin = nbins * ang / (2 * kPi);
                const int bin = floor_d(fbin - 0.5);
                const double rbin = fbin - bin - 0.5;
                sv0[wid][lane] = (1 - rbin) * mod * wgt;
                sv1[wid][lane] = (rbin)*mod * wgt;
                b0 = (bin + nbins) % nbins;
            }
        }
        bm[lane] = 0;
        if (lane < nbins - 32) bm[lane + 32] = 0;
        __syncwarp();
        const unsigned grp = __match_any_sync(0xffffffffu, b0);
        if (b0 != 99) bm[b0] = grp;  // all members of a group write the same value
        __syncwarp();
        // ---- phase B: lane owns bins lane and lane + 32; sample j adds v0 to bin b0_j and v1 to bin b0_j + 1 ----
        {
            const unsigned m0 = bm[lane], m1 = bm[(lane + nbins - 1) % nbins];
            for (unsigned m = m0 | m1; m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                h0 += ((m0 >> j) & 1u) ? sv0[wid][j] : sv1[wid][j];
            }
        }
        if (lane < nbins - 32) {
            const unsigned m0 = bm[lane + 32], m1 = bm[lane + 31];
            for (unsigned m = m0 | m1; m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                h1 += ((m0 >> j) & 1u) ? sv0[wid][j] : sv1[wid][j];
            }
        }
        __syncwarp();
    }
    double* hist = hsm[wid];
    hist[lane] = h0;
    if (lane < nbins - 32) hist[lane + 32] = h1;
    __syncwarp();
    // vl/sift.c:1000-1036.  The reference's in-place loop carries the OLD value of the previous bin (`prev`), i.e. every pass
    // computes new[i] = ((old[i-1] + old[i]) + old[i+1]) / 3 from the old histogram alone (circular): lane = bin, the same
    // operations in the same order per bin, 2 divisions per lane and pass instead of 36.  (The first version ran the serial
    // loop redundantly in every lane on a private 36-double copy: 216 divisions and 167 registers per thread.)
    const bool two = lane < nbins - 32;   // lanes 0..3 also own bins 32..35
    double c0 = h0, c1 = h1;
    for (int iter = 0; iter < 6; iter++) {
        const double a0 = hist[(lane + nbins - 1) % nbins], b0v = hist[lane + 1];   // lane + 1 <= 32 < nbins
        const double n0 = ((a0 + c0) + b0v) / 3.0;
        double n1 = 0.0;
        if (two) {
            const double a1 = hist[lane + 31], b1 = hist[(lane + 33) % nbins];
            n1 = ((a1 + c1) + b1) / 3.0;
        }
        __syncwarp();
        hist[lane] = c0 = n0;
        if (two) hist[lane + 32] = c1 = n1;
        __syncwarp();
    }
    double maxh = (c0 > 0.0) ? c0 : 0.0;                        // running maximum from 0 (vl/sift.c:1010-1012); max is order-free
    if (two) maxh = (maxh > c1) ? maxh : c1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double m = __shfl_xor_sync(0xffffffffu, maxh, o);
        maxh = (maxh > m) ? maxh : m;
    }
    // peaks in bin order, at most 4 (vl/sift.c:1015-1036)
    bool pk0, pk1 = false;
    double ang0, ang1 = 0.0;
    {
        const double hm = hist[(lane + nbins - 1) % nbins], hp = hist[lane + 1];
        pk0 = c0 > 0.8 * maxh && c0 > hm && c0 > hp;
        const double di = -0.5 * (hp - hm) / (hp + hm - 2 * c0);
        ang0 = 2 * kPi * (lane + di + 0.5) / nbins;
    }
    if (two) {
        const double hm = hist[lane + 31], hp = hist[(lane + 33) % nbins];
        pk1 = c1 > 0.8 * maxh && c1 > hm && c1 > hp;
        const double di = -0.5 * (hp - hm) / (hp + hm - 2 * c1);
        ang1 = 2 * kPi * ((lane + 32) + di + 0.5) / nbins;
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, pk0), m1 = __ballot_sync(0xffffffffu, pk1);
    const int n0 = __popc(m0), na = min(4, n0 + __popc(m1));
    if (lane == 0) nangles[ki] = na;
    if (lane < 4 && lane >= na) angles[ki * 4 + lane] = 0;
    if (pk0) {
        const int r = __popc(m0 & ((1u << lane) - 1u));
        if (r < 4) angles[ki * 4 + r] = ang0;
    }
    if (pk1) {
        const int r = n0 + __popc(m1 & ((1u << lane) - 1u));
        if (r < 4) angles[ki * 4 + r] = ang1;
    }
}
void launch_orient(const OctaveSet& os, const SiftConsts& sc, const double* expn_tab, const KeyIn* keys, int nkeys,
                   const int* order, int* nangles, double* angles, cudaStream_t st) {
    if (nkeys <= 0) return;
    KScope ks("sift.orient", st, 36.0 * nkeys);
    orient_kernel<<<div_up(nkeys, 4), 128, 0, st>>>(os, sc, expn_tab, keys, nkeys, order, nangles, angles);
    PB_KERNEL_CHECK();
}

// Descriptor.  Phase A: lane = sample (descriptor_sample, sift_device.cuh); the samples of the conservative
// per-row ranges are flattened so that all 32 lanes stay busy whatever the patch geometry.  Phase B: lane = (cell,
// orientation parity): lane 2c + p owns the four orientation bins p, p+2, p+4, p+6 of cell c; a sample touches one
// even and one odd orientation bin, so both lanes of a cell consume every sample that reaches the cell, in order.
constexpr int kDescRows = 192;   // rows of the patch handled per pass (taller patches take several passes)
struct DescWarpSmem {
    int rowstart[kDescRows + 1];
    int rowx0[kDescRows];
    // [cell][sample of the batch * 2 + orientation parity]: the addend of that sample for that cell.  Pitch 66, not 64: lane
    // 2c + p reads column 2q + p of row c, and with a pitch that is a multiple of 32 the sixteen cells would share a bank
    float vals[16][66];
    float hist[128];
};
// order (optional): slot -> job index, largest patch first (841 .. 13 k samples per descriptor; see orient_kernel)
#ifndef PB_DESCR_MINB
#define PB_DESCR_MINB 1
#endif
__global__ void __launch_bounds__(128, PB_DESCR_MINB) descr_kernel(OctaveSet os, SiftConsts sc, const double* __restrict__ expn_tab,
                                                    const KeyIn* __restrict__ keys, const DescJob* __restrict__ jobs,
                                                    int njobs, const int* __restrict__ order, float* __restrict__ descr,
                                                    int* __restrict__ written) {
    __shared__ double tab[257];
    __shared__ DescWarpSmem wsm[4];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) tab[i] = expn_tab[i];
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + wid;
    if (slot >= njobs) return;
    const int job = order ? order[slot] : slot;
    DescWarpSmem& S = wsm[wid];
    const DescJob j = jobs[job];
    const KeyIn k = keys[j.key];
    const DescFrame F = descriptor_frame(os.ov[k.oct], sc, 0, 0, k.is, k.x, k.y, k.sigma, os.xper[k.oct], j.angle, j.st0,
                                         j.ct0);
    if (!F.valid) {
        if (lane == 0) written[job] = 0;
        return;
    }
    const int cell = lane >> 1, par = lane & 1;
    const int cx = (cell & 3) - 2, cy = (cell >> 2) - 2;
    float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;   // orientation bins par, par+2, par+4, par+6 of my cell
    int ry0, ry1;
    descriptor_rows(F, &ry0, &ry1);
    const float2* gp = reinterpret_cast<const float2*>(F.pt);
    for (int rg = ry0; rg <= ry1; rg += kDescRows) {
        const int nr = (ry1 - rg + 1) < kDescRows ? (ry1 - rg + 1) : kDescRows;
        // conservative column range of every row (lanes = rows) + exclusive prefix of the row lengths
        int running = 0;
        for (int r0 = 0; r0 < nr; r0 += 32) {
            const int r = r0 + lane;
            int cnt = 0;
            if (r < nr) {
                int x0, x1;
                descriptor_row_range(F, rg + r, &x0, &x1);
                cnt = x1 >= x0 ? x1 - x0 + 1 : 0;
                S.rowx0[r] = x0;
            }
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (r < nr) S.rowstart[r + 1] = running + incl;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) S.rowstart[0] = 0;
        __syncwarp();
        const int total = running;
        int r = 0;
        for (int base = 0; base < total; base += 32) {
            // ---- phase A ----
            const int i = base + lane;
            unsigned cellmask = 0;
            int kk0 = 0, kk1 = 0;   // which of the lane-owned bins (parity 0 / parity 1) this sample feeds
            if (i < total) {
                while (i >= S.rowstart[r + 1]) ++r;
                const int dxi = S.rowx0[r] + (i - S.rowstart[r]), dyi = rg + r;
                const float2 g = gp[(long)(F.yi + dyi) * F.pitch + (F.xi + dxi)];
                const DescSample sm = descriptor_sample(F, tab, dxi, dyi, g.x, g.y);
                if (sm.active) {
                    // orientation bins bint and bint + 1: the even one goes to the parity-0 lane of a cell, the odd one to
                    // the parity-1 lane; within a lane the four owned bins are (bin >> 1)
                    const int sel0 = (sm.bint & 1) ? 1 : 0, sel1 = 1 - sel0;
                    kk0 = ((sm.bint + sel0) & 7) >> 1;
                    kk1 = ((sm.bint + sel1) & 7) >> 1;
                    const float a0 = sm.at[sel0], a1 = sm.at[sel1];
                    // the (up to) four cells (binx + dx, biny + dy) inside the grid: the sample's addends are written where
                    // the owning lanes will pick them up with ONE shared-memory read, and flagged in a 4x4 bit mask
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
                            const int ccx = sm.binx + dx, ccy = sm.biny + dy;
                            if (ccx >= -2 && ccx <= 1 && ccy >= -2 && ccy <= 1) {
                                const int c = (ccy + 2) * 4 + (ccx + 2);
                                const float w = sm.wxy[dx][dy];
                                S.vals[c][lane * 2] = w * a0;
                                S.vals[c][lane * 2 + 1] = w * a1;
                                cellmask |= 1u << c;
                            }
                        }
                }
            }
            __syncwarp();
            unsigned mine = 0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const unsigned m = __ballot_sync(0xffffffffu, (cellmask >> c) & 1u);
                if (c == cell) mine = m;
            }
            // every lane is also a sample: publish ITS bin selectors for both parities
            const unsigned k0a = __ballot_sync(0xffffffffu, kk0 & 1), k1a = __ballot_sync(0xffffffffu, kk0 >> 1);
            const unsigned k0b = __ballot_sync(0xffffffffu, kk1 & 1), k1b = __ballot_sync(0xffffffffu, kk1 >> 1);
            const unsigned klo = par ? k0b : k0a, khi = par ? k1b : k1a;
            // ---- phase B ----
            // lane 2c + p adds, in sample order, the addends of the samples that reach cell c: one read per addend
            const float* mv = &S.vals[cell][par];
            for (; mine; mine &= mine - 1) {
                const int q = __ffs(mine) - 1;
                const float v = mv[q * 2];
                const int kk = ((klo >> q) & 1u) | (((khi >> q) & 1u) << 1);
                // add-then-select: no branch on kk (lanes disagree on it; as an if-chain this loop cost 134 cycles per
                // addend, 40 % of the kernel), and no "+ 0" that could touch a signed zero
                const float t0 = h0 + v, t1 = h1 + v, t2 = h2 + v, t3 = h3 + v;
                h0 = kk == 0 ? t0 : h0;
                h1 = kk == 1 ? t1 : h1;
                h2 = kk == 2 ? t2 : h2;
                h3 = kk == 3 ? t3 : h3;
            }
            __syncwarp();
        }
    }
    // 128 bins to shared memory in descriptor order, then the two sequential L2 normalisations (vl/sift.c:1415-1436)
    float* hs = S.hist;
    hs[cell * 8 + par] = h0; hs[cell * 8 + par + 2] = h1; hs[cell * 8 + par + 4] = h2; hs[cell * 8 + par + 6] = h3;
    __syncwarp();
    float norm = 0.0f;
    for (int i = 0; i < 128; ++i) norm += hs[i] * hs[i];
    norm = fast_sqrt_f(norm) + kEpsF;
    const bool zero = sc.norm_thresh != 0 && (double)norm < sc.norm_thresh;
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        float q = hs[lane * 4 + t] / norm;
        if ((double)q > 0.2) q = (float)0.2;
        v[t] = q;
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; ++t) hs[lane * 4 + t] = v[t];
    __syncwarp();
    float norm2 = 0.0f;
    for (int i = 0; i < 128; ++i) norm2 += hs[i] * hs[i];
    norm2 = fast_sqrt_f(norm2) + kEpsF;
    float4 o = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(v[0] / norm2, v[1] / norm2, v[2] / norm2, v[3] / norm2);
    reinterpret_cast<float4*>(descr + (size_t)job * 128)[lane] = o;
    if (lane == 0) written[job] = 1;
}
void launch_descr(const OctaveSet& os, const SiftConsts& sc, const double* expn_tab, const KeyIn* keys,
                  const DescJob* jobs, int njobs, const int* order, float* descr, int* written, double patch_bytes,
                  cudaStream_t st) {
    if (njobs <= 0) return;
    KScope ks("sift.descr", st, patch_bytes + 512.0 * njobs);
    descr_kernel<<<div_up(njobs, 4), 128, 0, st>>>(os, sc, expn_tab, keys, jobs, njobs, order, descr, written);
    PB_KERNEL_CHECK();
}

}  // namespace pb
