// canvas_kernels.cu -- sm_100a kernels of the image-space stages: cylindrical projection (+gray), homography warp
// + canvas shift, multiband blend (recursive Gaussian, reduce, expand/blend/collapse) and the equalisation tail.
// All of them are HBM-bound byte/float streaming; arithmetic is in canvas_device.cuh (bit-exact bodies).
#include "canvas_kernels.h"
#include <type_traits>
#include <cmath>
#include "common.h"
#include "ktimer.h"

namespace pb {

// ---------------------------------------------------------------------------------------------------------
// projection + gray
// ---------------------------------------------------------------------------------------------------------
__global__ void project_gray_kernel(const u8* __restrict__ src, int w, int h, const float* __restrict__ ktab,
                                    u8* __restrict__ dst, float* __restrict__ gray32, int gray_pitch,
                                    u8* __restrict__ gray8) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t n = (size_t)w * h, o = (size_t)y * w + x;
    float sx, sy;
    u8 r = 0, g = 0, b = 0;
    if (project_source(x, y, w, h, ktab, &sx, &sy)) {
        r = bilinear_u8(src, w, h, sx, sy);
        g = bilinear_u8(src + n, w, h, sx, sy);
        b = bilinear_u8(src + 2 * n, w, h, sx, sy);
    }
    dst[o] = r; dst[n + o] = g; dst[2 * n + o] = b;
    const u8 gr = gray_u8(r, g, b);
    if (gray32) gray32[(size_t)y * gray_pitch + x] = (float)gr;
    if (gray8) gray8[o] = gr;
}
void launch_project_gray(const u8* src, int w, int h, const float* ktab, u8* dst_rgb, float* gray_f32, int gray_pitch,
                         u8* gray8, cudaStream_t st) {
    KScope ks("canvas.project_gray", st, 10.0 * w * h);
    dim3 b(64, 4), g(div_up(w, 64), div_up(h, 4));
    project_gray_kernel<<<g, b, 0, st>>>(src, w, h, ktab, dst_rgb, gray_f32, gray_pitch, gray8);
    PB_KERNEL_CHECK();
}

__global__ void gray_kernel(const u8* __restrict__ rgb, int w, int h, float* __restrict__ gray32, int gray_pitch,
                            u8* __restrict__ gray8) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t n = (size_t)w * h, o = (size_t)y * w + x;
    const u8 gr = gray_u8(rgb[o], rgb[n + o], rgb[2 * n + o]);
    if (gray32) gray32[(size_t)y * gray_pitch + x] = (float)gr;
    if (gray8) gray8[o] = gr;
}
void launch_gray(const u8* rgb, int w, int h, float* gray_f32, int gray_pitch, u8* gray8, cudaStream_t st) {
    KScope ks("canvas.gray", st, 7.0 * w * h);
    dim3 b(128, 2), g(div_up(w, 128), div_up(h, 2));
    gray_kernel<<<g, b, 0, st>>>(rgb, w, h, gray_f32, gray_pitch, gray8);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// BMP pixel area <-> planar RGB (the container step either side of the path: CImg::_load_bmp 24-bpp branch,
// CImg.h:48533-48546, and _save_bmp, :52614).  BMP rows are BGR, padded to 4 bytes, bottom-up unless height < 0.
// ---------------------------------------------------------------------------------------------------------
__global__ void bmp_to_planar_kernel(const u8* __restrict__ bmp, int stride, int w, int h, int bottom_up, u8* __restrict__ dst) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const u8* p = bmp + (size_t)(bottom_up ? h - 1 - y : y) * stride + 3 * x;
    const size_t n = (size_t)w * h, o = (size_t)y * w + x;
    dst[o] = p[2]; dst[n + o] = p[1]; dst[2 * n + o] = p[0];
}
__global__ void planar_to_bmp_kernel(const u8* __restrict__ src, int w, int h, int stride, u8* __restrict__ bmp) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;   // x == w .. covers the row padding
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= h) return;
    u8* row = bmp + (size_t)(h - 1 - y) * stride;
    const size_t n = (size_t)w * h, o = (size_t)y * w + x;
    if (x < w) { row[3 * x] = src[2 * n + o]; row[3 * x + 1] = src[n + o]; row[3 * x + 2] = src[o]; }
    else if (x == w) for (int k = 3 * w; k < stride; ++k) row[k] = 0;
}
void launch_bmp_to_planar(const u8* bmp, int stride, int w, int h, bool bottom_up, u8* dst, cudaStream_t st) {
    KScope ks("io.bmp_decode", st, 6.0 * w * h);
    dim3 b(64, 4), g(div_up(w, 64), div_up(h, 4));
    bmp_to_planar_kernel<<<g, b, 0, st>>>(bmp, stride, w, h, bottom_up ? 1 : 0, dst);
    PB_KERNEL_CHECK();
}
void launch_planar_to_bmp(const u8* src, int w, int h, int stride, u8* bmp, cudaStream_t st) {
    KScope ks("io.bmp_encode", st, 6.0 * w * h);
    dim3 b(64, 4), g(div_up(w + 1, 64), div_up(h, 4));
    planar_to_bmp_kernel<<<g, b, 0, st>>>(src, w, h, stride, bmp);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// warp + shift
// ---------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void warp_shift_kernel(const u8* __restrict__ src, int sw, int sh, const double* __restrict__ H8g,
                                  float offx, float offy, const u8* __restrict__ prev, int pw, int ph, int ioffx,
                                  int ioffy, u8* __restrict__ a, u8* __restrict__ b, int cw, int ch) {
    __shared__ double H8[8];
    if (threadIdx.x < 8 && threadIdx.y == 0) H8[threadIdx.x] = H8g ? H8g[threadIdx.x] : 0.0;
    __syncthreads();
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cw || y >= ch) return;
    const size_t cn = (size_t)cw * ch, o = (size_t)y * cw + x;
    if (a) {
        long s = warp_source(H8, x, y, offx, offy, sw, sh);
        const size_t sn = (size_t)sw * sh;
#pragma unroll
        for (int c = 0; c < NCH; ++c) a[c * cn + o] = s >= 0 ? src[c * sn + s] : (u8)0;
    }
    if (b) {
        int nx = x + ioffx, ny = y + ioffy;
        const bool in = nx >= 0 && nx < pw && ny >= 0 && ny < ph;
        const size_t pn = (size_t)pw * ph, s = in ? (size_t)ny * pw + nx : 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) b[c * cn + o] = in ? prev[c * pn + s] : (u8)0;
    }
}
void launch_warp_shift(const u8* src, int sw, int sh, const double* H8, float offx, float offy, const u8* prev, int pw,
                       int ph, int ioffx, int ioffy, u8* a, u8* b, int cw, int ch, cudaStream_t st, int nch) {
    KScope ks("canvas.warp_shift", st, 4.0 * nch * cw * ch);
    dim3 bl(64, 4), g(div_up(cw, 64), div_up(ch, 4));
    if (nch == 1) warp_shift_kernel<1><<<g, bl, 0, st>>>(src, sw, sh, H8, offx, offy, prev, pw, ph, ioffx, ioffy, a, b, cw, ch);
    else warp_shift_kernel<3><<<g, bl, 0, st>>>(src, sw, sh, H8, offx, offy, prev, pw, ph, ioffx, ioffy, a, b, cw, ch);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// seam statistics + level 0 planes
// ---------------------------------------------------------------------------------------------------------
// kAllChannels: the src/ex6 variant counts a pixel when ALL THREE channels are non-zero
// (src/ex6/ImageProcess.cpp:651-660); the root variant tests channel 0 only (ImageProcess.cpp:661-671, quirk Q4).
template <bool kAllChannels>
__global__ void seam_stats_kernel(const u8* __restrict__ a, const u8* __restrict__ b, int cw, int ch,
                                  int* __restrict__ stats) {
    __shared__ unsigned s[4];
    if (threadIdx.x < 4) s[threadIdx.x] = 0;
    __syncthreads();
    const int mid_y = ch / 2;
    unsigned sa = 0, na = 0, so = 0, no = 0;
    const size_t n = (size_t)cw * ch;
    for (int x = threadIdx.x; x < cw; x += blockDim.x) {
        const size_t o = (size_t)mid_y * cw + x;
        bool ina = a[o] != 0, inb = b[o] != 0;
        if (kAllChannels) {
            ina = ina && a[n + o] != 0 && a[2 * n + o] != 0;
            inb = inb && b[n + o] != 0 && b[2 * n + o] != 0;
        }
        if (ina) {
            sa += (unsigned)x; ++na;
            if (inb) { so += (unsigned)x; ++no; }
        }
    }
    atomicAdd(&s[0], sa); atomicAdd(&s[1], na); atomicAdd(&s[2], so); atomicAdd(&s[3], no);
    __syncthreads();
    if (threadIdx.x < 4) stats[threadIdx.x] = (int)s[threadIdx.x];
}
void launch_seam_stats(const u8* a, const u8* b, int cw, int ch, int* stats, bool all_channels, cudaStream_t st) {
    KScope ks("blend.seam_stats", st, (all_channels ? 6.0 : 2.0) * cw);
    if (all_channels) seam_stats_kernel<true><<<1, 1024, 0, st>>>(a, b, cw, ch, stats);
    else seam_stats_kernel<false><<<1, 1024, 0, st>>>(a, b, cw, ch, stats);
    PB_KERNEL_CHECK();
}

// kDoubleSeam: the src/ex6 variant keeps the seam position in double (src/ex6/ImageProcess.cpp:678-697); the root
// variant narrows both ratios to float first (ImageProcess.cpp:684-698).
template <bool kDoubleSeam, int NCH>
__global__ void level0_kernel(const u8* __restrict__ a, const u8* __restrict__ b, int cw, int ch,
                              const int* __restrict__ stats, float* __restrict__ G0, int* __restrict__ err_flag) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int sum_a_x = stats[0], width_mid_a = stats[1], sum_overlap_x = stats[2], width_mid_overlap = stats[3];
    if (width_mid_a == 0 || width_mid_overlap == 0) {
        if (x == 0 && y == 0) *err_flag = 1;
        return;
    }
    if (x >= cw || y >= ch) return;
    float m;
    if (kDoubleSeam) {   // sums are exact in double (the reference accumulates them in double)
        const double ratio = (double)(unsigned)sum_a_x / (double)width_mid_a;
        const double overlap_ratio = (double)(unsigned)sum_overlap_x / (double)width_mid_overlap;
        if (ratio < overlap_ratio) m = ((double)x < overlap_ratio) ? 1.0f : 0.0f;
        else m = (x >= (int)(overlap_ratio + 1)) ? 1.0f : 0.0f;
    } else {
        const float ratio = (float)(1.0 * (double)sum_a_x / (double)width_mid_a);
        const float overlap_ratio = (float)(1.0 * (double)sum_overlap_x / (double)width_mid_overlap);
        if (ratio < overlap_ratio) m = ((float)x < overlap_ratio) ? 1.0f : 0.0f;
        else m = (x >= (int)(overlap_ratio + 1)) ? 1.0f : 0.0f;
    }
    const size_t n = (size_t)cw * ch, o = (size_t)y * cw + x;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        G0[c * n + o] = (float)a[c * n + o];
        G0[(NCH + c) * n + o] = (float)b[c * n + o];
    }
    G0[2 * NCH * n + o] = m;
}
void launch_level0(const u8* a, const u8* b, int cw, int ch, const int* stats, float* G0, int* err_flag,
                   bool double_seam, cudaStream_t st, int nch) {
    KScope ks("blend.level0", st, (10.0 * nch + 4.0) * cw * ch);
    dim3 bl(128, 2), g(div_up(cw, 128), div_up(ch, 2));
    if (nch == 1) {
        if (double_seam) level0_kernel<true, 1><<<g, bl, 0, st>>>(a, b, cw, ch, stats, G0, err_flag);
        else level0_kernel<false, 1><<<g, bl, 0, st>>>(a, b, cw, ch, stats, G0, err_flag);
    } else {
        if (double_seam) level0_kernel<true, 3><<<g, bl, 0, st>>>(a, b, cw, ch, stats, G0, err_flag);
        else level0_kernel<false, 3><<<g, bl, 0, st>>>(a, b, cw, ch, stats, G0, err_flag);
    }
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// recursive Gaussian (CImg vanvliet order 0, Neumann / Triggs boundary; CImg.h:34905-34932)
//
// A line (image row for the x pass, image column for the y pass) is a strictly serial third-order recurrence in
// double precision: out[n] = ((in[n] + f1 out[n-1]) + f2 out[n-2]) + f3 out[n-3], rounded to float on store, then the
// same backwards.  The rounding order is the reference's, so the dependent chain per sample (DMUL + 3 DADD) cannot
// be shortened; what CAN be removed is everything else on the critical path -- global-memory latency and the
// transposition the x pass needs.  One CTA = two warps that share 32 lines:
//   * the CONSUMER warp (lane = line) runs nothing but the recurrence, in place on 32x32 tiles in shared memory;
//   * the PRODUCER warp streams the tiles: cp.async (LDGSTS) global -> shared, completion signalled on an mbarrier
//     (cp.async.mbarrier.arrive), NS tiles in flight, and copies finished tiles back to HBM with coalesced stores.
// Tile layout in shared memory is [element][line] with a 33-float pitch: conflict-free for the consumer (lanes =
// consecutive lines), for the x-pass producer (lanes = consecutive elements of one line, i.e. a transposing copy)
// and for the y-pass producer (lanes = consecutive lines of one element).
// ---------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
// wait with back-off: a warp that polls an mbarrier competes with the consumer's LDS / STS for the shared-memory pipe
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    for (;;) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (ok) break;
        __nanosleep(200);
    }
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr int kIirNSDeriche = 4;   // tiles in flight per CTA (4.2 KB each); Deriche keeps two rings (input and causal output)
constexpr int kIirPitch = 33;      // floats between consecutive elements of a tile
}  // namespace

// One full 32-sample tile of one line, in place in shared memory (t = the lane's column of the tile).  The bare
// DMUL + 3 DADD chain is 32 cycles per sample (tools/ubench/iir_step.cu); with the two float<->double conversions
// (F2F: ~19 cycles latency, ~10 issue cycles per warp on a unit shared by the SM, tools/ubench/f2f_lat.cu) this plain
// form measures 38 cycles in isolation and 62 in the kernel.  Integer-pipe conversions and hand software pipelining
// were both tried and were slower (87 and 51 cycles).
// (Conversions on the integer pipe -- tools/ubench/cvt_exact.cuh, bit-identical to F2F -- were measured inside this kernel in
// round 2: x pass 2.60 -> 3.20 (widening only) -> 4.12 ms (both), y pass 1.83 -> 1.99 -> 2.43 ms on a 17997 x 2268 blend.
// ncu: the XU pipe is at 19 % in both passes; the conversions are not the limiter.)
// PHASED (y pass): three phases over the 32 samples of the tile, all in registers -- widen (independent conversions,
// pipelined), the recurrence alone (DMUL + 3 DADD per sample, nothing else on the chain), narrow + store.  In the one-loop
// form ptxas places the F2F.F64.F32 of a sample (~19 cycles) right before the DADD that needs it, between two links of the
// chain; it also re-merges the phases when they are only written as separate loops, so the chain is made to DEPEND on all
// the conversions: `zero` is a kernel argument that is 0 at run time, OR-ed (masked) into the state before the first link.
template <bool FWD, bool PHASED, bool ROWMAJOR>
__device__ __forceinline__ void iir_tile32(float* t, double& v1, double& v2, double& v3, const IirCoef& c, int zero) {
    if (ROWMAJOR) {
        // x pass: the lane's 32 samples are contiguous in shared memory (line pitch 36 floats: 16-byte aligned, and the
        // eight lanes of a 128-bit wavefront cover all 32 banks): 8 LDS.128 + 8 STS.128 per tile instead of 32 + 32.  ncu
        // of the [element][line] form: the consumer spends ~40 % of its samples waiting to issue its LDS / STS behind the
        // loader's and the storer's traffic on the same MIO queue.
        float4* t4 = reinterpret_cast<float4*>(t);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int kk = FWD ? k : 7 - k;
            const float4 in = t4[kk];
            const float f[4] = {FWD ? in.x : in.w, FWD ? in.y : in.z, FWD ? in.z : in.y, FWD ? in.w : in.x};
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double v0 = (double)f[i];
                if (!FWD) v0 *= c.sum;
                v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
                o[i] = (float)v0;
                v3 = v2; v2 = v1; v1 = v0;
            }
            t4[kk] = FWD ? make_float4(o[0], o[1], o[2], o[3]) : make_float4(o[3], o[2], o[1], o[0]);
        }
        return;
    }
    if (PHASED) {
        double d[32];
        unsigned seen = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            d[e] = (double)t[(FWD ? e : 31 - e) * kIirPitch];
            if (!FWD) d[e] *= c.sum;
            seen |= (unsigned)__double2loint(d[e]);
        }
        v1 = __hiloint2double(__double2hiint(v1), __double2loint(v1) | (int)(seen & (unsigned)zero));
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            double v0 = d[e];
            v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
            d[e] = v0;
            v3 = v2; v2 = v1; v1 = v0;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) t[(FWD ? e : 31 - e) * kIirPitch] = (float)d[e];
        return;
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        float* p = t + (FWD ? e : 31 - e) * kIirPitch;
        double v0 = (double)(*p);
        if (!FWD) v0 *= c.sum;
        v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
        *p = (float)v0;
        v3 = v2; v2 = v1; v1 = v0;
    }
}

// kElemContig: true  = x pass (the 32 elements of a tile are contiguous in HBM; planes are contiguous, so line l
//                      starts at l * N)
//              false = y pass (the 32 lines of a tile are contiguous in HBM, elements are `elem_stride` apart; line l
//                      of plane p = l / lines_per_plane starts at p * plane_stride + l % lines_per_plane)
// Roles: consumer (recurrence), loader (HBM -> tile), storer (tile -> HBM).  A CTA has four warps, one per SM
// sub-partition; the role of a warp rotates with blockIdx.x so that the consumers of the CTAs resident on an SM
// spread over all four sub-partitions (with fixed roles every consumer lands on sub-partition 0 and they serialise
// on its FP64 pipe: measured 62 -> 153 cycles per sample going from 1 to 8 CTAs per SM).  The fourth warp exits.
// Barriers per stage: full (loader -> consumer), done (consumer -> storer), vacant (storer -> loader).
//
// Coef selects the recurrence: IirCoef = Van Vliet (fp64 state; the backward run filters the forward output in place,
// so pass 1 streams `dst` back in); DericheCoef = Deriche (fp32 state; causal and anticausal runs both filter the
// INPUT, so pass 1 streams `src` again TOGETHER with the causal output Y already in `dst`, and the consumer stores
// out = Y + yc, CImg.h:34797 -- src and dst must be distinct buffers).
template <bool kElemContig, class Coef, int NS, bool PHASE = false, bool ROWM = false>
__global__ void __launch_bounds__(128) iir_pipe_kernel(const float* src, float* dst, int N,   // src == dst is allowed (in-place y pass): no __restrict__
                                                      long nlines, int lines_per_plane, long plane_stride,
                                                      long elem_stride, Coef c, int zero) {
    constexpr bool kDeriche = std::is_same<Coef, DericheCoef>::value;
    constexpr int kIirNS = NS;
    constexpr bool kPhased = PHASE && !kDeriche;
    // Van Vliet x pass: tiles are [line][element] with a line pitch of 36 floats (see iir_tile32); every other form keeps
    // [element][line] with pitch 33.  kLS / kES: floats between consecutive lines / elements of a tile.
    constexpr bool kRowMajor = kElemContig && !kDeriche && ROWM;
    constexpr int kLS = kRowMajor ? 36 : 1, kES = kRowMajor ? 1 : kIirPitch;
    __shared__ __align__(16) float tiles[kIirNS][kRowMajor ? 32 * 36 : 32 * kIirPitch];
    // Deriche, pass 1: the causal output Y of the same tile, streamed back beside the input so that the consumer forms
    // out = Y + yc itself (a read-modify-write in the storer exposes one HBM latency per tile)
    __shared__ float ytiles[kDeriche ? kIirNS : 1][kDeriche ? 32 * kIirPitch : 1];
    __shared__ unsigned long long full[kIirNS], done[kIirNS], vacant[kIirNS];
    const int lane = threadIdx.x & 31;
    const int warp = ((threadIdx.x >> 5) + 4 - (blockIdx.x & 3)) & 3;   // role: 0 consumer, 1 loader, 2 storer, 3 idle
    const long line0 = (long)blockIdx.x * 32;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kIirNS; ++s) { mbar_init(&full[s], 32); mbar_init(&done[s], 1); mbar_init(&vacant[s], 1); }
    }
    __syncthreads();
    if (warp == 3) return;
    const int T = (N + 31) >> 5;   // tiles per line
    // job q of the CTA: q < T forward over tile q; q >= T backward over tile 2T-1-q.  Stage = q % NS; the k-th use
    // of a stage (k = q / NS) completes phase k of each of its barriers.
    if (warp == 0) {
        // ------------------------------------------ consumer ------------------------------------------
        if constexpr (kDeriche) {
            DericheFwd f{};
            DericheBwd b{};
            for (int q = 0; q < 2 * T; ++q) {
                const int s = q % kIirNS;
                mbar_wait(&full[s], (unsigned)((q / kIirNS) & 1));
                float* t = &tiles[s][lane];
                const bool fwd = q < T;
                const int tile = fwd ? q : 2 * T - 1 - q;
                const int ne = (N - tile * 32) < 32 ? (N - tile * 32) : 32;
                if (fwd) {
                    if (q == 0) f.init(t[0], c);
                    if (ne == 32) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) t[e * kIirPitch] = f.step(t[e * kIirPitch], c);
                    } else {
                        for (int e = 0; e < ne; ++e) t[e * kIirPitch] = f.step(t[e * kIirPitch], c);
                    }
                } else {
                    const float* ty = &ytiles[s][lane];
                    if (q == T) b.init(t[(ne - 1) * kIirPitch], c);
                    if (ne == 32) {   // out = Y + yc (CImg.h:34797)
#pragma unroll
                        for (int e = 31; e >= 0; --e) t[e * kIirPitch] = ty[e * kIirPitch] + b.step(t[e * kIirPitch], c);
                    } else {
                        for (int e = ne - 1; e >= 0; --e) t[e * kIirPitch] = ty[e * kIirPitch] + b.step(t[e * kIirPitch], c);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&done[s]);
            }
            return;
        } else {
        double v1 = 0, v2 = 0, v3 = 0, iplus = 0;
        for (int q = 0; q < 2 * T; ++q) {
            const int s = q % kIirNS;
            mbar_wait(&full[s], (unsigned)((q / kIirNS) & 1));
            float* t = &tiles[s][lane * kLS];
            const bool fwd = q < T;
            const int tile = fwd ? q : 2 * T - 1 - q;
            const int ne = (N - tile * 32) < 32 ? (N - tile * 32) : 32;
            if (fwd) {
                if (q == 0) v1 = v2 = v3 = (double)t[0] / c.sumsq;
                if (tile == T - 1) iplus = (double)t[(ne - 1) * kES];
                if (ne == 32) {
                    iir_tile32<true, kPhased, kRowMajor>(t, v1, v2, v3, c, zero);
                } else {
                    for (int e = 0; e < ne; ++e) {
                        double v0 = (double)t[e * kES];
                        v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
                        t[e * kES] = (float)v0;
                        v3 = v2; v2 = v1; v1 = v0;
                    }
                }
            } else {
                int e = ne - 1;
                if (q == T) {  // Triggs boundary on the last sample of the line
                    const double uplus = iplus / c.bnd, vplus = uplus / c.bnd;
                    const double unp = v1 - uplus, unp1 = v2 - uplus, unp2 = v3 - uplus;
                    const double v0 = (c.M[0] * unp + c.M[1] * unp1 + c.M[2] * unp2 + vplus) * c.sum;
                    const double n1 = (c.M[3] * unp + c.M[4] * unp1 + c.M[5] * unp2 + vplus) * c.sum;
                    const double n2 = (c.M[6] * unp + c.M[7] * unp1 + c.M[8] * unp2 + vplus) * c.sum;
                    t[e * kES] = (float)v0;
                    v3 = n2; v2 = n1; v1 = v0;
                    --e;
                }
                if (e == 31) {
                    iir_tile32<false, kPhased, kRowMajor>(t, v1, v2, v3, c, zero);
                } else {
                    for (; e >= 0; --e) {
                        double v0 = (double)t[e * kES];
                        v0 *= c.sum;
                        v0 += v1 * c.f1; v0 += v2 * c.f2; v0 += v3 * c.f3;
                        t[e * kES] = (float)v0;
                        v3 = v2; v2 = v1; v1 = v0;
                    }
                }
            }
            __syncwarp();                       // one arrival per warp: 32 arrivals on one mbarrier serialise
            if (lane == 0) mbar_arrive(&done[s]);
        }
        return;
        }
    }
    // ------------------------------------------ loader / storer ------------------------------------------
    // x pass: lane = element, loop over the CTA's lines; y pass: lane = line, loop over the tile's elements
    const int nl = (int)((nlines - line0) < 32 ? (nlines - line0) : 32);
    long mybase = 0;
    bool line_ok = true;
    if (kElemContig) {
        mybase = line0 * (long)N + lane;
    } else {
        const long l = line0 + lane;
        line_ok = l < nlines;
        const long p = (line_ok ? l : 0) / lines_per_plane;
        mybase = p * plane_stride + ((line_ok ? l : 0) - p * lines_per_plane);
    }
    // smem offset (floats) of this lane's first tile element and the stride between the elements it touches
    // x pass: lane = element, the loop runs over the lines; y pass: lane = line, the loop runs over the elements
    const int s_off = kElemContig ? lane * kES : lane * kLS;
    const int s_step = kElemContig ? kLS : kES;
    const long g_step = kElemContig ? (long)N : elem_stride;   // HBM stride between those elements
    if (warp == 1) {
        for (int pass = 0; pass < 2; ++pass) {
            const float* from = (pass == 0 || kDeriche) ? src : dst;
            if (pass == 1)   // the backward run reads the forward output: every forward tile must have reached HBM
                for (int j = (T - kIirNS > 0 ? T - kIirNS : 0); j < T; ++j)
                    mbar_wait_relaxed(&vacant[j % kIirNS], (unsigned)((j / kIirNS) & 1));
            for (int i = 0; i < T; ++i) {
                const int q = pass * T + i;
                const int s = q % kIirNS;
                if (q >= kIirNS) mbar_wait_relaxed(&vacant[s], (unsigned)(((q - kIirNS) / kIirNS) & 1));
                const int tile = pass == 0 ? i : T - 1 - i;
                const int e0 = tile * 32;
                float* sp = &tiles[s][s_off];
                const bool with_y = kDeriche && pass == 1;
                float* yp = &ytiles[kDeriche ? s : 0][kDeriche ? s_off : 0];
                if (kElemContig) {
                    if (e0 + lane < N) {
                        const float* g = from + mybase + e0;
                        const float* gy = dst + mybase + e0;
#pragma unroll 8
                        for (int r = 0; r < nl; ++r) { cp_async4(sp, g); sp += s_step; g += g_step; }
                        if (with_y) {
#pragma unroll 8
                            for (int r = 0; r < nl; ++r) { cp_async4(yp, gy); yp += s_step; gy += g_step; }
                        }
                    }
                } else if (line_ok) {
                    const int ne = (N - e0) < 32 ? (N - e0) : 32;
                    const float* g = from + mybase + (long)e0 * elem_stride;
                    const float* gy = dst + mybase + (long)e0 * elem_stride;
#pragma unroll 8
                    for (int e = 0; e < ne; ++e) { cp_async4(sp, g); sp += s_step; g += g_step; }
                    if (with_y) {
#pragma unroll 8
                        for (int e = 0; e < ne; ++e) { cp_async4(yp, gy); yp += s_step; gy += g_step; }
                    }
                }
                cp_async_arrive(&full[s]);   // one arrival per lane once ALL its copies above have landed
            }
        }
    } else {
        for (int pass = 0; pass < 2; ++pass)
            for (int i = 0; i < T; ++i) {
                const int q = pass * T + i;
                const int s = q % kIirNS;
                mbar_wait_relaxed(&done[s], (unsigned)((q / kIirNS) & 1));
                const int tile = pass == 0 ? i : T - 1 - i;
                const int e0 = tile * 32;
                const float* sp = &tiles[s][s_off];
                if (kElemContig) {
                    if (e0 + lane < N) {
                        float* g = dst + mybase + e0;
#pragma unroll 8
                        for (int r = 0; r < nl; ++r) { *g = *sp; sp += s_step; g += g_step; }

                    }
                } else if (line_ok) {
                    const int ne = (N - e0) < 32 ? (N - e0) : 32;
                    float* g = dst + mybase + (long)e0 * elem_stride;
#pragma unroll 8
                    for (int e = 0; e < ne; ++e) { *g = *sp; sp += s_step; g += g_step; }

                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&vacant[s]);   // release: the loader (same CTA) re-reads forward tiles after acquiring this
            }
    }
}

// x pass: row-major tiles, 128-bit shared accesses in the consumer; y pass: conversions ahead of the chain (iir_tile32).
// Measured on a 17997 x 2268 blend (11 levels): x 2.58 -> 2.38 ms, y 1.83 -> 1.71 ms; the other way round (phased x,
// either layout) is slower: 2.70 / 2.57 ms.
template <bool X>
static void launch_iir_pass(int grid, cudaStream_t st, const float* src, float* dst, int N, long nlines, int lpp, long plane,
                            long estride, const IirCoef& coef) {
    iir_pipe_kernel<X, IirCoef, 4, !X, X><<<grid, 128, 0, st>>>(src, dst, N, nlines, lpp, plane, estride, coef, 0);
}
void launch_iir_blur(const float* src, float* dst, int w, int h, int nplanes, const IirCoef& coef, cudaStream_t st) {
    const float* ysrc = src;
    const long plane = (long)w * h;
    if (w > 1) {
        const long nlines = (long)nplanes * h;
        KScope ks("blend.iir_x", st, 8.0 * nplanes * w * h);
        launch_iir_pass<true>(div_up(nlines, 32), st, src, dst, w, nlines, h, plane, 1L, coef);
        PB_KERNEL_CHECK();
        ysrc = dst;
    }
    if (h > 1) {
        const long nlines = (long)nplanes * w;
        KScope ks("blend.iir_y", st, 8.0 * nplanes * w * h);
        launch_iir_pass<false>(div_up(nlines, 32), st, ysrc, dst, h, nlines, w, plane, (long)w, coef);
        PB_KERNEL_CHECK();
        ysrc = dst;
    }
    if (ysrc != dst)
        PB_CUDA(cudaMemcpyAsync(dst, src, (size_t)nplanes * w * h * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

void launch_deriche_blur(const float* src, float* tmp, float* dst, int w, int h, int nplanes, const DericheCoef& coef,
                         cudaStream_t st) {
    const long plane = (long)w * h;
    const float* ysrc = src;
    if (w > 1) {
        const long nlines = (long)nplanes * h;
        float* xdst = h > 1 ? tmp : dst;
        KScope ks("blend.deriche", st, 16.0 * nplanes * w * h);
        iir_pipe_kernel<true, DericheCoef, kIirNSDeriche><<<div_up(nlines, 32), 128, 0, st>>>(src, xdst, w, nlines, h, plane, 1L, coef, 0);
        PB_KERNEL_CHECK();
        ysrc = xdst;
    }
    if (h > 1) {
        const long nlines = (long)nplanes * w;
        KScope ks("blend.deriche", st, 16.0 * nplanes * w * h);
        iir_pipe_kernel<false, DericheCoef, kIirNSDeriche><<<div_up(nlines, 32), 128, 0, st>>>(ysrc, dst, h, nlines, w, plane, (long)w, coef, 0);
        PB_KERNEL_CHECK();
        ysrc = dst;
    }
    if (ysrc != dst)
        PB_CUDA(cudaMemcpyAsync(dst, src, (size_t)nplanes * w * h * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

// ---------------------------------------------------------------------------------------------------------
// 2:1 reduce (moving average, x pass rounded to float then y pass) and linear expand
// ---------------------------------------------------------------------------------------------------------
// One thread = one output pixel of ALL planes: the moving-average taps (<= 3 x 3 for a 2:1 reduce; any count in
// general) and their weights are fetched once and reused by every plane.
// (A form with the plane count at compile time and all 9 x 7 tap loads of a thread issued before the first use was measured
// on a 17997 x 2268 blend: 1.24 ms against 0.99 ms for this loop over 11 launches -- its 109 registers halve the resident
// threads.)
__global__ void __launch_bounds__(256, 8) reduce_kernel(const float* __restrict__ src, int w, int h, int nplanes,
                                                    float* __restrict__ dst, int nw, int nh, DevMovAvg tx, DevMovAvg ty) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nw || y >= nh) return;
    const int xs0 = tx.start[x], xs1 = tx.start[x + 1], ys0 = ty.start[y], ys1 = ty.start[y + 1];
    const size_t plane = (size_t)w * h, nplane = (size_t)nw * nh;
    if (xs1 - xs0 <= 3 && ys1 - ys0 <= 3) {
        int sx[3], sy[3];
        float wx[3], wy[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const bool okx = xs0 + i < xs1, oky = ys0 + i < ys1;
            sx[i] = okx ? tx.src[xs0 + i] : 0; wx[i] = okx ? tx.wgt[xs0 + i] : 0.0f;
            sy[i] = oky ? ty.src[ys0 + i] : 0; wy[i] = oky ? ty.wgt[ys0 + i] : 0.0f;
        }
        const int nx = xs1 - xs0, ny = ys1 - ys0;
        for (int p = 0; p < nplanes; ++p) {
            const float* s = src + p * plane;
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (j < ny) {
                    const float* row = s + (size_t)sy[j] * w;
                    float r = 0.0f;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (i < nx) r += row[sx[i]] * wx[i];
                    r = r / tx.div;
                    acc += r * wy[j];
                }
            dst[p * nplane + (size_t)y * nw + x] = acc / ty.div;
        }
    } else {
        for (int p = 0; p < nplanes; ++p) {
            const float* s = src + p * plane;
            float acc = 0.0f;
            for (int j = ys0; j < ys1; ++j) {
                const float rowavg = movavg_sample(s + (size_t)ty.src[j] * w, 1, tx.start, tx.src, tx.wgt, x, tx.div);
                acc += rowavg * ty.wgt[j];
            }
            dst[p * nplane + (size_t)y * nw + x] = acc / ty.div;
        }
    }
}
void launch_reduce(const float* src, int w, int h, int nplanes, float* dst, int nw, int nh, DevMovAvg tx, DevMovAvg ty,
                   cudaStream_t st) {
    if (nw <= 0 || nh <= 0) return;
    KScope ks("blend.reduce", st, 4.0 * nplanes * ((double)w * h + (double)nw * nh));
    dim3 b(64, 4), g(div_up(nw, 64), div_up(nh, 4));
    reduce_kernel<<<g, b, 0, st>>>(src, w, h, nplanes, dst, nw, nh, tx, ty);
    PB_KERNEL_CHECK();
}

__device__ __forceinline__ float upsample_at(const float* __restrict__ src, int uw, int uh, const DevLinear& tx,
                                             const DevLinear& ty, int x, int y) {
    const int px = tx.pos[x], py = ty.pos[y];
    const double ax = tx.alpha[x], ay = ty.alpha[y];
    const float* r0 = src + (size_t)py * uw;
    const float v0 = linear_sample(r0, 1, uw, px, ax);
    const float v1 = (py < uh - 1) ? linear_sample(r0 + uw, 1, uw, px, ax) : v0;
    return (float)((1 - ay) * (double)v0 + ay * (double)v1);
}

__global__ void expand_kernel(const float* __restrict__ src, int w, int h, float* __restrict__ dst, int nw, int nh,
                              DevLinear tx, DevLinear ty) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int p = blockIdx.z;
    if (x >= nw || y >= nh) return;
    dst[((size_t)p * nh + y) * nw + x] = upsample_at(src + (size_t)p * w * h, w, h, tx, ty, x, y);
}
void launch_expand(const float* src, int w, int h, int nplanes, float* dst, int nw, int nh, DevLinear tx, DevLinear ty,
                   cudaStream_t st) {
    KScope ks("blend.expand", st, 4.0 * nplanes * ((double)w * h + (double)nw * nh));
    dim3 b(64, 4), g(div_up(nw, 64), div_up(nh, 4), nplanes);
    expand_kernel<<<g, b, 0, st>>>(src, w, h, dst, nw, nh, tx, ty);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// Laplacian blend + collapse of one level
// ---------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void collapse_kernel(const float* __restrict__ G, int w, int h, const float* __restrict__ Gup,
                                const float* __restrict__ Eup, int uw, int uh, DevLinear tx, DevLinear ty,
                                float* __restrict__ E, u8* __restrict__ out8) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t n = (size_t)w * h, o = (size_t)y * w + x, un = (size_t)uw * uh;
    const float m = G[2 * NCH * n + o];
    // the resampling position of this pixel is shared by the 3 NCH up-sampled planes
    int px = 0, py = 0, px1 = 0;
    double ax = 0, ay = 0;
    bool has_y1 = false;
    if (Gup) {
        px = tx.pos[x]; py = ty.pos[y]; ax = tx.alpha[x]; ay = ty.alpha[y];
        px1 = px < uw - 1 ? px + 1 : px;
        has_y1 = py < uh - 1;
    }
    auto up = [&](const float* __restrict__ plane) {   // == upsample_at(plane, uw, uh, tx, ty, x, y)
        const float* r0 = plane + (size_t)py * uw;
        const float a0 = r0[px], a1 = r0[px1];
        const float v0 = (float)((1 - ax) * (double)a0 + ax * (double)a1);
        float v1 = v0;
        if (has_y1) {
            const float b0 = r0[uw + px], b1 = r0[uw + px1];
            v1 = (float)((1 - ax) * (double)b0 + ax * (double)b1);
        }
        return (float)((1 - ay) * (double)v0 + ay * (double)v1);
    };
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float la = G[c * n + o], lb = G[(NCH + c) * n + o];
        float e;
        if (Gup) {
            la = la - up(Gup + c * un);
            lb = lb - up(Gup + (NCH + c) * un);
            const float bl = blend_px(la, lb, m);
            e = collapse_px(bl, up(Eup + c * un));
        } else {
            e = blend_px(la, lb, m);
        }
        if (out8) out8[c * n + o] = (u8)e;
        else E[c * n + o] = e;
    }
}
// Tiled form of the same level step.  CImg's linear resize interpolates along x, rounds to float, then interpolates along
// y (CImg.h:29641-29652); in collapse_kernel every output pixel redoes the two x-interpolations of its two source rows for
// all nine up-sampled planes, although at 2:1 each x-interpolated value is shared by four output rows.  Here a CTA
// (64 x 16 output pixels) first fills shared memory with the x-interpolated source rows it needs (nine planes, at most
// kCollapseSrcRows rows), then every pixel only does the y-interpolation and the blend.  Same operations on the same
// operands; ~15 instead of 27 double-precision interpolations per pixel.
// Occupancy: both this kernel and the reduce wait on loads (ncu: long_scoreboard 7.7 / 10.8 per issue, DRAM at 22-25 %), so
// they are capped at 48 / 32 registers (5 / 8 CTAs of 256 threads per SM, a few spilled words): collapse 1.31 -> 1.00 ms,
// reduce 0.99 -> 0.84 ms over the levels of a 17997 x 2268 blend (6 CTAs for the collapse: 1.10 ms).
constexpr int kCollapseTileH = 16, kCollapseSrcRows = 12;
template <int NCH>
__global__ void __launch_bounds__(256, 5) collapse_tiled_kernel(const float* __restrict__ G, int w, int h,
                                                             const float* __restrict__ Gup, const float* __restrict__ Eup,
                                                             int uw, int uh, DevLinear tx, DevLinear ty,
                                                             float* __restrict__ E, u8* __restrict__ out8) {
    __shared__ float X[3 * NCH][kCollapseSrcRows][64];
    const int x = blockIdx.x * 64 + threadIdx.x;
    const int yt0 = blockIdx.y * kCollapseTileH;
    const int yt1 = (yt0 + kCollapseTileH < h ? yt0 + kCollapseTileH : h) - 1;   // last output row of the tile
    const int r0 = ty.pos[yt0];
    int r1 = ty.pos[yt1] + 1;
    if (r1 > uh - 1) r1 = uh - 1;
    const int nrows = r1 - r0 + 1;   // <= kCollapseSrcRows, guaranteed by the launcher
    if (nrows > kCollapseSrcRows) __trap();
    const size_t n = (size_t)w * h, un = (size_t)uw * uh;
    const bool xin = x < w;
    // this thread's level-i values (7 planes x 4 rows) are requested first: they do not depend on the x-interpolation
    // below, which then hides their latency
    float gv[kCollapseTileH / 4][2 * NCH + 1];
    if (xin) {
#pragma unroll
        for (int j = 0; j < kCollapseTileH / 4; ++j) {
            const int y = yt0 + threadIdx.y + 4 * j;
            const size_t o = (size_t)(y <= yt1 ? y : yt1) * w + x;
#pragma unroll
            for (int p = 0; p < 2 * NCH + 1; ++p) gv[j][p] = G[p * n + o];
        }
    }
    if (xin) {
        const int px = tx.pos[x], px1 = px < uw - 1 ? px + 1 : px;
        const double ax = tx.alpha[x];
        for (int r = threadIdx.y; r < nrows; r += 4) {
            const size_t ro = (size_t)(r0 + r) * uw;
#pragma unroll
            for (int p = 0; p < 3 * NCH; ++p) {
                const float* plane = p < 2 * NCH ? Gup + (size_t)p * un : Eup + (size_t)(p - 2 * NCH) * un;
                const float a0 = plane[ro + px], a1 = plane[ro + px1];
                X[p][r][threadIdx.x] = (float)((1 - ax) * (double)a0 + ax * (double)a1);
            }
        }
    }
    __syncthreads();
    if (!xin) return;
#pragma unroll
    for (int j = 0; j < kCollapseTileH / 4; ++j) {
        const int y = yt0 + threadIdx.y + 4 * j;
        if (y > yt1) break;
        const int py = ty.pos[y], pr = py - r0;
        const bool has_y1 = py < uh - 1;
        const double ay = ty.alpha[y];
        const size_t o = (size_t)y * w + x;
        const float m = gv[j][2 * NCH];
        auto up = [&](int p) {
            const float v0 = X[p][pr][threadIdx.x];
            const float v1 = has_y1 ? X[p][pr + 1][threadIdx.x] : v0;
            return (float)((1 - ay) * (double)v0 + ay * (double)v1);
        };
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float la = gv[j][c] - up(c), lb = gv[j][NCH + c] - up(NCH + c);
            const float e = collapse_px(blend_px(la, lb, m), up(2 * NCH + c));
            if (out8) out8[c * n + o] = (u8)e;
            else E[c * n + o] = e;
        }
    }
}
void launch_collapse(const float* G_i, int w, int h, const float* G_up, const float* E_up, int uw, int uh, DevLinear tx,
                     DevLinear ty, float* E_out, u8* out_u8, cudaStream_t st, int nch) {
    KScope ks("blend.collapse", st, ((8.0 * nch + 4.0) + (out_u8 ? 1.0 : 4.0) * nch) * w * h + (8.0 * nch + 4.0 + 4.0 * nch) * uw * uh);
    // the tiled form needs the source rows of a 16-row tile to fit its shared-memory window: pos[y] advances by
    // fx = (uh - 1) / (h - 1) per output row (hostnum::linear_table), so a tile spans at most ceil(15 fx) + 3 source rows --
    // 11 for the pyramid's 2:1 steps; anything coarser takes the per-pixel kernel
    const double fx = h > 1 ? (uh - 1.0) / (h - 1.0) : 1e9;
    if (G_up && std::ceil((kCollapseTileH - 1) * fx) + 3 <= kCollapseSrcRows) {
        dim3 b(64, 4), g(div_up(w, 64), div_up(h, kCollapseTileH));
        if (nch == 1) collapse_tiled_kernel<1><<<g, b, 0, st>>>(G_i, w, h, G_up, E_up, uw, uh, tx, ty, E_out, out_u8);
        else collapse_tiled_kernel<3><<<g, b, 0, st>>>(G_i, w, h, G_up, E_up, uw, uh, tx, ty, E_out, out_u8);
    } else {
        dim3 b(64, 4), g(div_up(w, 64), div_up(h, 4));
        if (nch == 1) collapse_kernel<1><<<g, b, 0, st>>>(G_i, w, h, G_up, E_up, uw, uh, tx, ty, E_out, out_u8);
        else collapse_kernel<3><<<g, b, 0, st>>>(G_i, w, h, G_up, E_up, uw, uh, tx, ty, E_out, out_u8);
    }
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// equalisation tail
// ---------------------------------------------------------------------------------------------------------
__global__ void luma_hist_kernel(const u8* __restrict__ rgb, size_t n, int* __restrict__ hist) {
    __shared__ int sh[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[luma_bin(rgb[i], rgb[n + i], rgb[2 * n + i])], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
void launch_luma_hist(const u8* rgb, int w, int h, int* hist256, cudaStream_t st) {
    size_t n = (size_t)w * h;
    PB_CUDA(cudaMemsetAsync(hist256, 0, 256 * sizeof(int), st));
    KScope ks("tail.luma_hist", st, 3.0 * n);
    int blocks = div_up((long)n, 256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    luma_hist_kernel<<<blocks, 256, 0, st>>>(rgb, n, hist256);
    PB_KERNEL_CHECK();
}

__global__ void equalize_mix_kernel(const u8* __restrict__ rgb, size_t n, const int* __restrict__ lut,
                                    u8* __restrict__ out, double num, double den) {
    __shared__ int sl[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sl[i] = lut[i];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u8 r, g, b;
        equalize_mix_px(rgb[i], rgb[n + i], rgb[2 * n + i], sl, &r, &g, &b, num, den);
        out[i] = r; out[n + i] = g; out[2 * n + i] = b;
    }
}
void launch_equalize_mix(const u8* rgb, int w, int h, const int* lut256, u8* out, double num, double den,
                         cudaStream_t st) {
    size_t n = (size_t)w * h;
    KScope ks("tail.equalize_mix", st, 6.0 * n);
    int blocks = div_up((long)n, 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    equalize_mix_kernel<<<blocks, 256, 0, st>>>(rgb, n, lut256, out, num, den);
    PB_KERNEL_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// Reinhard colour transfer (transfer.cpp:4-13, 125-225)
// ---------------------------------------------------------------------------------------------------------
__global__ void lab_forward_kernel(const u8* __restrict__ rgb, size_t n, float* __restrict__ lab) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float L, a, b;
    rgb_to_lab((float)rgb[i], (float)rgb[n + i], (float)rgb[2 * n + i], &L, &a, &b);
    lab[i] = L; lab[n + i] = a; lab[2 * n + i] = b;
}
// The reference sums every plane in raster order into a FLOAT accumulator (transfer.cpp:129-160): the value depends on
// the order, so the sum is kept serial -- one warp per plane: the lanes fetch 32 consecutive samples with one coalesced
// load, then every lane walks them in order through shuffles (all lanes hold the same accumulator).  mode 0: sum of v;
// mode 1: sum of (v - mean)^2 with mean = stats[plane] (separately rounded difference, product and sum).
__global__ void __launch_bounds__(32) lab_serial_sum_kernel(const float* __restrict__ lab, size_t n, int mode,
                                                            const float* __restrict__ mean, float* __restrict__ out) {
    const float* p = lab + (size_t)blockIdx.x * n;
    const int lane = threadIdx.x;
    const float mu = mode ? mean[blockIdx.x] : 0.0f;
    float acc = 0.0f;
    for (size_t base = 0; base < n; base += 32) {
        const size_t i = base + lane;
        float v = i < n ? p[i] : 0.0f;
        if (mode) { const float d = v - mu; v = d * d; }
        const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
        if (cnt == 32) {
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += __shfl_sync(0xffffffffu, v, k);
        } else {
            for (int k = 0; k < cnt; ++k) acc += __shfl_sync(0xffffffffu, v, k);
        }
    }
    if (lane == 0) out[blockIdx.x] = acc;
}
// stats layout (floats): [0..2] sum src, [3..5] sum tem, [6..8] mean src, [9..11] mean tem, [12..14] sumsq src,
// [15..17] sumsq tem, [18..20] sd src, [21..23] sd tem
__global__ void lab_finish_stats_kernel(float* __restrict__ st, float n_src, float n_tem, int stage) {
    const int c = threadIdx.x;
    if (c >= 3) return;
    if (stage == 0) { st[6 + c] = st[c] / n_src; st[9 + c] = st[3 + c] / n_tem; }
    else { st[18 + c] = sqrtf(st[12 + c] / n_src); st[21 + c] = sqrtf(st[15 + c] / n_tem); }
}
__global__ void lab_inverse_kernel(const float* __restrict__ lab, size_t n, const float* __restrict__ st, u8* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = lab_match(lab[c * n + i], st[6 + c], st[18 + c], st[9 + c], st[21 + c]);
    float R, G, B;
    lab_to_rgb(v[0], v[1], v[2], &R, &G, &B);
    out[i] = (u8)R; out[n + i] = (u8)G; out[2 * n + i] = (u8)B;
}
void launch_color_transfer(const u8* src, int w, int h, const u8* tem, int tw, int th, float* lab_src, float* lab_tem,
                           float* stats24, u8* out, cudaStream_t st) {
    const size_t n = (size_t)w * h, nt = (size_t)tw * th;
    {
        KScope ks("transfer.lab", st, 15.0 * (n + nt));
        lab_forward_kernel<<<div_up((long)n, 256), 256, 0, st>>>(src, n, lab_src);
        lab_forward_kernel<<<div_up((long)nt, 256), 256, 0, st>>>(tem, nt, lab_tem);
        PB_KERNEL_CHECK();
    }
    {
        KScope ks("transfer.stats", st, 8.0 * 3 * (n + nt));
        lab_serial_sum_kernel<<<3, 32, 0, st>>>(lab_src, n, 0, nullptr, stats24);
        lab_serial_sum_kernel<<<3, 32, 0, st>>>(lab_tem, nt, 0, nullptr, stats24 + 3);
        lab_finish_stats_kernel<<<1, 32, 0, st>>>(stats24, (float)(w * h), (float)(tw * th), 0);
        lab_serial_sum_kernel<<<3, 32, 0, st>>>(lab_src, n, 1, stats24 + 6, stats24 + 12);
        lab_serial_sum_kernel<<<3, 32, 0, st>>>(lab_tem, nt, 1, stats24 + 9, stats24 + 15);
        lab_finish_stats_kernel<<<1, 32, 0, st>>>(stats24, (float)(w * h), (float)(tw * th), 1);
        PB_KERNEL_CHECK();
    }
    KScope ks("transfer.map", st, 15.0 * n);
    lab_inverse_kernel<<<div_up((long)n, 256), 256, 0, st>>>(lab_src, n, stats24, out);
    PB_KERNEL_CHECK();
}

}  // namespace pb
