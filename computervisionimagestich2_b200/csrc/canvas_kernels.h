// canvas_kernels.h -- launch interface of the image-space kernels (canvas_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "canvas_device.cuh"

namespace pb {

// Cylindrical projection fused with grayscale + u8->f32 (Projection.cpp:20-73, ImageProcess.cpp:27-51).
// src/dst_rgb: planar u8 [3][h][w]; gray_f32: pitched float image for the SIFT engine (may be null);
// gray_u8 (optional, [h][w]).  ktab: device copy of hostnum::cylinder_table (length min(w, h)).
void launch_project_gray(const u8* src, int w, int h, const float* ktab, u8* dst_rgb, float* gray_f32, int gray_pitch,
                         u8* gray_u8, cudaStream_t st);
// 24-bpp BMP pixel area (BGR rows padded to `stride`, bottom-up unless the header says otherwise) <-> planar RGB
void launch_bmp_to_planar(const u8* bmp, int stride, int w, int h, bool bottom_up, u8* dst, cudaStream_t st);
void launch_planar_to_bmp(const u8* src, int w, int h, int stride, u8* bmp, cudaStream_t st);
void launch_gray(const u8* rgb, int w, int h, float* gray_f32, int gray_pitch, u8* gray_u8, cudaStream_t st);

// a = warp(src image, H8, off) and b = shift(previous canvas, ioff) in one pass over the new canvas
// (ImageProcess.cpp:596-620).  H8: device pointer to 8 doubles.  Either output may be null.
// nch = 3: planar RGB throughout; nch = 1: one colour plane (src, prev, a, b are single planes) -- the colour planes of
// the canvas stages are independent, which is how a sharded job splits them over ranks.
void launch_warp_shift(const u8* src, int sw, int sh, const double* H8, float offx, float offy, const u8* prev, int pw,
                       int ph, int ioffx, int ioffy, u8* a, u8* b, int cw, int ch, cudaStream_t st, int nch = 3);

// Seam statistics of the middle row (ImageProcess.cpp:659-671): stats[4] = {sum_a_x, width_mid_a, sum_overlap_x,
// width_mid_overlap} (int32 wrap-around like the reference's int sums).
// all_channels: count a pixel only when all three channels are non-zero (src/ex6/ImageProcess.cpp:651-660).
void launch_seam_stats(const u8* a, const u8* b, int cw, int ch, int* stats, bool all_channels, cudaStream_t st);
// Level-0 float planes: G0[0..2] = a, G0[3..5] = b, G0[6] = seam mask (ImageProcess.cpp:678-698). err_flag set to 1
// when the middle row is empty (the reference would loop forever / divide by zero).
// double_seam: seam position kept in double (src/ex6/ImageProcess.cpp:678-697) instead of float.
// nch colour planes: G0[0..nch) = a, G0[nch..2 nch) = b, G0[2 nch] = mask.
void launch_level0(const u8* a, const u8* b, int cw, int ch, const int* stats, float* G0, int* err_flag,
                   bool double_seam, cudaStream_t st, int nch = 3);

// CImg get_blur(2, true, true) on `nplanes` planes [nplanes][h][w]: dst = blur(src) (x pass src->dst, y pass in place
// on dst; src may equal dst).
void launch_iir_blur(const float* src, float* dst, int w, int h, int nplanes, const IirCoef& coef, cudaStream_t st);
// CImg get_blur(2) (is_gaussian = false: Deriche order 0, CImg.h:34777-34869, 35111-35147; the src/ex6 pyramid):
// dst = blur(src), tmp = scratch of the same size; the three buffers must be distinct.
void launch_deriche_blur(const float* src, float* tmp, float* dst, int w, int h, int nplanes, const DericheCoef& coef,
                         cudaStream_t st);

// Moving-average 2:1 reduce (CImg resize type 2 via type 3): src [n][h][w] -> dst [n][nh][nw]; tables on device.
struct DevMovAvg { const int* start; const int* src; const float* wgt; float div; };  // div = source length (1 for identity)
void launch_reduce(const float* src, int w, int h, int nplanes, float* dst, int nw, int nh, DevMovAvg tx, DevMovAvg ty,
                   cudaStream_t st);
// Linear up-resize (CImg resize type 3): src [n][h][w] -> dst [n][nh][nw]
struct DevLinear { const int* pos; const double* alpha; };
void launch_expand(const float* src, int w, int h, int nplanes, float* dst, int nw, int nh, DevLinear tx, DevLinear ty,
                   cudaStream_t st);

// One level of Laplacian blend + collapse (ImageProcess.cpp:727-771):
//   E_i = clamp( blend(Ga_i - up(Ga_{i+1}), Gb_i - up(Gb_{i+1}), M_i) + up(E_{i+1}) )
// G_i: 7 planes [7][h][w] (a rgb, b rgb, mask); G_up: 7 planes of level i+1 (uw x uh) or null for the top level;
// E_up: 3 planes of level i+1 or null; E_out: 3 float planes, or out_u8 (level 0: truncating store, planar u8).
// With nch colour planes the level has 2 nch + 1 planes and E has nch.
void launch_collapse(const float* G_i, int w, int h, const float* G_up, const float* E_up, int uw, int uh, DevLinear tx,
                     DevLinear ty, float* E_out, u8* out_u8, cudaStream_t st, int nch = 3);

// Equalisation tail (equalization.cpp:74-131 + ImageProcess.cpp:237-268)
void launch_luma_hist(const u8* rgb, int w, int h, int* hist256, cudaStream_t st);
// luminance mix Y * num / den + Y_eq / den (19/20: ImageProcess.cpp:261; 5/6: src/ex6/ImageProcess.cpp:270)
void launch_equalize_mix(const u8* rgb, int w, int h, const int* lut256, u8* out, double num, double den,
                         cudaStream_t st);

// Reinhard l-alpha-beta colour transfer (the reference's class transfer, transfer.cpp): out = src recoloured to the
// per-channel mean / deviation of tem in l-alpha-beta space.  lab_src / lab_tem: scratch of 3 * w * h / 3 * tw * th
// floats; stats24: 24 floats of scratch.  The plane sums keep the reference's serial float order (one warp per plane).
void launch_color_transfer(const u8* src, int w, int h, const u8* tem, int tw, int th, float* lab_src, float* lab_tem,
                           float* stats24, u8* out, cudaStream_t st);

}  // namespace pb
