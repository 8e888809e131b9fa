/* pano_b200.h -- C ABI of libpano_b200.so, the B200 (sm_100a) implementation of the panorama-stitching hot path of
 * chensh236/ComputerVisionImageStich2.  Plain pointers and sizes only; every buffer named here is HOST memory unless
 * its name starts with d_.  All functions return 0 on success and a negative code on failure;
 * pano_b200_last_error() gives the message.  A context is bound to one CUDA device and one stream and must not be
 * used from two threads at once (distinct contexts may run concurrently).
 *
 * Image layout everywhere: CImg planar, data[x + y*W + c*W*H], c = 0,1,2 = R,G,B (CImg.h:48533-48546).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference tree).
 */
#ifndef PANO_B200_H
#define PANO_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pano_b200_ctx pano_b200_ctx;

/* VlSiftKeypoint (vl/sift.h:19-31), 32 bytes */
typedef struct pano_b200_keypoint {
    int o, ix, iy, is;
    float x, y, s, sigma;
} pano_b200_keypoint;

/* ImgPair (ImageProcess.h:43-47) */
typedef struct pano_b200_pair {
    pano_b200_keypoint src, dst;
} pano_b200_pair;

/* per-stage wall times of the last pano_b200_stitch call, milliseconds */
typedef struct pano_b200_times {
    double project, sift, table, match, ransac, warp, blend, tail, total;
    int64_t match_pairs_evaluated, sift_pixels;
    int n_match_calls, n_blends;
} pano_b200_times;

int pano_b200_create(int device, pano_b200_ctx** out);
void pano_b200_destroy(pano_b200_ctx* ctx);
const char* pano_b200_last_error(pano_b200_ctx* ctx);
int pano_b200_device_count(void);
void pano_b200_free(void* p); /* frees any buffer this library returned through an out-pointer */

/* ---- which caller of the hot path is reproduced.  The reference ships the same pipeline twice: the root
 *      ImageProcess.cpp (PANO_B200_PROFILE_ROOT, the default) and src/ex6/ImageProcess.cpp (PANO_B200_PROFILE_EX6:
 *      images in left-to-right order stitched as a fixed chain from image n/2, src/ex6/ImageProcess.cpp:147-160;
 *      canvas bounds from fewer corners, :230-243; 3-channel seam test and min(w,h) level count, :638-665; Deriche
 *      pyramid blur get_blur(2), :702-705; 5/6 : 1/6 luminance mix, :270).  ransac_seed is the value passed to srand
 *      at the top of ImageProcess::RANSAC: 666666 in the root variant (ImageProcess.cpp:397), time(0) in src/ex6
 *      (src/ex6/ImageProcess.cpp:403) -- the caller supplies it.  Applies to the pipeline entry points and to
 *      pano_b200_ransac / pano_b200_blend / pano_b200_equalize_mix / pano_b200_plan_canvas_ex. */
#define PANO_B200_PROFILE_ROOT 0
#define PANO_B200_PROFILE_EX6 1
int pano_b200_set_profile(pano_b200_ctx* ctx, int variant, unsigned ransac_seed);

/* ---- whole pipeline: `ImageProcess ip(dir, n)` + private member `result` (main.cpp:9, ImageProcess.cpp:3-8,
 *      ImageProcess.h:145).  imgs[i]: planar RGB of size w[i] x h[i].  *out is library-allocated (pano_b200_free). */
int pano_b200_stitch(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, int n, uint8_t** out,
                     int* out_w, int* out_h);
/* same, into a caller-provided (e.g. pinned) buffer of out_cap bytes */
int pano_b200_stitch_into(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, int n,
                          uint8_t* out, size_t out_cap, int* out_w, int* out_h);
/* BMP files in, BMP file out (the container step either side of the path: CImg<uchar>(file) at ImageProcess.cpp:16,
 * CImg.h:48395-48546, and CImg::save): files[i] / sizes[i] = the bytes of <dir><i+1>.bmp (uncompressed 24-bpp); the
 * BGR bottom-up padded rows are decoded to planar RGB, and the panorama encoded back, by GPU kernels.  *out_bmp is
 * library-allocated (pano_b200_free). */
int pano_b200_stitch_bmp(pano_b200_ctx* ctx, const uint8_t* const* files, const size_t* sizes, int n, uint8_t** out_bmp,
                         size_t* out_size);

/* ---- sharded jobs (one process per GPU; the exchange between ranks is the caller's, see
 *      computervisionimagestich2_b200/dist.py).  pano_b200_extract = the body of readFile for ONE image
 *      (ImageProcess.cpp:12-23): projected image (proj_out, w*h*3, may be NULL) + feature table, library-allocated.
 *      pano_b200_stitch_features = ImageProcess::matching (ImageProcess.cpp:101-271) on images whose projections and
 *      feature tables were computed elsewhere; match_idx (optional) is an nimg*nimg array of per-directed-pair match
 *      lists (entry [i*nimg+j] = getImgPair(imgs[i], imgs[j]) as nfeat[j] indices into table i, or -1; NULL entries
 *      are evaluated here). */
int pano_b200_extract(pano_b200_ctx* ctx, const uint8_t* rgb, int w, int h, uint8_t* proj_out, float** descr,
                      pano_b200_keypoint** keys, int* n);
int pano_b200_stitch_features(pano_b200_ctx* ctx, int nimg, const uint8_t* const* proj, const int* w, const int* h,
                              const float* const* descr, const pano_b200_keypoint* const* keys, const int* nfeat,
                              const int* const* match_idx, uint8_t** out, int* out_w, int* out_h);
/* Batched independent image pairs (BASELINE.json configs[4]).  imgs[2p], imgs[2p+1] = planar RGB host buffers of pair p.
 * Per pair: the body of readFile for both images (ImageProcess.cpp:12-23; all 2*npairs images of the call run
 * concurrently on the context's lanes), getImgPair in both directions (ImageProcess.cpp:273-351; all 2*npairs directed
 * problems in ONE launch) and RANSAC (ImageProcess.cpp:395-436) on every direction that reaches the adjacency threshold
 * of 20 matches (ImageProcess.cpp:128; all of them in ONE launch).  Direction 0 = getImgPair(a, b): for each feature of b
 * its match in a, coefficients map a -> b; direction 1 the reverse.  out[p].pair is left untouched. */
typedef struct pano_b200_pair_record {
    int64_t pair;
    int32_t nfeat[2], nmatch[2], has_h[2];
    double H[2][8];
} pano_b200_pair_record;
int pano_b200_pairs(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, int npairs,
                    pano_b200_pair_record* out);
/* the same with the 2 * npairs input images already resident in HBM (d_imgs[k] = device pointers, planar RGB) */
int pano_b200_pairs_staged(pano_b200_ctx* ctx, const uint8_t* const* d_imgs, const int* w, const int* h, int npairs,
                           pano_b200_pair_record* out);
/* Inputs staged in HBM once, then stitched any number of times with no host<->device pixel traffic
 * (device-resident throughput measurement); pano_b200_result_copy downloads the last result. */
int pano_b200_stage_images(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, int n);
int pano_b200_stitch_staged(pano_b200_ctx* ctx, int* out_w, int* out_h);
int pano_b200_result_copy(pano_b200_ctx* ctx, uint8_t* out);
/* the lines the reference prints to stdout: middle index, then "src dst" per stitched edge (ImageProcess.cpp:183,391) */
int pano_b200_stitch_log(pano_b200_ctx* ctx, char* dst, int cap);
int pano_b200_stitch_times(pano_b200_ctx* ctx, pano_b200_times* t);
int pano_b200_stitch_nfeatures(pano_b200_ctx* ctx, int image); /* size of imgs[i].features after the run, -1 if none */

/* ---- stages ---------------------------------------------------------------------------------------------------- */
/* Projection::imageProjection (Projection.cpp:20-73) [+ ImageProcess::toGrayScale (ImageProcess.cpp:27-40) when
 * out_gray != NULL].  Either output may be NULL. */
int pano_b200_project(pano_b200_ctx* ctx, const uint8_t* rgb, int w, int h, uint8_t* out_rgb, uint8_t* out_gray);
/* ImageProcess::toGrayScale (ImageProcess.cpp:27-40) */
int pano_b200_gray(pano_b200_ctx* ctx, const uint8_t* rgb, int w, int h, uint8_t* out_gray);

/* ImageProcess::siftAlgorithm (ImageProcess.cpp:44-99): gray u8 image -> feature table in std::map order (sorted by
 * descriptor, duplicates dropped).  *descr [n][128] floats and *keys [n] are library-allocated. */
int pano_b200_sift_features(pano_b200_ctx* ctx, const uint8_t* gray, int w, int h, float** descr,
                            pano_b200_keypoint** keys, int* n);

/* The raw vl_sift_* call sequence of siftAlgorithm (ImageProcess.cpp:55-92) on a float image with explicit
 * (noctaves, nlevels): every (keypoint, angle) descriptor in call order, nothing sorted or de-duplicated.
 * Out arrays are library-allocated: keys [n] (ix/iy as detected), angles [n], descr [n][128]; octave_nkeys[noctaves]
 * (caller-provided, may be NULL) receives the refined keypoint count per octave. */
int pano_b200_sift_raw(pano_b200_ctx* ctx, const float* image, int w, int h, int noctaves, int nlevels,
                       pano_b200_keypoint** keys, double** angles, float** descr, int* n, int* octave_nkeys);

/* Scale-space dump of one octave for parity tests: after pano_b200_sift_raw, copies the octave's GSS levels
 * ([nlevels+3][oh][ow] floats) and gradient map ([nlevels][oh][ow][2]) to host buffers (either may be NULL). */
int pano_b200_sift_octave_dims(pano_b200_ctx* ctx, int octave, int* ow, int* oh);
int pano_b200_sift_octave_dump(pano_b200_ctx* ctx, int octave, float* gss, float* grad);

/* ImageProcess::getImgPair (ImageProcess.cpp:273-351): A = database, B = queries, both feature tables in sorted order.
 * match_idx[b] = row of A matched by row b of B, or -1.  Returns the number of matches through *nmatches. */
int pano_b200_match(pano_b200_ctx* ctx, const float* descrA, int nA, const float* descrB, int nB, int* match_idx,
                    int* nmatches);
/* both directed problems of an image pair in one call, as the reference evaluates them for every stitched edge
 * (ImageProcess.cpp:177-178): idx_ab[b] = getImgPair(A, B) (database A, queries B, nB entries), idx_ba[a] =
 * getImgPair(B, A) (nA entries).  With the default matcher both come from ONE pass over the |A| x |B| SAD matrix. */
int pano_b200_match_pair(pano_b200_ctx* ctx, const float* descrA, int nA, const float* descrB, int nB, int* idx_ab,
                         int* idx_ba);

/* ImageProcess::RANSAC (ImageProcess.cpp:395-436): pairs (src -> dst) -> 8 bilinear coefficients in Homography
 * constructor order (x' = H[0] x + H[1] y + H[2] x y + H[3]; y' = H[4] x + H[5] y + H[6] x y + H[7]).
 * Optional outputs: counts[72] inlier count per hypothesis, hyps[72][8] fitted hypotheses (getHomographyMat,
 * ImageProcess.cpp:439-462), inliers[npairs] / *ninliers the winning inlier set (getInlinerIndex, :473-497). */
int pano_b200_ransac(pano_b200_ctx* ctx, const pano_b200_pair* pairs, int npairs, double* H8, int* counts,
                     double* hyps, int* inliers, int* ninliers);

/* Canvas sizing (ImageProcess.cpp:206-216, 532-594): bounds[4] = min_x, min_y, max_x, max_y (already merged with the
 * current result size), size[2] = new_width, new_height.  Host arithmetic only. */
int pano_b200_plan_canvas(int dst_w, int dst_h, const double* forward_H8, int result_w, int result_h, float* bounds,
                          int* size);
/* same with the profile explicit (src/ex6/ImageProcess.cpp:230-243, 545-580 for PANO_B200_PROFILE_EX6) */
int pano_b200_plan_canvas_ex(int variant, int dst_w, int dst_h, const double* forward_H8, int result_w, int result_h,
                             float* bounds, int* size);

/* warpingImageByHomography + movingImageByOffset (ImageProcess.cpp:596-620) in one pass over the new canvas.
 * src (sw x sh) is warped with H8 and float offsets into a; prev (pw x ph) is shifted by the int offsets into b.
 * Pass src = NULL or prev = NULL to run only one of them. */
int pano_b200_warp_shift(pano_b200_ctx* ctx, const uint8_t* src, int sw, int sh, const double* H8, float offx,
                         float offy, const uint8_t* prev, int pw, int ph, int ioffx, int ioffy, int cw, int ch,
                         uint8_t* a, uint8_t* b);

/* ImageProcess::blendTwoImages (ImageProcess.cpp:648-773) */
int pano_b200_blend(pano_b200_ctx* ctx, const uint8_t* a, const uint8_t* b, int w, int h, uint8_t* out);
/* equalization(tmp, 1) + the Y mix that closes ImageProcess::matching (equalization.cpp:74-131, ImageProcess.cpp:237-268) */
int pano_b200_equalize_mix(pano_b200_ctx* ctx, const uint8_t* rgb, int w, int h, uint8_t* out);
/* CImg<float>::get_blur(2, true, true) (CImg.h:35111-35147, vanvliet) and get_resize(nw, nh, 1, c, 3)
 * (CImg.h:29321-29700) on float planes [c][h][w]; the two CImg primitives blendTwoImages is made of. */
int pano_b200_cimg_blur2(pano_b200_ctx* ctx, const float* src, int w, int h, int c, float* dst);
/* CImg<float>::get_blur(2) = Deriche order 0 (CImg.h:34777-34869), the pyramid blur of src/ex6/ImageProcess.cpp:702-705 */
int pano_b200_cimg_blur2_deriche(pano_b200_ctx* ctx, const float* src, int w, int h, int c, float* dst);
int pano_b200_cimg_resize3(pano_b200_ctx* ctx, const float* src, int w, int h, int c, int nw, int nh, float* dst);

/* ---- uint8 / tensor-core matcher (north-star stage 3).  NOT part of the reference-parity path: the reference matches
 *      float descriptors under L1 (ImageProcess.cpp:273-351).  Descriptors are quantised with VLFeat's convention
 *      q = (uint8) min(512 x, 255) and compared under squared L2 with int32 accumulation on tcgen05 tensor cores;
 *      match_idx[b] = row of A with 4 d0 < d1, else -1; d01 (optional, [nB][3]) = d0, d1, nearest row. */
int pano_b200_quantize_u8(pano_b200_ctx* ctx, const float* descr, int n, uint8_t* out);
int pano_b200_match_u8(pano_b200_ctx* ctx, const uint8_t* descrA, int nA, const uint8_t* descrB, int nB, int* match_idx,
                       int* d01, int* nmatches);
/* times the matcher kernels alone on resident tables (CUDA events), milliseconds per repetition; descrA / descrB are
 * row-major host tables, or both NULL for uniform pseudo-random bytes */
int pano_b200_bench_match_u8(pano_b200_ctx* ctx, const uint8_t* descrA, int nA, const uint8_t* descrB, int nB, int reps,
                             float* ms_per_rep);

/* ---- sharded panorama job with a device-resident exchange (images shard over the GPUs of a box, one
 *      all-gather of descriptor blocks, directed matching problems dealt to the ranks, rank 0 stitches).  The reference has
 *      no counterpart: its readFile / getImgPair loops (ImageProcess.cpp:12-23, 117-137) are what is being partitioned.
 *      Pointers named d_* are DEVICE memory owned by the caller (the NCCL send / receive buffers); the library copies
 *      device-to-device on its own stream and returns when the copy is complete. -------------------------------------- */
int pano_b200_shard_begin(pano_b200_ctx* ctx, int n_global);            /* new job with image slots 0 .. n_global-1 */
/* readFile (projection + SIFT + table) of the n_local images this rank owns; slot[k] = global index of image k;
 * on_device != 0: imgs[k] are device pointers (inputs already staged in HBM) */
int pano_b200_shard_extract(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, const int* slot,
                            int n_local, int on_device);
/* copies image i's descriptor table [nfeat][128] f32 and projected planar RGB into device buffers, its keypoints into a
 * host buffer; any of the three may be NULL */
int pano_b200_shard_export(pano_b200_ctx* ctx, int i, float* d_descr_out, pano_b200_keypoint* keys_out, uint8_t* d_proj_out);
/* fills slot i from exchange buffers; d_proj may be NULL on ranks that do not stitch */
int pano_b200_shard_import(pano_b200_ctx* ctx, int i, int w, int h, int nfeat, const float* d_descr,
                           const pano_b200_keypoint* keys, const uint8_t* d_proj);
/* getImgPair(image I[k], image J[k]) for k < nprob in one batch; d_idx_out receives the lists back to back
 * (nfeat[J[k]] ints each) */
int pano_b200_shard_match(pano_b200_ctx* ctx, const int* I, const int* J, int nprob, int* d_idx_out);
/* match list of the directed problem (i, j), evaluated on another rank (host memory, nfeat[j] entries) */
int pano_b200_shard_preset(pano_b200_ctx* ctx, int i, int j, const int* idx, int n);
/* the sequential part (adjacency, order, RANSAC, warp, blend, equalisation) on the job's images with the preset lists;
 * out (optional, host) receives the panorama when out_cap is large enough (else -4) */
int pano_b200_shard_stitch(pano_b200_ctx* ctx, uint8_t* out, size_t out_cap, int* out_w, int* out_h);

/* ---- plane-sharded canvas stages.  The colour planes of warp / shift / blend are independent except for the seam
 *      statistics of plane 0 (16 bytes per stitched edge, ImageProcess.cpp:659-671) and the final equalisation, so up to
 *      three ranks that hold the same images and match lists can each carry one plane.  exchange() is called once per
 *      edge on every participant: is_source = 1 on the rank carrying plane 0 (stats4 holds the values to send), 0
 *      elsewhere (stats4 receives them); it returns 0 on success. */
typedef int (*pano_b200_seam_exchange)(void* user, int* stats4, int is_source);
int pano_b200_shard_stitch_planes(pano_b200_ctx* ctx, int first_plane, int nplanes, pano_b200_seam_exchange exchange,
                                  void* user, int* out_w, int* out_h);       /* everything but the equalisation tail */
int pano_b200_shard_plane_export(pano_b200_ctx* ctx, int k, uint8_t* d_out);   /* plane k of this rank's result (device) */
int pano_b200_shard_plane_import(pano_b200_ctx* ctx, int channel, const uint8_t* d_in);
int pano_b200_shard_tail(pano_b200_ctx* ctx, uint8_t* out, size_t out_cap, int* out_w, int* out_h);

/* ---- Reinhard l-alpha-beta colour transfer: the reference's `transfer tran(src, tem, out)` (transfer.cpp:4-13; its
 *      call site ImageProcess.cpp:180-182 is commented out in the reference, so this is a stage of its own).  src, tem,
 *      out: planar RGB uint8 in host memory; out has the size of src.  Parity: within 1 LSB of the reference (device
 *      log / pow are within 1 ulp of glibc's, not identical); the plane sums keep the reference's serial float order. */
int pano_b200_color_transfer(pano_b200_ctx* ctx, const uint8_t* src, int w, int h, const uint8_t* tem, int tw, int th,
                             uint8_t* out);

/* after pano_b200_bench_match_u8: *ms = ms per repetition of the SAME kernel with the epilogue reduced to releasing the
 * accumulators (TMA + UTCIMMA only: the tensor-pipe peak this tile shape can reach), *ksteps = K steps of 32 bytes per
 * MMA tile (4 descriptor steps + the norm-extension steps), i.e. int8 ops issued = 2 * 32 * ksteps * nA * nB */
int pano_b200_bench_match_u8_peak(pano_b200_ctx* ctx, float* ms, int* ksteps);

/* ---- measurement helpers --------------------------------------------------------------------------------------- */
void* pano_b200_alloc_pinned(size_t bytes);           /* page-locked host memory for timed host<->device copies */
void pano_b200_free_pinned(void* p);
/* number of concurrent per-image lanes (stream + SIFT engine + host thread) used by the pipeline; default 8 */
int pano_b200_set_lanes(pano_b200_ctx* ctx, int nlanes);
/* matcher of getImgPair (ImageProcess.cpp:273-351): PANO_B200_MATCH_PREFILTER (default) = rigorous two-level uint8
 * pre-filter (grouped lower bound on 32 group bytes per row, exact 128-byte SAD of the row pairs it cannot skip, certain
 * reject / certain accept) + exact float-L1 re-rank of the undecided queries, both directions of an image pair from one
 * pass; PANO_B200_MATCH_PREFILTER_FULLSAD = the same with the full SAD of every row pair; PANO_B200_MATCH_PREFILTER_ONEDIR
 * = one full-SAD pass per directed problem; PANO_B200_MATCH_FULL = exact float-L1 scan of every (query, row) pair.  All
 * return the reference's match lists bit for bit (the pre-filter never drops a row the exact rule needs and decides a
 * query without float arithmetic only where the reference's decision is certain). */
#define PANO_B200_MATCH_PREFILTER 0
#define PANO_B200_MATCH_FULL 1
#define PANO_B200_MATCH_PREFILTER_ONEDIR 2
#define PANO_B200_MATCH_PREFILTER_FULLSAD 3   /* pre-filter with the full 128-byte SAD of every row pair (no grouped bound) */
int pano_b200_set_match_mode(pano_b200_ctx* ctx, int mode);
/* pre-filter bookkeeping since the last reset: out[0] = queries, out[1] = queries the SAD pass could not reject,
 * out[2] = queries that fell back to the full scan (candidate list overflow), out[3] = directed problems, out[4] = image
 * pairs served by the symmetric pass; reset != 0 clears the counters after reading */
int pano_b200_match_stats(pano_b200_ctx* ctx, long long out[5], int reset);
/* the same counters plus out[5] = image pairs through the grouped pass, out[6] = exact SADs the grouped pass evaluated
 * (of out[5] pairs' |X| x |Y| row pairs), out[7] = queries it accepted with certainty, out[8] = batches redone with
 * the full SAD pass because its pair queue overflowed; fills min(n, 9) entries */
int pano_b200_match_stats_ex(pano_b200_ctx* ctx, long long* out, int n, int reset);
int pano_b200_flush_l2(pano_b200_ctx* ctx);           /* overwrite a 256 MB scratch buffer (2x L2) */
int pano_b200_timer_start(pano_b200_ctx* ctx);        /* CUDA events on the context's stream */
int pano_b200_timer_stop(pano_b200_ctx* ctx, float* ms);
void pano_b200_ktimer_enable(int on);                 /* per-kernel CUDA-event timing (adds 2 event records / launch) */
void pano_b200_ktimer_reset(void);
long pano_b200_ktimer_launches(void);                 /* kernels launched by this library since the last reset */
int pano_b200_ktimer_report(char* dst, int cap);      /* JSON {kernel: {launches, ms, bytes}} */

#ifdef __cplusplus
}
#endif
#endif
