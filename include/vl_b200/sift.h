/* vl_b200/sift.h -- drop-in replacement for VLFeat's vl/sift.h (the subset the stitcher binds), backed by
 * libpano_b200.so.  An application that includes this header instead of "vl/sift.h" and links libpano_b200.so instead
 * of libvl keeps calling
 *     vl_sift_new / vl_sift_process_first_octave / vl_sift_process_next_octave / vl_sift_detect /
 *     vl_sift_calc_keypoint_orientations / vl_sift_calc_keypoint_descriptor / vl_sift_delete
 * exactly as ImageProcess::siftAlgorithm does (ImageProcess.cpp:55-96); the scale space, detector, orientation and
 * descriptor stages then run as sm_100a kernels.
 *
 * ABI notes (reference: vl/sift.h:19-78):
 *  - VlSiftKeypoint and the leading part of VlSiftFilt keep VLFeat's field order and types, because callers read
 *    f->keys / f->nkeys directly (ImageProcess.cpp:64,66) and VLFeat's inline getters read the other fields.
 *  - `temp`, `octave`, `dog`, `grad` hold DEVICE addresses unless mirroring is switched on with
 *    vl_b200_sift_set_mirror(f, 1), in which case they point to host copies refreshed by every process_* / detect
 *    call (used by the parity tests; costs a device-to-host copy per call).
 *  - `keys` is library-owned host memory, valid until the next detect / process_* / delete call (as in VLFeat).
 *  - Return codes: VL_ERR_OK (0), VL_ERR_EOF (5) as in vl/generic.h:108-113; VL_ERR_BAD_ARG (3) for
 *    configurations outside the accelerated path (o_min != 0).
 */
#ifndef VL_B200_SIFT_H
#define VL_B200_SIFT_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VL_ERR_OK 0
#define VL_ERR_OVERFLOW 1
#define VL_ERR_ALLOC 2
#define VL_ERR_BAD_ARG 3
#define VL_ERR_IO 4
#define VL_ERR_EOF 5

typedef float vl_sift_pix;
#ifndef VL_B200_BASIC_TYPES
#define VL_B200_BASIC_TYPES
typedef unsigned long long vl_size;   /* vl/host.h:392 */
typedef unsigned long long vl_uindex; /* vl/host.h:394 */
typedef unsigned int vl_uint32;       /* vl/host.h:382 */
typedef vl_uint32 vl_type;            /* vl/generic.h:18 */
#define VL_TYPE_FLOAT 1               /* vl/generic.h:21 */
#define VL_TYPE_DOUBLE 2              /* vl/generic.h:22 */
#endif

typedef struct _VlSiftKeypoint {
    int o;
    int ix, iy, is;
    float x, y, s, sigma;
} VlSiftKeypoint;

typedef struct _VlSiftFilt {
    double sigman, sigma0, sigmak, dsigma0;
    int width, height, O, S, o_min, s_min, s_max, o_cur;
    vl_sift_pix *temp, *octave, *dog;
    int octave_width, octave_height;
    vl_sift_pix* gaussFilter;
    double gaussFilterSigma;
    vl_size gaussFilterWidth;
    VlSiftKeypoint* keys;
    int nkeys, keys_res;
    double peak_thresh, edge_thresh, norm_thresh, magnif, windowSize;
    vl_sift_pix* grad;
    int grad_o;
    /* ---- end of the VLFeat layout; private tail of the B200 implementation ---- */
    void* b200_impl;
} VlSiftFilt;

VlSiftFilt* vl_sift_new(int width, int height, int noctaves, int nlevels, int o_min);
void vl_sift_delete(VlSiftFilt* f);
int vl_sift_process_first_octave(VlSiftFilt* f, vl_sift_pix const* im);
int vl_sift_process_next_octave(VlSiftFilt* f);
void vl_sift_detect(VlSiftFilt* f);
int vl_sift_calc_keypoint_orientations(VlSiftFilt* f, double angles[4], VlSiftKeypoint const* k);
void vl_sift_calc_keypoint_descriptor(VlSiftFilt* f, vl_sift_pix* descr, VlSiftKeypoint const* k, double angle);

/* B200 extensions */
void vl_b200_sift_set_mirror(VlSiftFilt* f, int on);   /* keep host copies of octave / dog / grad (parity tests) */
void vl_b200_sift_set_device(int device);              /* device used by subsequent vl_sift_new calls (default 0) */

/* the getters / setters VLFeat defines inline (vl/sift.h:134-404) */
static inline int vl_sift_get_octave_index(VlSiftFilt const* f) { return f->o_cur; }
static inline int vl_sift_get_noctaves(VlSiftFilt const* f) { return f->O; }
static inline int vl_sift_get_octave_first(VlSiftFilt const* f) { return f->o_min; }
static inline int vl_sift_get_octave_width(VlSiftFilt const* f) { return f->octave_width; }
static inline int vl_sift_get_octave_height(VlSiftFilt const* f) { return f->octave_height; }
static inline int vl_sift_get_nlevels(VlSiftFilt const* f) { return f->S; }
static inline int vl_sift_get_nkeypoints(VlSiftFilt const* f) { return f->nkeys; }
static inline VlSiftKeypoint const* vl_sift_get_keypoints(VlSiftFilt const* f) { return f->keys; }
static inline double vl_sift_get_peak_thresh(VlSiftFilt const* f) { return f->peak_thresh; }
static inline double vl_sift_get_edge_thresh(VlSiftFilt const* f) { return f->edge_thresh; }
static inline double vl_sift_get_norm_thresh(VlSiftFilt const* f) { return f->norm_thresh; }
static inline double vl_sift_get_magnif(VlSiftFilt const* f) { return f->magnif; }
static inline double vl_sift_get_window_size(VlSiftFilt const* f) { return f->windowSize; }
static inline vl_sift_pix* vl_sift_get_octave(VlSiftFilt const* f, int s) {
    return f->octave + (size_t)f->octave_width * f->octave_height * (s - f->s_min);
}
static inline void vl_sift_set_peak_thresh(VlSiftFilt* f, double t) { f->peak_thresh = t; }
static inline void vl_sift_set_edge_thresh(VlSiftFilt* f, double t) { f->edge_thresh = t; }
static inline void vl_sift_set_norm_thresh(VlSiftFilt* f, double t) { f->norm_thresh = t; }
static inline void vl_sift_set_magnif(VlSiftFilt* f, double m) { f->magnif = m; }
static inline void vl_sift_set_window_size(VlSiftFilt* f, double x) { f->windowSize = x; }

#ifdef __cplusplus
}
#endif
#endif
