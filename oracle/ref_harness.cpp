// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin C-ABI harness over the *real* reference (chensh236/ComputerVisionImageStich2, root variant)
// so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can run the
// reference's own functions on buffers.  It is compiled by oracle/Makefile against the sources where they
// lie under /root/reference (nothing is copied); the only output is oracle/_ref/libpano_ref.so.
//
// Every entry point calls the reference's own method named in its comment; no arithmetic is restated here.
// Private members of ImageProcess are reached with the usual `#define private public` test hack.
//
// Mechanical adjustments made by the build recipe (see oracle/Makefile):
//   * the two `result.display()` calls (ImageProcess.cpp:233,270) are dropped by a sed pipe at compile time
//     (CImg throws CImgDisplayException when cimg_display==0);
//   * vl/mathop.c is compiled at -O0 (vl_get_vector_comparison_function_f lacks a `return`, mathop.c:470-487);
//   * everything is built with -ffp-contract=off (the pipeline is numerically chaotic under FMA contraction).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <queue>
#include <chrono>
#include <iostream>
#include <sstream>
#include <new>

#define private public
#include "ImageProcess.h"
#undef private

extern "C" {

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
namespace {

// An ImageProcess whose constructor has NOT run (the only constructor does the whole job from files).
struct Holder {
    alignas(ImageProcess) unsigned char buf[sizeof(ImageProcess)];
    ImageProcess *ip;
    Holder() {
        ip = reinterpret_cast<ImageProcess *>(buf);
        new (&ip->imgs) std::vector<Image>();
        new (&ip->result) CImg<unsigned char>();
        ip->stichingMat = 0;
    }
    ~Holder() {
        if (ip->stichingMat) {
            for (size_t i = 0; i < ip->imgs.size(); ++i) delete[] ip->stichingMat[i];
            delete[] ip->stichingMat;
        }
        ip->result.~CImg<unsigned char>();
        ip->imgs.~vector<Image>();
    }
};

CImg<unsigned char> wrap_u8(const uint8_t *p, int w, int h, int c) {
    CImg<unsigned char> img(w, h, 1, c);
    std::memcpy(img.data(), p, (size_t)w * h * c);
    return img;
}

struct SilenceStdout {
    std::streambuf *old;
    std::ostringstream sink;
    SilenceStdout() { old = std::cout.rdbuf(sink.rdbuf()); }
    ~SilenceStdout() { std::cout.rdbuf(old); }
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Image make_image(const float *descr, const VlSiftKeypoint *keys, int n) {
    Image im;
    for (int i = 0; i < n; ++i) {
        std::vector<float> d(descr + (size_t)i * 128, descr + (size_t)(i + 1) * 128);
        im.features.insert(std::pair<std::vector<float>, VlSiftKeypoint>(d, keys[i]));
    }
    return im;
}

int flatten(const std::map<std::vector<float>, VlSiftKeypoint> &f, float *descr, VlSiftKeypoint *keys, int cap) {
    int i = 0;
    for (auto it = f.begin(); it != f.end(); ++it, ++i) {
        if (i < cap) {
            if (descr) std::memcpy(descr + (size_t)i * 128, it->first.data(), 128 * sizeof(float));
            if (keys) keys[i] = it->second;
        }
    }
    return i;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// stage entry points (each calls the reference method named on the right)
// ---------------------------------------------------------------------------------------------------------------

// Projection::imageProjection (Projection.cpp:20-73).  planar u8 [c][y][x], 3 channels in and out.
int ref_project(const uint8_t *src, int w, int h, uint8_t *dst) {
    CImg<unsigned char> out = Projection::imageProjection(wrap_u8(src, w, h, 3));
    std::memcpy(dst, out.data(), (size_t)w * h * 3);
    return 0;
}

// ImageProcess::toGrayScale (ImageProcess.cpp:27-40)
int ref_gray(const uint8_t *src, int w, int h, uint8_t *dst) {
    Holder H;
    CImg<unsigned char> out = H.ip->toGrayScale(wrap_u8(src, w, h, 3));
    std::memcpy(dst, out.data(), (size_t)w * h);
    return 0;
}

// ImageProcess::siftAlgorithm (ImageProcess.cpp:44-99) on a 1-channel u8 image.
// Returns the number of features in the std::map (sorted by descriptor, de-duplicated); fills up to cap rows.
int ref_sift_features(const uint8_t *gray, int w, int h, float *descr, VlSiftKeypoint *keys, int cap) {
    Holder H;
    std::map<std::vector<float>, VlSiftKeypoint> f = H.ip->siftAlgorithm(wrap_u8(gray, w, h, 1));
    return flatten(f, descr, keys, cap);
}

// Raw VLFeat dump: runs the exact vl_sift_* call sequence of siftAlgorithm (ImageProcess.cpp:55-92) on a float
// image and records, per octave, the GSS, DoG and gradient buffers, the refined keypoints, their orientations and
// one descriptor per (keypoint, angle) in call order.
struct RefOctave {
    int w, h, nkeys, ndesc;
    std::vector<float> gss, dog, grad;
    std::vector<VlSiftKeypoint> keys;
    std::vector<int> nangles;
    std::vector<double> angles;     // [nkeys][4]
    std::vector<float> descr;       // [ndesc][128]
    std::vector<int> descr_key;     // [ndesc] -> key index
    std::vector<int> descr_written; // [ndesc] 0 if vl_sift_calc_keypoint_descriptor bailed out (sift.c:1321-1328)
};
struct RefSiftDump {
    std::vector<RefOctave> oct;
};

RefSiftDump *ref_sift_dump_run(const float *im, int w, int h, int noctaves, int nlevels, int o_min) {
    RefSiftDump *D = new RefSiftDump();
    VlSiftFilt *f = vl_sift_new(w, h, noctaves, nlevels, o_min);
    if (vl_sift_process_first_octave(f, im) != VL_ERR_EOF) {
        while (true) {
            vl_sift_detect(f);
            RefOctave O;
            O.w = f->octave_width;
            O.h = f->octave_height;
            size_t n = (size_t)O.w * O.h;
            int nl = f->s_max - f->s_min + 1;
            O.gss.assign(f->octave, f->octave + n * nl);
            O.dog.assign(f->dog, f->dog + n * (nl - 1));
            O.nkeys = f->nkeys;
            O.keys.assign(f->keys, f->keys + f->nkeys);
            O.nangles.resize(O.nkeys);
            O.angles.assign((size_t)O.nkeys * 4, 0.0);
            for (int i = 0; i < O.nkeys; ++i) {
                VlSiftKeypoint k = f->keys[i];
                double ang[4];
                int na = vl_sift_calc_keypoint_orientations(f, ang, &k);
                O.nangles[i] = na;
                for (int j = 0; j < na; ++j) {
                    O.angles[(size_t)i * 4 + j] = ang[j];
                    float d[128];
                    for (int q = 0; q < 128; ++q) d[q] = -1.0f; // sentinel: descriptors are >= 0 when written
                    vl_sift_calc_keypoint_descriptor(f, d, &k, ang[j]);
                    O.descr_written.push_back(d[0] >= 0.0f ? 1 : 0);
                    O.descr.insert(O.descr.end(), d, d + 128);
                    O.descr_key.push_back(i);
                }
            }
            O.ndesc = (int)O.descr_key.size();
            // gradient buffer is valid for levels s_min+1 .. s_max-2 once any orientation call ran (sift.c:792-876)
            int ng = (f->s_max - 2) - (f->s_min + 1) + 1;
            if (f->grad_o == f->o_cur && ng > 0) O.grad.assign(f->grad, f->grad + 2 * n * ng);
            D->oct.push_back(std::move(O));
            if (vl_sift_process_next_octave(f) == VL_ERR_EOF) break;
        }
    }
    vl_sift_delete(f);
    return D;
}
int ref_sift_dump_noctaves(RefSiftDump *D) { return (int)D->oct.size(); }
void ref_sift_dump_info(RefSiftDump *D, int o, int *w, int *h, int *nkeys, int *ndesc, int *has_grad) {
    RefOctave &O = D->oct[o];
    *w = O.w; *h = O.h; *nkeys = O.nkeys; *ndesc = O.ndesc; *has_grad = O.grad.empty() ? 0 : 1;
}
// what: 0 gss, 1 dog, 2 grad, 3 keys, 4 nangles, 5 angles, 6 descr, 7 descr_key, 8 descr_written
void ref_sift_dump_copy(RefSiftDump *D, int o, int what, void *dst) {
    RefOctave &O = D->oct[o];
    switch (what) {
    case 0: std::memcpy(dst, O.gss.data(), O.gss.size() * 4); break;
    case 1: std::memcpy(dst, O.dog.data(), O.dog.size() * 4); break;
    case 2: std::memcpy(dst, O.grad.data(), O.grad.size() * 4); break;
    case 3: std::memcpy(dst, O.keys.data(), O.keys.size() * sizeof(VlSiftKeypoint)); break;
    case 4: std::memcpy(dst, O.nangles.data(), O.nangles.size() * 4); break;
    case 5: std::memcpy(dst, O.angles.data(), O.angles.size() * 8); break;
    case 6: std::memcpy(dst, O.descr.data(), O.descr.size() * 4); break;
    case 7: std::memcpy(dst, O.descr_key.data(), O.descr_key.size() * 4); break;
    case 8: std::memcpy(dst, O.descr_written.data(), O.descr_written.size() * 4); break;
    }
}
void ref_sift_dump_free(RefSiftDump *D) { delete D; }

// ImageProcess::getImgPair (ImageProcess.cpp:273-351): A = database (kd-forest), B = queries.
// Inputs are feature tables (any order; the std::map re-sorts them).  Output pairs (src = A keypoint, dst = B keypoint).
int ref_match(const float *descA, const VlSiftKeypoint *keysA, int nA, const float *descB,
              const VlSiftKeypoint *keysB, int nB, VlSiftKeypoint *outA, VlSiftKeypoint *outB, int cap) {
    Holder H;
    Image A = make_image(descA, keysA, nA), B = make_image(descB, keysB, nB);
    std::vector<ImgPair> p = H.ip->getImgPair(A, B);
    for (size_t i = 0; i < p.size() && (int)i < cap; ++i) {
        outA[i] = p[i].src;
        outB[i] = p[i].dst;
    }
    return (int)p.size();
}

// The kd-forest exactly as getImgPair drives it (ImageProcess.cpp:280-327): 1 tree, float data, exact search; the K
// nearest neighbours of each query through vl_kdforestsearcher_query (per_query != 0) or vl_kdforest_query_with_array.
// idx [nq][K] (int64, -1 = none), dist [nq][K] (double, the VlKDForestNeighbor.distance values).
int ref_kdforest_query(const float *data, int n, int dim, int metric, const float *queries, int nq, int K,
                       int per_query, long long *idx, double *dist) {
    VlKDForest *forest = vl_kdforest_new(VL_TYPE_FLOAT, dim, 1, (VlVectorComparisonType)metric);
    if (!forest) return -1;
    vl_kdforest_build(forest, n, data);
    if (per_query) {
        VlKDForestSearcher *searcher = vl_kdforest_new_searcher(forest);
        std::vector<VlKDForestNeighbor> nb(K);
        for (int q = 0; q < nq; ++q) {
            vl_kdforestsearcher_query(searcher, nb.data(), K, queries + (size_t)q * dim);
            for (int k = 0; k < K; ++k) {
                idx[(size_t)q * K + k] = (long long)nb[k].index;
                dist[(size_t)q * K + k] = nb[k].distance;
            }
        }
        vl_kdforestsearcher_delete(searcher);
    } else {
        std::vector<vl_uint32> ix((size_t)nq * K);
        std::vector<float> ds((size_t)nq * K);
        vl_kdforest_query_with_array(forest, ix.data(), K, nq, ds.data(), queries);
        for (size_t i = 0; i < ix.size(); ++i) { idx[i] = ix[i] == (vl_uint32)-1 ? -1 : (long long)ix[i]; dist[i] = ds[i]; }
    }
    vl_kdforest_delete(forest);
    return 0;
}

// ImageProcess::RANSAC (ImageProcess.cpp:395-436): pairs (src -> dst); returns the 8 bilinear coefficients
// in Homography constructor order (x': a b c d ; y': e f g h).
int ref_ransac(const VlSiftKeypoint *src, const VlSiftKeypoint *dst, int n, double *H8) {
    Holder H;
    std::vector<ImgPair> pairs;
    for (int i = 0; i < n; ++i) pairs.push_back(ImgPair(src[i], dst[i]));
    Homography R = H.ip->RANSAC(pairs);
    H8[0] = R.H[0][0]; H8[1] = R.H[0][1]; H8[2] = R.H[0][2]; H8[3] = R.H[1][0];
    H8[4] = R.H[1][1]; H8[5] = R.H[1][2]; H8[6] = R.H[2][0]; H8[7] = R.H[2][1];
    return 0;
}

static Homography mkH(const double *h) { return Homography(h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]); }

// ImageProcess::getHomographyMat (ImageProcess.cpp:439-462) on exactly 4 pairs.
int ref_fit4(const VlSiftKeypoint *src, const VlSiftKeypoint *dst, double *H8) {
    Holder H;
    std::vector<ImgPair> pairs;
    for (int i = 0; i < 4; ++i) pairs.push_back(ImgPair(src[i], dst[i]));
    Homography R = H.ip->getHomographyMat(pairs);
    H8[0] = R.H[0][0]; H8[1] = R.H[0][1]; H8[2] = R.H[0][2]; H8[3] = R.H[1][0];
    H8[4] = R.H[1][1]; H8[5] = R.H[1][2]; H8[6] = R.H[2][0]; H8[7] = R.H[2][1];
    return 0;
}

// ImageProcess::getInlinerIndex (ImageProcess.cpp:473-497)
int ref_inliers(const VlSiftKeypoint *src, const VlSiftKeypoint *dst, int n, const double *H8, int *idx) {
    Holder H;
    std::vector<ImgPair> pairs;
    for (int i = 0; i < n; ++i) pairs.push_back(ImgPair(src[i], dst[i]));
    Homography R = mkH(H8);
    std::vector<int> in = H.ip->getInlinerIndex(pairs, R, std::set<int>());
    for (size_t i = 0; i < in.size(); ++i) idx[i] = in[i];
    return (int)in.size();
}

// ImageProcess::getInlinerHomography (ImageProcess.cpp:500-529)
int ref_refit(const VlSiftKeypoint *src, const VlSiftKeypoint *dst, int n, const int *idx, int nin, double *H8) {
    Holder H;
    std::vector<ImgPair> pairs;
    for (int i = 0; i < n; ++i) pairs.push_back(ImgPair(src[i], dst[i]));
    std::vector<int> in(idx, idx + nin);
    Homography R = H.ip->getInlinerHomography(pairs, in);
    H8[0] = R.H[0][0]; H8[1] = R.H[0][1]; H8[2] = R.H[0][2]; H8[3] = R.H[1][0];
    H8[4] = R.H[1][1]; H8[5] = R.H[1][2]; H8[6] = R.H[2][0]; H8[7] = R.H[2][1];
    return 0;
}

// canvas bounds: get{Min,Max}{X,Y}AfterWarping (ImageProcess.cpp:532-594) -> out[4] = minx,miny,maxx,maxy
int ref_warp_bounds(int w, int h, const double *H8, float *out) {
    Holder H;
    CImg<unsigned char> img(w, h, 1, 3, 0);
    Homography R = mkH(H8);
    out[0] = H.ip->getMinXAfterWarping(img, R);
    out[1] = H.ip->getMinYAfterWarping(img, R);
    out[2] = H.ip->getMaxXAfterWarping(img, R);
    out[3] = H.ip->getMaxYAfterWarping(img, R);
    return 0;
}

// ImageProcess::warpingImageByHomography (ImageProcess.cpp:596-606): dst is a zeroed (cw x ch x 3) canvas.
int ref_warp(const uint8_t *src, int w, int h, const double *H8, float offx, float offy, int cw, int ch, uint8_t *dst) {
    Holder H;
    CImg<unsigned char> s = wrap_u8(src, w, h, 3), d(cw, ch, 1, 3, 0);
    Homography R = mkH(H8);
    H.ip->warpingImageByHomography(s, d, R, offx, offy);
    std::memcpy(dst, d.data(), (size_t)cw * ch * 3);
    return 0;
}

// ImageProcess::movingImageByOffset (ImageProcess.cpp:608-620)
int ref_shift(const uint8_t *src, int w, int h, int offx, int offy, int cw, int ch, uint8_t *dst) {
    Holder H;
    CImg<unsigned char> s = wrap_u8(src, w, h, 3), d(cw, ch, 1, 3, 0);
    H.ip->movingImageByOffset(s, d, offx, offy);
    std::memcpy(dst, d.data(), (size_t)cw * ch * 3);
    return 0;
}

// ImageProcess::blendTwoImages (ImageProcess.cpp:648-773)
int ref_blend(const uint8_t *a, const uint8_t *b, int w, int h, uint8_t *out) {
    Holder H;
    CImg<unsigned char> r = H.ip->blendTwoImages(wrap_u8(a, w, h, 3), wrap_u8(b, w, h, 3));
    if (r.width() != w || r.height() != h || r.spectrum() != 3) return -1;
    std::memcpy(out, r.data(), (size_t)w * h * 3);
    return 0;
}

// CImg pieces used by blendTwoImages, for level-by-level parity: get_blur(2,true,true) on a float plane set
int ref_cimg_blur2(const float *src, int w, int h, int c, float *dst) {
    CImg<float> s(w, h, 1, c);
    std::memcpy(s.data(), src, (size_t)w * h * c * 4);
    CImg<float> r = s.get_blur(2, true, true);
    std::memcpy(dst, r.data(), (size_t)w * h * c * 4);
    return 0;
}
// get_resize(nw, nh, 1, c, 3)
int ref_cimg_resize3(const float *src, int w, int h, int c, int nw, int nh, float *dst) {
    CImg<float> s(w, h, 1, c);
    std::memcpy(s.data(), src, (size_t)w * h * c * 4);
    CImg<float> r = s.get_resize(nw, nh, 1, c, 3);
    if (r.width() != nw || r.height() != nh) return -1;
    std::memcpy(dst, r.data(), (size_t)nw * nh * c * 4);
    return 0;
}

// The post-processing tail of ImageProcess::matching (ImageProcess.cpp:237-268): equalization(tmp,1) + Y mix.
// That code is inline in matching(); it is reached here by running matching() on a single image (the BFS loop
// then has no edges and the tail runs on `result = imgs[0].projectedSrc`).
int ref_equalize_mix(const uint8_t *img, int w, int h, uint8_t *out) {
    Holder H;
    SilenceStdout q;
    Image im;
    im.projectedSrc = wrap_u8(img, w, h, 3);
    H.ip->imgs.push_back(im);
    H.ip->matching();
    std::memcpy(out, H.ip->result.data(), (size_t)w * h * 3);
    return 0;
}

// equalization(tmp, 1) alone (equalization.cpp:4-25, 74-131)
int ref_equalize(const uint8_t *img, int w, int h, uint8_t *out) {
    CImg<unsigned char> t = wrap_u8(img, w, h, 3);
    equalization e(t, 1);
    std::memcpy(out, t.data(), (size_t)w * h * 3);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// whole pipeline
// ---------------------------------------------------------------------------------------------------------------
struct RefPano {
    CImg<unsigned char> result;
    double t_features, t_matching;
    std::vector<int> nfeat;
    std::string log;
};

// The ImageProcess constructor body (ImageProcess.cpp:3-8, readFile :11-24) on in-memory planar RGB images.
RefPano *ref_stitch_mem(const uint8_t *const *imgs, const int *w, const int *h, int n) {
    RefPano *P = new RefPano();
    Holder H;
    SilenceStdout q;
    double t0 = now_s();
    for (int i = 0; i < n; ++i) {
        Image cur;
        cur.projectedSrc = Projection::imageProjection(wrap_u8(imgs[i], w[i], h[i], 3));
        cur.features = H.ip->siftAlgorithm(H.ip->toGrayScale(cur.projectedSrc));
        H.ip->imgs.push_back(cur);
        P->nfeat.push_back((int)cur.features.size());
    }
    double t1 = now_s();
    H.ip->matching();
    double t2 = now_s();
    P->t_features = t1 - t0;
    P->t_matching = t2 - t1;
    P->result = H.ip->result;
    P->log = q.sink.str();
    return P;
}

// The real constructor, from <dir>/<i>.bmp, i = 1..n (main.cpp:9)
RefPano *ref_stitch_dir(const char *dir, int n) {
    RefPano *P = new RefPano();
    SilenceStdout q;
    double t0 = now_s();
    {
        ImageProcess ip(std::string(dir), n);
        P->result = ip.result;
        for (size_t i = 0; i < ip.imgs.size(); ++i) P->nfeat.push_back((int)ip.imgs[i].features.size());
    }
    P->t_features = 0;
    P->t_matching = now_s() - t0;
    P->log = q.sink.str();
    return P;
}
void ref_pano_info(RefPano *P, int *w, int *h, double *t_features, double *t_matching) {
    *w = P->result.width(); *h = P->result.height();
    *t_features = P->t_features; *t_matching = P->t_matching;
}
void ref_pano_copy(RefPano *P, uint8_t *dst) { std::memcpy(dst, P->result.data(), P->result.size()); }
int ref_pano_nfeat(RefPano *P, int i) { return i < (int)P->nfeat.size() ? P->nfeat[i] : -1; }
int ref_pano_log(RefPano *P, char *dst, int cap) {
    int n = (int)P->log.size();
    if (n >= cap) n = cap - 1;
    std::memcpy(dst, P->log.data(), n);
    dst[n] = 0;
    return n;
}
void ref_pano_free(RefPano *P) { delete P; }

// CImg BMP load (CImg.h:48395) -> planar u8; returns 0 and fills w,h; dst may be NULL to query the size.
int ref_load_bmp(const char *path, int *w, int *h, uint8_t *dst) {
    CImg<unsigned char> img(path);
    *w = img.width(); *h = img.height();
    if (img.spectrum() != 3) return -1;
    if (dst) std::memcpy(dst, img.data(), img.size());
    return 0;
}
int ref_save_bmp(const char *path, const uint8_t *src, int w, int h) {
    wrap_u8(src, w, h, 3).save(path);
    return 0;
}

} // extern "C"
