// oracle/ref_ex6_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI harness over the reference's SECOND caller of the hot path, src/ex6 (fixed chain order, min(w,h) level
// count, 3-channel emptiness test, Deriche pyramid blur, 5/6 : 1/6 final mix).  Compiled by
// oracle/Makefile against the sources where they lie under /root/reference/src/ex6 into oracle/_ref/libpano_ref_ex6.so
// (a separate library: both variants define `class ImageProcess`).  No arithmetic is restated here; every entry point
// calls the reference's own method named in its comment.  Determinism adjustments: see ref_ex6_shim.h.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <chrono>
#include <iostream>
#include <sstream>
#include <new>

#define private public
#include "ImageProcess.h"
#undef private

static unsigned g_seed = 666666u;
extern "C" unsigned pano_ex6_seed(void) { return g_seed; }

namespace {

// An ImageProcess whose constructor has NOT run (the only constructor does the whole job from files).
struct Holder {
    alignas(ImageProcess) unsigned char buf[sizeof(ImageProcess)];
    ImageProcess *ip;
    Holder() {
        std::memset(buf, 0, sizeof buf);
        ip = reinterpret_cast<ImageProcess *>(buf);
        ip->imgs = 0;
        ip->picSum = 0;
        new (&ip->result) CImg<unsigned char>();
        new (&ip->YCbCrResult) CImg<float>();
        new (&ip->balanced) CImg<unsigned char>();
        new (&ip->YCbCrBalanced) CImg<float>();
        new (&ip->forward_H) Homography();
        new (&ip->backward_H) Homography();
    }
    void set_images(int n) {
        ip->imgs = new Image[n];
        ip->picSum = n;
    }
    ~Holder() {
        delete[] ip->imgs;
        ip->YCbCrBalanced.~CImg<float>();
        ip->balanced.~CImg<unsigned char>();
        ip->YCbCrResult.~CImg<float>();
        ip->result.~CImg<unsigned char>();
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

struct Ex6Pano {
    CImg<unsigned char> result;
    double t_features, t_matching;
    std::vector<int> nfeat;
    std::string log;
};

} // namespace

extern "C" {

// the value `time(0)` yields inside RANSAC (src/ex6/ImageProcess.cpp:403)
void ex6_set_seed(unsigned seed) { g_seed = seed; }

// ImageProcess::blend (src/ex6/ImageProcess.cpp:638-742)
int ex6_blend(const uint8_t *a, const uint8_t *b, int w, int h, uint8_t *out) {
    Holder H;
    CImg<unsigned char> r = H.ip->blend(wrap_u8(a, w, h, 3), wrap_u8(b, w, h, 3));
    if (r.width() != w || r.height() != h || r.spectrum() != 3) return -1;
    std::memcpy(out, r.data(), (size_t)w * h * 3);
    return 0;
}

// CImg<float>::get_blur(2) as called by blend (src/ex6/ImageProcess.cpp:702-705): boundary_conditions = true,
// is_gaussian = false -> Deriche order 0 along x then y (CImg.h:35111-35147, 34777-34869)
int ex6_cimg_blur2(const float *src, int w, int h, int c, float *dst) {
    CImg<float> s(w, h, 1, c);
    std::memcpy(s.data(), src, (size_t)w * h * c * 4);
    CImg<float> r = s.get_blur(2);
    std::memcpy(dst, r.data(), (size_t)w * h * c * 4);
    return 0;
}

// ImageProcess::RANSAC (src/ex6/ImageProcess.cpp:401-445) with the pinned seed
int ex6_ransac(const VlSiftKeypoint *src, const VlSiftKeypoint *dst, int n, double *H8) {
    Holder H;
    std::vector<ImgPair> pairs;
    for (int i = 0; i < n; ++i) pairs.push_back(ImgPair(src[i], dst[i]));
    Homography R = H.ip->RANSAC(pairs);
    H8[0] = R.H[0][0]; H8[1] = R.H[0][1]; H8[2] = R.H[0][2]; H8[3] = R.H[1][0];
    H8[4] = R.H[1][1]; H8[5] = R.H[1][2]; H8[6] = R.H[2][0]; H8[7] = R.H[2][1];
    return 0;
}

// The tail of ImageProcess::matching (src/ex6/ImageProcess.cpp:261-278, 283-317): equalization(balanced, 1) and the
// 5/6 : 1/6 luminance mix.  The code is inline in matching(); with two images the fixed chain has no edge to stitch
// (nextIndex[1] is empty, start index = 1), so matching() runs the tail on result = imgs[1].projectedSrc.
int ex6_tail(const uint8_t *img, int w, int h, uint8_t *out) {
    Holder H;
    SilenceStdout q;
    H.set_images(2);
    H.ip->imgs[1].projectedSrc = wrap_u8(img, w, h, 3);
    H.ip->matching();
    std::memcpy(out, H.ip->result.data(), (size_t)w * h * 3);
    return 0;
}

// The ImageProcess constructor body (src/ex6/ImageProcess.cpp:4-17, readFile_Single :27-42) on in-memory planar RGB
// images; returns NULL where the reference calls exit(1) (projected width > height, :35-38).
Ex6Pano *ex6_stitch_mem(const uint8_t *const *imgs, const int *w, const int *h, int n) {
    Ex6Pano *P = new Ex6Pano();
    Holder H;
    SilenceStdout q;
    double t0 = now_s();
    H.set_images(n);
    for (int i = 0; i < n; ++i) {
        Image cur;
        cur.projectedSrc = Projection::imageProjection(wrap_u8(imgs[i], w[i], h[i], 3));
        if (cur.projectedSrc.width() > cur.projectedSrc.height()) { delete P; return 0; }
        cur.features = ImageProcess::siftAlgorithm(ImageProcess::toGrayScale(cur.projectedSrc));
        H.ip->imgs[i] = cur;
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

// The real constructor, from <dir>/<i>.bmp, i = 1..n (src/ex6/main.cpp); also writes <dir>result.bmp (:15-16)
Ex6Pano *ex6_stitch_dir(const char *dir, int n) {
    Ex6Pano *P = new Ex6Pano();
    SilenceStdout q;
    double t0 = now_s();
    {
        ImageProcess ip(std::string(dir), n);
        P->result = ip.result;
        for (int i = 0; i < n; ++i) P->nfeat.push_back((int)ip.imgs[i].features.size());
    }
    P->t_features = 0;
    P->t_matching = now_s() - t0;
    P->log = q.sink.str();
    return P;
}
void ex6_pano_info(Ex6Pano *P, int *w, int *h, double *t_features, double *t_matching) {
    *w = P->result.width(); *h = P->result.height();
    *t_features = P->t_features; *t_matching = P->t_matching;
}
void ex6_pano_copy(Ex6Pano *P, uint8_t *dst) { std::memcpy(dst, P->result.data(), P->result.size()); }
int ex6_pano_nfeat(Ex6Pano *P, int i) { return i < (int)P->nfeat.size() ? P->nfeat[i] : -1; }
int ex6_pano_log(Ex6Pano *P, char *dst, int cap) {
    int n = (int)P->log.size();
    if (n >= cap) n = cap - 1;
    std::memcpy(dst, P->log.data(), n);
    dst[n] = 0;
    return n;
}
void ex6_pano_free(Ex6Pano *P) { delete P; }

} // extern "C"
