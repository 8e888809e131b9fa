// vl_sift_shim.cu -- the vl_sift_* C surface (include/vl_b200/sift.h) on top of the SIFT engine.
//
// VLFeat's API is chatty: one call per keypoint for orientations, one per (keypoint, angle) for descriptors
// (ImageProcess.cpp:66-86).  One kernel launch per call would be launch-bound, so vl_sift_detect runs the whole
// octave eagerly -- extrema, refinement, gradient map, every orientation and every descriptor -- and the per-keypoint
// calls become host look-ups keyed by the keypoint VALUE (the caller passes a modified copy, ImageProcess.cpp:67,84).
// A keypoint or angle that was not produced by vl_sift_detect falls back to a single-item kernel launch.
#include "../../include/vl_b200/sift.h"
#include "sift_engine.h"
#include <map>
#include <memory>
#include <cmath>
#include <tuple>
#include <cstring>
#include <cstdlib>

using namespace pb;

namespace {

int g_device = 0;

struct Impl {
    cudaStream_t st = nullptr;
    std::unique_ptr<SiftEngine> eng;
    int device = 0;
    int oi = 0;  // index of the current octave (0-based from o_min)
    bool mirror = false;
    bool configured = false;
    DevBuf<float> d_img, d_dog;
    std::vector<float> h_octave, h_dog, h_grad;
    // eager per-octave cache
    typedef std::tuple<uint32_t, uint32_t, uint32_t, int> KeyId;  // bits of x, y, sigma; is
    std::map<KeyId, int> index;
    std::vector<int> nangles;
    std::vector<double> angles;      // [nkeys][4]
    std::vector<int> desc_first;     // [nkeys] first descriptor row of the key
    std::vector<float> descr;        // [ndesc][128]
    std::vector<int> written;        // [ndesc]
};

uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
Impl::KeyId key_id(const VlSiftKeypoint* k) { return Impl::KeyId(fbits(k->x), fbits(k->y), fbits(k->sigma), k->is); }

SiftParams params_of(const VlSiftFilt* f) {
    SiftParams p;
    p.O = f->O; p.S = f->S; p.o_min = f->o_min;
    p.peak_thresh = f->peak_thresh; p.edge_thresh = f->edge_thresh; p.norm_thresh = f->norm_thresh;
    p.magnif = f->magnif; p.window_size = f->windowSize;
    return p;
}

void refresh_mirror(VlSiftFilt* f, Impl* I, bool with_dog_grad) {
    if (!I->mirror) return;
    OctaveBuf& ob = I->eng->octave(I->oi);
    const int nl = I->eng->nlevels();
    const size_t plane = (size_t)ob.w * ob.h;
    I->h_octave.resize(plane * nl);
    PB_CUDA(cudaMemcpy2DAsync(I->h_octave.data(), (size_t)ob.w * 4, ob.gss.p, (size_t)ob.pitch * 4, (size_t)ob.w * 4,
                              (size_t)ob.h * nl, cudaMemcpyDeviceToHost, I->st));
    if (with_dog_grad) {
        I->d_dog.ensure((size_t)ob.pitch * ob.h * (nl - 1));
        launch_dog(ob.view(nl), I->d_dog.p, I->st);
        I->h_dog.resize(plane * (nl - 1));
        PB_CUDA(cudaMemcpy2DAsync(I->h_dog.data(), (size_t)ob.w * 4, I->d_dog.p, (size_t)ob.pitch * 4, (size_t)ob.w * 4,
                                  (size_t)ob.h * (nl - 1), cudaMemcpyDeviceToHost, I->st));
        I->h_grad.resize(plane * 2 * (nl - 3));
        PB_CUDA(cudaMemcpy2DAsync(I->h_grad.data(), (size_t)ob.w * 8, ob.grad.p, (size_t)ob.pitch * 8, (size_t)ob.w * 8,
                                  (size_t)ob.h * (nl - 3), cudaMemcpyDeviceToHost, I->st));
    }
    PB_CUDA(cudaStreamSynchronize(I->st));
    f->octave = I->h_octave.data();
    if (with_dog_grad) { f->dog = I->h_dog.data(); f->grad = I->h_grad.data(); }
}

void point_to_device(VlSiftFilt* f, Impl* I) {
    if (I->mirror) return;
    OctaveBuf& ob = I->eng->octave(I->oi);
    f->octave = ob.gss.p;
    f->grad = ob.grad.p;
    f->temp = I->eng->temp().p;
    f->dog = nullptr;  // DoG is never materialised on the fast path (formed on the fly by the detector)
}

void clear_cache(Impl* I) {
    I->index.clear(); I->nangles.clear(); I->angles.clear(); I->desc_first.clear(); I->descr.clear(); I->written.clear();
}

}  // namespace

extern "C" {

void vl_b200_sift_set_device(int device) { g_device = device; }

VlSiftFilt* vl_sift_new(int width, int height, int noctaves, int nlevels, int o_min) {
    VlSiftFilt* f = (VlSiftFilt*)calloc(1, sizeof(VlSiftFilt));
    if (!f) return nullptr;
    // vl/sift.c:231-271
    if (noctaves < 0) {
        double l2 = log((double)(width < height ? width : height)) / 0.693147180559945;
        double v = floor(l2) - o_min - 3;
        noctaves = (int)(v > 1 ? v : 1);
    }
    f->width = width; f->height = height; f->O = noctaves; f->S = nlevels; f->o_min = o_min;
    f->s_min = -1; f->s_max = nlevels + 1; f->o_cur = o_min;
    f->sigman = 0.5;
    f->sigmak = pow(2.0, 1.0 / nlevels);
    f->sigma0 = 1.6 * f->sigmak;
    f->dsigma0 = f->sigma0 * sqrt(1.0 - 1.0 / (f->sigmak * f->sigmak));
    f->peak_thresh = 0.0; f->edge_thresh = 10.0; f->norm_thresh = 0.0; f->magnif = 3.0; f->windowSize = 2.0;
    f->grad_o = o_min - 1;
    Impl* I = new Impl();
    I->device = g_device;
    try {
        PB_CUDA(cudaSetDevice(I->device));
        PB_CUDA(cudaStreamCreateWithFlags(&I->st, cudaStreamNonBlocking));
        I->eng.reset(new SiftEngine(I->st));
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_sift_new (B200): %s\n", e.what());
        delete I;
        free(f);
        return nullptr;
    }
    f->b200_impl = I;
    return f;
}

void vl_sift_delete(VlSiftFilt* f) {
    if (!f) return;
    Impl* I = (Impl*)f->b200_impl;
    if (I) {
        cudaSetDevice(I->device);
        I->eng.reset();
        I->d_img.release();
        I->d_dog.release();
        if (I->st) cudaStreamDestroy(I->st);
        delete I;
    }
    if (f->keys) free(f->keys);
    free(f);
}

void vl_b200_sift_set_mirror(VlSiftFilt* f, int on) { ((Impl*)f->b200_impl)->mirror = on != 0; }

int vl_sift_process_first_octave(VlSiftFilt* f, vl_sift_pix const* im) {
    Impl* I = (Impl*)f->b200_impl;
    try {
        PB_CUDA(cudaSetDevice(I->device));
        f->o_cur = f->o_min;
        f->nkeys = 0;
        f->octave_width = f->o_min >= 0 ? f->width >> f->o_min : f->width << -f->o_min;
        f->octave_height = f->o_min >= 0 ? f->height >> f->o_min : f->height << -f->o_min;
        if (f->O == 0) return VL_ERR_EOF;
        if (f->o_min != 0) return VL_ERR_BAD_ARG;
        I->eng->configure(f->width, f->height, params_of(f));
        I->configured = true;
        I->oi = 0;
        clear_cache(I);
        const int pitch = align_up(f->width, 32);
        I->d_img.ensure((size_t)pitch * f->height);
        PB_CUDA(cudaMemcpy2DAsync(I->d_img.p, (size_t)pitch * 4, im, (size_t)f->width * 4, (size_t)f->width * 4,
                                  f->height, cudaMemcpyHostToDevice, I->st));
        I->eng->load_base_from_device(I->d_img.p, pitch);
        I->eng->build_octave(0);
        PB_CUDA(cudaStreamSynchronize(I->st));  // the caller may free `im` right away (ImageProcess.cpp:96)
        point_to_device(f, I);
        refresh_mirror(f, I, false);
        return VL_ERR_OK;
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_sift_process_first_octave (B200): %s\n", e.what());
        return VL_ERR_BAD_ARG;
    }
}

int vl_sift_process_next_octave(VlSiftFilt* f) {
    Impl* I = (Impl*)f->b200_impl;
    if (f->o_cur == f->o_min + f->O - 1) return VL_ERR_EOF;
    try {
        PB_CUDA(cudaSetDevice(I->device));
        I->oi += 1;
        f->o_cur += 1;
        f->nkeys = 0;
        f->octave_width = f->width >> f->o_cur;
        f->octave_height = f->height >> f->o_cur;
        clear_cache(I);
        I->eng->build_octave(I->oi);  // its base was written by the previous octave's blur of level s_best
        point_to_device(f, I);
        refresh_mirror(f, I, false);
        return VL_ERR_OK;
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_sift_process_next_octave (B200): %s\n", e.what());
        return VL_ERR_BAD_ARG;
    }
}

void vl_sift_detect(VlSiftFilt* f) {
    Impl* I = (Impl*)f->b200_impl;
    f->nkeys = 0;
    if (!I->configured) return;
    try {
        PB_CUDA(cudaSetDevice(I->device));
        clear_cache(I);
        // thresholds may have been changed through the setters since vl_sift_new
        SiftParams p = params_of(f);
        I->eng->set_thresholds(p);
        I->eng->detect_octave(I->oi);
        OctaveBuf& ob = I->eng->octave(I->oi);
        const int n = (int)ob.keys.size();
        if (n > f->keys_res) {
            f->keys_res = n + 500;  // vl/sift.c:580-590 grows by 500
            f->keys = (VlSiftKeypoint*)realloc(f->keys, (size_t)f->keys_res * sizeof(VlSiftKeypoint));
        }
        if (n) memcpy(f->keys, ob.keys.data(), (size_t)n * sizeof(VlSiftKeypoint));
        f->nkeys = n;
        f->grad_o = f->o_cur;
        // eager orientations + descriptors
        I->eng->orient_octave(I->oi);
        I->nangles = ob.h_nangles;
        I->angles = ob.h_angles;
        I->desc_first.assign(n, 0);
        std::vector<int> jk;
        std::vector<double> ja;
        for (int i = 0; i < n; ++i) {
            I->desc_first[i] = (int)jk.size();
            I->index[key_id(&f->keys[i])] = i;
            for (int j = 0; j < I->nangles[i]; ++j) { jk.push_back(i); ja.push_back(I->angles[(size_t)i * 4 + j]); }
        }
        I->descr.assign(jk.size() * 128, 0.f);
        I->written.assign(jk.size(), 0);
        if (!jk.empty()) I->eng->describe_octave(I->oi, jk, ja, nullptr, I->descr.data(), I->written.data());
        refresh_mirror(f, I, true);
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_sift_detect (B200): %s\n", e.what());
        f->nkeys = 0;
    }
}

int vl_sift_calc_keypoint_orientations(VlSiftFilt* f, double angles[4], VlSiftKeypoint const* k) {
    Impl* I = (Impl*)f->b200_impl;
    if (!I->configured || k->o != f->o_cur) return 0;  // vl/sift.c:936-937
    auto it = I->index.find(key_id(k));
    if (it != I->index.end()) {
        const int i = it->second, n = I->nangles[i];
        for (int j = 0; j < n; ++j) angles[j] = I->angles[(size_t)i * 4 + j];
        return n;
    }
    try {
        PB_CUDA(cudaSetDevice(I->device));
        std::vector<KeyIn> ki{KeyIn{k->x, k->y, k->sigma, (short)k->is, 0}};
        std::vector<int> na;
        std::vector<double> an;
        I->eng->orient_custom(I->oi, ki, na, an);
        for (int j = 0; j < na[0]; ++j) angles[j] = an[j];
        return na[0];
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_sift_calc_keypoint_orientations (B200): %s\n", e.what());
        return 0;
    }
}

void vl_sift_calc_keypoint_descriptor(VlSiftFilt* f, vl_sift_pix* descr, VlSiftKeypoint const* k, double angle) {
    Impl* I = (Impl*)f->b200_impl;
    if (!I->configured || k->o != f->o_cur) return;  // vl/sift.c:1321
    auto it = I->index.find(key_id(k));
    if (it != I->index.end()) {
        const int i = it->second;
        for (int j = 0; j < I->nangles[i]; ++j) {
            double a = I->angles[(size_t)i * 4 + j];
            if (memcmp(&a, &angle, sizeof a) == 0) {
                const int row = I->desc_first[i] + j;
                if (I->written[row]) memcpy(descr, &I->descr[(size_t)row * 128], 128 * sizeof(float));
                return;
            }
        }
    }
    try {
        PB_CUDA(cudaSetDevice(I->device));
        std::vector<KeyIn> ki{KeyIn{k->x, k->y, k->sigma, (short)k->is, 0}};
        std::vector<int> jk{0};
        std::vector<double> ja{angle};
        float d[128];
        int w = 0;
        I->eng->describe_octave(I->oi, jk, ja, &ki, d, &w);
        if (w) memcpy(descr, d, sizeof d);
    } catch (const std::exception& e) {
        fprintf(stderr, "vl_sift_calc_keypoint_descriptor (B200): %s\n", e.what());
    }
}

}  // extern "C"
