// sift_engine.cu -- see sift_engine.h
#include "sift_engine.h"
#include "host_numerics.h"
#include <algorithm>
#include <cmath>

namespace pb {

SiftEngine::SiftEngine(cudaStream_t st) : st_(st) {}
SiftEngine::~SiftEngine() {}

SiftConsts SiftEngine::consts() const {
    return SiftConsts{s_min_, s_max_, p_.S, p_.peak_thresh, p_.edge_thresh, p_.norm_thresh, p_.magnif, p_.window_size};
}

// vl/sift.c:125-141 -- evaluated on the host so that exp() is glibc's, as in the reference.
BlurTaps SiftEngine::make_taps(double sigma) {
    BlurTaps t;
    memset(&t, 0, sizeof t);
    t.W = hostnum::gaussian_taps(sigma, t.c, kMaxBlurW);
    return t;
}

double SiftEngine::presmooth_sigma_first() const {  // vl/sift.c:389-395
    double sa = sigma0_ * pow(sigmak_, s_min_);
    double sb = sigman_ * pow(2.0, -p_.o_min);
    return (sa > sb) ? sqrt(sa * sa - sb * sb) : 0.0;
}
double SiftEngine::presmooth_sigma_next() const {  // vl/sift.c:465-471 (powf, as in the reference)
    int s_best = std::min(s_min_ + p_.S, s_max_);
    double sa = sigma0_ * powf((float)sigmak_, (float)s_min_);
    double sb = sigma0_ * powf((float)sigmak_, (float)(s_best - p_.S));
    return (sa > sb) ? sqrt(sa * sa - sb * sb) : 0.0;
}
double SiftEngine::level_sigma(int s) const { return dsigma0_ * pow(sigmak_, s); }  // vl/sift.c:402, 478

void SiftEngine::configure(int w, int h, const SiftParams& p) {
    p_ = p;
    w_ = w;
    h_ = h;
    if (p.o_min < 0) throw std::runtime_error("o_min < 0 (up-sampled first octave) is not on the accelerated path");
    int O = p.O;
    if (O < 0) {  // vl/sift.c:231-233
        double l2 = log((double)std::min(w, h)) / 0.693147180559945;
        double v = floor(l2) - p.o_min - 3;
        O = (int)(v > 1 ? v : 1);
    }
    s_min_ = -1;
    s_max_ = p.S + 1;
    nlev_ = s_max_ - s_min_ + 1;
    sigman_ = 0.5;
    sigmak_ = pow(2.0, 1.0 / p.S);
    sigma0_ = 1.6 * sigmak_;
    dsigma0_ = sigma0_ * sqrt(1.0 - 1.0 / (sigmak_ * sigmak_));
    if ((int)oct_.size() != O) {
        oct_.clear();
        oct_.resize(O);
    }
    for (int i = 0; i < O; ++i) {
        OctaveBuf& ob = oct_[i];
        int o = p.o_min + i;
        ob.w = w >> o;
        ob.h = h >> o;
        ob.pitch = align_up(std::max(ob.w, 1), 32);
        size_t plane = (size_t)ob.pitch * std::max(ob.h, 1);
        ob.gss.ensure(plane * nlev_);
        ob.grad.ensure(plane * 2 * std::max(nlev_ - 3, 1));
        int cap = (int)std::max<size_t>(4096, (size_t)ob.w * ob.h / 16);
        if (cap > ob.cand_cap) ob.cand_cap = cap;
        ob.cand.ensure(ob.cand_cap);
        ob.refined.ensure(ob.cand_cap);
        ob.keys.clear();
    }
    temp_.ensure((size_t)oct_.empty() ? 1 : (size_t)oct_[0].pitch * std::max(oct_[0].h, 1));
    counts_.ensure(std::max(O, 1));
    h_counts_.ensure(std::max(O, 1));
    if (!tab_ready_) {  // vl/sift.c:56-63
        double tab[257];
        hostnum::expn_table(tab);
        expn_tab_.ensure(257);
        PB_CUDA(cudaMemcpyAsync(expn_tab_.p, tab, sizeof tab, cudaMemcpyHostToDevice, st_));
        PB_CUDA(cudaStreamSynchronize(st_));
        tab_ready_ = true;
    }
}

void SiftEngine::load_base_from_device(const float* d_img, int img_pitch) {
    OctaveBuf& ob = oct_[0];
    if (p_.o_min == 0) {
        launch_copy_f32(d_img, img_pitch, ob.gss.p, w_, h_, ob.pitch, st_);
    } else {  // vl/sift.c:379-381: every 2^o_min-th sample
        throw std::runtime_error("o_min > 0 not implemented");
    }
}

void SiftEngine::blur_level(int oi, int src_l, int dst_l, double sigma, bool seed_next) {
    OctaveBuf& ob = oct_[oi];
    if (ob.w < 1 || ob.h < 1) return;
    size_t plane = (size_t)ob.pitch * ob.h;
    BlurTaps taps = make_taps(sigma);
    float* ds = nullptr;
    int dsp = 0;
    if (seed_next && oi + 1 < (int)oct_.size() && oct_[oi + 1].w >= 1 && oct_[oi + 1].h >= 1) {
        ds = oct_[oi + 1].gss.p;
        dsp = oct_[oi + 1].pitch;
    }
    launch_blur(ob.gss.p + plane * src_l, temp_.p, ob.gss.p + plane * dst_l, ob.w, ob.h, ob.pitch, taps, ds, dsp, st_);
}

void SiftEngine::build_octave(int oi) {
    double pre = (oi == 0) ? presmooth_sigma_first() : presmooth_sigma_next();
    if (pre > 0) blur_level(oi, 0, 0, pre, false);
    int s_best = std::min(s_min_ + p_.S, s_max_);
    for (int s = s_min_ + 1; s <= s_max_; ++s)
        blur_level(oi, s - 1 - s_min_, s - s_min_, level_sigma(s), s == s_best);
}

void SiftEngine::seed_next_octave(int oi) {
    // fused into build_octave (the horizontal blur pass of level s_best also writes the 2:1 sub-sampled base)
    (void)oi;
}

static void finish_keys(const RefinedKey* r, int n, int o, double sigma0, int S, std::vector<VlKey>& keys) {
    // the reference refines candidates in detection (raster) order and compacts the good ones in place
    std::vector<int> idx;
    idx.reserve(n);
    for (int i = 0; i < n; ++i)
        if (r[i].good) idx.push_back(i);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) {
        const RefinedKey &A = r[a], &B = r[b];
        if (A.is0 != B.is0) return A.is0 < B.is0;
        if (A.iy0 != B.iy0) return A.iy0 < B.iy0;
        return A.ix0 < B.ix0;
    });
    double xper = pow(2.0, o);
    keys.resize(idx.size());
    for (size_t k = 0; k < idx.size(); ++k) {
        const RefinedKey& a = r[idx[k]];
        VlKey& q = keys[k];
        q.o = o; q.ix = a.ix; q.iy = a.iy; q.is = a.is;
        q.s = a.s; q.x = a.x; q.y = a.y;
        q.sigma = (float)(sigma0 * pow(2.0, a.sn / S) * xper);  // vl/sift.c:764
    }
}

void SiftEngine::detect_octave(int oi) {
    OctaveBuf& ob = oct_[oi];
    ob.keys.clear();
    ob.h_nangles.clear();
    ob.h_angles.clear();
    if (ob.w < 2 || ob.h < 2) return;
    OctaveView ov = ob.view(nlev_);
    SiftConsts sc = consts();
    int o = p_.o_min + oi;
    double xper = pow(2.0, o);
    launch_gradient(ov, sc, ob.grad.p, st_);
    int n = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        PB_CUDA(cudaMemsetAsync(counts_.p + oi, 0, sizeof(int), st_));
        launch_detect(ov, sc, ob.cand.p, counts_.p + oi, ob.cand_cap, st_);
        launch_refine(ov, sc, ob.cand.p, counts_.p + oi, ob.cand_cap, ob.refined.p, xper, st_);
        PB_CUDA(cudaMemcpyAsync(h_counts_.p + oi, counts_.p + oi, sizeof(int), cudaMemcpyDeviceToHost, st_));
        PB_CUDA(cudaStreamSynchronize(st_));
        n = h_counts_.p[oi];
        if (n <= ob.cand_cap) break;
        ob.cand_cap = n + n / 8;
        ob.cand.ensure(ob.cand_cap);
        ob.refined.ensure(ob.cand_cap);
    }
    if (n == 0) return;
    RefinedKey* hr = (RefinedKey*)h_stage_.ensure((size_t)n * sizeof(RefinedKey));
    PB_CUDA(cudaMemcpyAsync(hr, ob.refined.p, (size_t)n * sizeof(RefinedKey), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    finish_keys(hr, n, o, sigma0_, p_.S, ob.keys);
}

static void upload_keyin(OctaveBuf& ob, const std::vector<KeyIn>& ki, cudaStream_t st) {
    ob.keyin.ensure(ki.size());
    PB_CUDA(cudaMemcpyAsync(ob.keyin.p, ki.data(), ki.size() * sizeof(KeyIn), cudaMemcpyHostToDevice, st));
}

void SiftEngine::orient_custom(int oi, const std::vector<KeyIn>& keys, std::vector<int>& nang,
                               std::vector<double>& ang) {
    OctaveBuf& ob = oct_[oi];
    int n = (int)keys.size();
    nang.assign(n, 0);
    ang.assign((size_t)n * 4, 0.0);
    if (n == 0 || ob.w < 2 || ob.h < 2) return;
    upload_keyin(ob, keys, st_);
    ob.nangles.ensure(n);
    ob.angles.ensure((size_t)n * 4);
    int o = p_.o_min + oi;
    launch_orient(ob.view(nlev_), consts(), expn_tab_.p, o, ob.keyin.p, n, pow(2.0, o), ob.nangles.p, ob.angles.p, st_);
    PB_CUDA(cudaMemcpyAsync(nang.data(), ob.nangles.p, n * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaMemcpyAsync(ang.data(), ob.angles.p, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

void SiftEngine::orient_octave(int oi) {
    OctaveBuf& ob = oct_[oi];
    std::vector<KeyIn> ki(ob.keys.size());
    for (size_t i = 0; i < ki.size(); ++i) ki[i] = KeyIn{ob.keys[i].x, ob.keys[i].y, ob.keys[i].sigma, ob.keys[i].is};
    orient_custom(oi, ki, ob.h_nangles, ob.h_angles);
}

void SiftEngine::describe_octave(int oi, const std::vector<int>& key_idx, const std::vector<double>& ang,
                                 const std::vector<KeyIn>* custom_keys, float* out_descr, int* out_written) {
    OctaveBuf& ob = oct_[oi];
    int nj = (int)key_idx.size();
    if (nj == 0) return;
    if (ob.w < 2 || ob.h < 2) {
        for (int i = 0; i < nj; ++i) out_written[i] = 0;
        return;
    }
    if (custom_keys) upload_keyin(ob, *custom_keys, st_);
    std::vector<DescJob> jobs(nj);
    for (int i = 0; i < nj; ++i) {
        jobs[i].key = key_idx[i];
        jobs[i].pad = 0;
        jobs[i].angle = ang[i];
        jobs[i].st0 = sin(ang[i]);  // vl/sift.c:1308-1309, glibc
        jobs[i].ct0 = cos(ang[i]);
    }
    ob.jobs.ensure(nj);
    ob.descr.ensure((size_t)nj * 128);
    ob.written.ensure(nj);
    PB_CUDA(cudaMemcpyAsync(ob.jobs.p, jobs.data(), nj * sizeof(DescJob), cudaMemcpyHostToDevice, st_));
    int o = p_.o_min + oi;
    launch_descr(ob.view(nlev_), consts(), expn_tab_.p, o, ob.keyin.p, ob.jobs.p, nj, pow(2.0, o), ob.descr.p,
                 ob.written.p, st_);
    PB_CUDA(cudaMemcpyAsync(out_descr, ob.descr.p, (size_t)nj * 128 * sizeof(float), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaMemcpyAsync(out_written, ob.written.p, nj * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

void SiftEngine::extract(const float* d_img, int img_pitch, RawFeatures& out) {
    out = RawFeatures();
    const int O = (int)oct_.size();
    if (O == 0) return;
    SiftConsts sc = consts();
    // 1. whole pyramid + detection + refinement + gradient maps, no host round trip
    load_base_from_device(d_img, img_pitch);
    PB_CUDA(cudaMemsetAsync(counts_.p, 0, sizeof(int) * O, st_));
    for (int oi = 0; oi < O; ++oi) {
        OctaveBuf& ob = oct_[oi];
        ob.keys.clear();
        if (ob.w < 2 || ob.h < 2) continue;
        build_octave(oi);
        OctaveView ov = ob.view(nlev_);
        double xper = pow(2.0, p_.o_min + oi);
        launch_detect(ov, sc, ob.cand.p, counts_.p + oi, ob.cand_cap, st_);
        launch_refine(ov, sc, ob.cand.p, counts_.p + oi, ob.cand_cap, ob.refined.p, xper, st_);
        launch_gradient(ov, sc, ob.grad.p, st_);
    }
    PB_CUDA(cudaMemcpyAsync(h_counts_.p, counts_.p, sizeof(int) * O, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    // 2. refined candidates -> host (sigma needs pow), rare overflow handled by the octave-at-a-time path
    std::vector<int> cnt(O);
    size_t total = 0;
    for (int oi = 0; oi < O; ++oi) {
        cnt[oi] = h_counts_.p[oi];
        if (cnt[oi] > oct_[oi].cand_cap) {
            detect_octave(oi);  // grows the buffers and re-runs this octave synchronously
            cnt[oi] = -1;
        } else
            total += cnt[oi];
    }
    RefinedKey* hr = (RefinedKey*)h_stage_.ensure(std::max<size_t>(total, 1) * sizeof(RefinedKey));
    size_t off = 0;
    for (int oi = 0; oi < O; ++oi)
        if (cnt[oi] > 0) {
            PB_CUDA(cudaMemcpyAsync(hr + off, oct_[oi].refined.p, (size_t)cnt[oi] * sizeof(RefinedKey),
                                    cudaMemcpyDeviceToHost, st_));
            off += cnt[oi];
        }
    PB_CUDA(cudaStreamSynchronize(st_));
    off = 0;
    for (int oi = 0; oi < O; ++oi)
        if (cnt[oi] > 0) {
            finish_keys(hr + off, cnt[oi], p_.o_min + oi, sigma0_, p_.S, oct_[oi].keys);
            off += cnt[oi];
        }
    // 3. orientations of every octave, one round trip
    size_t nk_total = 0;
    for (int oi = 0; oi < O; ++oi) {
        OctaveBuf& ob = oct_[oi];
        int n = (int)ob.keys.size();
        out.noct_keys.push_back(n);
        ob.h_nangles.assign(n, 0);
        ob.h_angles.assign((size_t)n * 4, 0.0);
        if (n == 0) continue;
        nk_total += n;
        std::vector<KeyIn> ki(n);
        for (int i = 0; i < n; ++i) ki[i] = KeyIn{ob.keys[i].x, ob.keys[i].y, ob.keys[i].sigma, ob.keys[i].is};
        upload_keyin(ob, ki, st_);
        PB_CUDA(cudaStreamSynchronize(st_));  // ki is a pageable temporary
        ob.nangles.ensure(n);
        ob.angles.ensure((size_t)n * 4);
        int o = p_.o_min + oi;
        launch_orient(ob.view(nlev_), sc, expn_tab_.p, o, ob.keyin.p, n, pow(2.0, o), ob.nangles.p, ob.angles.p, st_);
        PB_CUDA(cudaMemcpyAsync(ob.h_nangles.data(), ob.nangles.p, n * sizeof(int), cudaMemcpyDeviceToHost, st_));
        PB_CUDA(cudaMemcpyAsync(ob.h_angles.data(), ob.angles.p, (size_t)n * 4 * sizeof(double),
                                cudaMemcpyDeviceToHost, st_));
    }
    PB_CUDA(cudaStreamSynchronize(st_));
    if (nk_total == 0) return;
    // 4. descriptors of every octave, one round trip (sin/cos on the host)
    std::vector<std::vector<DescJob>> jobs(O);
    size_t nj_total = 0;
    for (int oi = 0; oi < O; ++oi) {
        OctaveBuf& ob = oct_[oi];
        for (int i = 0; i < (int)ob.keys.size(); ++i)
            for (int j = 0; j < ob.h_nangles[i]; ++j) {
                double a = ob.h_angles[(size_t)i * 4 + j];
                jobs[oi].push_back(DescJob{i, 0, a, sin(a), cos(a)});
            }
        nj_total += jobs[oi].size();
    }
    std::vector<float> hd(nj_total * 128);
    std::vector<int> hw(nj_total);
    off = 0;
    for (int oi = 0; oi < O; ++oi) {
        OctaveBuf& ob = oct_[oi];
        int nj = (int)jobs[oi].size();
        if (nj == 0) continue;
        ob.jobs.ensure(nj);
        ob.descr.ensure((size_t)nj * 128);
        ob.written.ensure(nj);
        PB_CUDA(cudaMemcpyAsync(ob.jobs.p, jobs[oi].data(), nj * sizeof(DescJob), cudaMemcpyHostToDevice, st_));
        int o = p_.o_min + oi;
        launch_descr(ob.view(nlev_), sc, expn_tab_.p, o, ob.keyin.p, ob.jobs.p, nj, pow(2.0, o), ob.descr.p,
                     ob.written.p, st_);
        PB_CUDA(cudaMemcpyAsync(hd.data() + off * 128, ob.descr.p, (size_t)nj * 128 * sizeof(float),
                                cudaMemcpyDeviceToHost, st_));
        PB_CUDA(cudaMemcpyAsync(hw.data() + off, ob.written.p, nj * sizeof(int), cudaMemcpyDeviceToHost, st_));
        off += nj;
    }
    PB_CUDA(cudaStreamSynchronize(st_));
    // 5. assemble in the reference's insertion order, dropping descriptors the reference leaves unwritten
    out.descr.reserve(nj_total * 128);
    off = 0;
    for (int oi = 0; oi < O; ++oi) {
        OctaveBuf& ob = oct_[oi];
        for (size_t j = 0; j < jobs[oi].size(); ++j, ++off) {
            if (!hw[off]) { out.dropped_unwritten++; continue; }
            out.keys.push_back(ob.keys[jobs[oi][j].key]);
            out.angles.push_back(jobs[oi][j].angle);
            out.key_index.push_back(jobs[oi][j].key);
            out.descr.insert(out.descr.end(), hd.begin() + off * 128, hd.begin() + (off + 1) * 128);
        }
    }
    out.n = (int)out.keys.size();
}

}  // namespace pb
