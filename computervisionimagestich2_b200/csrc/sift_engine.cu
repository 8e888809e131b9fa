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
    // (level, row, column) of the detection packed into one integer: candidates are distinct pixels, so the order is total
    std::vector<std::pair<unsigned long long, int>> byraster;
    byraster.reserve(n);
    for (int i = 0; i < n; ++i)
        if (r[i].good)
            byraster.push_back({((unsigned long long)(unsigned)r[i].is0 << 48) | ((unsigned long long)(unsigned)r[i].iy0 << 24) |
                                    (unsigned long long)(unsigned)r[i].ix0, i});
    std::sort(byraster.begin(), byraster.end());
    std::vector<int> idx(byraster.size());
    for (size_t q = 0; q < byraster.size(); ++q) idx[q] = byraster[q].second;
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

OctaveSet SiftEngine::octave_set(int first, int count) const {
    OctaveSet os;
    memset(&os, 0, sizeof os);
    for (int i = 0; i < count && i < kMaxOctaveSet; ++i) {
        os.ov[i] = oct_[first + i].view(nlev_);
        os.xper[i] = pow(2.0, p_.o_min + first + i);
    }
    return os;
}

void SiftEngine::orient_custom(int oi, const std::vector<KeyIn>& keys, std::vector<int>& nang,
                               std::vector<double>& ang) {
    OctaveBuf& ob = oct_[oi];
    int n = (int)keys.size();
    nang.assign(n, 0);
    ang.assign((size_t)n * 4, 0.0);
    if (n == 0 || ob.w < 2 || ob.h < 2) return;
    std::vector<KeyIn> ki(keys);
    for (auto& k : ki) k.oct = 0;
    keyin_.ensure(n);
    nangles_.ensure(n);
    angles_.ensure((size_t)n * 4);
    PB_CUDA(cudaMemcpyAsync(keyin_.p, ki.data(), ki.size() * sizeof(KeyIn), cudaMemcpyHostToDevice, st_));
    launch_orient(octave_set(oi, 1), consts(), expn_tab_.p, keyin_.p, n, nullptr, nangles_.p, angles_.p, st_);
    PB_CUDA(cudaMemcpyAsync(nang.data(), nangles_.p, n * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaMemcpyAsync(ang.data(), angles_.p, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

void SiftEngine::orient_octave(int oi) {
    OctaveBuf& ob = oct_[oi];
    std::vector<KeyIn> ki(ob.keys.size());
    for (size_t i = 0; i < ki.size(); ++i)
        ki[i] = KeyIn{ob.keys[i].x, ob.keys[i].y, ob.keys[i].sigma, (short)ob.keys[i].is, 0};
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
    std::vector<KeyIn> ki;
    if (custom_keys) ki = *custom_keys;
    else {
        ki.resize(ob.keys.size());
        for (size_t i = 0; i < ki.size(); ++i)
            ki[i] = KeyIn{ob.keys[i].x, ob.keys[i].y, ob.keys[i].sigma, (short)ob.keys[i].is, 0};
    }
    for (auto& k : ki) k.oct = 0;
    keyin_.ensure(ki.size());
    PB_CUDA(cudaMemcpyAsync(keyin_.p, ki.data(), ki.size() * sizeof(KeyIn), cudaMemcpyHostToDevice, st_));
    std::vector<DescJob> jobs(nj);
    for (int i = 0; i < nj; ++i) {
        jobs[i].key = key_idx[i];
        jobs[i].pad = 0;
        jobs[i].angle = ang[i];
        jobs[i].st0 = sin(ang[i]);  // vl/sift.c:1308-1309, glibc
        jobs[i].ct0 = cos(ang[i]);
    }
    jobs_.ensure(nj);
    descr_.ensure((size_t)nj * 128);
    written_.ensure(nj);
    PB_CUDA(cudaMemcpyAsync(jobs_.p, jobs.data(), nj * sizeof(DescJob), cudaMemcpyHostToDevice, st_));
    launch_descr(octave_set(oi, 1), consts(), expn_tab_.p, keyin_.p, jobs_.p, nj, nullptr, descr_.p, written_.p, 0.0, st_);
    PB_CUDA(cudaMemcpyAsync(out_descr, descr_.p, (size_t)nj * 128 * sizeof(float), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaMemcpyAsync(out_written, written_.p, nj * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

// stable counting sort of item indices by descending radius (0..255)
static void order_by_radius_desc(const int* radius, int n, int* order) {
    int first[257] = {0};
    for (int i = 0; i < n; ++i) first[255 - radius[i] + 1]++;
    for (int b = 0; b < 256; ++b) first[b + 1] += first[b];
    for (int i = 0; i < n; ++i) order[first[255 - radius[i]]++] = i;
}

void SiftEngine::extract(const float* d_img, int img_pitch, RawFeatures& out, bool copy_descr) {
    out = RawFeatures();
    const int O = (int)oct_.size();
    if (O == 0) return;
    if (O > kMaxOctaveSet) throw std::runtime_error("more than 8 octaves per image are not supported by the batched path");
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
    PB_TRACE("sift.pyramid.issued");
    PB_CUDA(cudaMemcpyAsync(h_counts_.p, counts_.p, sizeof(int) * O, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    PB_TRACE("sift.pyramid.done");
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
    PB_TRACE("sift.keys.downloaded", (long)total);
    off = 0;
    for (int oi = 0; oi < O; ++oi)
        if (cnt[oi] > 0) {
            finish_keys(hr + off, cnt[oi], p_.o_min + oi, sigma0_, p_.S, oct_[oi].keys);
            off += cnt[oi];
        }
    // 3. orientations of the keypoints of all octaves: one launch, one round trip
    std::vector<int> kfirst(O + 1, 0);
    for (int oi = 0; oi < O; ++oi) {
        out.noct_keys.push_back((int)oct_[oi].keys.size());
        kfirst[oi + 1] = kfirst[oi] + (int)oct_[oi].keys.size();
    }
    const int nk = kfirst[O];
    if (nk == 0) return;
    const OctaveSet os = octave_set(0, O);
    KeyIn* hk = (KeyIn*)h_keyin_.ensure((size_t)nk * sizeof(KeyIn));
    for (int oi = 0; oi < O; ++oi) {
        const OctaveBuf& ob = oct_[oi];
        for (size_t i = 0; i < ob.keys.size(); ++i)
            hk[kfirst[oi] + i] = KeyIn{ob.keys[i].x, ob.keys[i].y, ob.keys[i].sigma, (short)ob.keys[i].is, (short)oi};
    }
    keyin_.ensure(nk);
    nangles_.ensure(nk);
    angles_.ensure((size_t)nk * 4);
    // largest windows first (both kernels: one warp per item, the launch ends with its slowest warp): a counting sort on
    // the integer window radius -- a comparison sort of the keys costs more host time than the ordering saves
    int* h_ok = (int*)h_order_k_.ensure((size_t)nk * sizeof(int));
    {
        std::vector<int> radius(nk);
        for (int i = 0; i < nk; ++i) {
            const double Wd = floor(3.0 * 1.5 * ((double)hk[i].sigma / os.xper[hk[i].oct]));   // vl/sift.c:928
            radius[i] = Wd > 255 ? 255 : (Wd > 1 ? (int)Wd : 1);
        }
        order_by_radius_desc(radius.data(), nk, h_ok);
    }
    order_k_.ensure(nk);
    PB_CUDA(cudaMemcpyAsync(order_k_.p, h_ok, (size_t)nk * sizeof(int), cudaMemcpyHostToDevice, st_));
    int* h_na = (int*)h_nang_.ensure((size_t)nk * sizeof(int));
    double* h_an = (double*)h_ang_.ensure((size_t)nk * 4 * sizeof(double));
    PB_CUDA(cudaMemcpyAsync(keyin_.p, hk, (size_t)nk * sizeof(KeyIn), cudaMemcpyHostToDevice, st_));
    PB_TRACE("sift.orient.launch", nk);
    launch_orient(os, sc, expn_tab_.p, keyin_.p, nk, order_k_.p, nangles_.p, angles_.p, st_);
    PB_CUDA(cudaMemcpyAsync(h_na, nangles_.p, (size_t)nk * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaMemcpyAsync(h_an, angles_.p, (size_t)nk * 4 * sizeof(double), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    PB_TRACE("sift.orient.done");
    for (int oi = 0; oi < O; ++oi) {
        OctaveBuf& ob = oct_[oi];
        const int n = (int)ob.keys.size();
        ob.h_nangles.assign(h_na + kfirst[oi], h_na + kfirst[oi] + n);
        ob.h_angles.assign(h_an + (size_t)kfirst[oi] * 4, h_an + (size_t)(kfirst[oi] + n) * 4);
    }
    // 4. descriptors of all octaves: one launch, one round trip (sin/cos on the host)
    size_t nj = 0;
    for (int i = 0; i < nk; ++i) nj += h_na[i];
    if (nj == 0) return;
    DescJob* hj = (DescJob*)h_jobs_.ensure(nj * sizeof(DescJob));
    {
        size_t q = 0;
        for (int i = 0; i < nk; ++i)
            for (int j = 0; j < h_na[i]; ++j) {
                const double a = h_an[(size_t)i * 4 + j];
                hj[q++] = DescJob{i, 0, a, sin(a), cos(a)};
            }
    }
    jobs_.ensure(nj);
    descr_.ensure(nj * 128);
    written_.ensure(nj);
    float* hd = (float*)h_descr_.ensure(nj * 128 * sizeof(float));
    int* hw = (int*)h_written_.ensure(nj * sizeof(int));
    PB_CUDA(cudaMemcpyAsync(jobs_.p, hj, nj * sizeof(DescJob), cudaMemcpyHostToDevice, st_));
    double patch_bytes = 0;   // sum_k (2 W_k + 1)^2 * 8 B of (modulus, angle) reads
    int* h_oj = (int*)h_order_j_.ensure(nj * sizeof(int));
    {
        std::vector<int> radius(nj);
        for (size_t q = 0; q < nj; ++q) {
            const KeyIn& kk = hk[hj[q].key];
            const double sbp = p_.magnif * ((double)kk.sigma / os.xper[kk.oct]);
            const double W = floor(1.4142135623730951 * sbp * 2.5 + 0.5);
            patch_bytes += (2 * W + 1) * (2 * W + 1) * 8.0;
            radius[q] = W > 255 ? 255 : (W > 0 ? (int)W : 0);
        }
        order_by_radius_desc(radius.data(), (int)nj, h_oj);
    }
    order_j_.ensure(nj);
    PB_CUDA(cudaMemcpyAsync(order_j_.p, h_oj, nj * sizeof(int), cudaMemcpyHostToDevice, st_));
    PB_TRACE("sift.descr.launch", (long)nj);
    launch_descr(os, sc, expn_tab_.p, keyin_.p, jobs_.p, (int)nj, order_j_.p, descr_.p, written_.p, patch_bytes, st_);
    PB_CUDA(cudaMemcpyAsync(hd, descr_.p, nj * 128 * sizeof(float), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaMemcpyAsync(hw, written_.p, nj * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    PB_TRACE("sift.descr.done");
    // 5. assemble in the reference's insertion order (octave, keypoint, angle), dropping descriptors the reference
    //    leaves unwritten
    if (copy_descr) out.descr.reserve(nj * 128);
    out.keys.reserve(nj);
    out.angles.reserve(nj);
    out.key_index.reserve(nj);
    out.dev_row.reserve(nj);
    out.d_descr = descr_.p;
    out.h_descr = hd;
    int oi = 0;
    for (size_t q = 0; q < nj; ++q) {
        if (!hw[q]) { out.dropped_unwritten++; continue; }
        const int gk = hj[q].key;
        while (gk >= kfirst[oi + 1]) ++oi;
        out.keys.push_back(oct_[oi].keys[gk - kfirst[oi]]);
        out.angles.push_back(hj[q].angle);
        out.key_index.push_back(gk - kfirst[oi]);
        if (copy_descr) out.descr.insert(out.descr.end(), hd + q * 128, hd + (q + 1) * 128);
        out.dev_row.push_back((int)q);
    }
    out.n = (int)out.keys.size();
}

}  // namespace pb
