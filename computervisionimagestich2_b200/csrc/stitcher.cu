// stitcher.cu -- see stitcher.h
#include "stitcher.h"
#include "stitch_host.h"
#include "host_numerics.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <queue>
#include <sstream>
#include <thread>

namespace pb {

bool trace_on() {
    static const bool on = [] { const char* e = getenv("PANO_B200_TRACE"); return e && *e && *e != '0'; }();
    return on;
}
void trace_point(const char* tag, long a) {
    static const auto t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const unsigned tid = (unsigned)(std::hash<std::thread::id>()(std::this_thread::get_id()) % 997u);
    if (a >= 0) fprintf(stderr, "trace %3u %10.3f %s %ld\n", tid, ms, tag, a);
    else fprintf(stderr, "trace %3u %10.3f %s\n", tid, ms, tag);
}

namespace {
struct WallTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
IirCoef make_iir(float sigma) {
    hostnum::VanVliet v = hostnum::vanvliet_coeffs(sigma);
    IirCoef c;
    c.f1 = v.filter[1]; c.f2 = v.filter[2]; c.f3 = v.filter[3];
    c.sumsq = v.filter[0]; c.sum = v.sum; c.bnd = v.bnd;
    for (int i = 0; i < 9; ++i) c.M[i] = v.M[i];
    return c;
}
DericheCoef make_deriche(float sigma) {
    hostnum::Deriche d = hostnum::deriche_coeffs(sigma);
    return DericheCoef{d.a0, d.a1, d.a2, d.a3, d.b1, d.b2, d.coefp, d.coefn};
}
}  // namespace

Stitcher::Stitcher(int device) : dev_(device) {
    PB_CUDA(cudaSetDevice(device));
    PB_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    PB_CUDA(cudaStreamCreateWithFlags(&rst_, cudaStreamNonBlocking));
    PB_CUDA(cudaEventCreateWithFlags(&ev_tables_, cudaEventDisableTiming));
    PB_CUDA(cudaEventCreateWithFlags(&ev_h8_, cudaEventDisableTiming));
    sift_.reset(new SiftEngine(st_));
}
Stitcher::~Stitcher() {
    cudaSetDevice(dev_);
    imgs_.clear();
    pool_.clear();
    sift_.reset();
    for (auto& L : lanes_) {
        L->eng.reset();
        if (L->st) cudaStreamDestroy(L->st);
        L->st = nullptr;
    }
    lanes_.clear();
    if (ev_tables_) cudaEventDestroy(ev_tables_);
    if (ev_h8_) cudaEventDestroy(ev_h8_);
    if (rst_) cudaStreamDestroy(rst_);
    if (st_) cudaStreamDestroy(st_);
}

void Stitcher::ensure_ktab(int short_side) {
    if (ktab_n_ == short_side) return;
    std::vector<float> k;
    hostnum::cylinder_table(short_side, k);
    ktab_.ensure(short_side);
    PB_CUDA(cudaMemcpyAsync(ktab_.p, k.data(), k.size() * sizeof(float), cudaMemcpyHostToDevice, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    ktab_n_ = short_side;
}

// ------------------------------------------------------------------------------------------------------------
// stages
// ------------------------------------------------------------------------------------------------------------
void Stitcher::project(const u8* rgb, int w, int h, u8* out_rgb, u8* out_gray) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t n = (size_t)w * h;
    in_rgb_.ensure(3 * n);
    a_.ensure(3 * n);
    tmp8_.ensure(n);
    ensure_ktab(std::min(w, h));
    PB_CUDA(cudaMemcpyAsync(in_rgb_.p, rgb, 3 * n, cudaMemcpyHostToDevice, st_));
    launch_project_gray(in_rgb_.p, w, h, ktab_.p, a_.p, nullptr, 0, tmp8_.p, st_);
    if (out_rgb) PB_CUDA(cudaMemcpyAsync(out_rgb, a_.p, 3 * n, cudaMemcpyDeviceToHost, st_));
    if (out_gray) PB_CUDA(cudaMemcpyAsync(out_gray, tmp8_.p, n, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

void Stitcher::gray(const u8* rgb, int w, int h, u8* out_gray) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t n = (size_t)w * h;
    in_rgb_.ensure(3 * n);
    tmp8_.ensure(n);
    PB_CUDA(cudaMemcpyAsync(in_rgb_.p, rgb, 3 * n, cudaMemcpyHostToDevice, st_));
    launch_gray(in_rgb_.p, w, h, nullptr, 0, tmp8_.p, st_);
    PB_CUDA(cudaMemcpyAsync(out_gray, tmp8_.p, n, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

void Stitcher::sift_raw_u8(const u8* gray8, int w, int h, const SiftParams& p, RawFeatures& out) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t n = (size_t)w * h;
    tmp8_.ensure(n);
    int pitch = align_up(w, 32);
    gray32_.ensure((size_t)pitch * h);
    PB_CUDA(cudaMemcpyAsync(tmp8_.p, gray8, n, cudaMemcpyHostToDevice, st_));
    launch_u8_to_f32(tmp8_.p, w, gray32_.p, w, h, pitch, st_);  // ImageProcess.cpp:47-51
    sift_->configure(w, h, p);
    sift_->extract(gray32_.p, pitch, out);
}
void Stitcher::sift_raw_f32(const float* img, int w, int h, const SiftParams& p, RawFeatures& out) {
    PB_CUDA(cudaSetDevice(dev_));
    int pitch = align_up(w, 32);
    gray32_.ensure((size_t)pitch * h);
    PB_CUDA(cudaMemcpy2DAsync(gray32_.p, (size_t)pitch * 4, img, (size_t)w * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice,
                              st_));
    sift_->configure(w, h, p);
    sift_->extract(gray32_.p, pitch, out);
}

// std::map<std::vector<float>, VlSiftKeypoint>::insert semantics (ImageProcess.cpp:57, 80-86): ordered by
// lexicographic descriptor comparison, an equal key keeps the FIRST inserted keypoint.
void Stitcher::build_table(const RawFeatures& raw, FeatureTable& t, std::vector<int>* sel, bool host_descr) {
    const int n = raw.n;
    auto less = [&raw](int a, int b) {
        const float *pa = raw.row(a), *pb_ = raw.row(b);
        for (int k = 0; k < 128; ++k) {
            if (pa[k] < pb_[k]) return true;
            if (pb_[k] < pa[k]) return false;
        }
        return false;
    };
    // The first two components, packed order-preservingly into one integer, decide almost every comparison; the full
    // lexicographic comparison only runs on ties of that prefix.  (Same strict weak order as `less`; ties between equal
    // descriptors keep insertion order, which decides which duplicate survives.)  The prefix travels with the index so
    // that the sort touches one small array.
    struct PI { uint64_t prefix; int i; };
    std::vector<PI> pi(n);
    auto ord = [](float f) {   // float -> uint32 with the same order (and -0 == +0)
        f += 0.0f;
        uint32_t u;
        memcpy(&u, &f, 4);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    };
    for (int i = 0; i < n; ++i) pi[i] = PI{((uint64_t)ord(raw.row(i)[0]) << 32) | ord(raw.row(i)[1]), i};
    std::stable_sort(pi.begin(), pi.end(), [&](const PI& a, const PI& b) {
        if (a.prefix != b.prefix) return a.prefix < b.prefix;
        return less(a.i, b.i);
    });
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = pi[i].i;
    t.descr.clear();
    t.keys.clear();
    if (host_descr) t.descr.reserve((size_t)n * 128);
    t.keys.reserve(n);
    if (sel) { sel->clear(); sel->reserve(n); }
    for (int i = 0; i < n; ++i) {
        if (i > 0 && !less(idx[i - 1], idx[i]) && !less(idx[i], idx[i - 1])) continue;  // duplicate key
        const int s = idx[i];
        if (sel) sel->push_back(s);
        if (host_descr) t.descr.insert(t.descr.end(), raw.row(s), raw.row(s) + 128);
        VlKey k = raw.keys[s];
        k.ix = (int)k.x;
        k.iy = (int)k.y;
        t.keys.push_back(k);
    }
    t.n = (int)t.keys.size();
    t.on_device = false; t.quantised = false;
}

void Stitcher::upload_table_on(FeatureTable& t, cudaStream_t st) {
    if (t.on_device) return;
    t.d_descr.ensure(std::max<size_t>((size_t)t.n * 128, 128));
    if (t.n > 0)
        PB_CUDA(cudaMemcpyAsync(t.d_descr.p, t.descr.data(), (size_t)t.n * 128 * sizeof(float), cudaMemcpyHostToDevice,
                                st));
    PB_CUDA(cudaStreamSynchronize(st));
    t.on_device = true; t.quantised = false;
}
void Stitcher::upload_table(FeatureTable& t) { upload_table_on(t, st_); }

void Stitcher::quantise_table(FeatureTable& t) {
    if (t.quantised) return;
    t.d_q8.ensure(std::max<size_t>((size_t)t.n * 32, 32));
    t.d_qe.ensure(std::max<size_t>(t.n, 1) + 1);        // [n] = the table's largest error bound
    int* h = h_qemax_.ensure(1);
    const size_t padded = match_group_pad_rows(t.n);
    t.d_g8.ensure(padded * 8);
    t.d_w16.ensure(padded);
    launch_sad_quantize(t.d_descr.p, t.n, t.d_q8.p, t.d_qe.p, t.d_qe.p + std::max(t.n, 1), st_, t.d_g8.p, t.d_w16.p);
    PB_CUDA(cudaMemcpyAsync(h, t.d_qe.p + std::max(t.n, 1), sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    t.qemax = *h;
    t.quantised = true;
}

// Several directed matching problems (A = database, B = queries) in ONE batch of launches.  out[k][b] = row of A
// matched by query row b, or -1 (ImageProcess.cpp:311-346).  Problems that come with their reverse (B, A) in the same
// batch share one pass over the SAD matrix (match_sad_sym_kernel).
void Stitcher::match_batch(const std::vector<std::pair<FeatureTable*, FeatureTable*>>& probs,
                           std::vector<std::vector<int>>& out, int* d_out) {
    PB_CUDA(cudaSetDevice(dev_));
    const int P = (int)probs.size();
    const bool pre = match_mode_ != 1;
    const bool sym = match_mode_ == 0 || match_mode_ == 3;
    const bool grouped = match_mode_ == 0 && !no_group_;   // pairs take the grouped pass (mode 3: the full SAD pass for both directions)
    const StageTimes tm_before = tm_;
    out.assign(P, std::vector<int>());
    std::vector<int> job_of(P, -1), prob_of;
    for (int k = 0; k < P; ++k) {
        FeatureTable &A = *probs[k].first, &B = *probs[k].second;
        upload_table(A);
        upload_table(B);
        out[k].assign(B.n, -1);
        if (B.n == 0 || A.n < 2) continue;  // the reference reads an unset second neighbour when NA < 2
        if (pre) { quantise_table(A); quantise_table(B); }
        job_of[k] = (int)prob_of.size();
        prob_of.push_back(k);
        tm_.match_pairs_evaluated += (long)A.n * B.n;
        tm_.n_match_calls++;
    }
    const int nj = (int)prob_of.size();
    if (d_out) {   // degenerate problems never reach a kernel: their lists are all -1
        size_t off = 0;
        for (int k = 0; k < P; ++k) {
            const size_t nb = out[k].size();
            if (job_of[k] < 0 && nb) PB_CUDA(cudaMemsetAsync(d_out + off, 0xff, nb * sizeof(int), st_));
            off += nb;
        }
        if (nj == 0) PB_CUDA(cudaStreamSynchronize(st_));
    }
    if (nj == 0) return;
    // pair every problem with its reverse where both are in the batch and both tables qualify for 16-bit bounds
    std::vector<int> partner(nj, -1);
    std::vector<int2> pairs;
    std::vector<int> singles;
    if (sym) {
        std::map<std::pair<const FeatureTable*, const FeatureTable*>, int> open;
        for (int q = 0; q < nj; ++q) {
            const FeatureTable* A = probs[prob_of[q]].first;
            const FeatureTable* B = probs[prob_of[q]].second;
            if (A->qemax > match_sym_err_cap() || B->qemax > match_sym_err_cap() || A == B) continue;
            if (grouped && (A->qemax > match_group_err_cap() || B->qemax > match_group_err_cap())) continue;
            auto it = open.find({B, A});
            if (it != open.end()) {
                partner[q] = it->second;
                partner[it->second] = q;
                pairs.push_back(make_int2(it->second, q));     // F = the earlier problem, R = this one
                open.erase(it);
            } else if (!open.count({A, B})) {
                open[{A, B}] = q;
            }
        }
    }
    for (int q = 0; q < nj; ++q)
        if (partner[q] < 0) singles.push_back(q);
    std::vector<char> is_r(nj, 0);
    for (auto& pr : pairs) is_r[pr.y] = 1;
    // scratch layout
    size_t npart = 0, nidx = 0, nspart = 0, nscratch = 0, ngroup = 0, gq_cap = 0;
    std::vector<size_t> poff(nj), ioff(nj), soff(nj), coff(nj), goff(nj, 0);
    int group_yblocks = 0;
    if (grouped)
        for (auto& pr : pairs) group_yblocks += match_group_yblocks(probs[prob_of[pr.x]].second->n);
    std::vector<int> nsplits(nj), snsplits(nj, 1);
    const int nlaunch_jobs = std::max<int>(1, (int)pairs.size() + (int)singles.size());
    for (int q = 0; q < nj; ++q) {
        FeatureTable &A = *probs[prob_of[q]].first, &B = *probs[prob_of[q]].second;
        nsplits[q] = match_num_splits(A.n, B.n);
        poff[q] = npart; ioff[q] = nidx;
        npart += (size_t)nsplits[q] * B.n;
        nidx += B.n;
        if (pre) {
            soff[q] = nspart; coff[q] = nscratch;
            if (grouped && partner[q] >= 0) {
                goff[q] = ngroup;                                  // statistics through atomics: no per-split partials
                ngroup += match_group_ints(B.n);
            } else if (is_r[q]) {
                nspart += (size_t)match_sym_yblocks(A.n) * B.n;   // one row of statistics per block of held rows of Y = A
            } else {
                snsplits[q] = match_sad_num_splits(A.n, B.n, nlaunch_jobs);
                nspart += (size_t)(snsplits[q] + 1) * B.n;         // attach may round the split count up by one
            }
            nscratch += match_prefilter_ints(B.n);
        }
    }
    partial_.ensure(npart);
    midx_.ensure(nidx);
    if (pre) {
        spartial_.ensure(std::max<size_t>(nspart, 1));
        mscratch_.ensure(nscratch);
        if (ngroup) {
            gscratch_.ensure(ngroup);
            PB_CUDA(cudaMemsetAsync(gscratch_.p, 0x7f, ngroup * sizeof(int), st_));
            // queue of the row pairs the grouped bound cannot skip: 0.07 - 0.2 % of the pairs on SIFT tables; room for 0.5 %
            double rowpairs = 0;
            for (auto& pr : pairs) rowpairs += (double)probs[prob_of[pr.x]].first->n * probs[prob_of[pr.x]].second->n;
            gq_cap = (size_t)std::min(std::max(rowpairs * 0.005, 262144.0), 1024.0 * 1024 * 1024);   // at most 8 GB
            gqueue_.ensure(gq_cap);
        }
        mcount_.ensure((size_t)4 * nj + 4);           // [4 * nj]: entries in the grouped pass's queue
        PB_CUDA(cudaMemsetAsync(mcount_.p, 0, ((size_t)4 * nj + 4) * sizeof(int), st_));
    }
    int* h = h_midx_.ensure(nidx + 4 * (size_t)nj);
    std::vector<MatchJob> jobs(nj);
    for (int q = 0; q < nj; ++q) {
        FeatureTable &A = *probs[prob_of[q]].first, &B = *probs[prob_of[q]].second;
        MatchJob J = make_match_job(A.d_descr.p, A.n, B.d_descr.p, B.n, partial_.p + poff[q], nsplits[q],
                                    midx_.p + ioff[q], nullptr);
        if (pre)
            match_prefilter_attach(J, A.d_q8.p, A.d_qe.p, B.d_q8.p, B.d_qe.p, snsplits[q], spartial_.p + soff[q],
                                   mscratch_.p + coff[q], mcount_.p + 4 * (size_t)q);
        if (grouped && partner[q] >= 0)
            match_group_attach(J, A.d_g8.p, A.d_w16.p, B.d_g8.p, B.d_w16.p, match_group_num_splits(A.n, group_yblocks),
                               gscratch_.p + goff[q]);
        jobs[q] = J;
    }
    if (!grouped)
        for (auto& pr : pairs) match_prefilter_pair(jobs[pr.x], jobs[pr.y]);
    // one upload: job table, pair list, single list
    const size_t jb = align_up((int)(jobs.size() * sizeof(MatchJob)), 16), pb = align_up((int)(pairs.size() * sizeof(int2)), 16);
    char* hj = h_mjobs_.ensure(jb + pb + singles.size() * sizeof(int) + 16);
    memcpy(hj, jobs.data(), jobs.size() * sizeof(MatchJob));
    if (!pairs.empty()) memcpy(hj + jb, pairs.data(), pairs.size() * sizeof(int2));
    if (!singles.empty()) memcpy(hj + jb + pb, singles.data(), singles.size() * sizeof(int));
    mjobs_.ensure(jb + pb + singles.size() * sizeof(int) + 16);
    PB_CUDA(cudaMemcpyAsync(mjobs_.p, hj, jb + pb + singles.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
    const MatchJob* dj = reinterpret_cast<const MatchJob*>(mjobs_.p);
    if (pre)
        launch_match_batch_prefilter(dj, reinterpret_cast<const MatchJob*>(hj), nj, reinterpret_cast<const int2*>(mjobs_.p + jb),
                                     reinterpret_cast<const int2*>(hj + jb), (int)pairs.size(),
                                     reinterpret_cast<const int*>(mjobs_.p + jb + pb), reinterpret_cast<const int*>(hj + jb + pb),
                                     (int)singles.size(), st_, grouped, gqueue_.p,
                                     reinterpret_cast<unsigned long long*>(mcount_.p + 4 * (size_t)nj), gq_cap);
    else launch_match_batch(dj, reinterpret_cast<const MatchJob*>(hj), nj, st_);
    PB_CUDA(cudaMemcpyAsync(h, midx_.p, sizeof(int) * nidx, cudaMemcpyDeviceToHost, st_));
    if (d_out) {
        size_t off = 0;
        for (int k = 0; k < P; ++k) {
            const size_t nb = out[k].size();
            if (job_of[k] >= 0)
                PB_CUDA(cudaMemcpyAsync(d_out + off, midx_.p + ioff[job_of[k]], nb * sizeof(int), cudaMemcpyDeviceToDevice, st_));
            off += nb;
        }
    }
    if (pre) PB_CUDA(cudaMemcpyAsync(h + nidx, mcount_.p, (size_t)4 * nj * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    if (grouped && !pairs.empty() && (h[nidx + 4 * (size_t)pairs[0].x + 3] >> 30)) {
        // the queue overflowed (tables on which the grouped bound skips little): the batch is redone with the full SAD pass
        tm_ = tm_before;
        no_group_ = true;
        try { match_batch(probs, out, d_out); } catch (...) { no_group_ = false; throw; }
        no_group_ = false;
        mstats_.group_overflow++;
        return;
    }
    for (int q = 0; q < nj; ++q) {
        const int k = prob_of[q];
        std::copy(h + ioff[q], h + ioff[q] + out[k].size(), out[k].begin());
        mstats_.problems++;
        mstats_.queries += (long long)out[k].size();
        if (pre) {
            mstats_.survivors += h[nidx + 4 * q];
            mstats_.overflow += h[nidx + 4 * q + 1];
            mstats_.group_exact += h[nidx + 4 * q + 2];
            mstats_.group_accepts += h[nidx + 4 * q + 3] & ((1 << 30) - 1);
        }
    }
    mstats_.sym_pairs += (long long)pairs.size();
    if (grouped) mstats_.group_pairs += (long long)pairs.size();
}

void Stitcher::match_idx(FeatureTable& A, FeatureTable& B, std::vector<int>& idx) {
    std::vector<std::pair<FeatureTable*, FeatureTable*>> probs{{&A, &B}};
    std::vector<std::vector<int>> out;
    match_batch(probs, out);
    idx.swap(out[0]);
}

void Stitcher::match(FeatureTable& A, FeatureTable& B, std::vector<KeyPair>& pairs) {
    std::vector<int> idx;
    match_idx(A, B, idx);
    pairs.clear();
    for (int b = 0; b < B.n; ++b)
        if (idx[b] >= 0) pairs.push_back(KeyPair{A.keys[idx[b]], B.keys[b]});
}

void Stitcher::quantize_u8(const float* descr, int n, u8* out) {
    PB_CUDA(cudaSetDevice(dev_));
    if (n <= 0) return;
    u8f_.ensure((size_t)n * 128);
    u8a_.ensure(u8_table_bytes(n));
    u8raw_.ensure((size_t)n * 128);
    PB_CUDA(cudaMemcpyAsync(u8f_.p, descr, (size_t)n * 512, cudaMemcpyHostToDevice, st_));
    launch_quantize_u8(u8f_.p, n, u8a_.p, st_);
    launch_relayout_u8(u8a_.p, n, u8raw_.p, false, st_);
    PB_CUDA(cudaMemcpyAsync(out, u8raw_.p, (size_t)n * 128, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

// uploads two row-major tables into the blocked device layout and prepares them (norms, extension columns)
void Stitcher::u8_upload(const u8* A, int nA, const u8* B, int nB, U8Table& TA, U8Table& TB) {
    u8raw_.ensure((size_t)std::max(nA, nB) * 128);
    u8a_.ensure(u8_table_bytes(nA)); u8b_.ensure(u8_table_bytes(nB));
    u8na_.ensure(nA); u8nb_.ensure(nB); u8scratch_.ensure(2);
    PB_CUDA(cudaMemsetAsync(u8a_.p, 0, u8_table_bytes(nA), st_));
    PB_CUDA(cudaMemsetAsync(u8b_.p, 0, u8_table_bytes(nB), st_));
    PB_CUDA(cudaMemcpyAsync(u8raw_.p, A, (size_t)nA * 128, cudaMemcpyHostToDevice, st_));
    launch_relayout_u8(u8raw_.p, nA, u8a_.p, true, st_);
    PB_CUDA(cudaMemcpyAsync(u8raw_.p, B, (size_t)nB * 128, cudaMemcpyHostToDevice, st_));
    launch_relayout_u8(u8raw_.p, nB, u8b_.p, true, st_);
    TA.blk = u8a_.p; TA.norm = u8na_.p; TA.n = nA;
    TB.blk = u8b_.p; TB.norm = u8nb_.p; TB.n = nB;
    u8_table_prepare(TA, u8scratch_.p, st_);
    u8_table_prepare(TB, u8scratch_.p, st_);
}

void Stitcher::match_u8(const u8* A, int nA, const u8* B, int nB, int* idx, int* d01) {
    PB_CUDA(cudaSetDevice(dev_));
    if (nB <= 0) return;
    if (nA <= 0) { for (int i = 0; i < nB; ++i) idx[i] = -1; return; }
    U8Table TA, TB;
    u8_upload(A, nA, B, nB, TA, TB);
    u8idx_.ensure(nB); u8d01_.ensure((size_t)nB * 3);
    const int ns = match_u8_num_splits(nA, nB);
    u8part_.ensure((size_t)ns * nB);
    launch_match_u8(TA, TB, u8part_.p, ns, u8idx_.p, u8d01_.p, st_);
    PB_CUDA(cudaMemcpyAsync(idx, u8idx_.p, sizeof(int) * nB, cudaMemcpyDeviceToHost, st_));
    if (d01) PB_CUDA(cudaMemcpyAsync(d01, u8d01_.p, sizeof(int) * 3 * nB, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

float Stitcher::bench_match_u8(const u8* A, int nA, const u8* B, int nB, int reps) {
    PB_CUDA(cudaSetDevice(dev_));
    std::vector<u8> ha, hb;
    if (!(A && B)) {   // uniform pseudo-random bytes
        ha.resize((size_t)nA * 128); hb.resize((size_t)nB * 128);
        unsigned x = 12345u;
        for (auto& v : ha) { x = x * 1664525u + 1013904223u; v = (u8)((x >> 24) & 63); }
        for (auto& v : hb) { x = x * 1664525u + 1013904223u; v = (u8)((x >> 24) & 63); }
        A = ha.data(); B = hb.data();
    }
    U8Table TA, TB;
    u8_upload(A, nA, B, nB, TA, TB);
    u8idx_.ensure(nB);
    const int ns = match_u8_num_splits(nA, nB);
    u8part_.ensure((size_t)ns * nB);
    launch_match_u8(TA, TB, u8part_.p, ns, u8idx_.p, nullptr, st_);   // warm-up
    PB_CUDA(cudaStreamSynchronize(st_));
    timer_start();
    for (int r = 0; r < reps; ++r) launch_match_u8(TA, TB, u8part_.p, ns, u8idx_.p, nullptr, st_);
    const float ms = timer_stop() / reps;
    // the same TMA + UTCIMMA stream with the epilogue reduced to releasing the accumulators: what the tensor pipe delivers
    // to THIS kernel's tile shape when nothing reads the results -- the in-run denominator for "of measured peak"
    launch_match_u8(TA, TB, u8part_.p, ns, u8idx_.p, nullptr, st_, true);
    PB_CUDA(cudaStreamSynchronize(st_));
    timer_start();
    for (int r = 0; r < reps; ++r) launch_match_u8(TA, TB, u8part_.p, ns, u8idx_.p, nullptr, st_, true);
    last_u8_mma_only_ms_ = timer_stop() / reps;
    last_u8_ksteps_ = 4 + TA.ext_steps;
    return ms;
}

bool Stitcher::ransac(const std::vector<const std::vector<KeyPair>*>& problems, std::vector<double>& H8s) {
    PB_CUDA(cudaSetDevice(dev_));
    const int P = (int)problems.size();
    const int iters = stitch::ransac_iterations();
    H8s.assign((size_t)P * 8, 0.0);
    std::vector<int> off(P + 1, 0), samples((size_t)P * iters * 4);
    int maxn = 0;
    for (int p = 0; p < P; ++p) {
        const int n = (int)problems[p]->size();
        off[p + 1] = off[p] + n;
        maxn = std::max(maxn, n);
        std::vector<int> s;
        if (!stitch::draw_samples(n, s, profile_.ransac_seed)) { err_ = "RANSAC needs at least 4 pairs"; return false; }
        std::copy(s.begin(), s.end(), samples.begin() + (size_t)p * iters * 4);
    }
    std::vector<KeyPair> all(off[P]);
    for (int p = 0; p < P; ++p) std::copy(problems[p]->begin(), problems[p]->end(), all.begin() + off[p]);
    const int words = div_up(maxn, 32);
    r_pairs_.ensure(all.size());
    r_off_.ensure(P + 1);
    r_samples_.ensure(samples.size());
    r_counts_.ensure((size_t)P * iters);
    r_masks_.ensure((size_t)P * iters * words);
    PB_CUDA(cudaMemcpyAsync(r_pairs_.p, all.data(), all.size() * sizeof(KeyPair), cudaMemcpyHostToDevice, rst_));
    PB_CUDA(cudaMemcpyAsync(r_off_.p, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, rst_));
    PB_CUDA(cudaMemcpyAsync(r_samples_.p, samples.data(), samples.size() * sizeof(int), cudaMemcpyHostToDevice, rst_));
    launch_ransac_score(r_pairs_.p, r_off_.p, P, r_samples_.p, iters, r_counts_.p, r_masks_.p, words, nullptr, rst_);
    std::vector<int> counts((size_t)P * iters);
    std::vector<unsigned> masks((size_t)P * iters * words);
    PB_CUDA(cudaMemcpyAsync(counts.data(), r_counts_.p, counts.size() * sizeof(int), cudaMemcpyDeviceToHost, rst_));
    PB_CUDA(cudaMemcpyAsync(masks.data(), r_masks_.p, masks.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, rst_));
    PB_CUDA(cudaStreamSynchronize(rst_));
    // selection + least-squares refit (SVD of an n x 4 system) per problem: independent, so the problems of a call -- the
    // two directions of an edge, or all directions of a batch of pairs -- run on host threads
    std::vector<char> okp(P, 1);
    auto finish = [&](int p) {
        const int n = (int)problems[p]->size();
        int best = stitch::select_hypothesis(counts.data() + (size_t)p * iters, iters);
        if (best < 0) { okp[p] = 0; return; }
        const unsigned* m = masks.data() + ((size_t)p * iters + best) * words;
        std::vector<int> inl;
        for (int i = 0; i < n; ++i)
            if (m[i >> 5] >> (i & 31) & 1u) inl.push_back(i);
        if (!stitch::refit(problems[p]->data(), inl, H8s.data() + (size_t)p * 8)) okp[p] = 0;
    };
    const int nthr = std::min(P, 8);
    if (nthr <= 1) {
        for (int p = 0; p < P; ++p) finish(p);
    } else {
        std::vector<std::thread> th;
        for (int t = 1; t < nthr; ++t)
            th.emplace_back([&, t] { for (int p = t; p < P; p += nthr) finish(p); });
        for (int p = 0; p < P; p += nthr) finish(p);
        for (auto& t : th) t.join();
    }
    for (int p = 0; p < P; ++p)
        if (!okp[p]) { err_ = "RANSAC found no inliers"; return false; }
    return true;
}

bool Stitcher::ransac_debug(const std::vector<KeyPair>& pairs, std::vector<int>& counts, std::vector<double>& hyps,
                            std::vector<int>& best_inliers, double* H8) {
    PB_CUDA(cudaSetDevice(dev_));
    const int iters = stitch::ransac_iterations();
    const int n = (int)pairs.size();
    std::vector<int> samples;
    if (!stitch::draw_samples(n, samples, profile_.ransac_seed)) return false;
    int off[2] = {0, n};
    const int words = div_up(n, 32);
    r_pairs_.ensure(n); r_off_.ensure(2); r_samples_.ensure(samples.size()); r_counts_.ensure(iters);
    r_masks_.ensure((size_t)iters * words); r_hyp_.ensure((size_t)iters * 8);
    PB_CUDA(cudaMemcpyAsync(r_pairs_.p, pairs.data(), n * sizeof(KeyPair), cudaMemcpyHostToDevice, st_));
    PB_CUDA(cudaMemcpyAsync(r_off_.p, off, sizeof off, cudaMemcpyHostToDevice, st_));
    PB_CUDA(cudaMemcpyAsync(r_samples_.p, samples.data(), samples.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
    launch_ransac_score(r_pairs_.p, r_off_.p, 1, r_samples_.p, iters, r_counts_.p, r_masks_.p, words, r_hyp_.p, st_);
    counts.resize(iters);
    hyps.resize((size_t)iters * 8);
    std::vector<unsigned> masks((size_t)iters * words);
    PB_CUDA(cudaMemcpyAsync(counts.data(), r_counts_.p, iters * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaMemcpyAsync(hyps.data(), r_hyp_.p, hyps.size() * sizeof(double), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaMemcpyAsync(masks.data(), r_masks_.p, masks.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    int best = stitch::select_hypothesis(counts.data(), iters);
    best_inliers.clear();
    if (best < 0) return false;
    const unsigned* m = masks.data() + (size_t)best * words;
    for (int i = 0; i < n; ++i)
        if (m[i >> 5] >> (i & 31) & 1u) best_inliers.push_back(i);
    return stitch::refit(pairs.data(), best_inliers, H8);
}

void Stitcher::warp_shift(const u8* src, int sw, int sh, const double* H8, float offx, float offy, const u8* prev, int pw,
                          int ph, int ioffx, int ioffy, int cw, int ch, u8* a_out, u8* b_out) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t cn = (size_t)cw * ch;
    a_.ensure(3 * cn);
    b_.ensure(3 * cn);
    H8_.ensure(8);
    if (src) {
        in_rgb_.ensure((size_t)3 * sw * sh);
        PB_CUDA(cudaMemcpyAsync(in_rgb_.p, src, (size_t)3 * sw * sh, cudaMemcpyHostToDevice, st_));
        PB_CUDA(cudaMemcpyAsync(H8_.p, H8, 8 * sizeof(double), cudaMemcpyHostToDevice, st_));
    }
    if (prev) {
        res_[0].ensure((size_t)3 * pw * ph);
        PB_CUDA(cudaMemcpyAsync(res_[0].p, prev, (size_t)3 * pw * ph, cudaMemcpyHostToDevice, st_));
    }
    launch_warp_shift(src ? in_rgb_.p : nullptr, sw, sh, H8_.p, offx, offy, prev ? res_[0].p : nullptr, pw, ph, ioffx,
                      ioffy, src ? a_.p : nullptr, prev ? b_.p : nullptr, cw, ch, st_);
    if (src && a_out) PB_CUDA(cudaMemcpyAsync(a_out, a_.p, 3 * cn, cudaMemcpyDeviceToHost, st_));
    if (prev && b_out) PB_CUDA(cudaMemcpyAsync(b_out, b_.p, 3 * cn, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

// ------------------------------------------------------------------------------------------------------------
// multiband blend (ImageProcess.cpp:648-773) on device buffers
// ------------------------------------------------------------------------------------------------------------
int Stitcher::blend_device(const u8* d_a, const u8* d_b, int cw, int ch, u8* d_out, bool defer_check, int nch,
                           bool own_stats) {
    const int NP = 2 * nch + 1;   // float planes per pyramid level: a, b (nch each) and the mask
    std::vector<int> lw, lh;
    const bool ex6 = profile_.ex6();
    const int L = stitch::blend_levels(cw, ch, lw, lh, ex6);
    if (L < 1) { err_ = "blend: canvas too small"; return -2; }
    // level storage: NP planes per level
    std::vector<size_t> goff(L + 1, 0);
    for (int i = 0; i < L; ++i) goff[i + 1] = goff[i] + (size_t)NP * lw[i] * lh[i];
    pyr_.ensure(goff[L]);
    tmpf_.ensure((size_t)NP * lw[0] * lh[0]);
    if (ex6) tmpf2_.ensure((size_t)NP * lw[0] * lh[0]);
    if (L > 1) {
        E_[0].ensure((size_t)nch * lw[1] * lh[1]);
        E_[1].ensure((size_t)nch * lw[1] * lh[1]);
    }
    // resampling tables for every level transition, packed and uploaded once
    std::vector<int> ti;
    std::vector<float> tf;
    std::vector<double> td;
    struct LevelTab { size_t mx_start, mx_src, mx_wgt, my_start, my_src, my_wgt, lx_pos, lx_alpha, ly_pos, ly_alpha; };
    std::vector<LevelTab> lt(L);
    auto push_mov = [&](int n, int m, size_t& s0, size_t& s1, size_t& s2) {
        s0 = ti.size();
        if (m == n) {  // identity (no resize along this axis)
            for (int i = 0; i <= m; ++i) ti.push_back(i);
            s1 = ti.size();
            for (int i = 0; i < m; ++i) ti.push_back(i);
            s2 = tf.size();
            for (int i = 0; i < m; ++i) tf.push_back(1.0f);
            return;
        }
        hostnum::MovAvgTable T = hostnum::movavg_table((unsigned)n, (unsigned)m);
        ti.insert(ti.end(), T.start.begin(), T.start.end());
        s1 = ti.size();
        ti.insert(ti.end(), T.src.begin(), T.src.end());
        s2 = tf.size();
        tf.insert(tf.end(), T.wgt.begin(), T.wgt.end());
    };
    auto push_lin = [&](int n, int m, size_t& s0, size_t& s1) {
        hostnum::LinearTable T;
        if (n == 1) { T.pos.assign(m, 0); T.alpha.assign(m, 0.0); }
        else if (n == m) { T.pos.resize(m); T.alpha.assign(m, 0.0); for (int i = 0; i < m; ++i) T.pos[i] = i; }
        else T = hostnum::linear_table((unsigned)n, (unsigned)m);
        s0 = ti.size();
        ti.insert(ti.end(), T.pos.begin(), T.pos.end());
        s1 = td.size();
        td.insert(td.end(), T.alpha.begin(), T.alpha.end());
    };
    for (int i = 0; i + 1 < L; ++i) {
        push_mov(lw[i], lw[i + 1], lt[i].mx_start, lt[i].mx_src, lt[i].mx_wgt);
        push_mov(lh[i], lh[i + 1], lt[i].my_start, lt[i].my_src, lt[i].my_wgt);
        push_lin(lw[i + 1], lw[i], lt[i].lx_pos, lt[i].lx_alpha);
        push_lin(lh[i + 1], lh[i], lt[i].ly_pos, lt[i].ly_alpha);
    }
    tab_i_.ensure(std::max<size_t>(ti.size(), 1));
    tab_f_.ensure(std::max<size_t>(tf.size(), 1));
    tab_d_.ensure(std::max<size_t>(td.size(), 1));
    stats_.ensure(8);
    // through pinned staging: an upload from pageable memory synchronises the stream, i.e. would wait for the warp kernel
    // issued just before.  The host may be a whole blend ahead of the GPU, so it first waits until the previous blend's
    // copies out of the staging buffers have executed (they are the first operations of that blend).
    PB_CUDA(cudaEventSynchronize(ev_tables_));
    if (!ti.empty()) {
        int* hp = h_tab_i_.ensure(ti.size());
        memcpy(hp, ti.data(), ti.size() * sizeof(int));
        PB_CUDA(cudaMemcpyAsync(tab_i_.p, hp, ti.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
    }
    if (!tf.empty()) {
        float* hp = h_tab_f_.ensure(tf.size());
        memcpy(hp, tf.data(), tf.size() * sizeof(float));
        PB_CUDA(cudaMemcpyAsync(tab_f_.p, hp, tf.size() * sizeof(float), cudaMemcpyHostToDevice, st_));
    }
    if (!td.empty()) {
        double* hp = h_tab_d_.ensure(td.size());
        memcpy(hp, td.data(), td.size() * sizeof(double));
        PB_CUDA(cudaMemcpyAsync(tab_d_.p, hp, td.size() * sizeof(double), cudaMemcpyHostToDevice, st_));
    }
    PB_CUDA(cudaEventRecord(ev_tables_, st_));
    PB_CUDA(cudaMemsetAsync(stats_.p, 0, 8 * sizeof(int), st_));
    int* err_flag = stats_.p + 4;
    if (defer_check) {
        blend_flag_.ensure(1);
        err_flag = blend_flag_.p;
    }

    const IirCoef coef = make_iir(2.0f);
    const DericheCoef dcoef = make_deriche(2.0f);
    // the seam statistics come from colour plane 0 (ImageProcess.cpp:659-671).  In a plane-sharded job only the rank that
    // carries plane 0 computes them; seam_exchange_ hands them to the others (16 bytes per edge).
    if (own_stats) launch_seam_stats(d_a, d_b, cw, ch, stats_.p, ex6, st_);
    if (seam_exchange_) {
        int* hs = h_stats_.ensure(4);
        if (own_stats) {
            PB_CUDA(cudaMemcpyAsync(hs, stats_.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st_));
            PB_CUDA(cudaStreamSynchronize(st_));
        }
        if (seam_exchange_(seam_user_, hs, own_stats ? 1 : 0) != 0) { err_ = "seam statistics exchange failed"; return -9; }
        if (!own_stats) PB_CUDA(cudaMemcpyAsync(stats_.p, hs, 4 * sizeof(int), cudaMemcpyHostToDevice, st_));
    }
    launch_level0(d_a, d_b, cw, ch, stats_.p, pyr_.p, err_flag, ex6, st_, nch);
    // REDUCE chain (ImageProcess.cpp:705-715)
    for (int i = 1; i < L; ++i) {
        if (ex6) launch_deriche_blur(pyr_.p + goff[i - 1], tmpf2_.p, tmpf_.p, lw[i - 1], lh[i - 1], NP, dcoef, st_);
        else launch_iir_blur(pyr_.p + goff[i - 1], tmpf_.p, lw[i - 1], lh[i - 1], NP, coef, st_);
        const LevelTab& t = lt[i - 1];
        DevMovAvg mx{tab_i_.p + t.mx_start, tab_i_.p + t.mx_src, tab_f_.p + t.mx_wgt,
                     lw[i] == lw[i - 1] ? 1.0f : (float)lw[i - 1]};
        DevMovAvg my{tab_i_.p + t.my_start, tab_i_.p + t.my_src, tab_f_.p + t.my_wgt,
                     lh[i] == lh[i - 1] ? 1.0f : (float)lh[i - 1]};
        launch_reduce(tmpf_.p, lw[i - 1], lh[i - 1], NP, pyr_.p + goff[i], lw[i], lh[i], mx, my, st_);
    }
    // Laplacian + blend + collapse, top-down (ImageProcess.cpp:727-771)
    const float* Eup = nullptr;
    int eb = 0;
    for (int i = L - 1; i >= 0; --i) {
        const float* Gi = pyr_.p + goff[i];
        DevLinear lx{nullptr, nullptr}, ly{nullptr, nullptr};
        const float* Gup = nullptr;
        int uw = 0, uh = 0;
        if (i < L - 1) {
            const LevelTab& t = lt[i];
            lx = DevLinear{tab_i_.p + t.lx_pos, tab_d_.p + t.lx_alpha};
            ly = DevLinear{tab_i_.p + t.ly_pos, tab_d_.p + t.ly_alpha};
            Gup = pyr_.p + goff[i + 1];
            uw = lw[i + 1]; uh = lh[i + 1];
        }
        if (i == 0) {
            launch_collapse(Gi, lw[0], lh[0], Gup, Eup, uw, uh, lx, ly, nullptr, d_out, st_, nch);
        } else {
            float* Eo = E_[eb].p;
            launch_collapse(Gi, lw[i], lh[i], Gup, Eup, uw, uh, lx, ly, Eo, nullptr, st_, nch);
            Eup = Eo;
            eb ^= 1;
        }
    }
    tm_.n_blends++;
    if (defer_check) return 0;
    int flag = 0;
    PB_CUDA(cudaMemcpyAsync(&flag, stats_.p + 4, sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    if (flag) { err_ = "blend: empty middle row (the reference does not terminate on this input)"; return -1; }
    return 0;
}

int Stitcher::check_blend_flag() {
    if (!blend_flag_.p) return 0;
    int flag = 0;
    PB_CUDA(cudaMemcpyAsync(&flag, blend_flag_.p, sizeof(int), cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    if (flag) { err_ = "blend: empty middle row (the reference does not terminate on this input)"; return -1; }
    return 0;
}

int Stitcher::blend(const u8* a, const u8* b, int cw, int ch, u8* out) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t cn = (size_t)3 * cw * ch;
    a_.ensure(cn); b_.ensure(cn); res_[0].ensure(cn);
    PB_CUDA(cudaMemcpyAsync(a_.p, a, cn, cudaMemcpyHostToDevice, st_));
    PB_CUDA(cudaMemcpyAsync(b_.p, b, cn, cudaMemcpyHostToDevice, st_));
    int rc = blend_device(a_.p, b_.p, cw, ch, res_[0].p);
    if (rc) return rc;
    PB_CUDA(cudaMemcpyAsync(out, res_[0].p, cn, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    return 0;
}

void Stitcher::equalize_mix_device(const u8* d_rgb, int w, int h, u8* d_out) {
    hist_.ensure(256);
    lut_.ensure(256);
    launch_luma_hist(d_rgb, w, h, hist_.p, st_);
    int hist[256], lut[256];
    PB_CUDA(cudaMemcpyAsync(hist, hist_.p, sizeof hist, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    {  // equalization.cpp:110-124
        double prob[256], sum[256];
        for (int i = 0; i < 256; ++i) prob[i] = (double)hist[i] / (double)(w * h);
        sum[0] = prob[0];
        lut[0] = (int)round(255.0 * sum[0]);
        for (int i = 1; i < 256; ++i) { sum[i] = sum[i - 1] + prob[i]; lut[i] = (int)round(255.0 * sum[i]); }
    }
    PB_CUDA(cudaMemcpyAsync(lut_.p, lut, sizeof lut, cudaMemcpyHostToDevice, st_));
    // luminance mix: 19/20 + 1/20 (ImageProcess.cpp:261) or 5/6 + 1/6 (src/ex6/ImageProcess.cpp:270)
    launch_equalize_mix(d_rgb, w, h, lut_.p, d_out, profile_.ex6() ? 5.0 : 19.0, profile_.ex6() ? 6.0 : 20.0, st_);
    PB_CUDA(cudaStreamSynchronize(st_));
}

void Stitcher::equalize_mix(const u8* rgb, int w, int h, u8* out) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t n = (size_t)3 * w * h;
    a_.ensure(n); b_.ensure(n);
    PB_CUDA(cudaMemcpyAsync(a_.p, rgb, n, cudaMemcpyHostToDevice, st_));
    equalize_mix_device(a_.p, w, h, b_.p);
    PB_CUDA(cudaMemcpyAsync(out, b_.p, n, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

// transfer tran(src, tem, out) (transfer.cpp:4-13), host buffers
void Stitcher::color_transfer(const u8* src, int w, int h, const u8* tem, int tw, int th, u8* out) {
    PB_CUDA(cudaSetDevice(dev_));
    const size_t n = (size_t)w * h, nt = (size_t)tw * th;
    a_.ensure(3 * n); b_.ensure(3 * nt); tmp8_.ensure(3 * n);
    pyr_.ensure(3 * n); tmpf_.ensure(std::max<size_t>(3 * nt, 64)); tab_f_.ensure(64);
    PB_CUDA(cudaMemcpyAsync(a_.p, src, 3 * n, cudaMemcpyHostToDevice, st_));
    PB_CUDA(cudaMemcpyAsync(b_.p, tem, 3 * nt, cudaMemcpyHostToDevice, st_));
    launch_color_transfer(a_.p, w, h, b_.p, tw, th, pyr_.p, tmpf_.p, tab_f_.p, tmp8_.p, st_);
    PB_CUDA(cudaMemcpyAsync(out, tmp8_.p, 3 * n, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

void Stitcher::cimg_blur2(const float* src, int w, int h, int c, float* dst) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t n = (size_t)w * h * c;
    tmpf_.ensure(n);
    PB_CUDA(cudaMemcpyAsync(tmpf_.p, src, n * 4, cudaMemcpyHostToDevice, st_));
    launch_iir_blur(tmpf_.p, tmpf_.p, w, h, c, make_iir(2.0f), st_);
    PB_CUDA(cudaMemcpyAsync(dst, tmpf_.p, n * 4, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

void Stitcher::cimg_blur2_deriche(const float* src, int w, int h, int c, float* dst) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t n = (size_t)w * h * c;
    tmpf_.ensure(n);
    tmpf2_.ensure(n);
    pyr_.ensure(n);
    PB_CUDA(cudaMemcpyAsync(pyr_.p, src, n * 4, cudaMemcpyHostToDevice, st_));
    launch_deriche_blur(pyr_.p, tmpf2_.p, tmpf_.p, w, h, c, make_deriche(2.0f), st_);
    PB_CUDA(cudaMemcpyAsync(dst, tmpf_.p, n * 4, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

void Stitcher::cimg_resize(const float* src, int w, int h, int c, int nw, int nh, float* dst) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t n = (size_t)w * h * c, nn = (size_t)nw * nh * c;
    tmpf_.ensure(n);
    pyr_.ensure(nn);
    PB_CUDA(cudaMemcpyAsync(tmpf_.p, src, n * 4, cudaMemcpyHostToDevice, st_));
    std::vector<int> ti;
    std::vector<float> tf;
    std::vector<double> td;
    if (nw <= w && nh <= h) {
        size_t o[6];
        auto push = [&](int a, int b, size_t& s0, size_t& s1, size_t& s2) {
            s0 = ti.size();
            if (a == b) {
                for (int i = 0; i <= b; ++i) ti.push_back(i);
                s1 = ti.size();
                for (int i = 0; i < b; ++i) ti.push_back(i);
                s2 = tf.size();
                for (int i = 0; i < b; ++i) tf.push_back(1.0f);
                return;
            }
            hostnum::MovAvgTable T = hostnum::movavg_table(a, b);
            ti.insert(ti.end(), T.start.begin(), T.start.end());
            s1 = ti.size();
            ti.insert(ti.end(), T.src.begin(), T.src.end());
            s2 = tf.size();
            tf.insert(tf.end(), T.wgt.begin(), T.wgt.end());
        };
        push(w, nw, o[0], o[1], o[2]);
        push(h, nh, o[3], o[4], o[5]);
        tab_i_.ensure(ti.size()); tab_f_.ensure(tf.size());
        PB_CUDA(cudaMemcpyAsync(tab_i_.p, ti.data(), ti.size() * 4, cudaMemcpyHostToDevice, st_));
        PB_CUDA(cudaMemcpyAsync(tab_f_.p, tf.data(), tf.size() * 4, cudaMemcpyHostToDevice, st_));
        PB_CUDA(cudaStreamSynchronize(st_));
        launch_reduce(tmpf_.p, w, h, c, pyr_.p, nw, nh, DevMovAvg{tab_i_.p + o[0], tab_i_.p + o[1], tab_f_.p + o[2], nw == w ? 1.0f : (float)w},
                      DevMovAvg{tab_i_.p + o[3], tab_i_.p + o[4], tab_f_.p + o[5], nh == h ? 1.0f : (float)h}, st_);
    } else if (nw >= w && nh >= h) {
        size_t o[4];
        auto push = [&](int a, int b, size_t& s0, size_t& s1) {
            hostnum::LinearTable T;
            if (a == 1) { T.pos.assign(b, 0); T.alpha.assign(b, 0.0); }
            else if (a == b) { T.pos.resize(b); T.alpha.assign(b, 0.0); for (int i = 0; i < b; ++i) T.pos[i] = i; }
            else T = hostnum::linear_table(a, b);
            s0 = ti.size();
            ti.insert(ti.end(), T.pos.begin(), T.pos.end());
            s1 = td.size();
            td.insert(td.end(), T.alpha.begin(), T.alpha.end());
        };
        push(w, nw, o[0], o[1]);
        push(h, nh, o[2], o[3]);
        tab_i_.ensure(ti.size()); tab_d_.ensure(td.size());
        PB_CUDA(cudaMemcpyAsync(tab_i_.p, ti.data(), ti.size() * 4, cudaMemcpyHostToDevice, st_));
        PB_CUDA(cudaMemcpyAsync(tab_d_.p, td.data(), td.size() * 8, cudaMemcpyHostToDevice, st_));
        PB_CUDA(cudaStreamSynchronize(st_));
        launch_expand(tmpf_.p, w, h, c, pyr_.p, nw, nh, DevLinear{tab_i_.p + o[0], tab_d_.p + o[1]},
                      DevLinear{tab_i_.p + o[2], tab_d_.p + o[3]}, st_);
    } else {
        throw std::runtime_error("cimg_resize: mixed shrink/grow is not on the stitching path");
    }
    PB_CUDA(cudaMemcpyAsync(dst, pyr_.p, nn * 4, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

// ------------------------------------------------------------------------------------------------------------
// pipeline
// ------------------------------------------------------------------------------------------------------------
// An image whose projection and feature table were computed elsewhere (another GPU of a sharded job).
void Stitcher::add_precomputed(const u8* proj_rgb, int w, int h, const float* descr, const VlKey* keys, int n) {
    PB_CUDA(cudaSetDevice(dev_));
    std::unique_ptr<Image> im;
    if (!pool_.empty()) { im = std::move(pool_.back()); pool_.pop_back(); }
    else im.reset(new Image());
    im->w = w; im->h = h;
    const size_t np = (size_t)w * h;
    im->proj.ensure(3 * np);
    PB_CUDA(cudaMemcpyAsync(im->proj.p, proj_rgb, 3 * np, cudaMemcpyHostToDevice, st_));
    im->feat.n = n;
    im->feat.descr.assign(descr, descr + (size_t)n * 128);
    im->feat.keys.assign(keys, keys + n);
    im->feat.on_device = false; im->feat.quantised = false;
    upload_table(im->feat);   // synchronises the stream: proj_rgb may be released by the caller
    imgs_.push_back(std::move(im));
}
// a preset list must have one entry per feature of image j, each -1 or a row of image i's table (it indexes A.keys)
bool Stitcher::preset_fits(const PresetMatch& pm) const {
    const int n = (int)imgs_.size();
    if (pm.i < 0 || pm.i >= n || pm.j < 0 || pm.j >= n || (int)pm.idx.size() != imgs_[pm.j]->feat.n) return false;
    const int na = imgs_[pm.i]->feat.n;
    for (int v : pm.idx)
        if (v < -1 || v >= na) return false;
    return true;
}
void Stitcher::preset_match(int i, int j, const int* idx, int nB) {
    PresetMatch pm;
    pm.i = i; pm.j = j;
    pm.idx.assign(idx, idx + nB);
    preset_.push_back(std::move(pm));
}
std::unique_ptr<Stitcher::Image> Stitcher::new_image() {
    std::unique_ptr<Image> im;
    if (!pool_.empty()) { im = std::move(pool_.back()); pool_.pop_back(); }
    else im.reset(new Image());
    im->w = im->h = 0;
    im->feat.n = 0;
    im->feat.keys.clear();
    im->feat.descr.clear();
    im->feat.on_device = false;
    im->feat.quantised = false;
    return im;
}

// ---- sharded job with a device-resident exchange (computervisionimagestich2_b200/dist.py) ------------------
// Every rank holds the job's image list 0..n-1; a rank fills the slots of the images it owns (shard_extract), exports
// their blocks into the exchange buffers (device memory owned by the caller: the NCCL send buffers), and imports the
// other ranks' blocks from the receive buffers.  Copies are device-to-device on the stitcher's stream.
void Stitcher::shard_begin(int n_global) {
    PB_CUDA(cudaSetDevice(dev_));
    clear();
    for (int i = 0; i < n_global; ++i) imgs_.push_back(new_image());
}
void Stitcher::shard_export(int i, float* d_descr_out, VlKey* keys_out, u8* d_proj_out) {
    PB_CUDA(cudaSetDevice(dev_));
    if (i < 0 || i >= (int)imgs_.size()) throw std::runtime_error("shard_export: image index out of range");
    Image& im = *imgs_[i];
    upload_table(im.feat);
    if (d_descr_out && im.feat.n > 0)
        PB_CUDA(cudaMemcpyAsync(d_descr_out, im.feat.d_descr.p, (size_t)im.feat.n * 512, cudaMemcpyDeviceToDevice, st_));
    if (keys_out && im.feat.n > 0) memcpy(keys_out, im.feat.keys.data(), (size_t)im.feat.n * sizeof(VlKey));
    if (d_proj_out) PB_CUDA(cudaMemcpyAsync(d_proj_out, im.proj.p, (size_t)3 * im.w * im.h, cudaMemcpyDeviceToDevice, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}
void Stitcher::shard_import(int i, int w, int h, int n, const float* d_descr, const VlKey* keys, const u8* d_proj) {
    PB_CUDA(cudaSetDevice(dev_));
    if (i < 0 || i >= (int)imgs_.size()) throw std::runtime_error("shard_import: image index out of range");
    Image& im = *imgs_[i];
    im.w = w; im.h = h;
    im.feat.n = n;
    im.feat.descr.clear();
    im.feat.keys.assign(keys, keys + n);
    im.feat.d_descr.ensure(std::max<size_t>((size_t)n * 128, 128));
    if (n > 0) PB_CUDA(cudaMemcpyAsync(im.feat.d_descr.p, d_descr, (size_t)n * 512, cudaMemcpyDeviceToDevice, st_));
    im.feat.on_device = true; im.feat.quantised = false;
    if (d_proj) {
        im.proj.ensure((size_t)3 * w * h);
        PB_CUDA(cudaMemcpyAsync(im.proj.p, d_proj, (size_t)3 * w * h, cudaMemcpyDeviceToDevice, st_));
    }
    PB_CUDA(cudaStreamSynchronize(st_));
}
// directed problems (I[k], J[k]) = getImgPair(imgs[I[k]], imgs[J[k]]): database I, queries J; the match lists are left
// in device memory, concatenated in problem order (nfeat[J[k]] ints each)
void Stitcher::shard_match(const int* I, const int* J, int nprob, int* d_idx_out) {
    PB_CUDA(cudaSetDevice(dev_));
    std::vector<std::pair<FeatureTable*, FeatureTable*>> probs;
    for (int k = 0; k < nprob; ++k) {
        if (I[k] < 0 || I[k] >= (int)imgs_.size() || J[k] < 0 || J[k] >= (int)imgs_.size())
            throw std::runtime_error("shard_match: image index out of range");
        probs.push_back({&imgs_[I[k]]->feat, &imgs_[J[k]]->feat});
    }
    std::vector<std::vector<int>> out;
    match_batch(probs, out, d_idx_out);
}

// readFile() body for one image with everything returned to the host (sharded jobs exchange these between ranks)
void Stitcher::extract(const u8* rgb, int w, int h, u8* proj_out, FeatureTable& t) {
    PB_CUDA(cudaSetDevice(dev_));
    const size_t np = (size_t)w * h;
    in_rgb_.ensure(3 * np);
    a_.ensure(3 * np);
    const int pitch = align_up(w, 32);
    gray32_.ensure((size_t)pitch * h);
    ensure_ktab(std::min(w, h));
    PB_CUDA(cudaMemcpyAsync(in_rgb_.p, rgb, 3 * np, cudaMemcpyHostToDevice, st_));
    launch_project_gray(in_rgb_.p, w, h, ktab_.p, a_.p, gray32_.p, pitch, nullptr, st_);
    if (proj_out) PB_CUDA(cudaMemcpyAsync(proj_out, a_.p, 3 * np, cudaMemcpyDeviceToHost, st_));
    SiftParams sp;
    RawFeatures raw;
    sift_->configure(w, h, sp);
    sift_->extract(gray32_.p, pitch, raw);
    build_table(raw, t);
    PB_CUDA(cudaStreamSynchronize(st_));
}

void Stitcher::clear() {
    preset_.clear();
    // keep the per-image HBM buffers (projected image, descriptor table) for the next job: cudaMalloc / cudaFree
    // cost milliseconds each and synchronise the device
    for (auto& im : imgs_) pool_.push_back(std::move(im));
    imgs_.clear();
    log_.clear();
    err_.clear();
    tm_ = StageTimes();
    rw_ = rh_ = 0;
}

// readFile body for one image (ImageProcess.cpp:12-23)
void Stitcher::add_image(const u8* rgb, int w, int h) {
    PB_CUDA(cudaSetDevice(dev_));
    size_t n = (size_t)w * h;
    in_rgb_.ensure(3 * n);
    PB_CUDA(cudaMemcpyAsync(in_rgb_.p, rgb, 3 * n, cudaMemcpyHostToDevice, st_));
    add_image_device(in_rgb_.p, w, h);
}

void Stitcher::add_image_device(const u8* d_rgb, int w, int h) {
    PB_CUDA(cudaSetDevice(dev_));
    WallTimer t0;
    std::unique_ptr<Image> im;
    if (!pool_.empty()) { im = std::move(pool_.back()); pool_.pop_back(); }
    else im.reset(new Image());
    im->w = w; im->h = h;
    size_t n = (size_t)w * h;
    im->proj.ensure(3 * n);
    int pitch = align_up(w, 32);
    gray32_.ensure((size_t)pitch * h);
    ensure_ktab(std::min(w, h));
    launch_project_gray(d_rgb, w, h, ktab_.p, im->proj.p, gray32_.p, pitch, nullptr, st_);
    PB_CUDA(cudaStreamSynchronize(st_));
    tm_.project += t0.ms();
    WallTimer t1;
    SiftParams sp;
    RawFeatures raw;
    sift_->configure(w, h, sp);
    sift_->extract(gray32_.p, pitch, raw);
    tm_.sift += t1.ms();
    tm_.sift_pixels += (long)n;
    WallTimer t2;
    build_table(raw, im->feat);
    upload_table(im->feat);
    tm_.table += t2.ms();
    imgs_.push_back(std::move(im));
}

// One lane's share of readFile(): images first, first + step, ...  Runs on its own host thread.
void Stitcher::lane_work(Lane& L, int first, int step, const u8* const* imgs, const int* w, const int* h, int n,
                         bool on_device, const int* slots) {
    try {
        PB_CUDA(cudaSetDevice(dev_));
        for (int i = first; i < n; i += step) {
            Image& im = *imgs_[slots[i]];
            const int iw = w[i], ih = h[i];
            im.w = iw; im.h = ih;
            const size_t np = (size_t)iw * ih;
            WallTimer t0;
            PB_TRACE("lane.image.begin", i);
            const u8* d_rgb = imgs[i];
            if (!on_device) {
                L.in_rgb.ensure(3 * np);
                PB_CUDA(cudaMemcpyAsync(L.in_rgb.p, imgs[i], 3 * np, cudaMemcpyHostToDevice, L.st));
                d_rgb = L.in_rgb.p;
            }
            im.proj.ensure(3 * np);
            const int pitch = align_up(iw, 32);
            L.gray32.ensure((size_t)pitch * ih);
            const int shortside = std::min(iw, ih);
            if (L.ktab_n != shortside) {
                std::vector<float> k;
                hostnum::cylinder_table(shortside, k);
                L.ktab.ensure(shortside);
                PB_CUDA(cudaMemcpyAsync(L.ktab.p, k.data(), k.size() * sizeof(float), cudaMemcpyHostToDevice, L.st));
                PB_CUDA(cudaStreamSynchronize(L.st));
                L.ktab_n = shortside;
            }
            launch_project_gray(d_rgb, iw, ih, L.ktab.p, im.proj.p, L.gray32.p, pitch, nullptr, L.st);
            L.t_project += t0.ms();
            WallTimer t1;
            SiftParams sp;
            RawFeatures raw;
            L.eng->configure(iw, ih, sp);
            L.eng->extract(L.gray32.p, pitch, raw, false);   // descriptors stay in the engine's pinned buffer
            L.t_sift += t1.ms();
            WallTimer t2;
            PB_TRACE("lane.table.begin", i);
            std::vector<int> sel;
            build_table(raw, im.feat, &sel, false);            // ... and the table's device copy is gathered below
            PB_TRACE("lane.table.sorted", i);
            if (raw.d_descr && !sel.empty()) {
                // the descriptors are still in the engine's device buffer: gather the sorted rows there instead of
                // sending the 1.2 MB table back over PCIe from pageable memory
                const int nt = (int)sel.size();
                int* hr = L.h_rows.ensure(nt);
                for (int q = 0; q < nt; ++q) hr[q] = raw.dev_row[sel[q]];
                L.rows.ensure(nt);
                im.feat.d_descr.ensure((size_t)nt * 128);
                PB_CUDA(cudaMemcpyAsync(L.rows.p, hr, (size_t)nt * sizeof(int), cudaMemcpyHostToDevice, L.st));
                launch_gather_rows128(raw.d_descr, L.rows.p, nt, im.feat.d_descr.p, L.st);
                PB_CUDA(cudaStreamSynchronize(L.st));   // the engine's buffer is re-used by the lane's next image
                im.feat.on_device = true; im.feat.quantised = false;
            } else {   // no descriptors at all
                im.feat.d_descr.ensure(128);
                im.feat.on_device = true; im.feat.quantised = false;
            }
            L.t_table += t2.ms();
            PB_TRACE("lane.image.end", i);
        }
    } catch (const std::exception& e) {
        L.err = e.what();
    }
}

void Stitcher::add_images(const u8* const* imgs, const int* w, const int* h, int n, bool on_device, const int* slots) {
    PB_CUDA(cudaSetDevice(dev_));
    WallTimer tw;
    if (n <= 0) return;
    const int nl = std::max(1, std::min(want_lanes_, n));
    while ((int)lanes_.size() < nl) {
        std::unique_ptr<Lane> L(new Lane());
        PB_CUDA(cudaStreamCreateWithFlags(&L->st, cudaStreamNonBlocking));
        L->eng.reset(new SiftEngine(L->st));
        lanes_.push_back(std::move(L));
    }
    // image i of this call goes to imgs_[slot[i]]: the slots given by the caller (sharded jobs: global image index,
    // shard_begin made the list) or n new slots appended to the list
    std::vector<int> own_slots;
    if (!slots) {
        const int base = (int)imgs_.size();
        for (int i = 0; i < n; ++i) {
            imgs_.push_back(new_image());
            own_slots.push_back(base + i);
        }
        slots = own_slots.data();
    } else {
        for (int i = 0; i < n; ++i)
            if (slots[i] < 0 || slots[i] >= (int)imgs_.size()) throw std::runtime_error("add_images: image slot out of range");
    }
    for (int k = 0; k < nl; ++k) { lanes_[k]->err.clear(); lanes_[k]->t_project = lanes_[k]->t_sift = lanes_[k]->t_table = 0; }
    if (nl == 1) {
        lane_work(*lanes_[0], 0, 1, imgs, w, h, n, on_device, slots);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < nl; ++k)
            th.emplace_back([this, k, nl, imgs, w, h, n, on_device, slots] { lane_work(*lanes_[k], k, nl, imgs, w, h, n, on_device, slots); });
        for (auto& t : th) t.join();
    }
    for (int k = 0; k < nl; ++k) {
        if (!lanes_[k]->err.empty()) throw std::runtime_error(lanes_[k]->err);
        tm_.project += lanes_[k]->t_project;   // per-lane times overlap; `features` is the wall time of the stage
        tm_.table += lanes_[k]->t_table;
    }
    for (int i = 0; i < n; ++i) tm_.sift_pixels += (long)w[i] * h[i];
    tm_.sift += tw.ms();
}

// Independent pairs: readFile of all 2K images on the lanes, the 2K directed matching problems in one launch, the RANSAC
// problems of the adjacent directions in one launch.  Nothing is stitched.
int Stitcher::pairs(const u8* const* imgs, const int* w, const int* h, int npairs, PairRecord* out, bool on_device) {
    PB_CUDA(cudaSetDevice(dev_));
    clear();
    if (npairs <= 0) return 0;
    add_images(imgs, w, h, 2 * npairs, on_device);
    std::vector<std::pair<FeatureTable*, FeatureTable*>> probs;
    for (int p = 0; p < npairs; ++p) {
        FeatureTable &A = imgs_[2 * p]->feat, &B = imgs_[2 * p + 1]->feat;
        probs.push_back({&A, &B});
        probs.push_back({&B, &A});
    }
    std::vector<std::vector<int>> idx;
    {
        WallTimer t;
        match_batch(probs, idx);
        tm_.match += t.ms();
    }
    std::vector<std::vector<KeyPair>> lists(probs.size());
    std::vector<const std::vector<KeyPair>*> problems;
    std::vector<int> prob_of;
    for (size_t q = 0; q < probs.size(); ++q) {
        const FeatureTable &S = *probs[q].first, &D = *probs[q].second;
        PairRecord& r = out[q / 2];
        const int d = (int)(q & 1);
        r.nfeat[d] = S.n;
        int c = 0;
        for (int v : idx[q]) c += v >= 0;
        r.nmatch[d] = c;
        r.has_h[d] = 0;
        for (int k = 0; k < 8; ++k) r.H[d][k] = 0.0;
        if (c < 20) continue;   // not adjacent (ImageProcess.cpp:128)
        lists[q].reserve(c);
        for (int b = 0; b < D.n; ++b)
            if (idx[q][b] >= 0) lists[q].push_back(KeyPair{S.keys[idx[q][b]], D.keys[b]});
        problems.push_back(&lists[q]);
        prob_of.push_back((int)q);
    }
    if (!problems.empty()) {
        WallTimer t;
        std::vector<double> H8s;
        if (!ransac(problems, H8s)) return -3;
        for (size_t k = 0; k < problems.size(); ++k) {
            PairRecord& r = out[prob_of[k] / 2];
            const int d = prob_of[k] & 1;
            r.has_h[d] = 1;
            std::copy(H8s.begin() + 8 * k, H8s.begin() + 8 * k + 8, r.H[d]);
        }
        tm_.ransac += t.ms();
    }
    return 0;
}

void Stitcher::stage_images(const u8* const* imgs, const int* w, const int* h, int n) {
    PB_CUDA(cudaSetDevice(dev_));
    staged_.clear();
    for (int i = 0; i < n; ++i) {
        std::unique_ptr<Staged> s(new Staged());
        s->w = w[i]; s->h = h[i];
        size_t bytes = (size_t)3 * w[i] * h[i];
        s->rgb.ensure(bytes);
        PB_CUDA(cudaMemcpyAsync(s->rgb.p, imgs[i], bytes, cudaMemcpyHostToDevice, st_));
        staged_.push_back(std::move(s));
    }
    PB_CUDA(cudaStreamSynchronize(st_));
}

int Stitcher::run_staged() {
    clear();
    const int n = (int)staged_.size();
    std::vector<const u8*> p(n);
    std::vector<int> w(n), h(n);
    for (int i = 0; i < n; ++i) { p[i] = staged_[i]->rgb.p; w[i] = staged_[i]->w; h[i] = staged_[i]->h; }
    add_images(p.data(), w.data(), h.data(), n, true);
    return run();
}

int Stitcher::stitch_bmp(const u8* const* files, const size_t* sizes, int n, u8** out, size_t* out_size) {
    PB_CUDA(cudaSetDevice(dev_));
    clear();
    auto rd32 = [](const u8* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
    staged_.clear();
    for (int i = 0; i < n; ++i) {
        const u8* f = files[i];
        if (sizes[i] < 54 || f[0] != 'B' || f[1] != 'M') { err_ = "not a BMP file"; return -7; }
        const uint32_t off = rd32(f + 10), hdr = rd32(f + 14);
        int w = (int)rd32(f + 18), h = (int)rd32(f + 22);
        const int bpp = f[28] | (f[29] << 8);
        const uint32_t comp = hdr >= 40 ? rd32(f + 30) : 0;
        const bool bottom_up = h > 0;
        if (h < 0) h = -h;
        const size_t stride = ((size_t)3 * w + 3) & ~(size_t)3;
        if (bpp != 24 || comp != 0 || w <= 0 || h <= 0 || sizes[i] < off + stride * h) {
            err_ = "only uncompressed 24-bpp BMP files are on the stitching path";
            return -7;
        }
        std::unique_ptr<Staged> s(new Staged());
        s->w = w; s->h = h;
        s->rgb.ensure((size_t)3 * w * h);
        // every file has its own raw buffer: the upload of file i + 1 (a blocking copy out of pageable memory) runs while
        // the decode kernel of file i executes, and nothing waits per file
        s->raw.ensure(stride * h);
        PB_CUDA(cudaMemcpyAsync(s->raw.p, f + off, stride * h, cudaMemcpyHostToDevice, st_));
        launch_bmp_to_planar(s->raw.p, (int)stride, w, h, bottom_up, s->rgb.p, st_);
        staged_.push_back(std::move(s));
    }
    PB_CUDA(cudaStreamSynchronize(st_));   // the lanes read the decoded images on their own streams
    for (auto& s : staged_) s->raw.release();
    const int rc = run_staged();
    if (rc) return rc;
    const int w = rw_, h = rh_;
    const size_t stride = ((size_t)3 * w + 3) & ~(size_t)3, total = 54 + stride * h;
    bmp_out_.ensure(stride * h);
    launch_planar_to_bmp(res_[cur_].p, w, h, (int)stride, bmp_out_.p, st_);
    u8* o = (u8*)malloc(total);
    if (!o) { err_ = "out of host memory"; return -8; }
    memset(o, 0, 54);
    auto wr32 = [](u8* p, uint32_t v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); };
    o[0] = 'B'; o[1] = 'M';
    wr32(o + 2, (uint32_t)total); wr32(o + 10, 54); wr32(o + 14, 40); wr32(o + 18, (uint32_t)w); wr32(o + 22, (uint32_t)h);
    o[26] = 1; o[28] = 24; wr32(o + 34, (uint32_t)(stride * h)); wr32(o + 38, 2835); wr32(o + 42, 2835);
    PB_CUDA(cudaMemcpyAsync(o + 54, bmp_out_.p, stride * h, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
    *out = o;
    *out_size = total;
    return 0;
}

void Stitcher::flush_l2() {
    PB_CUDA(cudaSetDevice(dev_));
    const size_t bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    flush_.ensure(bytes);
    PB_CUDA(cudaMemsetAsync(flush_.p, 1, bytes, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}
void Stitcher::timer_start() {
    PB_CUDA(cudaSetDevice(dev_));
    if (!ev0_) { PB_CUDA(cudaEventCreate(&ev0_)); PB_CUDA(cudaEventCreate(&ev1_)); }
    PB_CUDA(cudaEventRecord(ev0_, st_));
}
float Stitcher::timer_stop() {
    PB_CUDA(cudaEventRecord(ev1_, st_));
    PB_CUDA(cudaEventSynchronize(ev1_));
    float ms = 0;
    PB_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
    return ms;
}

// One edge of the stitch loop (ImageProcess.cpp:177-233; src/ex6/ImageProcess.cpp:195-258): match lists of both
// directions -> the larger one mirrored -> two RANSAC fits -> canvas plan -> warp + shift -> feature re-mapping ->
// blend.  s2d_idx / d2s_idx = getImgPair(imgs[src], imgs[dst]) / getImgPair(imgs[dst], imgs[src]) as row indices.
int Stitcher::stitch_edge(int src, int dst, int pre, const std::vector<int>& s2d_idx, const std::vector<int>& d2s_idx) {
    auto pairs_of = [&](int a, int b, const std::vector<int>& idx, std::vector<KeyPair>& out) {   // with the current keys
        out.clear();
        const FeatureTable &A = imgs_[a]->feat, &B = imgs_[b]->feat;
        for (int q = 0; q < B.n; ++q)
            if (idx[q] >= 0) out.push_back(KeyPair{A.keys[idx[q]], B.keys[q]});
    };
    std::vector<KeyPair> s2d, d2s;
    {
        WallTimer t;
        pairs_of(src, dst, s2d_idx, s2d);
        pairs_of(dst, src, d2s_idx, d2s);
        tm_.match += t.ms();
    }
    if (s2d.size() > d2s.size()) {
        d2s.clear();
        for (size_t q = 0; q < s2d.size(); ++q) d2s.push_back(KeyPair{s2d[q].dst, s2d[q].src});
    } else {
        s2d.clear();
        for (size_t q = 0; q < d2s.size(); ++q) s2d.push_back(KeyPair{d2s[q].dst, d2s[q].src});
    }
    std::vector<double> H;
    {
        WallTimer t;
        std::vector<const std::vector<KeyPair>*> probs{&d2s, &s2d};
        PB_TRACE("edge.ransac.begin", (long)d2s.size());
        if (!ransac(probs, H)) return -3;
        PB_TRACE("edge.ransac.end");
        tm_.ransac += t.ms();
    }
    const double* fwd = H.data();       // RANSAC(dstToSrcPair)
    const double* bwd = H.data() + 8;   // RANSAC(srcToDstPair)
    Image& D = *imgs_[dst];
    stitch::CanvasPlan cp = stitch::plan_canvas(D.w, D.h, fwd, rw_, rh_, profile_.ex6());
    if (cp.new_w <= 0 || cp.new_h <= 0 || (long)cp.new_w * cp.new_h > (1L << 31)) {
        err_ = "degenerate canvas";
        return -4;
    }
    const int nch = cplanes_;   // colour planes this stitcher carries through the canvas stages (3, or 1 in a plane-sharded job)
    const size_t cn = (size_t)nch * cp.new_w * cp.new_h;
    {
        WallTimer t;
        a_.ensure(cn);
        b_.ensure(cn);
        PB_CUDA(cudaEventSynchronize(ev_h8_));   // the previous edge's copy out of the staging buffer has executed
        double* hh = h_H8_.ensure(8);
        memcpy(hh, bwd, 8 * sizeof(double));
        PB_CUDA(cudaMemcpyAsync(H8_.p, hh, 8 * sizeof(double), cudaMemcpyHostToDevice, st_));
        PB_CUDA(cudaEventRecord(ev_h8_, st_));
        launch_warp_shift(D.proj.p + (size_t)cplane0_ * D.w * D.h, D.w, D.h, H8_.p, cp.min_x, cp.min_y, res_[cur_].p, rw_, rh_,
                          (int)cp.min_x, (int)cp.min_y, a_.p, b_.p, cp.new_w, cp.new_h, st_, nch);
        // no synchronisation here: the blend's host-side table construction overlaps this kernel (tm_.warp is then the
        // issue time only; the kernel's own time is in the per-kernel report)
        tm_.warp += t.ms();
    }
    stitch::update_features_by_homography(D.feat.keys.data(), D.feat.n, fwd, cp.min_x, cp.min_y);
    stitch::update_features_by_offset(imgs_[pre]->feat.keys.data(), imgs_[pre]->feat.n, (int)cp.min_x, (int)cp.min_y);
    {
        WallTimer t;
        res_[cur_ ^ 1].ensure(cn);
        int rc = blend_device(a_.p, b_.p, cp.new_w, cp.new_h, res_[cur_ ^ 1].p, true, nch, cplane0_ == 0);
        if (rc) return rc;
        cur_ ^= 1;
        rw_ = cp.new_w; rh_ = cp.new_h;
        tm_.blend += t.ms();
    }
    return 0;
}

// ImageProcess::matching of the src/ex6 variant (src/ex6/ImageProcess.cpp:147-260): images are assumed to be in
// left-to-right order; image i is adjacent to i-1 and i+1 (image n-1 lists no neighbour of its own, :152-157), the
// walk starts at image n/2 and visits each node's neighbours from the back of its list.  No adjacency discovery and no
// THRESHOLD test: only the 2(n-1) directed problems of the chain are matched, in one batched launch.
int Stitcher::run_chain(std::ostringstream& log) {
    const int n = (int)imgs_.size();
    if (n < 2) { err_ = "the ex6 order needs at least 2 images (the reference indexes imgs[1])"; return -7; }
    for (int i = 0; i < n; ++i)
        if (imgs_[i]->w > imgs_[i]->h) {   // src/ex6/ImageProcess.cpp:35-38: exit(1)
            err_ = "ex6: projected width > height (the reference exits on this input)";
            return -8;
        }
    std::vector<std::vector<int>> next(n);
    next[0].push_back(1);
    for (int i = 1; i < n - 1; i++) { next[i].push_back(i + 1); next[i].push_back(i - 1); }
    const int start = n / 2;
    // plan the walk, then evaluate every directed problem of its edges at once
    std::vector<std::pair<int, int>> edges;
    {
        std::vector<std::vector<int>> nx = next;
        std::queue<int> q;
        q.push(start);
        while (!q.empty()) {
            int src = q.front();
            q.pop();
            for (int i = (int)nx[src].size() - 1; i >= 0; i--) {
                int dst = nx[src][i];
                q.push(dst);
                nx[src].pop_back();
                for (auto it = nx[dst].begin(); it != nx[dst].end(); ++it)
                    if (*it == src) { nx[dst].erase(it); break; }
                edges.push_back({src, dst});
            }
        }
    }
    std::vector<std::vector<int>> fwd_idx(edges.size()), bwd_idx(edges.size());
    for (auto& pm : preset_)
        if (!preset_fits(pm)) {   // same contract as run(): a preset that does not fit is an error, never dropped
            err_ = "preset match list does not fit the images";
            return -6;
        }
    {
        WallTimer t;
        std::vector<std::pair<FeatureTable*, FeatureTable*>> probs;
        std::vector<std::vector<int>*> sink;
        auto want = [&](int a, int b, std::vector<int>& out) {
            for (auto& pm : preset_)
                if (pm.i == a && pm.j == b) { out = pm.idx; return; }
            probs.push_back({&imgs_[a]->feat, &imgs_[b]->feat});
            sink.push_back(&out);
        };
        for (size_t e = 0; e < edges.size(); ++e) {
            want(edges[e].first, edges[e].second, fwd_idx[e]);
            want(edges[e].second, edges[e].first, bwd_idx[e]);
        }
        if (!probs.empty()) {
            std::vector<std::vector<int>> out;
            match_batch(probs, out);
            for (size_t k = 0; k < out.size(); ++k) sink[k]->swap(out[k]);
        }
        tm_.match += t.ms();
    }
    {
        Image& s = *imgs_[start];
        rw_ = s.w; rh_ = s.h;
        cur_ = 0;
        res_[0].ensure((size_t)3 * rw_ * rh_);
        PB_CUDA(cudaMemcpyAsync(res_[0].p, s.proj.p, (size_t)3 * rw_ * rh_, cudaMemcpyDeviceToDevice, st_));
    }
    H8_.ensure(8);
    int pre = start;
    for (size_t e = 0; e < edges.size(); ++e) {
        const int src = edges[e].first, dst = edges[e].second;
        log << "src index:" << src << " dst index:" << dst << "\n";   // src/ex6/ImageProcess.cpp:182
        int rc = stitch_edge(src, dst, pre, fwd_idx[e], bwd_idx[e]);
        if (rc) return rc;
        pre = dst;
    }
    return 0;
}

int Stitcher::run() {
    PB_CUDA(cudaSetDevice(dev_));
    WallTimer ttot;
    const int n = (int)imgs_.size();
    if (n == 0) { err_ = "no images"; return -1; }
    std::ostringstream log;
    blend_flag_.ensure(1);
    PB_CUDA(cudaMemsetAsync(blend_flag_.p, 0, sizeof(int), st_));
    if (profile_.ex6()) {
        int rc = run_chain(log);
        if (rc) return rc;
        rc = check_blend_flag();
        if (rc) return rc;
        WallTimer t;
        res_[cur_ ^ 1].ensure((size_t)3 * rw_ * rh_);
        equalize_mix_device(res_[cur_].p, rw_, rh_, res_[cur_ ^ 1].p);
        cur_ ^= 1;
        tm_.tail += t.ms();
        log_ = log.str();
        tm_.total += ttot.ms();
        return 0;
    }
    std::vector<std::vector<char>> adj(n, std::vector<char>(n, 0));
    std::vector<std::vector<int>> next(n);
    std::vector<KeyPair> pairs;
    // adjacency discovery (ImageProcess.cpp:117-137).  The reference evaluates getImgPair(i, j) for i < j always and
    // for i > j only when (j, i) turned out not adjacent; the same set of directed problems is evaluated here, as two
    // batched launches.  Match INDICES depend only on the descriptor tables, so they are cached per directed pair
    // and reused by the stitching loop (which the reference re-evaluates, ImageProcess.cpp:177-178).
    std::vector<std::vector<std::vector<int>>> midx(n, std::vector<std::vector<int>>(n));
    std::vector<std::vector<char>> have(n, std::vector<char>(n, 0));
    for (auto& pm : preset_) {   // directed problems evaluated elsewhere (another GPU of a sharded job)
        if (!preset_fits(pm)) {
            err_ = "preset match list does not fit the images";
            return -6;
        }
        midx[pm.i][pm.j] = pm.idx;
        have[pm.i][pm.j] = 1;
    }
    auto run_wave = [&](const std::vector<std::pair<int, int>>& w0) {
        std::vector<std::pair<int, int>> w;
        for (auto& ij : w0)
            if (!have[ij.first][ij.second]) w.push_back(ij);
        if (w.empty()) return;
        std::vector<std::pair<FeatureTable*, FeatureTable*>> probs;
        for (auto& ij : w) probs.push_back({&imgs_[ij.first]->feat, &imgs_[ij.second]->feat});
        std::vector<std::vector<int>> out;
        match_batch(probs, out);
        for (size_t k = 0; k < w.size(); ++k) { midx[w[k].first][w[k].second].swap(out[k]); have[w[k].first][w[k].second] = 1; }
    };
    auto count_matches = [&](int i, int j) {
        int c = 0;
        for (int v : midx[i][j]) c += v >= 0;
        return c;
    };
    {
        WallTimer t;
        std::vector<std::pair<int, int>> wave;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) {
                wave.push_back({i, j});
                // the reverse problem shares the SAD matrix: with the symmetric pass it comes almost for free, and the
                // reference evaluates it anyway for every non-adjacent pair (below) and every tree edge (:177-178).
                // Only the lists the reference would compute are ever consulted.
                if (match_mode_ == 0) wave.push_back({j, i});
            }
        PB_TRACE("run.match.wave1.begin", (long)wave.size());
        run_wave(wave);
        PB_TRACE("run.match.wave1.end");
        wave.clear();
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < i; ++j)
                if (count_matches(j, i) < 20) wave.push_back({i, j});
        run_wave(wave);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                if (i == j) continue;
                if (adj[j][i]) { adj[i][j] = 1; next[i].push_back(j); continue; }
                if (count_matches(i, j) >= 20) { adj[i][j] = 1; next[i].push_back(j); }
            }
        tm_.match += t.ms();
    }
    const int start = stitch::middle_index(next, adj);
    log << start << "\n";
    int pre = start;
    std::queue<int> wait;
    wait.push(start);
    // result = imgs[start].projectedSrc
    {
        Image& s = *imgs_[start];
        rw_ = s.w; rh_ = s.h;
        cur_ = 0;
        res_[0].ensure((size_t)cplanes_ * rw_ * rh_);
        PB_CUDA(cudaMemcpyAsync(res_[0].p, s.proj.p + (size_t)cplane0_ * rw_ * rh_, (size_t)cplanes_ * rw_ * rh_,
                                cudaMemcpyDeviceToDevice, st_));
    }
    H8_.ensure(8);
    {   // the stitching order depends on the adjacency alone: plan it now and evaluate, in one launch, the directed
        // problems of the tree edges that the discovery stage skipped
        WallTimer t;
        std::vector<std::vector<char>> a2 = adj;
        std::queue<int> q2;
        q2.push(start);
        std::vector<std::pair<int, int>> wave;
        while (!q2.empty()) {
            int src = q2.front();
            q2.pop();
            for (int i = (int)next[src].size() - 1; i >= 0; i--) {
                int dst = next[src][i];
                if (!a2[src][dst]) continue;
                a2[src][dst] = a2[dst][src] = 0;
                q2.push(dst);
                if (!have[src][dst]) wave.push_back({src, dst});
                if (!have[dst][src]) wave.push_back({dst, src});
            }
        }
        run_wave(wave);
        tm_.match += t.ms();
    }
    while (!wait.empty()) {
        int src = wait.front();
        wait.pop();
        for (int i = (int)next[src].size() - 1; i >= 0; i--) {
            int dst = next[src][i];
            if (!adj[src][dst]) continue;
            adj[src][dst] = adj[dst][src] = 0;
            wait.push(dst);
            log << src << " " << dst << "\n";
            PB_TRACE("run.edge.begin", dst);
            int rc = stitch_edge(src, dst, pre, midx[src][dst], midx[dst][src]);
            PB_TRACE("run.edge.issued", dst);
            if (rc) return rc;
            pre = dst;
        }
    }
    {
        const int rc = check_blend_flag();
        PB_TRACE("run.edges.done");
        if (rc) return rc;
    }
    log_ = log.str();
    if (skip_tail_) {   // plane-sharded job: the equalisation needs all three planes (run_tail after plane_import)
        tm_.total += ttot.ms();
        return 0;
    }
    {
        WallTimer t;
        res_[cur_ ^ 1].ensure((size_t)3 * rw_ * rh_);
        equalize_mix_device(res_[cur_].p, rw_, rh_, res_[cur_ ^ 1].p);
        cur_ ^= 1;
        tm_.tail += t.ms();
    }
    tm_.total += ttot.ms();
    return 0;
}

// ---- plane-sharded canvas stages: the colour planes of warp / shift / blend are independent -----------
int Stitcher::run_planes(int first, int count, SeamExchange cb, void* user) {
    if (profile_.ex6()) { err_ = "plane sharding: the ex6 seam statistics need all three planes"; return -10; }
    if (!((count == 3 && first == 0) || (count == 1 && first >= 0 && first < 3))) { err_ = "plane sharding: planes must be 0..2"; return -10; }
    cplane0_ = first; cplanes_ = count;
    seam_exchange_ = cb; seam_user_ = user;
    skip_tail_ = true;
    int rc;
    try {
        rc = run();
    } catch (...) {
        cplane0_ = 0; cplanes_ = 3; seam_exchange_ = nullptr; seam_user_ = nullptr; skip_tail_ = false;
        throw;
    }
    seam_exchange_ = nullptr; seam_user_ = nullptr; skip_tail_ = false;
    planes_first_ = first; planes_count_ = count;
    cplane0_ = 0; cplanes_ = 3;
    return rc;
}
void Stitcher::plane_export(int k, u8* d_out) {
    PB_CUDA(cudaSetDevice(dev_));
    if (k < 0 || k >= planes_count_) throw std::runtime_error("plane_export: plane out of range");
    PB_CUDA(cudaMemcpyAsync(d_out, res_[cur_].p + (size_t)k * rw_ * rh_, (size_t)rw_ * rh_, cudaMemcpyDeviceToDevice, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}
void Stitcher::plane_import(int channel, const u8* d_in) {
    PB_CUDA(cudaSetDevice(dev_));
    if (channel < 0 || channel >= 3) throw std::runtime_error("plane_import: channel out of range");
    const size_t n = (size_t)rw_ * rh_;
    if (planes_count_ != 3) {   // first import: lay the canvas out as three planes, own planes in place
        res_[cur_ ^ 1].ensure(3 * n);
        PB_CUDA(cudaMemcpyAsync(res_[cur_ ^ 1].p + (size_t)planes_first_ * n, res_[cur_].p, (size_t)planes_count_ * n,
                                cudaMemcpyDeviceToDevice, st_));
        cur_ ^= 1;
        planes_first_ = 0; planes_count_ = 3;
    }
    PB_CUDA(cudaMemcpyAsync(res_[cur_].p + (size_t)channel * n, d_in, n, cudaMemcpyDeviceToDevice, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}
int Stitcher::run_tail() {
    PB_CUDA(cudaSetDevice(dev_));
    if (planes_count_ != 3) { err_ = "run_tail: the canvas does not hold all three planes"; return -10; }
    WallTimer t;
    res_[cur_ ^ 1].ensure((size_t)3 * rw_ * rh_);
    equalize_mix_device(res_[cur_].p, rw_, rh_, res_[cur_ ^ 1].p);
    cur_ ^= 1;
    tm_.tail += t.ms();
    return 0;
}

void Stitcher::copy_result(u8* dst) {
    PB_CUDA(cudaSetDevice(dev_));
    PB_CUDA(cudaMemcpyAsync(dst, res_[cur_].p, (size_t)3 * rw_ * rh_, cudaMemcpyDeviceToHost, st_));
    PB_CUDA(cudaStreamSynchronize(st_));
}

}  // namespace pb
