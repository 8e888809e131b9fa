// c_api.cu -- extern "C" entry points of libpano_b200.so (include/pano_b200.h).  Exceptions stop here.
#include "../../include/pano_b200.h"
#include "stitcher.h"
#include "stitch_host.h"
#include "ktimer.h"
#include <sstream>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace pb;

struct pano_b200_ctx {
    std::unique_ptr<Stitcher> st;
    std::string err;
    RawFeatures last_raw;
    FeatureTable match_a, match_b;   // pano_b200_match: the device tables persist between calls (cudaMalloc / cudaFree
                                     // cost more than the matching kernel at a few thousand features)
};

static_assert(sizeof(pano_b200_keypoint) == sizeof(VlKey), "keypoint ABI");
static_assert(sizeof(pano_b200_pair) == sizeof(KeyPair), "pair ABI");

#define PB_API_BEGIN try {
#define PB_API_END                                                     \
    }                                                                  \
    catch (const std::exception& e) {                                  \
        if (ctx) ctx->err = e.what();                                  \
        return -100;                                                   \
    }                                                                  \
    catch (...) {                                                      \
        if (ctx) ctx->err = "unknown exception";                       \
        return -101;                                                   \
    }

// results handed to the caller are malloc'ed (freed with pano_b200_free); an allocation failure becomes an error code
static void* xmalloc(size_t n) {
    void* p = malloc(n ? n : 1);
    if (!p) throw std::runtime_error("out of host memory");
    return p;
}

extern "C" {

int pano_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pano_b200_create(int device, pano_b200_ctx** out) {
    *out = nullptr;
    pano_b200_ctx* ctx = new pano_b200_ctx();
    try {
        ctx->st.reset(new Stitcher(device));
    } catch (const std::exception& e) {
        fprintf(stderr, "pano_b200_create: %s\n", e.what());
        delete ctx;
        return -100;
    }
    *out = ctx;
    return 0;
}
void pano_b200_destroy(pano_b200_ctx* ctx) { delete ctx; }
const char* pano_b200_last_error(pano_b200_ctx* ctx) {
    if (!ctx) return "no context";
    if (ctx->err.empty() && ctx->st) return ctx->st->error().c_str();
    return ctx->err.c_str();
}
void pano_b200_free(void* p) { free(p); }

int pano_b200_stitch(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, int n, uint8_t** out,
                     int* out_w, int* out_h) {
    PB_API_BEGIN
    ctx->err.clear();
    Stitcher& S = *ctx->st;
    S.clear();
    S.add_images(imgs, w, h, n, false);
    int rc = S.run();
    if (rc) { ctx->err = S.error(); return rc; }
    *out_w = S.result_width();
    *out_h = S.result_height();
    *out = (uint8_t*)xmalloc((size_t)3 * *out_w * *out_h);
    S.copy_result(*out);
    return 0;
    PB_API_END
}
int pano_b200_extract(pano_b200_ctx* ctx, const uint8_t* rgb, int w, int h, uint8_t* proj_out, float** descr,
                      pano_b200_keypoint** keys, int* n) {
    PB_API_BEGIN
    FeatureTable t;
    ctx->st->extract(rgb, w, h, proj_out, t);
    *n = t.n;
    *descr = (float*)xmalloc(std::max<size_t>((size_t)t.n * 128 * sizeof(float), 4));
    *keys = (pano_b200_keypoint*)xmalloc(std::max<size_t>((size_t)t.n * sizeof(VlKey), 4));
    memcpy(*descr, t.descr.data(), (size_t)t.n * 128 * sizeof(float));
    memcpy(*keys, t.keys.data(), (size_t)t.n * sizeof(VlKey));
    return 0;
    PB_API_END
}
int pano_b200_pairs(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, int npairs,
                    pano_b200_pair_record* out) {
    PB_API_BEGIN
    static_assert(sizeof(pano_b200_pair_record) == sizeof(Stitcher::PairRecord) && sizeof(pano_b200_pair_record) == 160,
                  "pair record layout");
    ctx->err.clear();
    int rc = ctx->st->pairs(imgs, w, h, npairs, reinterpret_cast<Stitcher::PairRecord*>(out));
    if (rc) ctx->err = ctx->st->error();
    return rc;
    PB_API_END
}
int pano_b200_pairs_staged(pano_b200_ctx* ctx, const uint8_t* const* d_imgs, const int* w, const int* h, int npairs,
                           pano_b200_pair_record* out) {
    PB_API_BEGIN
    ctx->err.clear();
    int rc = ctx->st->pairs(d_imgs, w, h, npairs, reinterpret_cast<Stitcher::PairRecord*>(out), true);
    if (rc) ctx->err = ctx->st->error();
    return rc;
    PB_API_END
}
int pano_b200_stitch_features(pano_b200_ctx* ctx, int nimg, const uint8_t* const* proj, const int* w, const int* h,
                              const float* const* descr, const pano_b200_keypoint* const* keys, const int* nfeat,
                              const int* const* match_idx, uint8_t** out, int* out_w, int* out_h) {
    PB_API_BEGIN
    ctx->err.clear();
    Stitcher& S = *ctx->st;
    S.clear();
    for (int i = 0; i < nimg; ++i)
        S.add_precomputed(proj[i], w[i], h[i], descr[i], reinterpret_cast<const VlKey*>(keys[i]), nfeat[i]);
    if (match_idx)
        for (int i = 0; i < nimg; ++i)
            for (int j = 0; j < nimg; ++j)
                if (i != j && match_idx[(size_t)i * nimg + j]) S.preset_match(i, j, match_idx[(size_t)i * nimg + j], nfeat[j]);
    int rc = S.run();
    if (rc) { ctx->err = S.error(); return rc; }
    *out_w = S.result_width();
    *out_h = S.result_height();
    *out = (uint8_t*)xmalloc((size_t)3 * *out_w * *out_h);
    S.copy_result(*out);
    return 0;
    PB_API_END
}
int pano_b200_shard_begin(pano_b200_ctx* ctx, int n_global) {
    PB_API_BEGIN
    if (n_global <= 0) return -1;
    ctx->err.clear();
    ctx->st->shard_begin(n_global);
    return 0;
    PB_API_END
}
int pano_b200_shard_extract(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, const int* slot,
                            int n_local, int on_device) {
    PB_API_BEGIN
    if (n_local < 0 || (n_local > 0 && (!imgs || !w || !h || !slot))) return -1;
    ctx->st->add_images(imgs, w, h, n_local, on_device != 0, slot);
    return 0;
    PB_API_END
}
int pano_b200_shard_export(pano_b200_ctx* ctx, int i, float* d_descr_out, pano_b200_keypoint* keys_out, uint8_t* d_proj_out) {
    PB_API_BEGIN
    ctx->st->shard_export(i, d_descr_out, reinterpret_cast<VlKey*>(keys_out), d_proj_out);
    return 0;
    PB_API_END
}
int pano_b200_shard_import(pano_b200_ctx* ctx, int i, int w, int h, int nfeat, const float* d_descr,
                           const pano_b200_keypoint* keys, const uint8_t* d_proj) {
    PB_API_BEGIN
    if (nfeat < 0 || w <= 0 || h <= 0 || (nfeat > 0 && (!d_descr || !keys))) return -1;
    ctx->st->shard_import(i, w, h, nfeat, d_descr, reinterpret_cast<const VlKey*>(keys), d_proj);
    return 0;
    PB_API_END
}
int pano_b200_shard_match(pano_b200_ctx* ctx, const int* I, const int* J, int nprob, int* d_idx_out) {
    PB_API_BEGIN
    if (nprob < 0 || (nprob > 0 && (!I || !J))) return -1;
    ctx->st->shard_match(I, J, nprob, d_idx_out);
    return 0;
    PB_API_END
}
int pano_b200_shard_preset(pano_b200_ctx* ctx, int i, int j, const int* idx, int n) {
    PB_API_BEGIN
    if (n < 0 || (n > 0 && !idx)) return -1;
    ctx->st->preset_match(i, j, idx, n);
    return 0;
    PB_API_END
}
int pano_b200_shard_stitch(pano_b200_ctx* ctx, uint8_t* out, size_t out_cap, int* out_w, int* out_h) {
    PB_API_BEGIN
    ctx->err.clear();
    Stitcher& S = *ctx->st;
    int rc = S.run();
    if (rc) { ctx->err = S.error(); return rc; }
    if (out_w) *out_w = S.result_width();
    if (out_h) *out_h = S.result_height();
    if (out) {
        if (out_cap < (size_t)3 * S.result_width() * S.result_height()) return -4;
        S.copy_result(out);
    }
    return 0;
    PB_API_END
}
int pano_b200_shard_stitch_planes(pano_b200_ctx* ctx, int first_plane, int nplanes, pano_b200_seam_exchange exchange,
                                  void* user, int* out_w, int* out_h) {
    PB_API_BEGIN
    ctx->err.clear();
    Stitcher& S = *ctx->st;
    int rc = S.run_planes(first_plane, nplanes, exchange, user);
    if (rc) { ctx->err = S.error(); return rc; }
    if (out_w) *out_w = S.result_width();
    if (out_h) *out_h = S.result_height();
    return 0;
    PB_API_END
}
int pano_b200_shard_plane_export(pano_b200_ctx* ctx, int k, uint8_t* d_out) {
    PB_API_BEGIN
    ctx->st->plane_export(k, d_out);
    return 0;
    PB_API_END
}
int pano_b200_shard_plane_import(pano_b200_ctx* ctx, int channel, const uint8_t* d_in) {
    PB_API_BEGIN
    ctx->st->plane_import(channel, d_in);
    return 0;
    PB_API_END
}
int pano_b200_shard_tail(pano_b200_ctx* ctx, uint8_t* out, size_t out_cap, int* out_w, int* out_h) {
    PB_API_BEGIN
    ctx->err.clear();
    Stitcher& S = *ctx->st;
    int rc = S.run_tail();
    if (rc) { ctx->err = S.error(); return rc; }
    if (out_w) *out_w = S.result_width();
    if (out_h) *out_h = S.result_height();
    if (out) {
        if (out_cap < (size_t)3 * S.result_width() * S.result_height()) return -4;
        S.copy_result(out);
    }
    return 0;
    PB_API_END
}
int pano_b200_color_transfer(pano_b200_ctx* ctx, const uint8_t* src, int w, int h, const uint8_t* tem, int tw, int th,
                             uint8_t* out) {
    PB_API_BEGIN
    if (!src || !tem || !out || w <= 0 || h <= 0 || tw <= 0 || th <= 0) return -1;
    ctx->st->color_transfer(src, w, h, tem, tw, th, out);
    return 0;
    PB_API_END
}
int pano_b200_bench_match_u8_peak(pano_b200_ctx* ctx, float* ms, int* ksteps) {
    PB_API_BEGIN
    if (ms) *ms = ctx->st->last_u8_mma_only_ms_;
    if (ksteps) *ksteps = ctx->st->last_u8_ksteps_;
    return 0;
    PB_API_END
}
int pano_b200_stitch_bmp(pano_b200_ctx* ctx, const uint8_t* const* files, const size_t* sizes, int n, uint8_t** out_bmp,
                         size_t* out_size) {
    PB_API_BEGIN
    ctx->err.clear();
    int rc = ctx->st->stitch_bmp(files, sizes, n, out_bmp, out_size);
    if (rc) ctx->err = ctx->st->error();
    return rc;
    PB_API_END
}
int pano_b200_stage_images(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, int n) {
    PB_API_BEGIN
    ctx->st->stage_images(imgs, w, h, n);
    return 0;
    PB_API_END
}
int pano_b200_stitch_staged(pano_b200_ctx* ctx, int* out_w, int* out_h) {
    PB_API_BEGIN
    int rc = ctx->st->run_staged();
    if (rc) { ctx->err = ctx->st->error(); return rc; }
    if (out_w) *out_w = ctx->st->result_width();
    if (out_h) *out_h = ctx->st->result_height();
    return 0;
    PB_API_END
}
int pano_b200_result_copy(pano_b200_ctx* ctx, uint8_t* out) {
    PB_API_BEGIN
    ctx->st->copy_result(out);
    return 0;
    PB_API_END
}
int pano_b200_stitch_into(pano_b200_ctx* ctx, const uint8_t* const* imgs, const int* w, const int* h, int n,
                          uint8_t* out, size_t out_cap, int* out_w, int* out_h) {
    PB_API_BEGIN
    ctx->err.clear();
    Stitcher& S = *ctx->st;
    S.clear();
    S.add_images(imgs, w, h, n, false);
    int rc = S.run();
    if (rc) { ctx->err = S.error(); return rc; }
    *out_w = S.result_width();
    *out_h = S.result_height();
    if ((size_t)3 * *out_w * *out_h > out_cap) { ctx->err = "output buffer too small"; return -5; }
    S.copy_result(out);
    return 0;
    PB_API_END
}
void* pano_b200_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void pano_b200_free_pinned(void* p) { if (p) cudaFreeHost(p); }
int pano_b200_set_profile(pano_b200_ctx* ctx, int variant, unsigned ransac_seed) {
    PB_API_BEGIN
    if (variant != PANO_B200_PROFILE_ROOT && variant != PANO_B200_PROFILE_EX6) {
        ctx->err = "unknown profile";
        return -1;
    }
    pb::stitch::Profile p;
    p.variant = variant;
    p.ransac_seed = ransac_seed;
    ctx->st->set_profile(p);
    return 0;
    PB_API_END
}
int pano_b200_set_lanes(pano_b200_ctx* ctx, int nlanes) {
    PB_API_BEGIN
    ctx->st->set_lanes(nlanes);
    return 0;
    PB_API_END
}
int pano_b200_set_match_mode(pano_b200_ctx* ctx, int mode) {
    PB_API_BEGIN
    if (mode != PANO_B200_MATCH_PREFILTER && mode != PANO_B200_MATCH_FULL && mode != PANO_B200_MATCH_PREFILTER_ONEDIR &&
        mode != PANO_B200_MATCH_PREFILTER_FULLSAD)
        return -1;
    ctx->st->set_match_mode(mode);
    return 0;
    PB_API_END
}
int pano_b200_match_stats(pano_b200_ctx* ctx, long long out[5], int reset) {
    PB_API_BEGIN
    const MatchStats& m = ctx->st->match_stats();
    if (out) { out[0] = m.queries; out[1] = m.survivors; out[2] = m.overflow; out[3] = m.problems; out[4] = m.sym_pairs; }
    if (reset) ctx->st->reset_match_stats();
    return 0;
    PB_API_END
}
int pano_b200_match_stats_ex(pano_b200_ctx* ctx, long long* out, int n, int reset) {
    PB_API_BEGIN
    const MatchStats& m = ctx->st->match_stats();
    const long long v[9] = {m.queries, m.survivors, m.overflow, m.problems, m.sym_pairs, m.group_pairs, m.group_exact,
                            m.group_accepts, m.group_overflow};
    for (int i = 0; out && i < n && i < 9; ++i) out[i] = v[i];
    if (reset) ctx->st->reset_match_stats();
    return 0;
    PB_API_END
}
int pano_b200_flush_l2(pano_b200_ctx* ctx) {
    PB_API_BEGIN
    ctx->st->flush_l2();
    return 0;
    PB_API_END
}
int pano_b200_timer_start(pano_b200_ctx* ctx) {
    PB_API_BEGIN
    ctx->st->timer_start();
    return 0;
    PB_API_END
}
int pano_b200_timer_stop(pano_b200_ctx* ctx, float* ms) {
    PB_API_BEGIN
    *ms = ctx->st->timer_stop();
    return 0;
    PB_API_END
}
void pano_b200_ktimer_enable(int on) { KTimer::get().enable(on != 0); }
void pano_b200_ktimer_reset(void) { KTimer::get().reset(); }
long pano_b200_ktimer_launches(void) { return KTimer::get().total_launches(); }
int pano_b200_ktimer_report(char* dst, int cap) {
    if (!dst || cap <= 0) return -1;
    try {
    std::ostringstream o;
    o << "{";
    bool first = true;
    for (auto& kv : KTimer::get().snapshot()) {
        if (!first) o << ", ";
        first = false;
        o << "\"" << kv.first << "\": {\"launches\": " << kv.second.launches << ", \"ms\": " << kv.second.ms
          << ", \"bytes\": " << kv.second.bytes << "}";
    }
    o << "}";
    std::string s = o.str();
    int n = (int)s.size();
    if (n >= cap) n = cap - 1;
    memcpy(dst, s.data(), n);
    dst[n] = 0;
    return n;
    } catch (...) {
        dst[0] = 0;
        return -100;
    }
}

int pano_b200_stitch_log(pano_b200_ctx* ctx, char* dst, int cap) {
    if (!ctx || !dst || cap <= 0) return -1;
    PB_API_BEGIN
    const std::string& s = ctx->st->log();
    int n = (int)s.size();
    if (n >= cap) n = cap - 1;
    memcpy(dst, s.data(), n);
    dst[n] = 0;
    return n;
    PB_API_END
}
int pano_b200_stitch_times(pano_b200_ctx* ctx, pano_b200_times* t) {
    if (!ctx || !t) return -1;
    const StageTimes& s = ctx->st->times();
    t->project = s.project; t->sift = s.sift; t->table = s.table; t->match = s.match; t->ransac = s.ransac;
    t->warp = s.warp; t->blend = s.blend; t->tail = s.tail; t->total = s.total;
    t->match_pairs_evaluated = s.match_pairs_evaluated; t->sift_pixels = s.sift_pixels;
    t->n_match_calls = s.n_match_calls; t->n_blends = s.n_blends;
    return 0;
}
int pano_b200_stitch_nfeatures(pano_b200_ctx* ctx, int image) {
    if (!ctx || image < 0 || image >= ctx->st->num_images()) return -1;
    return ctx->st->features(image).n;
}

int pano_b200_quantize_u8(pano_b200_ctx* ctx, const float* descr, int n, uint8_t* out) {
    PB_API_BEGIN
    ctx->st->quantize_u8(descr, n, out);
    return 0;
    PB_API_END
}
int pano_b200_match_u8(pano_b200_ctx* ctx, const uint8_t* descrA, int nA, const uint8_t* descrB, int nB, int* match_idx,
                       int* d01, int* nmatches) {
    PB_API_BEGIN
    ctx->st->match_u8(descrA, nA, descrB, nB, match_idx, d01);
    int c = 0;
    for (int b = 0; b < nB; ++b) c += match_idx[b] >= 0;
    if (nmatches) *nmatches = c;
    return 0;
    PB_API_END
}
int pano_b200_bench_match_u8(pano_b200_ctx* ctx, const uint8_t* descrA, int nA, const uint8_t* descrB, int nB, int reps,
                             float* ms_per_rep) {
    PB_API_BEGIN
    *ms_per_rep = ctx->st->bench_match_u8(descrA, nA, descrB, nB, reps);
    return 0;
    PB_API_END
}

int pano_b200_project(pano_b200_ctx* ctx, const uint8_t* rgb, int w, int h, uint8_t* out_rgb, uint8_t* out_gray) {
    PB_API_BEGIN
    ctx->st->project(rgb, w, h, out_rgb, out_gray);
    return 0;
    PB_API_END
}
int pano_b200_gray(pano_b200_ctx* ctx, const uint8_t* rgb, int w, int h, uint8_t* out_gray) {
    PB_API_BEGIN
    ctx->st->gray(rgb, w, h, out_gray);
    return 0;
    PB_API_END
}

int pano_b200_sift_features(pano_b200_ctx* ctx, const uint8_t* gray, int w, int h, float** descr,
                            pano_b200_keypoint** keys, int* n) {
    PB_API_BEGIN
    RawFeatures raw;
    ctx->st->sift_raw_u8(gray, w, h, SiftParams(), raw);
    FeatureTable t;
    Stitcher::build_table(raw, t);
    *n = t.n;
    *descr = (float*)xmalloc(std::max<size_t>((size_t)t.n * 128 * sizeof(float), 4));
    *keys = (pano_b200_keypoint*)xmalloc(std::max<size_t>((size_t)t.n * sizeof(VlKey), 4));
    memcpy(*descr, t.descr.data(), (size_t)t.n * 128 * sizeof(float));
    memcpy(*keys, t.keys.data(), (size_t)t.n * sizeof(VlKey));
    return 0;
    PB_API_END
}

int pano_b200_sift_raw(pano_b200_ctx* ctx, const float* image, int w, int h, int noctaves, int nlevels,
                       pano_b200_keypoint** keys, double** angles, float** descr, int* n, int* octave_nkeys) {
    PB_API_BEGIN
    SiftParams p;
    p.O = noctaves;
    p.S = nlevels;
    RawFeatures& raw = ctx->last_raw;
    ctx->st->sift_raw_f32(image, w, h, p, raw);
    *n = raw.n;
    *keys = (pano_b200_keypoint*)xmalloc(std::max<size_t>((size_t)raw.n * sizeof(VlKey), 4));
    *angles = (double*)malloc(std::max<size_t>((size_t)raw.n * sizeof(double), 8));
    *descr = (float*)xmalloc(std::max<size_t>((size_t)raw.n * 128 * sizeof(float), 4));
    memcpy(*keys, raw.keys.data(), (size_t)raw.n * sizeof(VlKey));
    memcpy(*angles, raw.angles.data(), (size_t)raw.n * sizeof(double));
    memcpy(*descr, raw.descr.data(), (size_t)raw.n * 128 * sizeof(float));
    if (octave_nkeys)
        for (size_t i = 0; i < raw.noct_keys.size(); ++i) octave_nkeys[i] = raw.noct_keys[i];
    return 0;
    PB_API_END
}

int pano_b200_sift_octave_dims(pano_b200_ctx* ctx, int octave, int* ow, int* oh) {
    SiftEngine& E = ctx->st->sift_engine();
    if (octave < 0 || octave >= E.noctaves()) return -1;
    *ow = E.octave(octave).w;
    *oh = E.octave(octave).h;
    return 0;
}
int pano_b200_sift_octave_dump(pano_b200_ctx* ctx, int octave, float* gss, float* grad) {
    PB_API_BEGIN
    SiftEngine& E = ctx->st->sift_engine();
    if (octave < 0 || octave >= E.noctaves()) return -1;
    OctaveBuf& ob = E.octave(octave);
    cudaStream_t st = E.stream();
    const int nl = E.nlevels();
    if (gss)
        PB_CUDA(cudaMemcpy2DAsync(gss, (size_t)ob.w * 4, ob.gss.p, (size_t)ob.pitch * 4, (size_t)ob.w * 4,
                                  (size_t)ob.h * nl, cudaMemcpyDeviceToHost, st));
    if (grad)
        PB_CUDA(cudaMemcpy2DAsync(grad, (size_t)ob.w * 8, ob.grad.p, (size_t)ob.pitch * 8, (size_t)ob.w * 8,
                                  (size_t)ob.h * (nl - 3), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    return 0;
    PB_API_END
}

int pano_b200_match(pano_b200_ctx* ctx, const float* descrA, int nA, const float* descrB, int nB, int* match_idx,
                    int* nmatches) {
    PB_API_BEGIN
    FeatureTable &A = ctx->match_a, &B = ctx->match_b;
    A.n = nA; A.descr.assign(descrA, descrA + (size_t)nA * 128); A.on_device = false; A.quantised = false;
    B.n = nB; B.descr.assign(descrB, descrB + (size_t)nB * 128); B.on_device = false; B.quantised = false;
    std::vector<int> idx;
    ctx->st->match_idx(A, B, idx);
    int c = 0;
    for (int b = 0; b < nB; ++b) { match_idx[b] = idx[b]; c += idx[b] >= 0; }
    if (nmatches) *nmatches = c;
    return 0;
    PB_API_END
}

int pano_b200_match_pair(pano_b200_ctx* ctx, const float* descrA, int nA, const float* descrB, int nB, int* idx_ab,
                         int* idx_ba) {
    PB_API_BEGIN
    if (nA < 0 || nB < 0 || (nA > 0 && !descrA) || (nB > 0 && !descrB)) return -1;
    FeatureTable &A = ctx->match_a, &B = ctx->match_b;
    A.n = nA; A.descr.assign(descrA, descrA + (size_t)nA * 128); A.on_device = false; A.quantised = false;
    B.n = nB; B.descr.assign(descrB, descrB + (size_t)nB * 128); B.on_device = false; B.quantised = false;
    std::vector<std::pair<FeatureTable*, FeatureTable*>> probs{{&A, &B}, {&B, &A}};
    std::vector<std::vector<int>> out;
    ctx->st->match_batch(probs, out);
    if (idx_ab) std::copy(out[0].begin(), out[0].end(), idx_ab);
    if (idx_ba) std::copy(out[1].begin(), out[1].end(), idx_ba);
    return 0;
    PB_API_END
}

int pano_b200_ransac(pano_b200_ctx* ctx, const pano_b200_pair* pairs, int npairs, double* H8, int* counts,
                     double* hyps, int* inliers, int* ninliers) {
    PB_API_BEGIN
    std::vector<KeyPair> p((const KeyPair*)pairs, (const KeyPair*)pairs + npairs);
    std::vector<int> c, inl;
    std::vector<double> hy;
    if (!ctx->st->ransac_debug(p, c, hy, inl, H8)) { ctx->err = "RANSAC failed (fewer than 4 pairs or no inliers)"; return -3; }
    if (counts) memcpy(counts, c.data(), c.size() * sizeof(int));
    if (hyps) memcpy(hyps, hy.data(), hy.size() * sizeof(double));
    if (inliers) memcpy(inliers, inl.data(), inl.size() * sizeof(int));
    if (ninliers) *ninliers = (int)inl.size();
    return 0;
    PB_API_END
}

int pano_b200_plan_canvas(int dst_w, int dst_h, const double* forward_H8, int result_w, int result_h, float* bounds,
                          int* size) {
    stitch::CanvasPlan p = stitch::plan_canvas(dst_w, dst_h, forward_H8, result_w, result_h);
    bounds[0] = p.min_x; bounds[1] = p.min_y; bounds[2] = p.max_x; bounds[3] = p.max_y;
    size[0] = p.new_w; size[1] = p.new_h;
    return 0;
}

int pano_b200_plan_canvas_ex(int variant, int dst_w, int dst_h, const double* forward_H8, int result_w, int result_h,
                             float* bounds, int* size) {
    stitch::CanvasPlan p =
        stitch::plan_canvas(dst_w, dst_h, forward_H8, result_w, result_h, variant == PANO_B200_PROFILE_EX6);
    bounds[0] = p.min_x; bounds[1] = p.min_y; bounds[2] = p.max_x; bounds[3] = p.max_y;
    size[0] = p.new_w; size[1] = p.new_h;
    return 0;
}

int pano_b200_warp_shift(pano_b200_ctx* ctx, const uint8_t* src, int sw, int sh, const double* H8, float offx,
                         float offy, const uint8_t* prev, int pw, int ph, int ioffx, int ioffy, int cw, int ch,
                         uint8_t* a, uint8_t* b) {
    PB_API_BEGIN
    ctx->st->warp_shift(src, sw, sh, H8, offx, offy, prev, pw, ph, ioffx, ioffy, cw, ch, a, b);
    return 0;
    PB_API_END
}

int pano_b200_blend(pano_b200_ctx* ctx, const uint8_t* a, const uint8_t* b, int w, int h, uint8_t* out) {
    PB_API_BEGIN
    int rc = ctx->st->blend(a, b, w, h, out);
    if (rc) ctx->err = ctx->st->error();
    return rc;
    PB_API_END
}
int pano_b200_equalize_mix(pano_b200_ctx* ctx, const uint8_t* rgb, int w, int h, uint8_t* out) {
    PB_API_BEGIN
    ctx->st->equalize_mix(rgb, w, h, out);
    return 0;
    PB_API_END
}
int pano_b200_cimg_blur2(pano_b200_ctx* ctx, const float* src, int w, int h, int c, float* dst) {
    PB_API_BEGIN
    ctx->st->cimg_blur2(src, w, h, c, dst);
    return 0;
    PB_API_END
}
int pano_b200_cimg_blur2_deriche(pano_b200_ctx* ctx, const float* src, int w, int h, int c, float* dst) {
    PB_API_BEGIN
    ctx->st->cimg_blur2_deriche(src, w, h, c, dst);
    return 0;
    PB_API_END
}
int pano_b200_cimg_resize3(pano_b200_ctx* ctx, const float* src, int w, int h, int c, int nw, int nh, float* dst) {
    PB_API_BEGIN
    ctx->st->cimg_resize(src, w, h, c, nw, nh, dst);
    return 0;
    PB_API_END
}

}  // extern "C"
