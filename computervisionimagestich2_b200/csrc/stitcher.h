// stitcher.h -- the host orchestrator of the GPU panorama path: the B200 counterpart of class ImageProcess
// (ImageProcess.h:77-146).  One Stitcher = one CUDA device + one stream + grow-only HBM workspaces.
//
//   readFile()  (ImageProcess.cpp:11-24)   -> add_image(): H2D, projection+gray kernel, SIFT engine, feature table
//   matching()  (ImageProcess.cpp:101-271) -> run(): all-pairs adjacency, middle image, BFS stitch loop, tail
// Every stage is also reachable on its own (host buffers in / out) for the parity tests and stage benchmarks.
#pragma once
#include <memory>
#include <string>
#include <sstream>
#include <vector>
#include "common.h"
#include "types.h"
#include "sift_engine.h"
#include "match_kernels.h"
#include "match_i8_kernels.h"
#include "canvas_kernels.h"
#include "stitch_host.h"

namespace pb {

struct FeatureTable {   // == std::map<std::vector<float>, VlSiftKeypoint> flattened in key order (ImageProcess.h:54)
    int n = 0;
    std::vector<float> descr;   // [n][128], lexicographically sorted, duplicates removed (first insertion kept)
    std::vector<VlKey> keys;    // ix, iy = (int)x, (int)y (ImageProcess.cpp:84-85); coordinates move during stitching
    DevBuf<float> d_descr;      // device copy of descr
    bool on_device = false;
    // quantised copy for the matcher's pre-filter (match_device.cuh): 128 bytes per row + the row's error bound
    DevBuf<unsigned> d_q8;
    DevBuf<int> d_qe;
    bool quantised = false;
    int qemax = 0;              // largest error bound of the table (decides whether the 16-bit symmetric pass applies)
    // grouped pass: 32 group bytes per row and w16, both readable in whole 64-row tiles (match_group_pad_rows)
    DevBuf<unsigned> d_g8;
    DevBuf<unsigned short> d_w16;
};

struct StageTimes {  // milliseconds, CUDA events on the stitcher's stream (host work between kernels included)
    double project = 0, sift = 0, table = 0, match = 0, ransac = 0, warp = 0, blend = 0, tail = 0, total = 0;
    long match_pairs_evaluated = 0;   // sum of NA*NB over all getImgPair calls
    long sift_pixels = 0;
    int n_match_calls = 0, n_blends = 0;
};

struct MatchStats {   // pre-filter bookkeeping since the last clear(): queries, survivors of the SAD pass, full-scan fallbacks
    long long queries = 0, survivors = 0, overflow = 0, problems = 0, sym_pairs = 0;
    long long group_pairs = 0, group_exact = 0;   // image pairs through the grouped pass; exact SADs it evaluated
    long long group_accepts = 0;                  // queries accepted with certainty by the grouped pass (no float arithmetic)
    long long group_overflow = 0;                 // batches redone with the full SAD pass because the pair queue overflowed
};

class Stitcher {
  public:
    explicit Stitcher(int device);
    ~Stitcher();
    cudaStream_t stream() const { return st_; }
    // which caller of the hot path is reproduced (root ImageProcess.cpp or src/ex6) and the RANSAC seed; applies to
    // run(), ransac(), blend(), equalize_mix()
    void set_profile(const stitch::Profile& p) { profile_ = p; }
    const stitch::Profile& profile() const { return profile_; }
    SiftEngine& sift_engine() { return *sift_; }

    // ---- stages (host in / host out) ---------------------------------------------------------------------
    void project(const u8* rgb, int w, int h, u8* out_rgb, u8* out_gray);
    void gray(const u8* rgb, int w, int h, u8* out_gray);
    void sift_raw_u8(const u8* gray8, int w, int h, const SiftParams& p, RawFeatures& out);
    void sift_raw_f32(const float* img, int w, int h, const SiftParams& p, RawFeatures& out);
    // sel (optional): raw feature index of every table row; host_descr = false leaves t.descr empty (the pipeline only
    // needs the device copy, which it gathers from the engine's buffer)
    static void build_table(const RawFeatures& raw, FeatureTable& t, std::vector<int>* sel = nullptr, bool host_descr = true);
    void upload_table(FeatureTable& t);
    // idx[b] = row of A matched by query row b of B, or -1 (ImageProcess.cpp:311-346)
    void match_idx(FeatureTable& A, FeatureTable& B, std::vector<int>& idx);
    // d_out (optional, device): the match lists concatenated in problem order (B.n ints each, -1 rows for the
    // degenerate problems); the host copies in `out` are always filled
    void match_batch(const std::vector<std::pair<FeatureTable*, FeatureTable*>>& probs, std::vector<std::vector<int>>& out,
                     int* d_out = nullptr);
    void match(FeatureTable& A, FeatureTable& B, std::vector<KeyPair>& pairs);
    // several RANSAC problems in one launch; returns false for a problem the reference cannot solve (<4 pairs ...)
    bool ransac(const std::vector<const std::vector<KeyPair>*>& problems, std::vector<double>& H8s);
    bool ransac_debug(const std::vector<KeyPair>& pairs, std::vector<int>& counts, std::vector<double>& hyps,
                      std::vector<int>& best_inliers, double* H8);
    void warp_shift(const u8* src, int sw, int sh, const double* H8, float offx, float offy, const u8* prev, int pw,
                    int ph, int ioffx, int ioffy, int cw, int ch, u8* a_out, u8* b_out);
    int blend(const u8* a, const u8* b, int cw, int ch, u8* out);          // host buffers
    void equalize_mix(const u8* rgb, int w, int h, u8* out);               // host buffers
    // uint8 / tcgen05 matcher (north-star stage 3; not on the reference-parity path): host tables in, host results out
    void quantize_u8(const float* descr, int n, u8* out);
    void match_u8(const u8* A, int nA, const u8* B, int nB, int* idx, int* d01);
    // ms per repetition of the matcher kernels on resident tables (A, B row-major host tables, or NULL = uniform bytes)
    float bench_match_u8(const u8* A, int nA, const u8* B, int nB, int reps);
    // set by bench_match_u8: ms per repetition of the MMA-only run (tensor-pipe peak of the kernel's own instruction
    // stream) and the K steps of 32 bytes each MMA tile runs (4 descriptor steps + the norm extension)
    float last_u8_mma_only_ms_ = 0;
    int last_u8_ksteps_ = 4;
    // Reinhard colour transfer of `src` towards `tem` (the reference's class transfer, dead code there: SURVEY 8f rank 3)
    void color_transfer(const u8* src, int w, int h, const u8* tem, int tw, int th, u8* out);
    void cimg_blur2(const float* src, int w, int h, int c, float* dst);    // get_blur(2,true,true), host buffers
    void cimg_blur2_deriche(const float* src, int w, int h, int c, float* dst);   // get_blur(2) (src/ex6), host buffers
    void cimg_resize(const float* src, int w, int h, int c, int nw, int nh, float* dst);

    // ---- pipeline ---------------------------------------------------------------------------------------
    void clear();
    void add_image(const u8* rgb, int w, int h);     // planar RGB host buffer
    void add_image_device(const u8* d_rgb, int w, int h);   // planar RGB already resident in HBM
    // readFile() for all images at once (ImageProcess.cpp:12-23): images are independent until matching, so they
    // are processed concurrently, image i on lane i % nlanes (one CUDA stream + SIFT engine + host thread per lane).
    // on_device: imgs[i] are HBM pointers (staged inputs) instead of host buffers.
    // slots (optional): imgs_[slots[i]] receives image i (the slots must exist: shard_begin); default = n appended slots
    void add_images(const u8* const* imgs, const int* w, const int* h, int n, bool on_device, const int* slots = nullptr);
    void set_lanes(int n) { want_lanes_ = n < 1 ? 1 : n; }
    // matcher: 0 = rigorous uint8 pre-filter + exact float re-rank, both directions of an image pair from one SAD pass
    // (default); 1 = full exact float scan (the round-1 kernel, kept as the cross-check); 2 = pre-filter, one SAD pass
    // per directed problem.  All give the reference's match lists bit for bit.
    void set_match_mode(int m) { match_mode_ = m; }
    int match_mode() const { return match_mode_; }
    const MatchStats& match_stats() const { return mstats_; }
    void reset_match_stats() { mstats_ = MatchStats(); }
    // ---- BMP files in / BMP file out: decode and encode run on the GPU ----------------------------
    // files[i]: the bytes of an uncompressed 24-bpp BMP.  Returns 0 and the encoded panorama (malloc'ed), or < 0.
    int stitch_bmp(const u8* const* files, const size_t* sizes, int n, u8** out, size_t* out_size);
    // ---- sharded jobs (features / matches computed on other GPUs) --------------------------------
    void extract(const u8* rgb, int w, int h, u8* proj_out, FeatureTable& t);    // one image -> host projection + table
    void add_precomputed(const u8* proj_rgb, int w, int h, const float* descr, const VlKey* keys, int n);
    void preset_match(int i, int j, const int* idx, int nB);   // getImgPair(imgs[i], imgs[j]) indices, evaluated elsewhere
    // device-resident exchange of a sharded job (see stitcher.cu): all pointers named d_* are device memory
    void shard_begin(int n_global);
    void shard_export(int i, float* d_descr_out, VlKey* keys_out, u8* d_proj_out);
    void shard_import(int i, int w, int h, int n, const float* d_descr, const VlKey* keys, const u8* d_proj);
    void shard_match(const int* I, const int* J, int nprob, int* d_idx_out);
    // plane-sharded canvas stages: run() on `count` colour planes starting at `first` (3 planes from 0 = everything but
    // the tail; 1 plane = a third of the pixel work).  cb hands the 16-byte seam statistics of plane 0 to the stitchers
    // that do not carry it: called once per edge on every participant, is_source = 1 where stats4 holds the values.
    typedef int (*SeamExchange)(void* user, int* stats4, int is_source);
    int run_planes(int first, int count, SeamExchange cb, void* user);
    void plane_export(int k, u8* d_out);              // plane k of this stitcher's result -> device buffer
    void plane_import(int channel, const u8* d_in);   // colour plane `channel` computed elsewhere
    int run_tail();                                   // equalisation + mix on the assembled canvas
    int image_width(int i) const { return imgs_[i]->w; }
    int image_height(int i) const { return imgs_[i]->h; }
    // ---- batched independent pairs (BASELINE configs[4]): imgs[2p], imgs[2p+1]; see pano_b200_pairs ----------------
    struct PairRecord { long long pair; int nfeat[2], nmatch[2], has_h[2]; double H[2][8]; };
    int pairs(const u8* const* imgs, const int* w, const int* h, int npairs, PairRecord* out, bool on_device = false);
    // inputs staged in HBM once (outside any timed region), then stitched any number of times
    void stage_images(const u8* const* imgs, const int* w, const int* h, int n);
    int run_staged();
    void flush_l2();
    void timer_start();
    float timer_stop();
    int run();                                       // 0 ok; fills result
    int result_width() const { return rw_; }
    int result_height() const { return rh_; }
    void copy_result(u8* dst);                        // planar RGB
    const std::string& log() const { return log_; }
    const StageTimes& times() const { return tm_; }
    int num_images() const { return (int)imgs_.size(); }
    const FeatureTable& features(int i) const { return imgs_[i]->feat; }
    const std::string& error() const { return err_; }

  private:
    struct Image {
        int w = 0, h = 0;
        DevBuf<u8> proj;   // projected planar RGB
        FeatureTable feat;
    };
    // defer_check: inside run() the empty-middle-row flag accumulates in blend_flag_ and is read once after the last edge
    // (no host round trip per blend); the stage API checks at once
    // nch: colour planes in d_a / d_b / d_out (3, or 1 in a plane-sharded job); own_stats: this stitcher carries plane 0
    // and computes the seam statistics itself
    int blend_device(const u8* d_a, const u8* d_b, int cw, int ch, u8* d_out, bool defer_check = false, int nch = 3,
                     bool own_stats = true);
    int check_blend_flag();
    void equalize_mix_device(const u8* d_rgb, int w, int h, u8* d_out);
    void ensure_ktab(int short_side);

    int run_chain(std::ostringstream& log);   // the src/ex6 stitch order
    int stitch_edge(int src, int dst, int pre, const std::vector<int>& s2d_idx, const std::vector<int>& d2s_idx);
    int dev_;
    stitch::Profile profile_;
    cudaStream_t st_ = nullptr;
    // RANSAC of an edge depends on features only, not on the canvas: it runs on its own stream so that the host never
    // waits for the previous blend; the two events guard the pinned staging buffers the host then re-writes ahead of
    // the GPU (resampling tables, warp coefficients)
    cudaStream_t rst_ = nullptr;
    cudaEvent_t ev_tables_ = nullptr, ev_h8_ = nullptr;
    PinBuf<double> h_H8_;
    std::unique_ptr<SiftEngine> sift_;
    std::vector<std::unique_ptr<Image>> imgs_;
    std::vector<std::unique_ptr<Image>> pool_;   // recycled Image objects (their HBM buffers stay allocated)
    // workspaces
    DevBuf<u8> in_rgb_, a_, b_, res_[2], tmp8_;
    DevBuf<float> gray32_, ktab_;
    int ktab_n_ = 0;
    DevBuf<Top2> partial_;
    DevBuf<SadStat> spartial_;
    DevBuf<int> mscratch_, mcount_, gscratch_;
    DevBuf<unsigned long long> gqueue_;
    bool no_group_ = false;
    PinBuf<int> h_qemax_;
    int match_mode_ = 0;
    MatchStats mstats_;
    void quantise_table(FeatureTable& t);
    DevBuf<int> midx_;
    PinBuf<int> h_midx_;
    DevBuf<char> mjobs_;        // job table + pair list + single list of a matching batch
    DevBuf<u8> u8a_, u8b_, u8raw_;
    DevBuf<int> u8na_, u8nb_, u8idx_, u8d01_, u8scratch_;
    DevBuf<U8Top3> u8part_;
    void u8_upload(const u8* A, int nA, const u8* B, int nB, U8Table& TA, U8Table& TB);
    DevBuf<float> u8f_;
    PinBuf<char> h_mjobs_;
    DevBuf<KeyPair> r_pairs_;
    DevBuf<int> r_off_, r_samples_, r_counts_;
    DevBuf<unsigned> r_masks_;
    DevBuf<double> r_hyp_, H8_;
    DevBuf<float> pyr_, tmpf_, tmpf2_, E_[2];
    DevBuf<int> tab_i_, stats_, hist_, lut_, blend_flag_;
    PinBuf<int> h_tab_i_;      // pinned staging of the resampling tables: their upload must not synchronise the stream
    PinBuf<float> h_tab_f_;
    PinBuf<double> h_tab_d_;
    DevBuf<float> tab_f_;
    DevBuf<double> tab_d_;
    int cur_ = 0, rw_ = 0, rh_ = 0;
    int cplane0_ = 0, cplanes_ = 3;            // colour planes carried through warp / shift / blend
    int planes_first_ = 0, planes_count_ = 3;  // what res_[cur_] holds after run_planes
    bool skip_tail_ = false;
    SeamExchange seam_exchange_ = nullptr;
    void* seam_user_ = nullptr;
    PinBuf<int> h_stats_;
    struct Lane {   // per-worker resources for concurrent feature extraction
        cudaStream_t st = nullptr;
        std::unique_ptr<SiftEngine> eng;
        DevBuf<u8> in_rgb;
        DevBuf<float> gray32, ktab;
        DevBuf<int> rows;          // device rows of the sorted table (gather on the device instead of re-uploading it)
        PinBuf<int> h_rows;
        int ktab_n = 0;
        double t_project = 0, t_sift = 0, t_table = 0;
        std::string err;
    };
    void lane_work(Lane& L, int first, int step, const u8* const* imgs, const int* w, const int* h, int n, bool on_device,
                   const int* slots);
    std::unique_ptr<Image> new_image();
    void upload_table_on(FeatureTable& t, cudaStream_t st);
    struct PresetMatch { int i, j; std::vector<int> idx; };
    bool preset_fits(const PresetMatch& pm) const;
    std::vector<PresetMatch> preset_;
    std::vector<std::unique_ptr<Lane>> lanes_;
    int want_lanes_ = 8;   // 8 x 4K job: 97.5 / 93.1 / 88.9 / 86.6 / 84.3 ms with 2 / 3 / 4 / 6 / 8 lanes (the host phases of one image overlap the kernels of the others)
    struct Staged { int w, h; DevBuf<u8> rgb, raw; };   // raw: the BMP pixel area as uploaded (stitch_bmp), freed after decoding
    std::vector<std::unique_ptr<Staged>> staged_;
    DevBuf<u8> bmp_out_;
    DevBuf<char> flush_;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    std::string log_, err_;
    StageTimes tm_;
};

}  // namespace pb
