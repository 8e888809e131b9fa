// sift_engine.h -- host side of the SIFT stage: owns the per-octave HBM buffers of one image, sequences the kernels
// of sift_kernels.cu on one CUDA stream, and evaluates on the host (glibc) the few transcendentals the reference
// evaluates with libm: Gaussian taps (exp), keypoint sigma (pow), descriptor frame (sin/cos), expn table (exp).
//
// Replaces, as one batched pass per image, the call sequence of ImageProcess::siftAlgorithm
// (ImageProcess.cpp:44-99): vl_sift_new / process_first_octave / {detect, calc_keypoint_orientations,
// calc_keypoint_descriptor}* / process_next_octave.
#pragma once
#include <vector>
#include <cstdint>
#include "common.h"
#include "types.h"
#include "sift_kernels.h"

namespace pb {

struct SiftParams {
    int O = 4, S = 2, o_min = 0;          // ImageProcess.h:15-16, ImageProcess.cpp:55
    double peak_thresh = 0.0, edge_thresh = 10.0, norm_thresh = 0.0, magnif = 3.0, window_size = 2.0;  // vl/sift.c:267-271
};

// Raw output of one image, in the reference's insertion order: octave, keypoint (raster order of the detected
// extremum), angle.
struct RawFeatures {
    std::vector<VlKey> keys;        // one per descriptor (keypoint repeated per angle), ix/iy as detected
    std::vector<double> angles;     // angle of each descriptor
    std::vector<int> key_index;     // index of the keypoint inside its octave
    std::vector<float> descr;       // [n][128]
    std::vector<int> noct_keys;     // refined keypoints per octave
    int n = 0;
    int dropped_unwritten = 0;      // (keypoint, angle) pairs whose descriptor the reference leaves unwritten (quirk Q3)
    // extract() only: the descriptors are still in the engine's device buffer (until its next extract on this stream);
    // descriptor i is row dev_row[i] of d_descr.  Lets the caller assemble a device table without a host round trip.
    const float* d_descr = nullptr;
    std::vector<int> dev_row;
    // extract(..., copy_descr = false): `descr` stays empty and the descriptors are read in place from the engine's pinned
    // download buffer (same lifetime as d_descr): descriptor i = h_descr + dev_row[i] * 128
    const float* h_descr = nullptr;
    const float* row(int i) const { return descr.empty() ? h_descr + (size_t)dev_row[i] * 128 : descr.data() + (size_t)i * 128; }
};

struct OctaveBuf {
    int w = 0, h = 0, pitch = 0;
    DevBuf<float> gss, grad;
    DevBuf<Cand> cand;
    DevBuf<RefinedKey> refined;
    int cand_cap = 0;
    // host mirrors of the per-octave results
    std::vector<VlKey> keys;
    std::vector<int> h_nangles;
    std::vector<double> h_angles;
    OctaveView view(int nlevels) const { return OctaveView{w, h, pitch, nlevels, gss.p, grad.p}; }
};

class SiftEngine {
  public:
    explicit SiftEngine(cudaStream_t st);
    ~SiftEngine();
    cudaStream_t stream() const { return st_; }

    // Geometry + constants for an image; (re)allocates octave buffers.  Mirrors vl_sift_new (vl/sift.c:217-279).
    void configure(int w, int h, const SiftParams& p);
    void set_thresholds(const SiftParams& p) {
        p_.peak_thresh = p.peak_thresh; p_.edge_thresh = p.edge_thresh; p_.norm_thresh = p.norm_thresh;
        p_.magnif = p.magnif; p_.window_size = p.window_size;
    }

    // --- batched path ------------------------------------------------------------------------------------
    // d_img: device float image, rows img_pitch floats apart, values 0..255.
    void extract(const float* d_img, int img_pitch, RawFeatures& out, bool copy_descr = true);
    // d_descr_out (optional): device buffer receiving the raw descriptors [n][128] in RawFeatures order.

    // --- octave-at-a-time path (vl_sift_* shim) ----------------------------------------------------------
    void load_base_from_device(const float* d_img, int img_pitch);  // level s_min of octave o_min
    void build_octave(int oi);            // blur chain of octave index oi (0-based from o_min); base must be present
    void seed_next_octave(int oi);        // write base of octave oi+1 from level s_best of octave oi
    void detect_octave(int oi);           // extrema + refinement + host sigma -> oct_[oi].keys ; also gradient map
    void orient_octave(int oi);           // -> oct_[oi].h_nangles / h_angles
    // descriptors for (key, angle) jobs of octave oi; out_descr [njobs][128], out_written [njobs]
    void describe_octave(int oi, const std::vector<int>& key_idx, const std::vector<double>& ang,
                         const std::vector<KeyIn>* custom_keys, float* out_descr, int* out_written);
    void orient_custom(int oi, const std::vector<KeyIn>& keys, std::vector<int>& nang, std::vector<double>& ang);

    int noctaves() const { return (int)oct_.size(); }
    OctaveBuf& octave(int oi) { return oct_[oi]; }
    int nlevels() const { return nlev_; }
    const SiftParams& params() const { return p_; }
    SiftConsts consts() const;
    DevBuf<float>& temp() { return temp_; }

    // host helpers (public for tests)
    static BlurTaps make_taps(double sigma);
    double presmooth_sigma_first() const;   // 0 if none
    double presmooth_sigma_next() const;    // 0 if none
    double level_sigma(int s) const;

  private:
    void blur_level(int oi, int src_l, int dst_l, double sigma, bool seed_next);
    cudaStream_t st_;
    SiftParams p_;
    int w_ = 0, h_ = 0, nlev_ = 0, s_min_ = -1, s_max_ = 3;
    double sigman_, sigma0_, sigmak_, dsigma0_;
    std::vector<OctaveBuf> oct_;
    DevBuf<float> temp_;
    DevBuf<double> expn_tab_;
    DevBuf<int> counts_;
    PinBuf<int> h_counts_;
    PinBuf<char> h_stage_;
    // keypoints / jobs / results of all octaves of the current image (one launch per stage)
    DevBuf<KeyIn> keyin_;
    DevBuf<int> nangles_, written_;
    DevBuf<double> angles_;
    DevBuf<DescJob> jobs_;
    DevBuf<float> descr_;
    PinBuf<char> h_keyin_, h_nang_, h_ang_, h_jobs_, h_descr_, h_written_, h_order_k_, h_order_j_;
    DevBuf<int> order_k_, order_j_;   // issue orders of the orientation / descriptor launches (largest window first)
    OctaveSet octave_set(int first, int count) const;
    bool tab_ready_ = false;
};

}  // namespace pb
