// tests/emul/emul.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A g++ build of the product's __host__ __device__ per-item bodies (csrc/*_device.cuh) and host numerics
// (csrc/host_numerics.h), driven by plain loops.  It lets the CPU-only test tier check the arithmetic that the CUDA
// kernels execute against the compiled reference (oracle/_ref) without a GPU.  The product library never links or
// calls this file; the GPU tier checks the real kernels through the C-ABI.
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../computervisionimagestich2_b200/csrc/sift_device.cuh"
#include "serial_reference.inc"
#include "../../computervisionimagestich2_b200/csrc/match_device.cuh"
#include "../../computervisionimagestich2_b200/csrc/canvas_device.cuh"
#include "../../computervisionimagestich2_b200/csrc/host_numerics.h"
#include "../../computervisionimagestich2_b200/csrc/stitch_host.h"

using namespace pb;


struct EmulOct {
    int w, h, pitch;
    std::vector<float> gss, grad;
    std::vector<VlKey> keys;
    std::vector<int> nangles;
    std::vector<double> angles;
    std::vector<float> descr;
    std::vector<int> descr_key, descr_written;
};
struct EmulSift { std::vector<EmulOct> oct; };
static int g_check_serial = 1;
static int g_serial_mismatch = 0;

static void blur_plane(const float* src, float* tmp, float* dst, int w, int h, int pitch, double sigma) {
    float c[129];
    int W = hostnum::gaussian_taps(sigma, c, 64);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) tmp[(size_t)y * pitch + x] = blur_sample(src + x, pitch, h, y, c, W);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) dst[(size_t)y * pitch + x] = blur_sample(tmp + (size_t)y * pitch, 1, w, x, c, W);
}

extern "C" {

EmulSift* emul_sift_run(const float* im, int w, int h, int O, int S) {
    EmulSift* E = new EmulSift();
    const int s_min = -1, s_max = S + 1, nlev = s_max - s_min + 1;
    const double sigman = 0.5, sigmak = pow(2.0, 1.0 / S), sigma0 = 1.6 * sigmak,
                 dsigma0 = sigma0 * sqrt(1.0 - 1.0 / (sigmak * sigmak));
    SiftConsts sc{s_min, s_max, S, 0.0, 10.0, 0.0, 3.0, 2.0};
    double tab[257];
    hostnum::expn_table(tab);
    std::vector<float> tmp;
    for (int o = 0; o < O; ++o) {
        E->oct.emplace_back();
        EmulOct& ob = E->oct.back();
        ob.w = w >> o; ob.h = h >> o; ob.pitch = (ob.w + 31) / 32 * 32;
        size_t plane = (size_t)ob.pitch * ob.h;
        ob.gss.assign(plane * nlev, -7.0f);
        ob.grad.assign(plane * 2 * (nlev - 3), -7.0f);
        tmp.assign(plane, 0.f);
        if (o == 0) {
            for (int y = 0; y < h; ++y) memcpy(&ob.gss[(size_t)y * ob.pitch], im + (size_t)y * w, w * sizeof(float));
            double sa = sigma0 * pow(sigmak, s_min), sb = sigman * pow(2.0, 0);
            if (sa > sb) blur_plane(ob.gss.data(), tmp.data(), ob.gss.data(), ob.w, ob.h, ob.pitch, sqrt(sa * sa - sb * sb));
        } else {
            EmulOct& pr = E->oct[o - 1];
            int s_best = std::min(s_min + S, s_max);
            const float* src = &pr.gss[(size_t)(s_best - s_min) * pr.pitch * pr.h];
            for (int y = 0; y < ob.h; ++y)
                for (int x = 0; x < ob.w; ++x) ob.gss[(size_t)y * ob.pitch + x] = src[(size_t)(2 * y) * pr.pitch + 2 * x];
            double sa = sigma0 * powf((float)sigmak, (float)s_min), sb = sigma0 * powf((float)sigmak, (float)(s_best - S));
            if (sa > sb) blur_plane(ob.gss.data(), tmp.data(), ob.gss.data(), ob.w, ob.h, ob.pitch, sqrt(sa * sa - sb * sb));
        }
        for (int s = s_min + 1; s <= s_max; ++s)
            blur_plane(&ob.gss[plane * (s - 1 - s_min)], tmp.data(), &ob.gss[plane * (s - s_min)], ob.w, ob.h, ob.pitch,
                       dsigma0 * pow(sigmak, s));
        OctaveView ov{ob.w, ob.h, ob.pitch, nlev, ob.gss.data(), ob.grad.data()};
        double xper = pow(2.0, o);
        for (int l = 0; l < nlev - 3; ++l)
            for (int y = 0; y < ob.h; ++y)
                for (int x = 0; x < ob.w; ++x) {
                    float* g = &ob.grad[(((size_t)l * ob.h + y) * ob.pitch + x) * 2];
                    gradient_at(&ob.gss[plane * (1 + l)], ob.w, ob.h, ob.pitch, x, y, g, g + 1);
                }
        for (int l = 1; l <= nlev - 3; ++l)
            for (int y = 1; y < ob.h - 1; ++y)
                for (int x = 1; x < ob.w - 1; ++x)
                    if (is_extremum(ov, x, y, l, sc.peak_thresh)) {
                        RefinedKey r = refine_key(ov, sc, x, y, l + s_min, xper);
                        if (!r.good) continue;
                        VlKey k;
                        k.o = o; k.ix = r.ix; k.iy = r.iy; k.is = r.is; k.x = r.x; k.y = r.y; k.s = r.s;
                        k.sigma = (float)(sigma0 * pow(2.0, r.sn / S) * xper);
                        ob.keys.push_back(k);
                    }
        for (size_t i = 0; i < ob.keys.size(); ++i) {
            const VlKey& k = ob.keys[i];
            double hist[36], ang[4] = {0, 0, 0, 0};
            int na = orientations_of(ov, sc, tab, o, k.o, k.is, k.x, k.y, k.sigma, xper, hist, 1, ang);
            ob.nangles.push_back(na);
            for (int j = 0; j < 4; ++j) ob.angles.push_back(j < na ? ang[j] : 0.0);
            for (int j = 0; j < na; ++j) {
                float fh[128], d[128];
                for (int q = 0; q < 128; ++q) d[q] = -1.0f;
                // the sample-parallel formulation the CUDA kernel runs: conservative row / column ranges, per-sample
                // records (phase A), bin owners adding their contributions in sample order (phase B)
                DescFrame F = descriptor_frame(ov, sc, o, k.o, k.is, k.x, k.y, k.sigma, xper, ang[j], sin(ang[j]), cos(ang[j]));
                int wr = F.valid;
                if (wr) {
                    for (int q = 0; q < 128; ++q) fh[q] = 0.f;
                    int ry0, ry1;
                    descriptor_rows(F, &ry0, &ry1);
                    for (int dyi = ry0; dyi <= ry1; ++dyi) {
                        int x0, x1;
                        descriptor_row_range(F, dyi, &x0, &x1);
                        const float* row = F.pt + 2 * ((long)(F.yi + dyi) * F.pitch);
                        for (int dxi = x0; dxi <= x1; ++dxi) {
                            DescSample S = descriptor_sample(F, tab, dxi, dyi, row[2 * (F.xi + dxi)], row[2 * (F.xi + dxi) + 1]);
                            if (!S.active) continue;
                            for (int cell = 0; cell < 16; ++cell) {
                                const int dbx = (cell & 3) - 2 - S.binx, dby = (cell >> 2) - 2 - S.biny;
                                if (dbx < 0 || dbx > 1 || dby < 0 || dby > 1) continue;
                                for (int par = 0; par < 2; ++par) {
                                    const int sel = ((S.bint & 1) == par) ? 0 : 1;
                                    fh[cell * 8 + ((S.bint + sel) & 7)] += S.wxy[dbx][dby] * S.at[sel];
                                }
                            }
                        }
                    }
                    descriptor_finish(sc, fh, 1, d);
                }
                if (g_check_serial) {  // cross-check against the serial restatement
                    float fh2[128], d2[128];
                    for (int q = 0; q < 128; ++q) d2[q] = -1.0f;
                    int wr2 = descriptor_of(ov, sc, tab, o, k.o, k.is, k.x, k.y, k.sigma, xper, ang[j], sin(ang[j]), cos(ang[j]), fh2, 1, d2);
                    if (wr2 != wr || memcmp(d, d2, sizeof d) != 0) g_serial_mismatch++;
                }
                ob.descr.insert(ob.descr.end(), d, d + 128);
                ob.descr_key.push_back((int)i);
                ob.descr_written.push_back(wr);
            }
        }
    }
    return E;
}
int emul_serial_mismatches() { return g_serial_mismatch; }
int emul_sift_noctaves(EmulSift* E) { return (int)E->oct.size(); }
void emul_sift_info(EmulSift* E, int o, int* w, int* h, int* pitch, int* nkeys, int* ndesc) {
    EmulOct& O = E->oct[o];
    *w = O.w; *h = O.h; *pitch = O.pitch; *nkeys = (int)O.keys.size(); *ndesc = (int)O.descr_key.size();
}
// what: 0 gss (pitched), 2 grad (pitched), 3 keys, 4 nangles, 5 angles, 6 descr, 7 descr_key, 8 descr_written
void emul_sift_copy(EmulSift* E, int o, int what, void* dst) {
    EmulOct& O = E->oct[o];
    switch (what) {
    case 0: memcpy(dst, O.gss.data(), O.gss.size() * 4); break;
    case 2: memcpy(dst, O.grad.data(), O.grad.size() * 4); break;
    case 3: memcpy(dst, O.keys.data(), O.keys.size() * sizeof(VlKey)); break;
    case 4: memcpy(dst, O.nangles.data(), O.nangles.size() * 4); break;
    case 5: memcpy(dst, O.angles.data(), O.angles.size() * 8); break;
    case 6: memcpy(dst, O.descr.data(), O.descr.size() * 4); break;
    case 7: memcpy(dst, O.descr_key.data(), O.descr_key.size() * 4); break;
    case 8: memcpy(dst, O.descr_written.data(), O.descr_written.size() * 4); break;
    }
}
void emul_sift_free(EmulSift* E) { delete E; }

}  // extern "C"

#include "emul_rest.inc"
