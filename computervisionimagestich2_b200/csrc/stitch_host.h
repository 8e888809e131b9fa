// stitch_host.h -- host-side control logic of the stitcher that involves no pixels (plain C++, no CUDA):
// RANSAC sample drawing and model selection, the least-squares refit, stitch-order discovery, canvas sizing and
// keypoint re-mapping.  Mirrors ImageProcess.cpp:353-436, 500-594, 622-640 of the reference.
#pragma once
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <set>
#include "types.h"
#include "host_numerics.h"
#include "ransac_device.cuh"

namespace pb {
namespace stitch {

// Which caller of the hot path is reproduced.  kRoot = the reference's root ImageProcess.cpp; kEx6 = its second
// caller src/ex6/ImageProcess.cpp (same kernels, different control flow and constants):
//                         kRoot                                   kEx6
//   stitch order          all-pairs adjacency + middle image      fixed chain 0-1-...-n-1 from image n/2 (:147-160)
//   RANSAC seed           srand(666666)                           srand(time(0)) -> the caller passes the seed
//   canvas bounds         min / max over the 4 warped corners     min over 2, max_x over 3 corners (:230-243, 545-580)
//   seam statistics       channel 0 != 0, float ratios            all 3 channels != 0, double ratios (:651-697)
//   pyramid levels        floor(log2(max(w, h)))                  floor(log2(min(w, h))) (:662-665)
//   pyramid blur          get_blur(2,true,true) = Van Vliet       get_blur(2) = Deriche (:702-705)
//   final luminance mix   19/20 : 1/20                            5/6 : 1/6 (:270)
struct Profile {
    enum Variant { kRoot = 0, kEx6 = 1 };
    int variant = kRoot;
    unsigned ransac_seed = 666666u;
    bool ex6() const { return variant == kEx6; }
};

// ImageProcess.cpp:398: k = ceil(log(1 - 0.99) / log(1 - 0.5^4)) = 72
inline int ransac_iterations() { return (int)std::ceil(std::log(1 - 0.99) / std::log(1 - std::pow(0.5, 4))); }

// ImageProcess.cpp:397, 409-418: srand(666666), then per iteration 4 distinct rand() % n by rejection.
// glibc rand() is part of the contract (same libc on the reference side): srand / rand are srandom_r / random_r on a
// 128-byte (TYPE_3) state, reproduced here on a private state so that concurrent contexts do not share libc's
// global generator.  idx: [iters][4] in draw order.
inline bool draw_samples(int npairs, std::vector<int>& idx, unsigned seed = 666666u) {
    if (npairs < 4) return false;  // the reference loops forever (quirk Q8)
    const int iters = ransac_iterations();
    idx.resize((size_t)iters * 4);
    struct random_data rd;
    char statebuf[128];
    std::memset(&rd, 0, sizeof rd);
    std::memset(statebuf, 0, sizeof statebuf);
    initstate_r(seed, statebuf, sizeof statebuf, &rd);
    auto rand = [&rd]() { int32_t r; random_r(&rd, &r); return (int)r; };
    for (int k = 0; k < iters; ++k) {
        std::set<int> chosen;
        for (int i = 0; i < 4; ++i) {
            int index = rand() % npairs;
            while (chosen.find(index) != chosen.end()) index = rand() % npairs;
            chosen.insert(index);
            idx[(size_t)k * 4 + i] = index;
        }
    }
    return true;
}

// ImageProcess.cpp:427: the FIRST hypothesis with a strictly larger inlier set wins.
inline int select_hypothesis(const int* counts, int iters) {
    int best = -1, bestc = 0;
    for (int k = 0; k < iters; ++k)
        if (counts[k] > bestc) { bestc = counts[k]; best = k; }
    return best;
}

// ImageProcess.cpp:500-529: least squares on the inliers (SVD pseudo-inverse; plain LU when exactly 4 inliers).
inline bool refit(const KeyPair* pairs, const std::vector<int>& inl, double* H8) {
    const int n = (int)inl.size();
    if (n == 0) return false;
    if (n == 4) {
        float sx[4], sy[4], dx[4], dy[4];
        for (int i = 0; i < 4; ++i) {
            const KeyPair& p = pairs[inl[i]];
            sx[i] = p.src.x; sy[i] = p.src.y; dx[i] = p.dst.x; dy[i] = p.dst.y;
        }
        fit4(sx, sy, dx, dy, H8);
        return true;
    }
    hostnum::Mat A(4, n);
    std::vector<double> b(n);
    for (int i = 0; i < n; ++i) {
        const KeyPair& p = pairs[inl[i]];
        A(0, i) = (double)p.src.x;
        A(1, i) = (double)p.src.y;
        A(2, i) = (double)p.src.x * p.src.y;
        A(3, i) = 1.0;
        b[i] = (double)p.dst.x;
    }
    const hostnum::Mat P = hostnum::pinv(A);   // one SVD for both right-hand sides: same bits as two get_solve calls
    hostnum::pinv_apply(P, b, H8);
    for (int i = 0; i < n; ++i) b[i] = (double)pairs[inl[i]].dst.y;
    hostnum::pinv_apply(P, b, H8 + 4);
    return true;
}

// Whole RANSAC on the host (test emulator only; the product scores hypotheses on the GPU).
inline bool ransac_host(const KeyPair* pairs, int n, double* H8, std::vector<int>* inliers_out = nullptr,
                        unsigned seed = 666666u) {
    std::vector<int> idx;
    if (!draw_samples(n, idx, seed)) return false;
    const int iters = ransac_iterations();
    std::vector<int> best;
    for (int k = 0; k < iters; ++k) {
        float sx[4], sy[4], dx[4], dy[4];
        for (int i = 0; i < 4; ++i) {
            const KeyPair& p = pairs[idx[(size_t)k * 4 + i]];
            sx[i] = p.src.x; sy[i] = p.src.y; dx[i] = p.dst.x; dy[i] = p.dst.y;
        }
        double H[8];
        fit4(sx, sy, dx, dy, H);
        std::vector<int> in;
        for (int i = 0; i < n; ++i)
            if (is_inlier(H, pairs[i].src.x, pairs[i].src.y, pairs[i].dst.x, pairs[i].dst.y)) in.push_back(i);
        if (in.size() > best.size()) best = in;
    }
    if (inliers_out) *inliers_out = best;
    return refit(pairs, best, H8);
}

// ImageProcess.cpp:353-393, including the index-vs-position comparison of the original (quirk Q5).
inline int middle_index(const std::vector<std::vector<int>>& next_index, const std::vector<std::vector<char>>& adj) {
    const int n = (int)next_index.size();
    int edge = 0;
    for (int i = 0; i < n; i++)
        if (next_index[i].size() == 1) { edge = i; break; }
    int next_one = edge;
    std::vector<int> que;
    for (int index = 0; index < n; index++) {
        if (que.empty()) que.push_back(edge);
        for (int i = 0; i < n; i++) {
            if (next_one == i) continue;
            bool flag = true;
            if (adj[next_one][i]) {
                for (int j = 0; j < (int)que.size(); j++)
                    if (i == j) { flag = false; break; }
                if (!flag) continue;
                if (i != edge) que.push_back(i);
                next_one = i;
                break;
            }
        }
    }
    return que[que.size() / 2];
}

// ImageProcess.cpp:532-594 + 206-216
struct CanvasPlan {
    float min_x, min_y, max_x, max_y;
    int new_w, new_h;
};
// ex6: src/ex6/ImageProcess.cpp:230-243, 545-580 look at fewer corners -- min_x over (0,h-1),(0,0); min_y over
// (w-1,0),(0,0); max_x over (0,0),(w-1,0),(w-1,h-1); max_y over all four.
inline CanvasPlan plan_canvas(int dw, int dh, const double* fwd, int res_w, int res_h, bool ex6 = false) {
    const float cx[4] = {0.f, (float)(dw - 1), 0.f, (float)(dw - 1)};
    const float cy[4] = {0.f, 0.f, (float)(dh - 1), (float)(dh - 1)};
    const bool in_minx[4] = {true, !ex6, true, !ex6}, in_miny[4] = {true, true, !ex6, !ex6};
    const bool in_maxx[4] = {true, true, !ex6, true};
    float minx = warp_x(fwd, cx[0], cy[0]), maxx = minx, miny = warp_y(fwd, cx[0], cy[0]), maxy = miny;
    for (int i = 1; i < 4; ++i) {
        float x = warp_x(fwd, cx[i], cy[i]), y = warp_y(fwd, cx[i], cy[i]);
        if (in_minx[i] && x < minx) minx = x;
        if (in_maxx[i] && x > maxx) maxx = x;
        if (in_miny[i] && y < miny) miny = y;
        if (y > maxy) maxy = y;
    }
    CanvasPlan p;
    p.min_x = (minx < 0) ? minx : 0;
    p.min_y = (miny < 0) ? miny : 0;
    p.max_x = (maxx >= (float)res_w) ? maxx : (float)res_w;
    p.max_y = (maxy >= (float)res_h) ? maxy : (float)res_h;
    p.new_w = (int)std::ceil(p.max_x - p.min_x);
    p.new_h = (int)std::ceil(p.max_y - p.min_y);
    return p;
}

// ImageProcess.cpp:622-640
inline void update_features_by_homography(VlKey* k, int n, const double* H8, float offx, float offy) {
    for (int i = 0; i < n; ++i) {
        float cx = k[i].x, cy = k[i].y;
        k[i].x = warp_x(H8, cx, cy) - offx;
        k[i].y = warp_y(H8, cx, cy) - offy;
        k[i].ix = (int)k[i].x;
        k[i].iy = (int)k[i].y;
    }
}
inline void update_features_by_offset(VlKey* k, int n, int offx, int offy) {
    for (int i = 0; i < n; ++i) {
        k[i].x -= offx;
        k[i].y -= offy;
        k[i].ix = (int)k[i].x;
        k[i].iy = (int)k[i].y;
    }
}

// blendTwoImages level geometry (ImageProcess.cpp:675-676, 706-707): level_num = floor(log2(max(w,h))),
// dims halve (integer division) per level; a level with a zero dimension and everything above it is empty (quirk Q7).
// ex6: level count from the SHORTER side (src/ex6/ImageProcess.cpp:662-665), so no level is ever empty.
inline int blend_levels(int w, int h, std::vector<int>& lw, std::vector<int>& lh, bool ex6 = false) {
    int max_len = ex6 ? (w < h ? w : h) : (w >= h ? w : h);
    int level_num = (int)std::floor(std::log2((double)max_len));
    lw.clear(); lh.clear();
    int cw = w, ch = h;
    for (int i = 0; i < level_num; ++i) {
        if (cw <= 0 || ch <= 0) break;
        lw.push_back(cw); lh.push_back(ch);
        cw /= 2; ch /= 2;
    }
    return (int)lw.size();  // number of non-empty levels
}

}  // namespace stitch
}  // namespace pb
