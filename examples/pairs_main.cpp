// examples/pairs_main.cpp -- C++ host of the batched-pairs entry (BASELINE.json configs[4]): the neighbours of a numbered
// BMP set, (1,2), (2,3), ..., as independent pairs through ONE pano_b200_pairs call.  Per pair and direction it prints
// the feature counts, the number of getImgPair matches (ImageProcess.cpp:273-351) and, where the direction is adjacent
// (>= 20 matches, ImageProcess.cpp:128), the eight RANSAC coefficients (ImageProcess.cpp:395-436, 465-471).
//   usage: pairs_main <dir-with-trailing-slash> <n>        e.g.  pairs_main ../oracle/_ref/data/Input/ 4
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "pano_b200/ImageProcess.h"   // PlanarImage::load_bmp, pano_b200.h

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: pairs_main <dir/> <n>\n"); return 2; }
    const std::string dir = argv[1];
    const int n = atoi(argv[2]);
    if (n < 2) { fprintf(stderr, "need at least two images\n"); return 2; }
    std::vector<pano_b200::PlanarImage> imgs;
    try {
        for (int i = 0; i < n; ++i) imgs.push_back(pano_b200::PlanarImage::load_bmp(dir + std::to_string(i + 1) + ".bmp"));
    } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 2; }
    const int npairs = n - 1;
    std::vector<const uint8_t*> ptr;
    std::vector<int> w, h;
    for (int p = 0; p < npairs; ++p)
        for (int k = 0; k < 2; ++k) {
            ptr.push_back(imgs[p + k].data());
            w.push_back(imgs[p + k].width());
            h.push_back(imgs[p + k].height());
        }
    pano_b200_ctx* ctx = nullptr;
    if (pano_b200_create(0, &ctx) != 0) { fprintf(stderr, "pano_b200_create failed (no CUDA device?)\n"); return 1; }
    std::vector<pano_b200_pair_record> rec(npairs);
    const int rc = pano_b200_pairs(ctx, ptr.data(), w.data(), h.data(), npairs, rec.data());
    if (rc != 0) { fprintf(stderr, "pano_b200_pairs: %d %s\n", rc, pano_b200_last_error(ctx)); pano_b200_destroy(ctx); return 1; }
    for (int p = 0; p < npairs; ++p)
        for (int d = 0; d < 2; ++d) {
            const int a = d ? p + 1 : p, b = d ? p : p + 1;
            printf("%d %d nfeat %d matches %d", a, b, rec[p].nfeat[d], rec[p].nmatch[d]);
            if (rec[p].has_h[d])
                for (int k = 0; k < 8; ++k) printf(" %.17g", rec[p].H[d][k]);
            printf("\n");
        }
    pano_b200_destroy(ctx);
    return 0;
}
