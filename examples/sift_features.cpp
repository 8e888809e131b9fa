// examples/sift_features.cpp -- ImageProcess::siftAlgorithm (ImageProcess.cpp:44-99) verbatim in structure, compiled
// against include/vl_b200/sift.h instead of vl/sift.h: the call sequence vl_sift_new / process_first_octave /
// {detect, calc_keypoint_orientations, calc_keypoint_descriptor}* / process_next_octave, reading f->keys / f->nkeys
// directly, inserting into std::map<std::vector<float>, VlSiftKeypoint>.  Input: a raw 8-bit gray image file.
// Output: one line "<n> <fnv1a64 of descriptors> <fnv1a64 of keypoints>" in map order.
//   g++ -std=c++11 -Iinclude examples/sift_features.cpp -Lcomputervisionimagestich2_b200 -lpano_b200 -o sift_features
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
extern "C" {
#include "vl_b200/sift.h"
}
#define NOTAVES_NUM 4
#define LEVEL_NUM 2
#define DESCRIPTOR_SUM 128
static uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}
int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: sift_features gray.raw width height\n"); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    std::vector<unsigned char> gray((size_t)w * h);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(gray.data(), 1, gray.size(), f) != gray.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    vl_sift_pix* imageData = new vl_sift_pix[(size_t)w * h];
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) imageData[i * w + j] = gray[(size_t)i * w + j];
    VlSiftFilt* siftFilt = vl_sift_new(w, h, NOTAVES_NUM, LEVEL_NUM, 0);
    if (!siftFilt) { fprintf(stderr, "vl_sift_new failed\n"); return 1; }
    std::map<std::vector<float>, VlSiftKeypoint> features;
    if (vl_sift_process_first_octave(siftFilt, imageData) != VL_ERR_EOF) {
        while (true) {
            vl_sift_detect(siftFilt);
            VlSiftKeypoint* pKeyPoint = siftFilt->keys;
            for (int i = 0; i < siftFilt->nkeys; i++) {
                VlSiftKeypoint tmpKeyPoint = *pKeyPoint;
                pKeyPoint++;
                double angles[4];
                int angleCount = vl_sift_calc_keypoint_orientations(siftFilt, angles, &tmpKeyPoint);
                for (int j = 0; j < angleCount; j++) {
                    double tmpAngle = angles[j];
                    vl_sift_pix descriptors[DESCRIPTOR_SUM];
                    vl_sift_calc_keypoint_descriptor(siftFilt, descriptors, &tmpKeyPoint, tmpAngle);
                    std::vector<float> des;
                    for (int k = 0; k < DESCRIPTOR_SUM; k++) des.push_back(descriptors[k]);
                    tmpKeyPoint.ix = tmpKeyPoint.x;
                    tmpKeyPoint.iy = tmpKeyPoint.y;
                    features.insert(std::pair<std::vector<float>, VlSiftKeypoint>(des, tmpKeyPoint));
                }
            }
            if (vl_sift_process_next_octave(siftFilt) == VL_ERR_EOF) break;
        }
    }
    vl_sift_delete(siftFilt);
    delete[] imageData;
    uint64_t hd = 0xcbf29ce484222325ULL, hk = 0xcbf29ce484222325ULL;
    for (auto& kv : features) {
        hd = fnv(hd, kv.first.data(), kv.first.size() * sizeof(float));
        hk = fnv(hk, &kv.second, sizeof(VlSiftKeypoint));
    }
    printf("%zu %016llx %016llx\n", features.size(), (unsigned long long)hd, (unsigned long long)hk);
    return 0;
}
