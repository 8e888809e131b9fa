// oracle/standalone_main.cpp -- TEST INFRASTRUCTURE.  The reference's pipeline with a plain main and no stage harness:
// ImageProcess(dir, n) exactly as main.cpp:9 calls it, then the FNV-1a64 (SURVEY.md 8c convention) and the byte
// count of the private member `result`.  Built by oracle/standalone_hash.sh from the sources under /root/reference
// (only the two headless result.display() calls are dropped, and vl/mathop.c is compiled at -O0, unpatched).
#define private public
#include "ImageProcess.h"
#undef private
#include <cstdint>
#include <cstdio>

static uint64_t fnv1a64(const unsigned char* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
// printed by the injected statement right before the equalisation tail (ImageProcess.cpp:237): the blended canvas
extern "C" void standalone_note_blend(const unsigned char* p, int w, int h) {
    printf("blend_only %dx%d fnv1a64=%016llx\n", w, h, (unsigned long long)fnv1a64(p, (size_t)3 * w * h));
}
int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <dir/> <n> [raw-out]\n", argv[0]); return 2; }
    ImageProcess ip(argv[1], atoi(argv[2]));
    const CImg<unsigned char>& r = ip.result;
    printf("panorama %dx%d fnv1a64=%016llx\n", r.width(), r.height(), (unsigned long long)fnv1a64(r.data(), r.size()));
    if (argc > 3) { FILE* f = fopen(argv[3], "wb"); fwrite(r.data(), 1, r.size(), f); fclose(f); }
    return 0;
}
