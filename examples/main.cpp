// examples/main.cpp -- the reference's main.cpp (main.cpp:3-11) with the B200 drop-in: the only differences are the
// include, the directory taken from argv instead of being hard-coded, and saving / hashing the (now public) result.
//   g++ -std=c++11 -Iinclude examples/main.cpp -Lcomputervisionimagestich2_b200 -lpano_b200 -o pano_main
#include "pano_b200/ImageProcess.h"
#include <cstdio>
int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : "../../Input/";
    int n = argc > 2 ? atoi(argv[2]) : 4;
    if (!dir.empty() && dir.back() != '/') dir += '/';
    try {
        ImageProcess ip(dir, n);
        uint64_t h = 0xcbf29ce484222325ULL;
        const size_t bytes = (size_t)3 * ip.result.width() * ip.result.height();
        for (size_t i = 0; i < bytes; ++i) { h ^= ip.result.data()[i]; h *= 0x100000001b3ULL; }
        printf("panorama %dx%d fnv1a64 %016llx\n", ip.result.width(), ip.result.height(), (unsigned long long)h);
        if (argc > 3) ip.result.save_bmp(argv[3]);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
