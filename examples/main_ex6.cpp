// examples/main_ex6.cpp -- the reference's src/ex6/main.cpp (src/ex6/main.cpp:3-13) with the B200 drop-in.  The
// original reads the data set name and the image count from stdin and prefixes "../../"; here they may also come from
// argv (directory as given).  The panorama is written to <dir>result.bmp by the constructor, as in the original.
//   g++ -std=c++11 -Iinclude examples/main_ex6.cpp -Lcomputervisionimagestich2_b200 -lpano_b200 -o pano_main_ex6
#include "pano_b200/ex6/ImageProcess.h"
#include <cstdio>
#include <iostream>
int main(int argc, char** argv) {
    std::string dir;
    int n = 0;
    if (argc > 2) {
        dir = argv[1];
        n = atoi(argv[2]);
    } else {
        std::string file;
        std::cin >> file;
        dir = "../../" + file + "/";
        std::cout << "Please input the sum of the images" << std::endl;
        std::cin >> n;
    }
    if (!dir.empty() && dir.back() != '/') dir += '/';
    try {
        ImageProcess ip(dir, n);
        printf("panorama %dx%d\n", ip.result.width(), ip.result.height());
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
