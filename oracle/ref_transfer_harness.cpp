// oracle/ref_transfer_harness.cpp -- TEST INFRASTRUCTURE.  C entry point around the reference's Reinhard colour transfer
// (class transfer, transfer.cpp:4-13), compiled from /root/reference by oracle/Makefile (target ref_transfer).
#include "transfer.h"
#include <cstring>
extern "C" int ref_color_transfer(const unsigned char* src, int w, int h, const unsigned char* tem, int tw, int th,
                                  unsigned char* out) {
    CImg<unsigned char> s(src, w, h, 1, 3), t(tem, tw, th, 1, 3), o;
    transfer tr(s, t, o);
    if (o.width() != w || o.height() != h || o.spectrum() != 3) return -1;
    memcpy(out, o.data(), (size_t)3 * w * h);
    return 0;
}
