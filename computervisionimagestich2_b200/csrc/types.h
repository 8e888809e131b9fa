// types.h -- plain data types shared by host code, kernels and the C ABI (no CUDA dependency).
#pragma once
#include <cstdint>

namespace pb {

struct VlKey {  // == VlSiftKeypoint (vl/sift.h:19-31)
    int o, ix, iy, is;
    float x, y, s, sigma;
};
static_assert(sizeof(VlKey) == 32, "VlSiftKeypoint layout");

struct KeyPair {  // == ImgPair (ImageProcess.h:43-47)
    VlKey src, dst;
};

}  // namespace pb
