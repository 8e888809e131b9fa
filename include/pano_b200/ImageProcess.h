// pano_b200/ImageProcess.h -- drop-in for the reference's `ImageProcess` class (ImageProcess.h:77-146) on top of the
// C ABI of libpano_b200.so.  Same constructor as the reference: ImageProcess(std::string dir, int n) loads
// <dir>1.bmp .. <dir>n.bmp (ImageProcess.cpp:16), stitches them and leaves the panorama in `result`.  The reference
// keeps `result` private and only display()s it (ImageProcess.cpp:233,270); here it is public and mimics the part of
// CImg<unsigned char> callers use: width() / height() / spectrum() / data() / operator()(x, y, c), planar layout
// data[x + y*W + c*W*H] (CImg.h:48533-48546), plus save_bmp().  Header-only, C++11, no CImg dependency.
//
// The reference's second copy of the class, src/ex6/ImageProcess.h, is served by pano_b200/ex6/ImageProcess.h, which
// includes this file with PANO_B200_IMAGEPROCESS_EX6 defined: same constructor, images stitched as a left-to-right
// chain, RANSAC seeded with time(0) (src/ex6/ImageProcess.cpp:403; define PANO_B200_RANSAC_SEED to pin it), elapsed
// time printed and the panorama saved as <dir>result.bmp (src/ex6/ImageProcess.cpp:12-16).
#ifndef PANO_B200_IMAGEPROCESS_H
#define PANO_B200_IMAGEPROCESS_H
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <chrono>
#include <stdexcept>
#include <string>
#include <vector>
#include "../pano_b200.h"

namespace pano_b200 {

struct PlanarImage {
    int w = 0, h = 0;
    std::vector<unsigned char> px;  // [3][h][w]
    int width() const { return w; }
    int height() const { return h; }
    int spectrum() const { return 3; }
    const unsigned char* data() const { return px.data(); }
    unsigned char operator()(int x, int y, int c) const { return px[(size_t)x + (size_t)y * w + (size_t)c * w * h]; }

    // 24-bpp uncompressed BMP, bottom-up BGR rows padded to 4 bytes (what CImg::_load_bmp reads, CImg.h:48395-48546)
    static PlanarImage load_bmp(const std::string& path) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + path);
        std::vector<unsigned char> buf;
        unsigned char tmp[1 << 16];
        size_t n;
        while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
        fclose(f);
        if (buf.size() < 54 || buf[0] != 'B' || buf[1] != 'M') throw std::runtime_error(path + ": not a BMP");
        auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, &buf[o], 4); return v; };
        auto i32 = [&](size_t o) { int32_t v; memcpy(&v, &buf[o], 4); return v; };
        const uint32_t off = u32(10);
        int w = i32(18), h = i32(22);
        const int bpp = buf[28] | (buf[29] << 8);
        if (bpp != 24 || u32(30) != 0) throw std::runtime_error(path + ": only uncompressed 24-bpp BMP is supported");
        const bool bottom_up = h > 0;
        if (h < 0) h = -h;
        const size_t stride = ((size_t)3 * w + 3) & ~(size_t)3;
        if (buf.size() < off + stride * h) throw std::runtime_error(path + ": truncated");
        PlanarImage im;
        im.w = w; im.h = h;
        im.px.resize((size_t)3 * w * h);
        const size_t plane = (size_t)w * h;
        for (int y = 0; y < h; ++y) {
            const unsigned char* row = &buf[off + stride * (bottom_up ? h - 1 - y : y)];
            for (int x = 0; x < w; ++x) {
                im.px[(size_t)y * w + x] = row[3 * x + 2];
                im.px[plane + (size_t)y * w + x] = row[3 * x + 1];
                im.px[2 * plane + (size_t)y * w + x] = row[3 * x];
            }
        }
        return im;
    }
    void save_bmp(const std::string& path) const {
        const size_t stride = ((size_t)3 * w + 3) & ~(size_t)3, plane = (size_t)w * h;
        std::vector<unsigned char> out(54 + stride * h, 0);
        out[0] = 'B'; out[1] = 'M';
        auto put = [&](size_t o, uint32_t v) { memcpy(&out[o], &v, 4); };
        put(2, (uint32_t)out.size()); put(10, 54); put(14, 40); put(18, (uint32_t)w); put(22, (uint32_t)h);
        out[26] = 1; out[28] = 24; put(34, (uint32_t)(stride * h)); put(38, 2835); put(42, 2835);
        for (int y = 0; y < h; ++y) {
            unsigned char* row = &out[54 + stride * (h - 1 - y)];
            for (int x = 0; x < w; ++x) {
                row[3 * x + 2] = px[(size_t)y * w + x];
                row[3 * x + 1] = px[plane + (size_t)y * w + x];
                row[3 * x] = px[2 * plane + (size_t)y * w + x];
            }
        }
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot write " + path);
        fwrite(out.data(), 1, out.size(), f);
        fclose(f);
    }
};

}  // namespace pano_b200

class ImageProcess {
  public:
    // ImageProcess.cpp:3-8: readFile(dir, n); matching();
    ImageProcess(std::string dir, int n, int device = 0) {
        std::vector<pano_b200::PlanarImage> imgs;
        for (int i = 0; i < n; ++i) imgs.push_back(pano_b200::PlanarImage::load_bmp(dir + std::to_string(i + 1) + ".bmp"));
        std::vector<const uint8_t*> p;
        std::vector<int> w, h;
        for (auto& im : imgs) { p.push_back(im.data()); w.push_back(im.w); h.push_back(im.h); }
        pano_b200_ctx* ctx = nullptr;
        if (pano_b200_create(device, &ctx) != 0) throw std::runtime_error("pano_b200_create failed (no CUDA device? there is no CPU fallback)");
#ifdef PANO_B200_IMAGEPROCESS_EX6
        const auto t0 = std::chrono::steady_clock::now();
#ifdef PANO_B200_RANSAC_SEED
        pano_b200_set_profile(ctx, PANO_B200_PROFILE_EX6, (unsigned)(PANO_B200_RANSAC_SEED));
#else
        pano_b200_set_profile(ctx, PANO_B200_PROFILE_EX6, (unsigned)time(0));   // src/ex6/ImageProcess.cpp:403
#endif
#endif
        uint8_t* out = nullptr;
        int ow = 0, oh = 0;
        const int rc = pano_b200_stitch(ctx, p.data(), w.data(), h.data(), n, &out, &ow, &oh);
        if (rc != 0) {
            std::string msg = pano_b200_last_error(ctx);
            pano_b200_destroy(ctx);
            throw std::runtime_error("pano_b200_stitch failed: " + msg);
        }
        std::vector<char> log(1 << 16);
        pano_b200_stitch_log(ctx, log.data(), (int)log.size());
        fputs(log.data(), stdout);  // the reference prints the middle index and each "src dst" edge (ImageProcess.cpp:183,391)
        result.w = ow; result.h = oh;
        result.px.assign(out, out + (size_t)3 * ow * oh);
        pano_b200_free(out);
        pano_b200_destroy(ctx);
#ifdef PANO_B200_IMAGEPROCESS_EX6
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("costs:%gs\n", secs);            // src/ex6/ImageProcess.cpp:12-13
        result.save_bmp(dir + "result.bmp");    // src/ex6/ImageProcess.cpp:15-16
#endif
    }
    pano_b200::PlanarImage result;
};
#endif
