// oracle/ref_ex6_shim.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Force-included (-include) in front of the reference's src/ex6/ImageProcess.cpp when oracle/Makefile compiles the
// `ex6` variant of the reference.  That variant is not deterministic as shipped:
//   * RANSAC seeds libc with srand(time(0))                                   (src/ex6/ImageProcess.cpp:403);
//   * the two RANSAC calls of one edge run on two std::threads that share libc's rand() state (:221-226).
// To have a reproducible oracle, the two are pinned WITHOUT touching any arithmetic:
//   * `time(x)` inside that translation unit returns the seed chosen by the harness (pano_ex6_seed);
//   * `thread` inside that translation unit is a class that runs its function in the constructor (so the calls are
//     serialised in program order; each RANSAC call re-seeds, hence the order does not matter).
// The header of the variant is included first, with its private section opened for the harness, so that the two
// macros only rewrite the .cpp body.
#pragma once
#define private public
#include "ImageProcess.h"
#undef private

extern "C" unsigned pano_ex6_seed(void);

struct pano_ex6_sync_thread {
    template <class F, class... A>
    explicit pano_ex6_sync_thread(F f, A... a) { f(a...); }
    void join() {}
};
#define thread pano_ex6_sync_thread
#define time(x) pano_ex6_seed()
