// common.h -- error handling and small RAII helpers for device / pinned memory.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace pb {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s:%d: %s: %s", file, line, what, cudaGetErrorString(e));
        throw CudaError(buf);
    }
}
#define PB_CUDA(x) ::pb::cuda_check((x), #x, __FILE__, __LINE__)
#define PB_KERNEL_CHECK() ::pb::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

// Host-side timeline for diagnostics: PANO_B200_TRACE=1 makes every PB_TRACE(tag) print "trace <thread> <ms> <tag>" on
// stderr (ms since the first trace point of the process).  Off: one predictable branch.
bool trace_on();
void trace_point(const char* tag, long a = -1);
#define PB_TRACE(...) do { if (::pb::trace_on()) ::pb::trace_point(__VA_ARGS__); } while (0)

inline int div_up(long a, long b) { return (int)((a + b - 1) / b); }
inline int align_up(int a, int b) { return (a + b - 1) / b * b; }

// Grow-only device buffer (cudaMalloc'ed; contents are NOT preserved across a grow).
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    T* ensure(size_t n) {
        if (n > cap) {
            if (p) PB_CUDA(cudaFree(p));
            p = nullptr;
            size_t want = n + n / 8 + 64;
            PB_CUDA(cudaMalloc(&p, want * sizeof(T)));
            cap = want;
        }
        return p;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Grow-only pinned host buffer.
template <class T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    T* ensure(size_t n) {
        if (n > cap) {
            if (p) PB_CUDA(cudaFreeHost(p));
            p = nullptr;
            size_t want = n + n / 8 + 64;
            PB_CUDA(cudaMallocHost(&p, want * sizeof(T)));
            cap = want;
        }
        return p;
    }
};

}  // namespace pb
