// ktimer.h -- per-kernel launch counters and (optional) CUDA-event timing, used by bench.py for `gpu_launches` and for
// the live per-kernel durations behind the roofline numbers.  Counters are always on (one relaxed increment per
// launch); event timing only when enabled (pano_b200_ktimer_enable), because it adds two event records per launch.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace pb {

struct KStat {
    long launches = 0;
    double ms = 0;       // resolved event time
    double bytes = 0;    // algorithmic bytes (or ops) the launcher declared
};

class KTimer {
  public:
    static KTimer& get() { static KTimer k; return k; }
    void enable(bool on) { std::lock_guard<std::mutex> g(m_); enabled_ = on; }
    bool enabled() const { return enabled_; }
    void reset() {
        std::lock_guard<std::mutex> g(m_);
        resolve_locked();
        stats_.clear();
    }
    // called by KScope
    void begin(const char* name, cudaStream_t st, double bytes, cudaEvent_t* e0, cudaEvent_t* e1) {
        std::lock_guard<std::mutex> g(m_);
        KStat& s = stats_[name];
        s.launches++;
        s.bytes += bytes;
        *e0 = *e1 = nullptr;
        if (!enabled_) return;
        *e0 = new_event();
        *e1 = new_event();
        cudaEventRecord(*e0, st);
    }
    void end(const char* name, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1) {
        if (!e0) return;
        cudaEventRecord(e1, st);
        std::lock_guard<std::mutex> g(m_);
        pending_.push_back(Pending{name, e0, e1});
    }
    // synchronises the pending events and returns a snapshot
    std::map<std::string, KStat> snapshot() {
        std::lock_guard<std::mutex> g(m_);
        resolve_locked();
        return stats_;
    }
    long total_launches() {
        std::lock_guard<std::mutex> g(m_);
        long n = 0;
        for (auto& kv : stats_) n += kv.second.launches;
        return n;
    }

  private:
    struct Pending { const char* name; cudaEvent_t e0, e1; };
    cudaEvent_t new_event() {
        if (!free_.empty()) { cudaEvent_t e = free_.back(); free_.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void resolve_locked() {
        for (auto& p : pending_) {
            cudaEventSynchronize(p.e1);
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) stats_[p.name].ms += ms;
            free_.push_back(p.e0);
            free_.push_back(p.e1);
        }
        pending_.clear();
    }
    std::mutex m_;
    bool enabled_ = false;
    std::map<std::string, KStat> stats_;
    std::vector<Pending> pending_;
    std::vector<cudaEvent_t> free_;
};

struct KScope {
    const char* name;
    cudaStream_t st;
    cudaEvent_t e0, e1;
    KScope(const char* n, cudaStream_t s, double bytes = 0) : name(n), st(s) { KTimer::get().begin(n, s, bytes, &e0, &e1); }
    ~KScope() { KTimer::get().end(name, st, e0, e1); }
};

}  // namespace pb
