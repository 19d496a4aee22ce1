// Small persistent host thread pool used by the host-buffer entry points to widen / scatter rows of a result
// while later chunks are still crossing PCIe.  Host-side plumbing only: no rasterization happens here.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace rgpu {

class HostPool {
public:
    explicit HostPool(unsigned n) {
        if (n < 1) n = 1;
        for (unsigned i = 0; i < n; i++) workers_.emplace_back([this] { run(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    unsigned size() const { return (unsigned)workers_.size(); }
    void submit(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> l(m_);
            q_.push_back(std::move(f));
            pending_++;
        }
        cv_.notify_one();
    }
    void wait() {
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
    }

private:
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                f = std::move(q_.front());  // first in, first out: tasks that wait for a copy are queued in copy order
                q_.pop_front();
            }
            f();
            {
                std::lock_guard<std::mutex> l(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::deque<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    size_t pending_ = 0;
    bool stop_ = false;
};

}  // namespace rgpu

// host_simd.cpp: streaming loops with the widest vector unit of the CPU (run-time dispatch)
namespace rgpu {
const char* host_simd_name();
void expand_alpha_simd(const float* alpha, const float colour[4], float* out, size_t n_px);  // out[4 i + k] = colour[k] * alpha[i]
void widen_row_simd(const float* src, double* dst, size_t n);                               // dst[i] = (double)src[i]
// one row of a run-coded image (compact.cu) -> pixels; returns the literals consumed; fence afterwards
size_t expand_runs_f32(const unsigned char* cls, size_t n_segs, size_t width, const float* lits, float* dst);
size_t expand_runs_f64(const unsigned char* cls, size_t n_segs, size_t width, const float* lits, double* dst);
void host_store_fence();
}  // namespace rgpu
