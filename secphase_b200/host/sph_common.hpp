// sph_common.hpp -- shared helpers of libsecphase_host: thread-local error text, a small
// fork-join worker pool (the reference's tpool.c is a job queue over read groups; here the
// workers only inflate BGZF blocks and copy record fields, the scoring jobs went to the GPU).
#pragma once
#include <condition_variable>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace sph {

void set_error(const char *fmt, ...);
const char *last_error();

class WorkerPool {
public:
    explicit WorkerPool(int n_threads);
    ~WorkerPool();
    int size() const { return (int) threads_.size() + 1; }
    // Runs fn(i) for i in [0,n); the calling thread takes part; returns when all are done.
    // Safe to call from several threads at once.
    void parallel_for(int64_t n, const std::function<void(int64_t)> &fn);

private:
    struct Job {
        const std::function<void(int64_t)> *fn;
        int64_t n;
        int64_t next = 0;  // guarded by mu_
        int64_t done = 0;  // guarded by mu_
        std::condition_variable cv;
    };
    void worker();
    bool run_one(std::unique_lock<std::mutex> &lk);
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Job *> jobs_;
    bool stop_ = false;
};

}  // namespace sph
