#include "sph_common.hpp"

#include "../../include/secphase_host.h"

namespace sph {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char *last_error() { return g_err.c_str(); }

WorkerPool::WorkerPool(int n_threads) {
    for (int i = 1; i < n_threads; i++) threads_.emplace_back([this] { worker(); });
}

WorkerPool::~WorkerPool() {
    {
        std::lock_guard<std::mutex> g(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
}

// Takes one index of the oldest job that still has unclaimed indices and runs it.
bool WorkerPool::run_one(std::unique_lock<std::mutex> &lk) {
    Job *j = nullptr;
    for (Job *c : jobs_)
        if (c->next < c->n) {
            j = c;
            break;
        }
    if (!j) return false;
    int64_t i = j->next++;
    lk.unlock();
    (*j->fn)(i);
    lk.lock();
    if (++j->done == j->n) j->cv.notify_all();
    return true;
}

void WorkerPool::worker() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
        if (run_one(lk)) continue;
        if (stop_) return;
        cv_.wait(lk);
    }
}

void WorkerPool::parallel_for(int64_t n, const std::function<void(int64_t)> &fn) {
    if (n <= 0) return;
    if (n == 1 || threads_.empty()) {
        for (int64_t i = 0; i < n; i++) fn(i);
        return;
    }
    Job job;
    job.fn = &fn;
    job.n = n;
    std::unique_lock<std::mutex> lk(mu_);
    jobs_.push_back(&job);
    cv_.notify_all();
    while (job.next < job.n) {  // the caller works on its own job only
        int64_t i = job.next++;
        lk.unlock();
        fn(i);
        lk.lock();
        ++job.done;
    }
    while (job.done < job.n) job.cv.wait(lk);
    for (auto it = jobs_.begin(); it != jobs_.end(); ++it)
        if (*it == &job) {
            jobs_.erase(it);
            break;
        }
}

}  // namespace sph

extern "C" const char *sph_last_error(void) { return sph::last_error(); }
