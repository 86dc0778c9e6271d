// Minimal persistent worker pool for the host Delaunay stage: parallel_for over independent jobs.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace svb {

class ThreadPool {
   public:
    explicit ThreadPool(int n_threads) { resize(n_threads); }
    ~ThreadPool() { stop(); }

    int size() const { return (int)workers_.size() + 1; }  // the calling thread also works

    void resize(int n_threads) {
        stop();
        if (n_threads < 1) n_threads = 1;
        quit_ = false;
        for (int i = 0; i < n_threads - 1; i++) workers_.emplace_back([this, i] { loop(i + 1); });
    }

    // fn(job, worker) for job in [0, n); returns when all jobs are done.  worker in [0, size()).
    void parallel_for(int n, const std::function<void(int, int)> &fn) {
        if (n <= 0) return;
        if (workers_.empty() || n == 1) {
            for (int i = 0; i < n; i++) fn(i, 0);
            return;
        }
        {
            std::unique_lock<std::mutex> lk(mu_);
            done_cv_.wait(lk, [this] { return active_ == 0; });  // late wakers of the previous call have left
            fn_ = &fn;
            n_.store(n);
            next_.store(0);
            pending_ = n;
            generation_++;
        }
        cv_.notify_all();
        work(0);
        std::unique_lock<std::mutex> lk(mu_);
        // also wait until every worker has left work(): no straggler may touch the next call's job counter
        done_cv_.wait(lk, [this] { return pending_ == 0 && active_ == 0; });
        fn_ = nullptr;
    }

   private:
    void work(int worker) {
        while (true) {
            const int i = next_.fetch_add(1);
            if (i >= n_.load()) break;
            (*fn_)(i, worker);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
    void loop(int worker) {
        unsigned long seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return quit_ || generation_ != seen; });
                if (quit_) return;
                seen = generation_;
                active_++;
            }
            work(worker);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--active_ == 0) done_cv_.notify_all();
            }
        }
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
        workers_.clear();
    }

    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int, int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    std::atomic<int> n_{0};
    int pending_ = 0;
    int active_ = 0;
    unsigned long generation_ = 0;
    bool quit_ = false;
};

}  // namespace svb
