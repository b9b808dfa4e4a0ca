// A handful of persistent host threads per context for the O(1)-size serial tails that run in parallel: the Horner
// chains of the MSMs of one batch (L and R of an IPA round; A_I, A_O, S; the five T commitments) -- ~0.1 ms each.
// Round 1 spawned a std::thread per chain and call (thread creation inside a 0.25 ms MSM); the workers are
// created once, on first use, and sleep on a condition variable in between.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace bp {

class HostWorkers {
public:
    explicit HostWorkers(int max_workers) : max_(max_workers) {}
    ~HostWorkers() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return (int)threads_.size(); }
    // runs fn(0), ..., fn(njobs - 1); fn(0) on the calling thread, the rest on the workers (njobs - 1 <= max_workers)
    void run(int njobs, const std::function<void(int)>& fn) {
        if (njobs <= 1) {
            if (njobs == 1) fn(0);
            return;
        }
        // as many workers as this batch needs (at most max_), woken one per job: a context that only ever runs L/R pairs
        // keeps a single helper thread, and nobody is woken for nothing on a host whose cores are busy with other provers
        while ((int)threads_.size() < njobs - 1 && (int)threads_.size() < max_) threads_.emplace_back([this] { loop(); });
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            next_ = 1;
            end_ = njobs;
            pending_ = njobs - 1;
        }
        for (int j = 1; j < njobs; j++) cv_work_.notify_one();
        fn(0);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void loop() {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_work_.wait(lk, [this] { return stop_ || next_ < end_; });
            if (stop_) return;
            const int job = next_++;
            const std::function<void(int)>* fn = fn_;
            lk.unlock();
            (*fn)(job);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_one();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)>* fn_ = nullptr;
    int next_ = 0, end_ = 0, pending_ = 0;
    bool stop_ = false;
    int max_ = 0;
};

}  // namespace bp
