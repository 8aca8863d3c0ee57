// host_pool.hpp -- the host threads of the library, started once.
//
// Every host phase (the flattener's passes, the planner's tasks, the copies out of pinned memory) is a short
// parallel loop, a millisecond or two of work; starting and joining 16-32 std::threads for each of them costs as
// much as the loop.  The pool keeps the threads parked on a condition variable and hands them one loop at a time:
// run(n, fn, max_threads) executes fn(k) for k in [0, n), tasks taken dynamically, the caller working too, and
// returns when all are done.  A loop started from inside a pool task, or while another caller's loop is running,
// runs inline on the calling thread -- never a deadlock, never more threads than cores.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace pcs {

class HostPool {
 public:
  static HostPool& get() {
    static HostPool pool;
    return pool;
  }

  // fn(k) for k in [0, n) on at most max_threads threads (the caller included).  fn must not throw.
  void run(size_t n, const std::function<void(size_t)>& fn, unsigned max_threads) {
    if (n == 0) return;
    const unsigned want = static_cast<unsigned>(std::min<size_t>(n, std::max(1u, max_threads)));
    std::unique_lock<std::mutex> busy(job_mutex_, std::try_to_lock);
    if (want <= 1 || in_task() || !busy.owns_lock()) {
      for (size_t k = 0; k < n; ++k) fn(k);
      return;
    }
    grow(want - 1);
    {
      std::lock_guard<std::mutex> lock(m_);
      fn_ = &fn;
      n_ = n;
      next_.store(0, std::memory_order_relaxed);
      helpers_ = std::min<unsigned>(want - 1, static_cast<unsigned>(workers_.size()));
      pending_ = helpers_;
      ++generation_;
    }
    wake_.notify_all();
    work(fn, n);
    std::unique_lock<std::mutex> lock(m_);
    done_.wait(lock, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  HostPool() = default;
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lock(m_);
      stop_ = true;
      ++generation_;
    }
    wake_.notify_all();
    for (auto& t : workers_) t.join();
  }
  HostPool(const HostPool&) = delete;
  HostPool& operator=(const HostPool&) = delete;

  static bool& in_task() {
    thread_local bool flag = false;
    return flag;
  }

  void work(const std::function<void(size_t)>& fn, size_t n) {
    in_task() = true;
    for (size_t k = next_.fetch_add(1, std::memory_order_relaxed); k < n; k = next_.fetch_add(1, std::memory_order_relaxed)) fn(k);
    in_task() = false;
  }

  void grow(unsigned n_workers) {
    const unsigned cap = std::max(1u, std::min(64u, std::thread::hardware_concurrency())) - 1u;
    n_workers = std::min(n_workers, std::max(cap, 1u));
    std::lock_guard<std::mutex> lock(m_);
    while (workers_.size() < n_workers) {
      const unsigned index = static_cast<unsigned>(workers_.size());
      const uint64_t born = generation_;
      workers_.emplace_back([this, index, born] { loop(index, born); });
    }
  }

  void loop(unsigned index, uint64_t seen) {
    for (;;) {
      const std::function<void(size_t)>* fn = nullptr;
      size_t n = 0;
      {
        std::unique_lock<std::mutex> lock(m_);
        wake_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        if (index >= helpers_) continue;  // this loop does not need so many threads
        fn = fn_;
        n = n_;
      }
      work(*fn, n);
      {
        std::lock_guard<std::mutex> lock(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }

  std::mutex job_mutex_;  // one loop at a time
  std::mutex m_;
  std::condition_variable wake_, done_;
  std::vector<std::thread> workers_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t n_ = 0;
  std::atomic<size_t> next_{0};
  unsigned helpers_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

}  // namespace pcs
