// srb_workers.h -- host threads that issue the CUDA work of a multi-GPU context's devices side by side.
//
// The C ABI is called from ONE thread (the reference's solver is one process and one thread,
// irls_map_solver.cpp:192-265).  A device-resident solve on G devices issues about a dozen runtime calls per
// device and line-search step; issued one device after the other from that one thread they cost more than the
// kernels they start (cfg3 on 8 GPUs: ~100 calls per step against 13 us of tile kernel per device).  DeviceWorkers
// keeps G - 1 helper threads: run(count, job) executes job(0) on the calling thread and job(1 .. count-1) on the
// helpers and returns when all are done -- a fork/join "round" whose end is also the ordering point between
// devices (e.g. every device has RECORDED its event before any device WAITS on it in the next round).
// While a solve is running the helpers spin on a generation counter (a round costs about a microsecond);
// between solves they sleep on a condition variable.  Plain C++11, no CUDA: unit-tested on the CPU
// (tests/test_workers.py).
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace srb {

class DeviceWorkers {
 public:
  explicit DeviceWorkers(int helpers) : helpers_(helpers < 0 ? 0 : helpers) {
    threads_.reserve(helpers_);
    try {
      for (int i = 0; i < helpers_; ++i) threads_.emplace_back([this, i] { loop(i + 1); });
    } catch (...) {  // the system refused a thread: release the ones that started, let the caller fall back
      shutdown();
      throw;
    }
  }
  ~DeviceWorkers() { shutdown(); }
  DeviceWorkers(const DeviceWorkers&) = delete;
  DeviceWorkers& operator=(const DeviceWorkers&) = delete;

  int capacity() const { return helpers_ + 1; }
  // helpers spin (true: a solve is running, rounds follow each other closely) or sleep (false) between rounds
  void set_hot(bool hot) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      hot_.store(hot, std::memory_order_release);
    }
    cv_.notify_all();
  }
  // job(i) for i in [0, count), count <= capacity(): i = 0 here, the rest on the helpers; returns when all are done.
  // Not re-entrant; always called from the thread that owns the context.
  void run(int count, const std::function<void(int)>& job) {
    if (count > capacity()) count = capacity();
    if (count <= 1 || helpers_ == 0) {
      if (count >= 1) job(0);
      return;
    }
    job_ = &job;
    count_ = count;
    pending_.store(helpers_, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);  // a helper about to sleep must not miss the new generation
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    job(0);
    for (unsigned spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins) relax(spins);
    job_ = nullptr;
  }

 private:
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_.store(true, std::memory_order_release);
    }
    cv_.notify_all();
    for (std::thread& t : threads_) t.join();
    threads_.clear();
  }
  static void relax(unsigned spins) {
#if defined(__x86_64__) || defined(__i386__)
    if (spins < 4096) {
      __builtin_ia32_pause();
      return;
    }
#endif
    (void)spins;
    std::this_thread::yield();
  }
  void loop(int index) {
    unsigned long long seen = 0;
    for (;;) {
      for (unsigned spins = 0; hot_.load(std::memory_order_acquire) && !stop_.load(std::memory_order_acquire) &&
                               gen_.load(std::memory_order_acquire) == seen; ++spins)
        relax(spins);
      if (gen_.load(std::memory_order_acquire) == seen) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] {
          return stop_.load(std::memory_order_acquire) || gen_.load(std::memory_order_acquire) != seen ||
                 hot_.load(std::memory_order_acquire);
        });
      }
      if (stop_.load(std::memory_order_acquire)) return;
      if (gen_.load(std::memory_order_acquire) == seen) continue;  // woken to spin: no round yet
      ++seen;  // rounds are strictly sequential: run() does not start one before the last has been joined
      if (index < count_) (*job_)(index);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }

  const int helpers_;
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<unsigned long long> gen_{0};
  std::atomic<int> pending_{0};
  std::atomic<bool> stop_{false}, hot_{false};
  const std::function<void(int)>* job_ = nullptr;
  int count_ = 0;
};

}  // namespace srb
