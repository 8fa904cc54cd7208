// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// Restatement of src/util/IndexThreadReduce.h:64-193: a fixed pool of worker threads that pull
// [first,end) in chunks of `stepSize` (dynamic assignment) and reduce a per-call `stats` vector;
// a worker that got no chunk still runs the callback once with (0,0) — the reference (ab)uses this
// to run setZero once per thread (EnergyFunctional.cpp:199-201).
#pragma once
#include <algorithm>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace orc {

struct Stats10 { double v[10]; };

class ThreadReduce {
 public:
  using Fn = std::function<void(int, int, Stats10 *, int)>;
  explicit ThreadReduce(int nthreads) : n_(nthreads) {
    isDone_.assign(n_, 0); gotOne_.assign(n_, 0);
    running_ = true; nextIndex_ = 0; maxIndex_ = 0; stepSize_ = 1; generation_ = 0;
    for (int i = 0; i < n_; i++) { isDone_[i] = 0; gotOne_[i] = 1; }
    for (int i = 0; i < n_; i++) workers_.emplace_back([this, i] { workerLoop(i); });
  }
  ~ThreadReduce() {
    { std::unique_lock<std::mutex> lk(m_); running_ = false; generation_++; }
    todo_.notify_all();
    for (auto &t : workers_) t.join();
  }
  int threads() const { return n_; }
  // IndexThreadReduce.h:64-121
  void reduce(const Fn &fn, int first, int end, int stepSize = 0) {
    memset(&stats, 0, sizeof(stats));
    if (stepSize == 0) stepSize = ((end - first) + n_ - 1) / n_;
    std::unique_lock<std::mutex> lk(m_);
    fn_ = fn; nextIndex_ = first; maxIndex_ = end; stepSize_ = stepSize;
    for (int i = 0; i < n_; i++) { isDone_[i] = 0; gotOne_[i] = 0; }
    generation_++;
    todo_.notify_all();
    done_.wait(lk, [this] { for (int i = 0; i < n_; i++) if (!isDone_[i]) return false; return true; });
    nextIndex_ = 0; maxIndex_ = 0;
  }
  Stats10 stats;

 private:
  void workerLoop(int idx) {  // IndexThreadReduce.h:147-193
    std::unique_lock<std::mutex> lk(m_);
    long seen = 0;
    while (running_) {
      int todo = 0; bool got = false;
      if (nextIndex_ < maxIndex_) { todo = nextIndex_; nextIndex_ += stepSize_; got = true; }
      if (got) {
        int hi = std::min(todo + stepSize_, maxIndex_);
        lk.unlock();
        Stats10 s; memset(&s, 0, sizeof(s));
        fn_(todo, hi, &s, idx);
        lk.lock();
        gotOne_[idx] = 1;
        for (int i = 0; i < 10; i++) stats.v[i] += s.v[i];
      } else {
        if (!gotOne_[idx]) {
          lk.unlock();
          Stats10 s; memset(&s, 0, sizeof(s));
          fn_(0, 0, &s, idx);
          lk.lock();
          gotOne_[idx] = 1;
          for (int i = 0; i < 10; i++) stats.v[i] += s.v[i];
        }
        isDone_[idx] = 1;
        done_.notify_all();
        seen = generation_;
        todo_.wait(lk, [this, seen] { return generation_ != seen; });
      }
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::vector<char> isDone_, gotOne_;
  std::mutex m_;
  std::condition_variable todo_, done_;
  int nextIndex_, maxIndex_, stepSize_;
  long generation_;
  bool running_;
  Fn fn_;
};

}  // namespace orc
