// chan.hpp — a Go channel for the C++ host side: unbuffered rendezvous by default, close(), `recv(v)`
// returning false once closed and drained.  The reference's seam is three such channels
// (reference raftpipe.go:3-7; all unbuffered: raft.go:65-66, server/main.go:30).
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <stdexcept>
#include <utility>

namespace raftsql {

struct ChanClosed : std::runtime_error {
  ChanClosed() : std::runtime_error("send on closed channel") {}
};

template <class T>
class Chan {
 public:
  explicit Chan(size_t buffered = 0) : cap_(buffered) {}

  // Blocks until a receiver has taken the value (unbuffered) or there is room (buffered).
  // `stop`, when given, aborts the wait: `select { case ch <- v: case <-stopc: }` (reference raft.go:89-93);
  // an aborted send withdraws its value — as in Go, the receiver never sees it — and returns false.
  bool send(T v, const std::atomic<bool> *stop = nullptr) {
    std::unique_lock<std::mutex> lk(mu_);
    if (closed_) throw ChanClosed();
    const unsigned long my = ++seq_;
    q_.emplace_back(my, std::move(v));
    cv_.notify_all();
    if (cap_ && q_.size() <= cap_) return true;
    while (queued(my)) {
      if (closed_) {
        withdraw(my);
        throw ChanClosed();
      }
      if (stop && stop->load()) {
        withdraw(my);
        return false;
      }
      cv_.wait_for(lk, std::chrono::milliseconds(20));
    }
    return true;
  }

  // false: closed and drained.  timeout_ms < 0: wait for ever; on timeout throws std::runtime_error.
  bool recv(T &out, long timeout_ms = -1) {
    std::unique_lock<std::mutex> lk(mu_);
    auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms < 0 ? 0 : timeout_ms);
    while (q_.empty()) {
      if (closed_) return false;
      if (timeout_ms >= 0 && std::chrono::steady_clock::now() >= deadline) throw std::runtime_error("recv timed out");
      cv_.wait_for(lk, std::chrono::milliseconds(20));
    }
    out = std::move(q_.front().second);
    q_.pop_front();
    cv_.notify_all();
    return true;
  }

  // `select { case v, ok := <-ch: ... default: }`: 1 = took a value, 0 = nothing offered right now,
  // -1 = closed and drained.
  int try_recv(T &out) {
    std::lock_guard<std::mutex> lk(mu_);
    if (!q_.empty()) {
      out = std::move(q_.front().second);
      q_.pop_front();
      cv_.notify_all();
      return 1;
    }
    return closed_ ? -1 : 0;
  }

  void close() {
    std::lock_guard<std::mutex> lk(mu_);
    closed_ = true;
    cv_.notify_all();
  }
  bool closed() {
    std::lock_guard<std::mutex> lk(mu_);
    return closed_;
  }

 private:
  // the queue is a handful of values at most (one per blocked sender): linear scans are fine
  bool queued(unsigned long id) const {
    for (const auto &p : q_)
      if (p.first == id) return true;
    return false;
  }
  void withdraw(unsigned long id) {
    for (auto it = q_.begin(); it != q_.end(); ++it)
      if (it->first == id) {
        q_.erase(it);
        return;
      }
  }

  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::pair<unsigned long, T>> q_;
  size_t cap_;
  bool closed_ = false;
  unsigned long seq_ = 0;
};

}  // namespace raftsql
