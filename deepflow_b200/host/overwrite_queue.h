/*
 * overwrite_queue.h — the droplet queue the ingester's decoders read from, mirrored in
 * C++ because this image has no Go toolchain.  Same contract as
 *   OverwriteQueue            server/libs/queue/queue.go:43-262
 *   QueueReader / QueueWriter server/libs/queue/interface.go:26-41
 * fixed power-of-two capacity, Put never blocks and overwrites the OLDEST pending items
 * when full (calling `release` on each overwritten non-nil item, queue.go:130-166),
 * Get/Gets block while the queue is empty, an optional flush indicator puts a nil item
 * every `flush_ms` (queue.go:103-111) which consumers treat as "flush now".
 */
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace dfhost {

class OverwriteQueue {
 public:
  using Item = void*;
  using Release = std::function<void(Item)>;

  OverwriteQueue(size_t size, Release release = nullptr, int flush_ms = 0) : release_(std::move(release)) {
    size_t s = 1;
    while (s < size) s <<= 1;  // queue.go:92-97
    size_ = s;
    items_.assign(s, nullptr);
    if (flush_ms > 0) {
      ticker_ = std::thread([this, flush_ms] {
        while (!closed_.load()) {
          std::this_thread::sleep_for(std::chrono::milliseconds(flush_ms));
          Item nil = nullptr;
          Put(&nil, 1);
        }
      });
    }
  }
  ~OverwriteQueue() { Close(); }
  void Close() {
    closed_.store(true);
    if (ticker_.joinable()) ticker_.join();
  }

  size_t Len() const { return pending_; }
  size_t Size() const { return size_; }
  uint64_t Overwritten() const { return overwritten_; }

  // false = "Requested size is larger than capacity" (OverflowError)
  bool Put(const Item* items, size_t n) {
    if (n > size_) return false;
    std::unique_lock<std::mutex> g(mu_);
    size_t free_size = size_ - pending_;
    if (n > free_size) {
      if (release_) {  // queue.go:147-160: everything from the oldest pending item up to the new write end
        size_t from = first_index(), to = write_cursor_ + n;
        if (to > size_) to &= size_ - 1;
        if (from <= to) {
          release_range(from, to);
        } else {
          release_range(from, size_);
          release_range(0, to);
        }
      }
      overwritten_ += n - free_size;
    }
    for (size_t i = 0; i < n; i++) items_[(write_cursor_ + i) & (size_ - 1)] = items[i];
    in_ += n;
    pending_ = pending_ + n < size_ ? pending_ + n : size_;
    write_cursor_ = (write_cursor_ + n) & (size_ - 1);
    g.unlock();
    cv_.notify_one();
    return true;
  }
  bool Put(Item item) { return Put(&item, 1); }

  Item Get() {  // blocks while empty
    std::unique_lock<std::mutex> g(mu_);
    cv_.wait(g, [this] { return pending_ > 0; });
    size_t f = first_index();
    Item it = items_[f];
    items_[f] = nullptr;
    pending_--;
    out_++;
    return it;
  }
  // fills out[0..cap) with up to `cap` items, blocks while empty; cap == 0 returns 0 once an item is pending
  size_t Gets(Item* out, size_t cap) {
    std::unique_lock<std::mutex> g(mu_);
    cv_.wait(g, [this] { return pending_ > 0; });
    size_t n = cap < pending_ ? cap : pending_;
    size_t f = first_index();
    for (size_t i = 0; i < n; i++) {
      size_t k = (f + i) & (size_ - 1);
      out[i] = items_[k];
      items_[k] = nullptr;
    }
    pending_ -= n;
    out_ += n;
    return n;
  }

 private:
  size_t first_index() const { return (write_cursor_ + size_ - pending_) & (size_ - 1); }
  void release_range(size_t a, size_t b) {
    for (size_t i = a; i < b; i++)
      if (items_[i]) release_(items_[i]);  // nil = flush indicator
  }
  std::vector<Item> items_;
  size_t size_ = 0, write_cursor_ = 0, pending_ = 0;
  uint64_t in_ = 0, out_ = 0, overwritten_ = 0;
  Release release_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<bool> closed_{false};
  std::thread ticker_;
};

}  // namespace dfhost
