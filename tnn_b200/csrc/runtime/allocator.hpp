// Offset allocators of the host runtime.  They manage OFFSETS inside slabs that the caller
// backs with real memory (a device slab from the CUDA caching allocator, a pinned host slab,
// or an NVLink peer-mapped symmetric slab), so one implementation serves every memory kind.
//
//  * DELAllocator  -- double-ended slab allocator: allocations bump from the LEFT or the RIGHT
//    end of a slab; flip() switches sides so layer i's output and layer i+1's output ping-pong
//    (reference: include/device/del_allocator_v2.hpp:73-336).  Freed ranges coalesce; empty
//    slabs are merged into one super-slab on the next growth.
//  * PoolAllocator -- size-keyed free list, a cached block is reused when it is at most 4x the
//    request (reference: include/device/pool_allocator.hpp:27-120).
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>
#include <vector>

namespace tnn {

struct Block {
  int slab;
  uint64_t offset;
  uint64_t size;
};

class DELAllocator {
 public:
  explicit DELAllocator(uint64_t slab_bytes = 64ull << 20, uint64_t align = 256)
      : slab_bytes_(slab_bytes), align_(align) {}

  // returns {slab index, offset}; a new slab (index == previous slab count) must be backed by
  // the caller with at least slab_size(index) bytes.
  Block allocate(uint64_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    bytes = round_up(std::max<uint64_t>(bytes, 1));
    for (int pass = 0; pass < 2; ++pass) {
      for (int s = 0; s < (int)slabs_.size(); ++s) {
        Slab& sl = slabs_[s];
        // free list first (first fit)
        for (auto it = sl.free.begin(); it != sl.free.end(); ++it) {
          if (it->second >= bytes) {
            uint64_t off = right_ ? it->first + it->second - bytes : it->first;
            uint64_t fo = it->first, fs = it->second;
            sl.free.erase(it);
            if (fs > bytes) {
              if (right_) sl.free[fo] = fs - bytes;
              else sl.free[fo + bytes] = fs - bytes;
            }
            sl.live[off] = bytes;
            in_use_ += bytes;
            peak_ = std::max(peak_, in_use_);
            return Block{s, off, bytes};
          }
        }
        // bump region between left and right cursors
        if (sl.right - sl.left >= bytes) {
          uint64_t off;
          if (right_) {
            sl.right -= bytes;
            off = sl.right;
          } else {
            off = sl.left;
            sl.left += bytes;
          }
          sl.live[off] = bytes;
          in_use_ += bytes;
          peak_ = std::max(peak_, in_use_);
          return Block{s, off, bytes};
        }
      }
      if (pass == 0) grow(bytes);
    }
    throw std::runtime_error("DELAllocator: allocation failed");
  }

  void free(int slab, uint64_t offset) {
    std::lock_guard<std::mutex> g(mu_);
    Slab& sl = slabs_.at(slab);
    auto it = sl.live.find(offset);
    if (it == sl.live.end()) throw std::runtime_error("DELAllocator: double free / unknown block");
    uint64_t size = it->second;
    sl.live.erase(it);
    in_use_ -= size;
    // return to a cursor if adjacent, else to the coalescing free list
    uint64_t lo = offset, hi = offset + size;
    auto nxt = sl.free.lower_bound(lo);
    if (nxt != sl.free.end() && nxt->first == hi) {
      hi += nxt->second;
      nxt = sl.free.erase(nxt);
    }
    if (nxt != sl.free.begin()) {
      auto prv = std::prev(nxt);
      if (prv->first + prv->second == lo) {
        lo = prv->first;
        sl.free.erase(prv);
      }
    }
    if (hi == sl.left) {
      sl.left = lo;
    } else if (lo == sl.right) {
      sl.right = hi;
    } else {
      sl.free[lo] = hi - lo;
    }
    // cursors may now touch free-list ranges
    bool moved = true;
    while (moved) {
      moved = false;
      if (!sl.free.empty()) {
        auto last = std::prev(sl.free.end());
        if (last->first + last->second == sl.left) {
          sl.left = last->first;
          sl.free.erase(last);
          moved = true;
          continue;
        }
        auto first = sl.free.begin();
        if (first->first == sl.right) {
          sl.right = first->first + first->second;
          sl.free.erase(first);
          moved = true;
        }
      }
    }
  }

  void flip() { right_ = !right_; }
  void set_side(bool right) { right_ = right; }
  bool side() const { return right_; }
  void reserve(uint64_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    uint64_t have = 0;
    for (auto& s : slabs_) have += s.size;
    if (have < bytes) grow(bytes - have);
  }
  // drop every allocation (slabs stay)
  void clear() {
    std::lock_guard<std::mutex> g(mu_);
    for (auto& s : slabs_) {
      s.left = 0;
      s.right = s.size;
      s.live.clear();
      s.free.clear();
    }
    in_use_ = 0;
  }
  int num_slabs() const { return (int)slabs_.size(); }
  uint64_t slab_size(int i) const { return slabs_.at(i).size; }
  uint64_t total_allocated() const { return in_use_; }
  uint64_t peak_allocated() const { return peak_; }
  void reset_peak() { peak_ = in_use_; }
  uint64_t total_reserved() const {
    uint64_t t = 0;
    for (auto& s : slabs_) t += s.size;
    return t;
  }

 private:
  struct Slab {
    uint64_t size;
    uint64_t left, right;                 // bump cursors
    std::map<uint64_t, uint64_t> free;    // offset -> size (holes)
    std::map<uint64_t, uint64_t> live;    // offset -> size
  };
  uint64_t round_up(uint64_t b) const { return (b + align_ - 1) / align_ * align_; }
  void grow(uint64_t need) {
    uint64_t sz = std::max(slab_bytes_, round_up(need));
    slabs_.push_back(Slab{sz, 0, sz, {}, {}});
  }
  uint64_t slab_bytes_, align_;
  bool right_ = false;
  std::vector<Slab> slabs_;
  uint64_t in_use_ = 0, peak_ = 0;
  std::mutex mu_;
};

class PoolAllocator {
 public:
  // returns a cached block id >= 0 when one fits (size <= 4x request), else -1 (caller
  // allocates a fresh block and registers it with add()).
  int64_t acquire(uint64_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = free_.lower_bound(bytes);
    if (it != free_.end() && it->first <= bytes * 4) {
      int64_t id = it->second;
      free_.erase(it);
      cached_bytes_ -= sizes_[id];
      return id;
    }
    return -1;
  }
  int64_t add(uint64_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    sizes_.push_back(bytes);
    total_bytes_ += bytes;
    return (int64_t)sizes_.size() - 1;
  }
  void release(int64_t id) {
    std::lock_guard<std::mutex> g(mu_);
    free_.emplace(sizes_.at(id), id);
    cached_bytes_ += sizes_[id];
  }
  uint64_t block_size(int64_t id) const { return sizes_.at(id); }
  uint64_t total_bytes() const { return total_bytes_; }
  uint64_t cached_bytes() const { return cached_bytes_; }

 private:
  std::multimap<uint64_t, int64_t> free_;
  std::vector<uint64_t> sizes_;
  uint64_t total_bytes_ = 0, cached_bytes_ = 0;
  std::mutex mu_;
};

}  // namespace tnn
