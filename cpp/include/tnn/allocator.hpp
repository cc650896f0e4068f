// Allocators (reference: include/device/iallocator.hpp:15-30, device_allocator.hpp, pool_allocator.hpp:27-120,
// del_allocator_v2.hpp:73-336).
//   DeviceAllocator  pass-through (aligned host memory, optionally pinned; cudaMalloc on a GPU)
//   PoolAllocator    size-keyed free lists, a cached block is reused when it is <= 4x the request
//   DELAllocatorV2   the executor's WORKSPACE allocator: large slabs carved from BOTH ends.  flip() toggles the end
//                    new blocks come from, so the output of layer i (one end) and the output of layer i+1 (other
//                    end) never interleave and a freed activation immediately re-opens the bump region of its end.
//                    The allocation sequence of a training step is deterministic, so every step sees the same
//                    addresses -- the property CUDA-graph capture of the step relies on.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "runtime/symmetric_memory.hpp"
#include "tnn/device.hpp"

namespace tnn {

class IAllocator {
 public:
  virtual ~IAllocator() = default;
  virtual void* allocate(size_t bytes) = 0;
  virtual void deallocate(void* p) = 0;
  virtual const Device& device() const = 0;
  virtual size_t bytes_in_use() const { return 0; }
  virtual size_t bytes_reserved() const { return 0; }
};

// Owning handle of one allocation; freed through its allocator when the last reference drops.
struct Storage {
  void* ptr = nullptr;
  size_t bytes = 0;
  IAllocator* owner = nullptr;          // null: memory is borrowed (not freed here)
  std::shared_ptr<void> keep_alive;     // parent storage of a borrowed sub-range
  Device dev;
  ~Storage() {
    if (owner && ptr) owner->deallocate(ptr);
  }
};
using StoragePtr = std::shared_ptr<Storage>;

class DeviceAllocator : public IAllocator {
 public:
  explicit DeviceAllocator(Device d, bool pinned_host = false) : dev_(d), pinned_(pinned_host) {}
  void* allocate(size_t bytes) override;
  void deallocate(void* p) override;
  const Device& device() const override { return dev_; }
  size_t bytes_in_use() const override { return in_use_; }
  static DeviceAllocator& instance(const Device& d);      // per-device singleton
  static DeviceAllocator& pinned();                       // cudaMallocHost-backed host allocator

 private:
  Device dev_;
  bool pinned_;
  std::mutex mu_;
  std::map<void*, size_t> live_;
  size_t in_use_ = 0;
};
inline IAllocator& HostAllocator() { return DeviceAllocator::instance(getHost()); }
inline IAllocator& GPUAllocator(int i = 0) { return DeviceAllocator::instance(getGPU(i)); }

class PoolAllocator : public IAllocator {
 public:
  explicit PoolAllocator(IAllocator& backing) : backing_(backing) {}
  ~PoolAllocator() override;
  void* allocate(size_t bytes) override;
  void deallocate(void* p) override;
  const Device& device() const override { return backing_.device(); }
  size_t bytes_in_use() const override { return in_use_; }
  size_t bytes_reserved() const override { return reserved_; }
  void release_cached();                                   // return every cached block to the backing allocator
  static PoolAllocator& instance(const Device& d, flowHandle_t flow = defaultFlowHandle);

 private:
  IAllocator& backing_;
  std::mutex mu_;
  std::multimap<size_t, void*> free_;
  std::map<void*, size_t> size_of_;
  size_t in_use_ = 0, reserved_ = 0;
};

class DELAllocatorV2 : public IAllocator {
 public:
  enum Side { LEFT = 0, RIGHT = 1 };
  explicit DELAllocatorV2(IAllocator& backing, size_t slab_bytes = (size_t)64 << 20, size_t alignment = 256);
  ~DELAllocatorV2() override;
  static std::unique_ptr<DELAllocatorV2> create(const Device& d, flowHandle_t flow = defaultFlowHandle);

  void* allocate(size_t bytes) override;
  void deallocate(void* p) override;
  const Device& device() const override { return backing_.device(); }
  size_t bytes_in_use() const override { return in_use_; }
  size_t bytes_reserved() const override { return reserved_; }

  void flip() { side_ = side_ == LEFT ? RIGHT : LEFT; }
  void set_side(Side s) { side_ = s; }
  Side side() const { return side_; }
  void reserve(size_t bytes);          // make sure ONE slab can serve `bytes` without growing later
  void clear();                        // drop every slab (only legal with no live block)
  size_t num_slabs() const { return slabs_.size(); }
  size_t peak_in_use() const { return peak_; }
  size_t total_allocated() const { return in_use_; }
  bool poison_on_free = false;         // debug: fill freed blocks (host slabs only) with 0xDD

 private:
  struct Slab {
    char* base = nullptr;
    size_t size = 0;
    size_t left_top = 0;               // [0, left_top) belongs to the left end
    size_t right_bottom = 0;           // [right_bottom, size) belongs to the right end
    std::map<size_t, size_t> holes[2]; // freed blocks strictly inside an end: offset -> bytes
    size_t live = 0;
  };
  struct Block {
    int slab;
    size_t off, bytes;
    Side side;
  };
  bool try_slab(Slab& s, int idx, size_t n, void** out);
  void grow(size_t need);
  void release_block(Slab& s, const Block& b);

  IAllocator& backing_;
  size_t slab_bytes_, align_;
  Side side_ = LEFT;
  std::vector<Slab> slabs_;
  std::map<void*, Block> live_;
  std::mutex mu_;
  size_t in_use_ = 0, reserved_ = 0, peak_ = 0;
};

// One rank's view of a SymmetricMemory group (runtime/symmetric_memory.hpp: cuMemCreate + peer mappings + NVLS multicast
// window): blocks come out of the rank's slab, so peer(p, q) / multicast(p) translate a block to the same offset in rank
// q's slab / in the switch-reduced window.  This is the reference's "registered slab" allocator (ibv_allocator.hpp) with
// NVLink instead of a NIC on the other side.
class SymmetricAllocator : public IAllocator {
 public:
  SymmetricAllocator(std::shared_ptr<SymmetricMemory> group, int rank) : group_(std::move(group)), rank_(rank), dev_(DeviceType::GPU, group_->device_of(rank)) {}
  void* allocate(size_t bytes) override { return group_->allocate(rank_, bytes); }
  void deallocate(void* p) override { group_->deallocate(rank_, p); }
  const Device& device() const override { return dev_; }
  size_t bytes_in_use() const override { return group_->free_list(rank_).in_use(); }
  size_t bytes_reserved() const override { return group_->bytes(); }
  int rank() const { return rank_; }
  SymmetricMemory& group() { return *group_; }
  void* peer(const void* p, int other) const { return group_->peer(rank_, p, other); }
  void* multicast(const void* p) const { return group_->multicast(rank_, p); }      // nullptr without NVLS

 private:
  std::shared_ptr<SymmetricMemory> group_;
  int rank_;
  Device dev_;
};

}  // namespace tnn
