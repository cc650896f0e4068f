// Host runtime: devices / flows, allocators, tensors.  No torch, no Python (SURVEY 2.2 "C++ host runtime").
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <unordered_map>

#include "kernels/gemm_xt.cuh"
#include "kernels/simt_ops.cuh"
#include "kernels/umma_gemm.cuh"
#include "runtime/cpu_math.hpp"
#include "tnn/allocator.hpp"
#include "tnn/device.hpp"
#include "tnn/tensor.hpp"

namespace tnn {

// ===================================================================================== Device / Flow
size_t Device::total_memory() const {
  if (is_gpu()) {
    size_t f = 0, t = 0;
    set_current();
    TNN_CUDA(cudaMemGetInfo(&f, &t));
    return t;
  }
  FILE* fp = fopen("/proc/meminfo", "r");
  size_t kb = 0;
  if (fp) {
    char line[256];
    while (fgets(line, sizeof line, fp))
      if (sscanf(line, "MemTotal: %zu kB", &kb) == 1) break;
    fclose(fp);
  }
  return kb * 1024;
}
size_t Device::free_memory() const {
  if (is_gpu()) {
    size_t f = 0, t = 0;
    set_current();
    TNN_CUDA(cudaMemGetInfo(&f, &t));
    return f;
  }
  FILE* fp = fopen("/proc/meminfo", "r");
  size_t kb = 0;
  if (fp) {
    char line[256];
    while (fgets(line, sizeof line, fp))
      if (sscanf(line, "MemAvailable: %zu kB", &kb) == 1) break;
    fclose(fp);
  }
  return kb * 1024;
}

Event::Event(bool timing) {
  cudaError_t e = cudaEventCreateWithFlags(&ev_, timing ? cudaEventDefault : cudaEventDisableTiming);
  if (e != cudaSuccess) ev_ = nullptr;       // CPU-only box: events degrade to no-ops
}
Event::~Event() {
  if (ev_) cudaEventDestroy(ev_);
}
void Event::synchronize() const {
  if (ev_) TNN_CUDA(cudaEventSynchronize(ev_));
}
float Event::elapsed_ms(const Event& a, const Event& b) {
  float ms = 0.f;
  if (a.ev_ && b.ev_) TNN_CUDA(cudaEventElapsedTime(&ms, a.ev_, b.ev_));
  return ms;
}

Flow::Flow(Device dev, flowHandle_t id) : dev_(dev), id_(id) {
  if (dev_.is_gpu()) {
    dev_.set_current();
    TNN_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  }
}
Flow::~Flow() {
  if (stream_) cudaStreamDestroy(stream_);
}
void Flow::synchronize() const {
  if (stream_) TNN_CUDA(cudaStreamSynchronize(stream_));
}
void Flow::record(Event& e) const {
  if (stream_ && e.get()) TNN_CUDA(cudaEventRecord(e.get(), stream_));
}
void Flow::wait(const Event& e) const {
  if (stream_ && e.get()) TNN_CUDA(cudaStreamWaitEvent(stream_, e.get(), 0));
}
void Flow::wait(const Flow& other) const {
  if (!stream_ || !other.stream_ || &other == this) return;
  // a fresh event per edge: legal inside stream capture (becomes a graph dependency)
  cudaEvent_t ev;
  TNN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  TNN_CUDA(cudaEventRecord(ev, other.stream_));
  TNN_CUDA(cudaStreamWaitEvent(stream_, ev, 0));
  TNN_CUDA(cudaEventDestroy(ev));
}

DeviceManager::DeviceManager() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    n = 0;
    cudaGetLastError();
  }
  for (int i = 0; i < n; ++i) gpus_.emplace_back(DeviceType::GPU, i);
}
DeviceManager& DeviceManager::getInstance() {
  // intentionally leaked (as every singleton of the runtime): tensors owned by other statics may be released during
  // process teardown and must still find their allocator, flows and device manager alive
  static DeviceManager* m = new DeviceManager();
  return *m;
}
const Device& DeviceManager::getGPU(int i) const {
  if (i < 0 || i >= (int)gpus_.size()) throw std::out_of_range("DeviceManager: no GPU " + std::to_string(i) + " (" + std::to_string(gpus_.size()) + " visible)");
  return gpus_[(size_t)i];
}
Device DeviceManager::getDevice(const std::string& spec) const {
  std::string s = spec;
  for (auto& c : s) c = (char)toupper((unsigned char)c);
  int id = 0;
  const size_t colon = s.find(':');
  if (colon != std::string::npos) {
    id = atoi(s.c_str() + colon + 1);
    s = s.substr(0, colon);
  }
  if (s == "CPU" || s == "HOST") return host_;
  if (s == "GPU" || s == "CUDA") return getGPU(id);
  throw std::invalid_argument("DeviceManager: bad device spec '" + spec + "'");
}
Device DeviceManager::getDevice(DeviceType t, int id) const { return t == DeviceType::GPU ? getGPU(id) : host_; }
Flow& DeviceManager::getFlow(const Device& d, flowHandle_t h) {
  std::lock_guard<std::mutex> g(mu_);
  for (auto& f : flows_)
    if (f->device() == d && f->handle() == h) return *f;
  flows_.push_back(std::make_unique<Flow>(d, h));
  return *flows_.back();
}
std::vector<Device> DeviceManager::devices() const {
  std::vector<Device> v{host_};
  v.insert(v.end(), gpus_.begin(), gpus_.end());
  return v;
}

static thread_local std::unordered_map<int, flowHandle_t> tl_flow;   // device key -> current handle
static int dev_key(const Device& d) { return d.is_gpu() ? d.id() : -1; }
Flow& current_flow(const Device& d) {
  auto it = tl_flow.find(dev_key(d));
  return DeviceManager::getInstance().getFlow(d, it == tl_flow.end() ? defaultFlowHandle : it->second);
}
FlowGuard::FlowGuard(const Device& d, flowHandle_t h) : dev_(d) {
  auto it = tl_flow.find(dev_key(d));
  prev_ = it == tl_flow.end() ? defaultFlowHandle : it->second;
  tl_flow[dev_key(d)] = h;
}
FlowGuard::~FlowGuard() { tl_flow[dev_key(dev_)] = prev_; }

// ===================================================================================== allocators
void* DeviceAllocator::allocate(size_t bytes) {
  if (bytes == 0) bytes = 1;
  void* p = nullptr;
  if (dev_.is_gpu()) {
    dev_.set_current();
    TNN_CUDA(cudaMalloc(&p, bytes));
  } else if (pinned_) {
    TNN_CUDA(cudaMallocHost(&p, bytes));
  } else {
    if (posix_memalign(&p, 256, (bytes + 255) / 256 * 256) != 0) throw std::bad_alloc();
  }
  std::lock_guard<std::mutex> g(mu_);
  live_[p] = bytes;
  in_use_ += bytes;
  return p;
}
void DeviceAllocator::deallocate(void* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = live_.find(p);
    if (it != live_.end()) {
      in_use_ -= it->second;
      live_.erase(it);
    }
  }
  if (dev_.is_gpu()) cudaFree(p);
  else if (pinned_) cudaFreeHost(p);
  else free(p);
}
DeviceAllocator& DeviceAllocator::instance(const Device& d) {
  static std::mutex mu;
  static auto* all = new std::map<int, DeviceAllocator*>();
  std::lock_guard<std::mutex> g(mu);
  auto& a = (*all)[dev_key(d)];
  if (!a) a = new DeviceAllocator(d);
  return *a;
}
DeviceAllocator& DeviceAllocator::pinned() {
  static DeviceAllocator* a = new DeviceAllocator(getHost(), DeviceManager::getInstance().has_gpu());
  return *a;
}

PoolAllocator::~PoolAllocator() { release_cached(); }
void* PoolAllocator::allocate(size_t bytes) {
  bytes = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = free_.lower_bound(bytes);
    if (it != free_.end() && it->first <= 4 * bytes) {     // reuse only blocks <= 4x the request (pool_allocator.hpp:104)
      void* p = it->second;
      in_use_ += it->first;
      free_.erase(it);
      return p;
    }
  }
  void* p = backing_.allocate(bytes);
  std::lock_guard<std::mutex> g(mu_);
  size_of_[p] = bytes;
  in_use_ += bytes;
  reserved_ += bytes;
  return p;
}
void PoolAllocator::deallocate(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> g(mu_);
  auto it = size_of_.find(p);
  if (it == size_of_.end()) return;
  in_use_ -= it->second;
  free_.emplace(it->second, p);
}
void PoolAllocator::release_cached() {
  std::lock_guard<std::mutex> g(mu_);
  for (auto& kv : free_) {
    reserved_ -= kv.first;
    size_of_.erase(kv.second);
    backing_.deallocate(kv.second);
  }
  free_.clear();
}
PoolAllocator& PoolAllocator::instance(const Device& d, flowHandle_t flow) {
  static std::mutex mu;
  static auto* all = new std::map<std::pair<int, flowHandle_t>, PoolAllocator*>();
  std::lock_guard<std::mutex> g(mu);
  auto& a = (*all)[{dev_key(d), flow}];
  if (!a) a = new PoolAllocator(DeviceAllocator::instance(d));
  return *a;
}

DELAllocatorV2::DELAllocatorV2(IAllocator& backing, size_t slab_bytes, size_t alignment)
    : backing_(backing), slab_bytes_(slab_bytes), align_(alignment) {}
DELAllocatorV2::~DELAllocatorV2() {
  for (auto& s : slabs_) backing_.deallocate(s.base);
}
std::unique_ptr<DELAllocatorV2> DELAllocatorV2::create(const Device& d, flowHandle_t) {
  return std::make_unique<DELAllocatorV2>(DeviceAllocator::instance(d));
}
void DELAllocatorV2::grow(size_t need) {
  Slab s;
  // slabs double so that a long step ends up in a handful of large slabs
  size_t want = std::max(slab_bytes_, need);
  if (!slabs_.empty()) want = std::max(want, std::min<size_t>(slabs_.back().size * 2, (size_t)4 << 30));
  s.size = (want + align_ - 1) / align_ * align_;
  s.base = static_cast<char*>(backing_.allocate(s.size));
  s.left_top = 0;
  s.right_bottom = s.size;
  reserved_ += s.size;
  slabs_.push_back(std::move(s));
}
bool DELAllocatorV2::try_slab(Slab& s, int idx, size_t n, void** out) {
  auto& holes = s.holes[side_];
  for (auto it = holes.begin(); it != holes.end(); ++it) {
    if (it->second >= n) {                         // first fit inside this end's freed blocks
      const size_t off = it->first, sz = it->second;
      holes.erase(it);
      size_t at = off;
      if (side_ == LEFT) {
        if (sz > n) holes[off + n] = sz - n;
      } else {
        at = off + sz - n;                         // keep the remainder on the low side of a right-end hole
        if (sz > n) holes[off] = sz - n;
      }
      *out = s.base + at;
      live_[*out] = Block{idx, at, n, side_};
      ++s.live;
      return true;
    }
  }
  if (s.right_bottom - s.left_top >= n) {
    size_t at;
    if (side_ == LEFT) {
      at = s.left_top;
      s.left_top += n;
    } else {
      s.right_bottom -= n;
      at = s.right_bottom;
    }
    *out = s.base + at;
    live_[*out] = Block{idx, at, n, side_};
    ++s.live;
    return true;
  }
  return false;
}
void* DELAllocatorV2::allocate(size_t bytes) {
  const size_t n = (std::max<size_t>(bytes, 1) + align_ - 1) / align_ * align_;
  std::lock_guard<std::mutex> g(mu_);
  void* p = nullptr;
  for (size_t i = 0; i < slabs_.size(); ++i)
    if (try_slab(slabs_[i], (int)i, n, &p)) break;
  if (!p) {
    grow(n);
    if (!try_slab(slabs_.back(), (int)slabs_.size() - 1, n, &p)) throw std::bad_alloc();
  }
  in_use_ += n;
  peak_ = std::max(peak_, in_use_);
  return p;
}
void DELAllocatorV2::release_block(Slab& s, const Block& b) {
  if (poison_on_free && !backing_.device().is_gpu()) std::memset(s.base + b.off, 0xDD, b.bytes);
  auto& holes = s.holes[b.side];
  if (b.side == LEFT) {
    if (b.off + b.bytes == s.left_top) {
      s.left_top = b.off;
      // freed blocks that now touch the top of the bump region fold back into it
      while (!holes.empty()) {
        auto last = std::prev(holes.end());
        if (last->first + last->second != s.left_top) break;
        s.left_top = last->first;
        holes.erase(last);
      }
      return;
    }
  } else {
    if (b.off == s.right_bottom) {
      s.right_bottom = b.off + b.bytes;
      while (!holes.empty()) {
        auto first = holes.begin();
        if (first->first != s.right_bottom) break;
        s.right_bottom = first->first + first->second;
        holes.erase(first);
      }
      return;
    }
  }
  // interior block: insert and coalesce with its neighbours
  size_t off = b.off, sz = b.bytes;
  auto next = holes.lower_bound(off);
  if (next != holes.begin()) {
    auto prev = std::prev(next);
    if (prev->first + prev->second == off) {
      off = prev->first;
      sz += prev->second;
      holes.erase(prev);
    }
  }
  next = holes.lower_bound(off + sz);
  if (next != holes.end() && next->first == off + sz) {
    sz += next->second;
    holes.erase(next);
  }
  holes[off] = sz;
}
void DELAllocatorV2::deallocate(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.find(p);
  if (it == live_.end()) return;
  const Block b = it->second;
  live_.erase(it);
  Slab& s = slabs_[(size_t)b.slab];
  --s.live;
  in_use_ -= b.bytes;
  release_block(s, b);
}
void DELAllocatorV2::reserve(size_t bytes) {
  std::lock_guard<std::mutex> g(mu_);
  for (auto& s : slabs_)
    if (s.right_bottom - s.left_top >= bytes) return;
  // no slab can take it: when nothing is live, merge everything into ONE slab of the combined size
  if (live_.empty() && !slabs_.empty()) {
    size_t total = 0;
    for (auto& s : slabs_) {
      total += s.size;
      backing_.deallocate(s.base);
    }
    slabs_.clear();
    reserved_ = 0;
    bytes = std::max(bytes, total);
  }
  const size_t keep = slab_bytes_;
  slab_bytes_ = std::max(slab_bytes_, bytes);
  grow(bytes);
  slab_bytes_ = keep;
}
void DELAllocatorV2::clear() {
  std::lock_guard<std::mutex> g(mu_);
  if (!live_.empty()) throw std::runtime_error("DELAllocatorV2::clear with " + std::to_string(live_.size()) + " live blocks");
  for (auto& s : slabs_) backing_.deallocate(s.base);
  slabs_.clear();
  reserved_ = 0;
  in_use_ = 0;
}

// ===================================================================================== Tensor
std::string shape_to_string(const Shape& s) {
  std::string o = "[";
  for (size_t i = 0; i < s.size(); ++i) o += (i ? ", " : "") + std::to_string(s[i]);
  return o + "]";
}
static cudaStream_t stream_of(const Device& d) { return current_flow(d).stream(); }

Tensor make_tensor(IAllocator& alloc, DType_t dt, const Shape& shape) {
  auto st = std::make_shared<Storage>();
  st->bytes = std::max<size_t>(shape_numel(shape) * get_dtype_size(dt), 1);
  st->ptr = alloc.allocate(st->bytes);
  st->owner = &alloc;
  st->dev = alloc.device();
  return std::make_shared<TensorImpl>(dt, shape, std::move(st), 0, &alloc);
}
Tensor make_tensor(DType_t dt, const Shape& shape, const Device& d) {
  return make_tensor(d.is_gpu() ? static_cast<IAllocator&>(PoolAllocator::instance(d)) : static_cast<IAllocator&>(DeviceAllocator::instance(d)), dt, shape);
}
Tensor make_tensor_from(const std::vector<float>& values, const Shape& shape, const Device& d) {
  if (values.size() != shape_numel(shape)) throw std::invalid_argument("make_tensor_from: size mismatch");
  Tensor h = make_tensor(DType_t::FP32, shape, getHost());
  std::memcpy(h->data(), values.data(), values.size() * 4);
  return d.is_gpu() ? h->to_device(d) : h;
}
Tensor make_tensor_view(void* ptr, DType_t dt, const Shape& shape, const Device& d, std::shared_ptr<void> owner) {
  auto st = std::make_shared<Storage>();
  st->ptr = ptr;
  st->bytes = shape_numel(shape) * get_dtype_size(dt);
  st->owner = nullptr;
  st->keep_alive = std::move(owner);
  st->dev = d;
  return std::make_shared<TensorImpl>(dt, shape, std::move(st), 0, nullptr);
}
Tensor create_like(const TensorImpl& t, IAllocator* alloc) {
  IAllocator* a = alloc ? alloc : t.allocator();
  return a ? make_tensor(*a, t.data_type(), t.shape()) : make_tensor(t.data_type(), t.shape(), t.device());
}

Tensor TensorImpl::span(size_t offset_elems, const Shape& shape) const {
  const size_t es = get_dtype_size(dtype_);
  if (offset_ + (offset_elems + shape_numel(shape)) * es > storage_->bytes)
    throw std::out_of_range("Tensor::span: view " + shape_to_string(shape) + " at " + std::to_string(offset_elems) + " exceeds storage");
  return std::make_shared<TensorImpl>(dtype_, shape, storage_, offset_ + offset_elems * es, alloc_);
}
void TensorImpl::share_from(const TensorImpl& o) {
  dtype_ = o.dtype_;
  shape_ = o.shape_;
  storage_ = o.storage_;
  offset_ = o.offset_;
  alloc_ = o.alloc_;
}
Tensor TensorImpl::slice_batch(size_t begin, size_t count) const {
  if (shape_.empty() || begin + count > shape_[0]) throw std::out_of_range("Tensor::slice_batch");
  Shape s = shape_;
  s[0] = count;
  return span(begin * stride(0), s);
}

static void raw_copy(void* dst, const Device& dd, const void* src, const Device& sd, size_t bytes) {
  if (bytes == 0) return;
  if (!dd.is_gpu() && !sd.is_gpu()) {
    std::memmove(dst, src, bytes);
    return;
  }
  const Device& g = dd.is_gpu() ? dd : sd;
  g.set_current();
  cudaStream_t s = stream_of(g);
  const cudaMemcpyKind k = dd.is_gpu() ? (sd.is_gpu() ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice) : cudaMemcpyDeviceToHost;
  TNN_CUDA(cudaMemcpyAsync(dst, src, bytes, k, s));
  // pageable host memory on either side: the call is complete only after the stream drains
  if (k != cudaMemcpyDeviceToDevice) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &st);
    if (st == cudaStreamCaptureStatusNone) TNN_CUDA(cudaStreamSynchronize(s));
  }
}
void TensorImpl::copy_from(const TensorImpl& src) {
  if (src.dtype_ != dtype_ || src.size() != size())
    throw std::invalid_argument("Tensor::copy_from: dtype / size mismatch (" + std::string(dtype_to_string(src.dtype_)) + shape_to_string(src.shape_) + " -> " + dtype_to_string(dtype_) + shape_to_string(shape_) + ")");
  raw_copy(data(), device(), src.data(), src.device(), nbytes());
}
void TensorImpl::copy_from_host(const void* src, size_t bytes) { raw_copy(data(), device(), src, getHost(), std::min(bytes, nbytes())); }
void TensorImpl::copy_to_host(void* dst, size_t bytes) const { raw_copy(dst, getHost(), data(), device(), std::min(bytes, nbytes())); }
Tensor TensorImpl::to_device(const Device& d, IAllocator* alloc) const {
  Tensor t = alloc ? make_tensor(*alloc, dtype_, shape_) : make_tensor(dtype_, shape_, d);
  t->copy_from(*this);
  return t;
}
Tensor TensorImpl::to_host() const { return to_device(getHost()); }
Tensor TensorImpl::clone(IAllocator* alloc) const {
  Tensor t = create_like(*this, alloc);
  t->copy_from(*this);
  return t;
}
void TensorImpl::resize(const Shape& shape) {
  const size_t need = shape_numel(shape) * get_dtype_size(dtype_);
  if (offset_ + need > storage_->bytes) {
    Tensor t = alloc_ ? make_tensor(*alloc_, dtype_, shape) : make_tensor(dtype_, shape, device());
    storage_ = t->storage_;
    offset_ = 0;
  }
  shape_ = shape;
}

// ---- host helpers: read / write any float dtype through double
static double host_get(const void* p, DType_t dt, size_t i) {
  switch (dt) {
    case DType_t::FP32: return static_cast<const float*>(p)[i];
    case DType_t::FP64: return static_cast<const double*>(p)[i];
    case DType_t::BF16: return bf16_to_float(static_cast<const bf16_t*>(p)[i]);
    case DType_t::FP16: return fp16_to_float(static_cast<const fp16_t*>(p)[i]);
    case DType_t::INT32_T: return static_cast<const int32_t*>(p)[i];
    case DType_t::INT64_T: return (double)static_cast<const int64_t*>(p)[i];
    case DType_t::SIZE_T: return (double)static_cast<const uint64_t*>(p)[i];
    case DType_t::UINT8_T: case DType_t::BYTE: case DType_t::BOOL: return static_cast<const uint8_t*>(p)[i];
    default: throw std::runtime_error("host_get: dtype");
  }
}
static void host_set(void* p, DType_t dt, size_t i, double v) {
  switch (dt) {
    case DType_t::FP32: static_cast<float*>(p)[i] = (float)v; break;
    case DType_t::FP64: static_cast<double*>(p)[i] = v; break;
    case DType_t::BF16: static_cast<bf16_t*>(p)[i] = float_to_bf16((float)v); break;
    case DType_t::FP16: static_cast<fp16_t*>(p)[i] = float_to_fp16((float)v); break;
    case DType_t::INT32_T: static_cast<int32_t*>(p)[i] = (int32_t)v; break;
    case DType_t::INT64_T: static_cast<int64_t*>(p)[i] = (int64_t)v; break;
    case DType_t::SIZE_T: static_cast<uint64_t*>(p)[i] = (uint64_t)v; break;
    case DType_t::UINT8_T: case DType_t::BYTE: case DType_t::BOOL: static_cast<uint8_t*>(p)[i] = (uint8_t)v; break;
    default: throw std::runtime_error("host_set: dtype");
  }
}
// floating types of the templated GPU tensor ops (tensor_ops.cu) and their XtDType codes
static bool gpu_float(DType_t dt) { return dt == DType_t::FP32 || dt == DType_t::BF16 || dt == DType_t::FP16 || dt == DType_t::FP64; }
static int xt_code(DType_t dt) {
  switch (dt) {
    case DType_t::FP32: return XT_F32;
    case DType_t::FP16: return XT_F16;
    case DType_t::FP64: return XT_F64;
    default: return XT_BF16;
  }
}

void TensorImpl::fill(double v) {
  const size_t n = size();
  if (is_gpu()) {
    device().set_current();
    cudaStream_t s = stream_of(device());
    if (v == 0.0) {
      TNN_CUDA(cudaMemsetAsync(data(), 0, nbytes(), s));
    } else if (dtype_ == DType_t::FP32) {
      fill_f32(data_as<float>(), (long long)n, (float)v, s);
    } else {
      Tensor h = make_tensor(dtype_, shape_, getHost());
      h->fill(v);
      copy_from(*h);
    }
    return;
  }
  for (size_t i = 0; i < n; ++i) host_set(data(), dtype_, i, v);
}
void TensorImpl::fill_random_uniform(double lo, double hi, uint64_t seed) {
  if (is_gpu() && dtype_ == DType_t::FP32) {
    device().set_current();
    philox_uniform_f32(data_as<float>(), (long long)size(), (float)lo, (float)hi, seed, stream_of(device()));
    return;
  }
  if (is_gpu()) {
    Tensor h = make_tensor(dtype_, shape_, getHost());
    h->fill_random_uniform(lo, hi, seed);
    copy_from(*h);
    return;
  }
  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<double> d(lo, hi);
  for (size_t i = 0, n = size(); i < n; ++i) host_set(data(), dtype_, i, d(gen));
}
void TensorImpl::fill_random_normal(double mean, double stddev, uint64_t seed) {
  if (is_gpu() && dtype_ == DType_t::FP32) {
    device().set_current();
    philox_normal_f32(data_as<float>(), (long long)size(), (float)mean, (float)stddev, seed, stream_of(device()));
    return;
  }
  if (is_gpu()) {
    Tensor h = make_tensor(dtype_, shape_, getHost());
    h->fill_random_normal(mean, stddev, seed);
    copy_from(*h);
    return;
  }
  std::mt19937_64 gen(seed);
  std::normal_distribution<double> d(mean, stddev);
  for (size_t i = 0, n = size(); i < n; ++i) host_set(data(), dtype_, i, d(gen));
}

static void binary_inplace(TensorImpl& a, const TensorImpl* b, double scalar, int op) {
  if (b && (b->size() != a.size() || b->data_type() != a.data_type())) throw std::invalid_argument("Tensor arithmetic: shape / dtype mismatch");
  const size_t n = a.size();
  if (a.is_gpu()) {
    if (!gpu_float(a.data_type())) throw std::runtime_error("GPU tensor arithmetic supports FP32 / BF16 / FP16 / FP64");
    a.device().set_current();
    tensor_binary(a.data(), b ? b->data() : nullptr, (float)scalar, a.data(), (long long)n, op, xt_code(a.data_type()), stream_of(a.device()));
    return;
  }
  for (size_t i = 0; i < n; ++i) {
    const double x = host_get(a.data(), a.data_type(), i), y = b ? host_get(b->data(), b->data_type(), i) : scalar;
    double r = 0;
    switch (op) {
      case TOP_ADD: r = x + y; break;
      case TOP_SUB: r = x - y; break;
      case TOP_MUL: r = x * y; break;
      case TOP_DIV: r = x / y; break;
      default: throw std::runtime_error("binary op");
    }
    host_set(a.data(), a.data_type(), i, r);
  }
}
void TensorImpl::add(const TensorImpl& o) { binary_inplace(*this, &o, 0, TOP_ADD); }
void TensorImpl::sub(const TensorImpl& o) { binary_inplace(*this, &o, 0, TOP_SUB); }
void TensorImpl::mul(const TensorImpl& o) { binary_inplace(*this, &o, 0, TOP_MUL); }
void TensorImpl::div(const TensorImpl& o) { binary_inplace(*this, &o, 0, TOP_DIV); }
void TensorImpl::add_scalar(double s) { binary_inplace(*this, nullptr, s, TOP_ADD); }
void TensorImpl::mul_scalar(double s) { binary_inplace(*this, nullptr, s, TOP_MUL); }
void TensorImpl::axpy(double a, const TensorImpl& x) {
  if (x.size() != size() || x.dtype_ != dtype_) throw std::invalid_argument("Tensor::axpy: mismatch");
  if (is_gpu()) {
    if (dtype_ != DType_t::FP32) throw std::runtime_error("GPU axpy: FP32");
    device().set_current();
    axpy_f32(data_as<float>(), x.data_as<float>(), (float)a, (long long)size(), stream_of(device()));
    return;
  }
  for (size_t i = 0, n = size(); i < n; ++i) host_set(data(), dtype_, i, host_get(data(), dtype_, i) + a * host_get(x.data(), dtype_, i));
}

static double reduce_host(const TensorImpl& t, int mode, double mu) {
  double acc = mode == 3 ? INFINITY : mode == 4 ? -INFINITY : 0;
  for (size_t i = 0, n = t.size(); i < n; ++i) {
    const double v = host_get(t.data(), t.data_type(), i);
    switch (mode) {
      case 0: acc += v; break;
      case 2: acc += (v - mu) * (v - mu); break;
      case 3: acc = std::min(acc, v); break;
      case 4: acc = std::max(acc, v); break;
    }
  }
  return acc;
}
static double reduce_any(const TensorImpl& t, int mode, double mu) {
  if (t.is_gpu() && gpu_float(t.data_type()) && mode <= 2) {
    t.device().set_current();
    Tensor out = make_tensor(DType_t::FP32, {1}, t.device());
    tensor_reduce(t.data(), nullptr, (float)mu, (long long)t.size(), mode, xt_code(t.data_type()), out->data_as<float>(), stream_of(t.device()));
    float h = 0;
    out->copy_to_host(&h, 4);
    return h;
  }
  if (t.is_gpu()) return reduce_host(*t.to_host(), mode, mu);
  return reduce_host(t, mode, mu);
}
double TensorImpl::sum() const { return reduce_any(*this, 0, 0); }
double TensorImpl::variance() const { return reduce_any(*this, 2, mean()) / (double)size(); }
double TensorImpl::min() const { return reduce_any(*this, 3, 0); }
double TensorImpl::max() const { return reduce_any(*this, 4, 0); }

void TensorImpl::save(std::ostream& os) const {
  const uint32_t dt = (uint32_t)dtype_;
  const uint64_t nd = shape_.size();
  os.write(reinterpret_cast<const char*>(&dt), 4);
  os.write(reinterpret_cast<const char*>(&nd), 8);
  for (size_t d : shape_) {
    const uint64_t v = d;
    os.write(reinterpret_cast<const char*>(&v), 8);
  }
  if (is_gpu()) {
    Tensor h = to_host();
    os.write(static_cast<const char*>(h->data()), (std::streamsize)nbytes());
  } else {
    os.write(static_cast<const char*>(data()), (std::streamsize)nbytes());
  }
}
static void read_header(std::istream& is, DType_t* dt, Shape* shape) {
  uint32_t d = 0;
  uint64_t nd = 0;
  is.read(reinterpret_cast<char*>(&d), 4);
  is.read(reinterpret_cast<char*>(&nd), 8);
  if (!is || d >= (uint32_t)DType_t::UNKNOWN || nd > 16) throw std::runtime_error("tensor record: bad header");
  shape->resize(nd);
  for (uint64_t i = 0; i < nd; ++i) {
    uint64_t v = 0;
    is.read(reinterpret_cast<char*>(&v), 8);
    (*shape)[i] = (size_t)v;
  }
  *dt = (DType_t)d;
}
Tensor load(std::istream& is, IAllocator& alloc) {
  DType_t dt;
  Shape shape;
  read_header(is, &dt, &shape);
  Tensor h = make_tensor(dt, shape, getHost());
  is.read(static_cast<char*>(h->data()), (std::streamsize)h->nbytes());
  if (!is) throw std::runtime_error("tensor record: truncated payload");
  if (alloc.device().is_gpu()) return h->to_device(alloc.device(), &alloc);
  return h;
}
void load_into(std::istream& is, TensorImpl& dst) {
  DType_t dt;
  Shape shape;
  read_header(is, &dt, &shape);
  if (dt != dst.data_type() || shape_numel(shape) != dst.size())
    throw std::runtime_error("load_into: record is " + std::string(dtype_to_string(dt)) + shape_to_string(shape) + ", tensor is " + dtype_to_string(dst.data_type()) + shape_to_string(dst.shape()));
  Tensor h = make_tensor(dt, dst.shape(), getHost());
  is.read(static_cast<char*>(h->data()), (std::streamsize)h->nbytes());
  if (!is) throw std::runtime_error("tensor record: truncated payload");
  dst.copy_from(*h);
}

Tensor dtype_cast(const TensorImpl& t, DType_t to, IAllocator* alloc) {
  Tensor out = alloc ? make_tensor(*alloc, to, t.shape()) : make_tensor(to, t.shape(), t.device());
  if (to == t.data_type()) {
    out->copy_from(t);
    return out;
  }
  if (t.is_gpu()) {
    t.device().set_current();
    cudaStream_t s = stream_of(t.device());
    if (t.data_type() == DType_t::FP32 && to == DType_t::BF16) { cast_f32_to_bf16(t.data_as<float>(), out->data(), (long long)t.size(), s); return out; }
    if (t.data_type() == DType_t::BF16 && to == DType_t::FP32) { cast_bf16_to_f32(t.data(), out->data_as<float>(), (long long)t.size(), s); return out; }
    if (gpu_float(t.data_type()) && gpu_float(to)) {   // any pair of FP16 / BF16 / FP32 / FP64 on the device
      tensor_cast_any(t.data(), xt_code(t.data_type()), out->data(), xt_code(to), (long long)t.size(), s);
      return out;
    }
    Tensor h = dtype_cast(*t.to_host(), to);
    out->copy_from(*h);
    return out;
  }
  for (size_t i = 0, n = t.size(); i < n; ++i) host_set(out->data(), to, i, host_get(t.data(), t.data_type(), i));
  return out;
}

Tensor matmul(const TensorImpl& a, const TensorImpl& b, bool a_mn, bool b_mn, double alpha) {
  if (a.shape().size() != 2 || b.shape().size() != 2 || a.data_type() != b.data_type() || a.device() != b.device())
    throw std::invalid_argument("matmul: two 2-D tensors of one dtype on one device");
  const size_t M = a_mn ? a.shape()[1] : a.shape()[0], K = a_mn ? a.shape()[0] : a.shape()[1];
  const size_t N = b_mn ? b.shape()[1] : b.shape()[0], Kb = b_mn ? b.shape()[0] : b.shape()[1];
  if (K != Kb) throw std::invalid_argument("matmul: inner dimensions differ");
  const DType_t dt = a.data_type();
  Tensor out = make_tensor(dt, {M, N}, a.device());
  const long long lda = (long long)a.shape()[1], ldb = (long long)b.shape()[1];
  if (a.is_gpu()) {
    if (!gpu_float(dt)) throw std::runtime_error("matmul: GPU tensors are BF16 / FP16 / FP32 / FP64");
    a.device().set_current();
    cudaStream_t s = stream_of(a.device());
    if (dt == DType_t::FP64) {
      gemm_f64(a.data_as<double>(), b.data_as<double>(), (int)M, (int)N, (int)K, lda, ldb, a_mn, b_mn, alpha, nullptr, out->data_as<double>(),
               (long long)N, false, s);
    } else if (dt == DType_t::BF16) {
      Epilogue e;
      e.out = out->data();
      e.ldc = (long long)N;
      e.alpha = (float)alpha;
      gemm_bf16(a.data(), b.data(), (int)M, (int)N, (int)K, lda, ldb, a_mn, b_mn, e, 1, s);
    } else {
      XtEpilogue e;
      e.out = out->data();
      e.ldc = (long long)N;
      e.alpha = (float)alpha;
      if (dt == DType_t::FP16) gemm_f16(a.data(), b.data(), (int)M, (int)N, (int)K, lda, ldb, a_mn, b_mn, e, s);
      else gemm_tf32(a.data(), b.data(), (int)M, (int)N, (int)K, lda, ldb, a_mn, b_mn, e, s);
    }
    return out;
  }
  if (dt == DType_t::FP32) {
    cpu_sgemm(a.data_as<float>(), b.data_as<float>(), out->data_as<float>(), (int64_t)M, (int64_t)N, (int64_t)K, a_mn, !b_mn, (float)alpha, 0.f,
              (int64_t)lda, (int64_t)ldb, (int64_t)N);
    return out;
  }
  for (size_t m = 0; m < M; ++m)
    for (size_t n = 0; n < N; ++n) {
      double acc = 0;
      for (size_t k = 0; k < K; ++k)
        acc += host_get(a.data(), dt, a_mn ? k * (size_t)lda + m : m * (size_t)lda + k) * host_get(b.data(), dt, b_mn ? k * (size_t)ldb + n : n * (size_t)ldb + k);
      host_set(out->data(), dt, m * N + n, alpha * acc);
    }
  return out;
}

std::vector<float> to_float_vector(const TensorImpl& t) {
  Tensor h = t.is_gpu() ? t.to_host() : nullptr;
  const TensorImpl& s = h ? *h : t;
  std::vector<float> v(s.size());
  for (size_t i = 0; i < v.size(); ++i) v[i] = (float)host_get(s.data(), s.data_type(), i);
  return v;
}
double max_abs_diff(const TensorImpl& a, const TensorImpl& b) {
  const auto x = to_float_vector(a), y = to_float_vector(b);
  if (x.size() != y.size()) return INFINITY;
  double m = 0;
  for (size_t i = 0; i < x.size(); ++i) {
    const double d = std::fabs((double)x[i] - (double)y[i]);
    if (!(d <= m)) m = d;         // NaN propagates
  }
  return m;
}
std::string TensorImpl::head(size_t n) const {
  const auto v = to_float_vector(*span(0, {std::min(n, size())}));
  std::ostringstream os;
  os << dtype_to_string(dtype_) << shape_to_string(shape_) << "@" << device().name() << " {";
  for (size_t i = 0; i < v.size(); ++i) os << (i ? ", " : "") << v[i];
  os << (size() > n ? ", ...}" : "}");
  return os.str();
}
void TensorImpl::print_data(size_t n) const { std::cout << head(n) << std::endl; }

static Tensor binary_new(const Tensor& a, const Tensor& b, int op) {
  Tensor o = a->clone();
  binary_inplace(*o, b.get(), 0, op);
  return o;
}
Tensor operator+(const Tensor& a, const Tensor& b) { return binary_new(a, b, TOP_ADD); }
Tensor operator-(const Tensor& a, const Tensor& b) { return binary_new(a, b, TOP_SUB); }
Tensor operator*(const Tensor& a, const Tensor& b) { return binary_new(a, b, TOP_MUL); }
Tensor operator/(const Tensor& a, const Tensor& b) { return binary_new(a, b, TOP_DIV); }

}  // namespace tnn
