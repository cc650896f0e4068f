// Tensor: N-d row-major, dtype-erased, contiguous (views are contiguous spans), living on a Device
// (reference: include/tensor/tensor.hpp:41-610, tensor_factory.hpp:12-159).  GPU arithmetic runs in the
// hand-written kernels of csrc/kernels/tensor_ops.cu on the device's current Flow; CPU arithmetic in plain loops
// (the test oracle).  File record = {u32 dtype, u64 ndims, u64 dims[], raw bytes} (tensor.hpp:585-606).
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <iosfwd>
#include <memory>
#include <string>
#include <vector>

#include "tnn/allocator.hpp"
#include "tnn/device.hpp"
#include "tnn/type.hpp"

namespace tnn {

class TensorImpl;
using Tensor = std::shared_ptr<TensorImpl>;
using ConstTensor = std::shared_ptr<const TensorImpl>;
using Shape = std::vector<size_t>;

// The Flow kernels of the calling thread are launched on (per device).  FlowGuard switches it for a scope
// (e.g. weight gradients on the side flow).
Flow& current_flow(const Device& d);
class FlowGuard {
 public:
  FlowGuard(const Device& d, flowHandle_t h);
  ~FlowGuard();

 private:
  Device dev_;
  flowHandle_t prev_;
};

inline size_t shape_numel(const Shape& s) {
  size_t n = 1;
  for (size_t d : s) n *= d;
  return n;
}
std::string shape_to_string(const Shape& s);

class TensorImpl : public std::enable_shared_from_this<TensorImpl> {
 public:
  TensorImpl(DType_t dt, Shape shape, StoragePtr storage, size_t offset_bytes, IAllocator* alloc)
      : dtype_(dt), shape_(std::move(shape)), storage_(std::move(storage)), offset_(offset_bytes), alloc_(alloc) {}

  // ---- geometry
  DType_t data_type() const { return dtype_; }
  const Shape& shape() const { return shape_; }
  size_t dims() const { return shape_.size(); }
  size_t dimension(size_t i) const { return shape_.at(i); }
  size_t stride(size_t i) const {
    size_t s = 1;
    for (size_t k = i + 1; k < shape_.size(); ++k) s *= shape_[k];
    return s;
  }
  size_t size() const { return shape_numel(shape_); }
  size_t nbytes() const { return size() * get_dtype_size(dtype_); }
  size_t capacity() const { return storage_ ? (storage_->bytes - offset_) / get_dtype_size(dtype_) : 0; }
  const Device& device() const { return storage_->dev; }
  DeviceType device_type() const { return storage_->dev.device_type(); }
  bool is_gpu() const { return storage_->dev.is_gpu(); }
  IAllocator* allocator() const { return alloc_; }
  const StoragePtr& storage() const { return storage_; }
  size_t offset_bytes() const { return offset_; }

  // ---- raw access
  void* data() { return static_cast<char*>(storage_->ptr) + offset_; }
  const void* data() const { return static_cast<const char*>(storage_->ptr) + offset_; }
  void* data_ptr() { return data(); }
  template <class T> T* data_as() { return static_cast<T*>(data()); }
  template <class T> const T* data_as() const { return static_cast<const T*>(data()); }
  // element access (host tensors only)
  template <class T>
  T& at(std::initializer_list<size_t> idx) {
    size_t off = 0, k = 0;
    for (size_t i : idx) off += i * stride(k++);
    return data_as<T>()[off];
  }

  // ---- views (share storage)
  Tensor span(size_t offset_elems, const Shape& shape) const;
  Tensor reshape(const Shape& shape) const { return span(0, shape); }
  Tensor share() const { return span(0, shape_); }
  void share_from(const TensorImpl& other);                 // become a view of `other`
  // first-dimension slice [begin, begin + count)
  Tensor slice_batch(size_t begin, size_t count) const;

  // ---- movement
  Tensor to_device(const Device& d, IAllocator* alloc = nullptr) const;
  Tensor to_host() const;
  Tensor clone(IAllocator* alloc = nullptr) const;
  void copy_from(const TensorImpl& src);                    // same dtype + numel; any device pair
  void copy_to(TensorImpl& dst) const { dst.copy_from(*this); }
  void copy_from_host(const void* src, size_t bytes);
  void copy_to_host(void* dst, size_t bytes) const;
  void resize(const Shape& shape);                          // reallocates when capacity is too small
  void ensure(const Shape& shape) { if (shape != shape_) resize(shape); }

  // ---- fill
  void fill(double v);
  void zero() { fill(0.0); }
  void fill_random_uniform(double lo, double hi, uint64_t seed);
  void fill_random_normal(double mean, double stddev, uint64_t seed);

  // ---- arithmetic (in place, same shape & dtype; FP32 / BF16 on the GPU, FP32 / FP64 on the CPU)
  void add(const TensorImpl& o);
  void sub(const TensorImpl& o);
  void mul(const TensorImpl& o);
  void div(const TensorImpl& o);
  void add_scalar(double s);
  void sub_scalar(double s) { add_scalar(-s); }
  void mul_scalar(double s);
  void div_scalar(double s) { mul_scalar(1.0 / s); }
  void axpy(double a, const TensorImpl& x);                 // this += a * x

  // ---- reductions (synchronise and return on the host)
  double sum() const;
  double mean() const { return sum() / (double)size(); }
  double variance() const;
  double min() const;
  double max() const;

  // ---- I/O
  void save(std::ostream& os) const;
  std::string head(size_t n = 8) const;
  void print_data(size_t n = 32) const;

 private:
  DType_t dtype_;
  Shape shape_;
  StoragePtr storage_;
  size_t offset_;
  IAllocator* alloc_;
};

// ---- factory (tensor_factory.hpp)
Tensor make_tensor(IAllocator& alloc, DType_t dt, const Shape& shape);
Tensor make_tensor(DType_t dt, const Shape& shape, const Device& d);
inline Tensor make_tensor(DType_t dt, const Shape& shape) { return make_tensor(dt, shape, getHost()); }
template <class T>
Tensor make_tensor(const Shape& shape, const Device& d = getHost()) { return make_tensor(dtype_of<T>(), shape, d); }
Tensor make_tensor_from(const std::vector<float>& values, const Shape& shape, const Device& d = getHost());
// wraps memory owned by someone else (kept alive by `owner`)
Tensor make_tensor_view(void* ptr, DType_t dt, const Shape& shape, const Device& d, std::shared_ptr<void> owner = nullptr);
Tensor create_like(const TensorImpl& t, IAllocator* alloc = nullptr);
Tensor dtype_cast(const TensorImpl& t, DType_t to, IAllocator* alloc = nullptr);
Tensor load(std::istream& is, IAllocator& alloc);
void load_into(std::istream& is, TensorImpl& dst);          // checks dtype + shape (tensor_factory.hpp:95-128)

Tensor operator+(const Tensor& a, const Tensor& b);
Tensor operator-(const Tensor& a, const Tensor& b);
Tensor operator*(const Tensor& a, const Tensor& b);
Tensor operator/(const Tensor& a, const Tensor& b);

// C[M, N] = alpha * A B^T-ish for 2-D tensors of one floating dtype: a is [M, K] (or [K, M] with a_mn), b is [N, K] (or [K, N]
// with b_mn).  GPU: BF16 -> tcgen05 kind::f16 (umma_gemm.cu), FP16 -> kind::f16 with fp16 operands, FP32 -> kind::tf32,
// FP64 -> DFMA kernel (gemm_xt.cu); the reference picks one cuBLAS routine per dtype (src/math/cuda/cublas_gemm.cu:20-164).
// CPU: FP32 through the packed SGEMM, the other dtypes through a double-precision loop.
Tensor matmul(const TensorImpl& a, const TensorImpl& b, bool a_mn = false, bool b_mn = false, double alpha = 1.0);

// largest |a - b| (host compare; bf16 / fp16 / fp32 / fp64)
double max_abs_diff(const TensorImpl& a, const TensorImpl& b);
std::vector<float> to_float_vector(const TensorImpl& t);

}  // namespace tnn
