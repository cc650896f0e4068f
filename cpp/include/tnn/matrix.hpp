// Matrix<T>: the legacy 2-D value type (reference: include/matrix/matrix.hpp:25-230 -- rows x cols, copy = deep copy, no
// copy-assignment, +, -, scalar *, /, matrix product, clone / reshape / resize, fills).  Here it is a thin value wrapper over a
// 2-D tnn::Tensor: storage, device placement and the arithmetic kernels are the Tensor's (CPU loops or the templated GPU tensor
// ops), the product is tnn::matmul (packed SGEMM on the CPU, the tcgen05 / tf32 / DFMA GEMM of the element type on the GPU).
#pragma once
#include <stdexcept>

#include "tnn/tensor.hpp"

namespace tnn {

template <typename T = float>
class Matrix {
 public:
  explicit Matrix(const Device& device = getHost()) : dev_(device) {}      // empty (0 x 0) matrix: no storage
  Matrix(size_t rows, size_t cols, const Device& device = getHost()) : dev_(device) {
    if (rows * cols) {
      t_ = make_tensor(dtype_of<T>(), {rows, cols}, device);
      t_->fill(0.0);
    }
  }
  // adopts a COPY of `data` (rows * cols elements of T on any device)
  Matrix(size_t rows, size_t cols, const TensorImpl& data, const Device& device = getHost()) : Matrix(rows, cols, device) {
    if (data.size() != rows * cols || data.data_type() != dtype_of<T>()) throw std::invalid_argument("Matrix: data does not hold rows*cols elements of T");
    t_->copy_from(data);
  }
  Matrix(const Matrix& o) : t_(o.t_ ? o.t_->clone() : nullptr), dev_(o.dev_) {}     // value semantics: a copy is a deep copy
  Matrix(Matrix&& o) noexcept = default;
  Matrix& operator=(const Matrix&) = delete;                              // as in the reference: assign by move or clone()
  Matrix& operator=(Matrix&& o) noexcept = default;

  size_t rows() const { return t_ ? t_->shape()[0] : 0; }
  size_t cols() const { return t_ ? t_->shape()[1] : 0; }
  size_t size() const { return t_ ? t_->size() : 0; }
  const Device& device() const { return dev_; }
  T* data() { return t_->template data_as<T>(); }
  const T* data() const { return t_->template data_as<T>(); }
  const Tensor& tensor() const { return t_; }                             // the underlying 2-D tensor (shared, not copied)
  T& operator()(size_t r, size_t c) { return t_->template at<T>({r, c}); }         // host matrices only
  T operator()(size_t r, size_t c) const { return t_->template at<T>({r, c}); }

  Matrix clone() const { return Matrix(*this); }
  Matrix reshape(size_t r, size_t c) const {
    if (r * c != size()) throw std::invalid_argument("Total number of elements must remain the same for reshape.");
    Matrix m(r, c, device());
    if (t_) m.t_->copy_from(*t_);
    return m;
  }
  // reallocates unless the shape is unchanged; contents are NOT preserved
  void resize(size_t r, size_t c) {
    if (r == rows() && c == cols()) return;
    t_ = r * c ? make_tensor(dtype_of<T>(), {r, c}, device()) : nullptr;
    if (t_) t_->fill(0.0);
  }
  Matrix to(const Device& d) const {
    Matrix m(d);
    if (t_) m.t_ = t_->device() == d ? t_->clone() : t_->to_device(d);
    return m;
  }

  void fill(T v) { if (t_) t_->fill((double)v); }
  void fill_random_uniform(T range, uint64_t seed = 0x9E3779B97F4A7C15ull) { if (t_) t_->fill_random_uniform(-(double)range, (double)range, seed); }
  void fill_random_normal(T mean, T stddev, uint64_t seed = 0x9E3779B97F4A7C15ull) { if (t_) t_->fill_random_normal((double)mean, (double)stddev, seed); }

  Matrix operator+(const Matrix& o) const { Matrix m(*this); m += o; return m; }
  Matrix operator-(const Matrix& o) const { Matrix m(*this); m -= o; return m; }
  Matrix& operator+=(const Matrix& o) { same(o, "addition"); if (t_) t_->add(*o.t_); return *this; }
  Matrix& operator-=(const Matrix& o) { same(o, "subtraction"); if (t_) t_->sub(*o.t_); return *this; }
  Matrix operator*(T s) const { Matrix m(*this); m *= s; return m; }
  Matrix& operator*=(T s) { if (t_) t_->mul_scalar((double)s); return *this; }
  Matrix operator/(T s) const { Matrix m(*this); m /= s; return m; }
  Matrix& operator/=(T s) {
    if (s == T(0)) throw std::invalid_argument("Division by zero.");
    if (t_) t_->mul_scalar(1.0 / (double)s);
    return *this;
  }
  // matrix product [rows, k] x [k, cols]
  Matrix operator*(const Matrix& o) const {
    if (cols() != o.rows()) throw std::invalid_argument("Matrix dimensions must match for multiplication.");
    Matrix m(device());
    if (t_ && o.t_) m.t_ = matmul(*t_, *o.t_, false, true);                // B is [K, N]: MN-major for the GEMM
    return m;
  }

 private:
  void same(const Matrix& o, const char* what) const {
    if (rows() != o.rows() || cols() != o.cols()) throw std::invalid_argument(std::string("Matrix dimensions must match for ") + what + ".");
  }
  Tensor t_;
  Device dev_ = getHost();
};
template <typename T>
Matrix<T> operator*(T s, const Matrix<T>& m) { return m * s; }

}  // namespace tnn
