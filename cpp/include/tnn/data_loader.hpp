// Data loaders (reference: include/data_loading/*.hpp, src/data_loading/data_loading.cpp:17-95): BaseDataLoader API
// (load_data / get_batch / reset / shuffle / size / get_data_shape / set_seed), a synthetic loader, CIFAR-10/100 binary
// records and MNIST CSV (NHWC, pixels / 255), DataLoaderFactory.  Batches are produced in PINNED host memory on a GPU
// build so that the trainer's H2D copies are asynchronous.
#pragma once
#include <cstdint>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "tnn/tensor.hpp"

namespace tnn {

class BaseDataLoader {
 public:
  virtual ~BaseDataLoader() = default;
  virtual bool load_data(const std::string& path) = 0;
  // fills x [B, ...] FP32 and y [B] INT32; returns false at the end of the epoch
  virtual bool get_batch(size_t batch_size, Tensor& x, Tensor& y);
  virtual void reset() { cursor_ = 0; }
  virtual void shuffle();
  virtual size_t size() const { return n_; }
  virtual Shape get_data_shape() const { return sample_shape_; }
  virtual size_t num_classes() const { return classes_; }
  void set_seed(uint64_t s) { rng_.seed(s); }
  void set_pinned(bool p) { pinned_ = p; }

 protected:
  virtual void fill_sample(size_t index, float* dst, int32_t* label) = 0;
  size_t n_ = 0, cursor_ = 0, classes_ = 0;
  Shape sample_shape_;
  std::vector<uint32_t> order_;
  std::mt19937_64 rng_{12345};
  bool pinned_ = false;
};

// Gaussian blobs whose mean depends on the class: learnable, deterministic, no files
class SyntheticDataLoader : public BaseDataLoader {
 public:
  SyntheticDataLoader(const Shape& sample_shape, size_t classes, size_t n, uint64_t seed = 1);
  bool load_data(const std::string& = "") override { return true; }

 protected:
  void fill_sample(size_t index, float* dst, int32_t* label) override;
  uint64_t seed_;
};

// CIFAR-10 (1 label byte + 3072 pixel bytes) / CIFAR-100 (coarse + fine label bytes + 3072) binary records, CHW -> HWC
class CIFARDataLoader : public BaseDataLoader {
 public:
  explicit CIFARDataLoader(bool cifar100) : c100_(cifar100) { sample_shape_ = {32, 32, 3}; classes_ = cifar100 ? 100 : 10; }
  bool load_data(const std::string& path) override;     // one .bin file (call repeatedly to add files)

 protected:
  void fill_sample(size_t index, float* dst, int32_t* label) override;
  bool c100_;
  std::vector<uint8_t> records_;
};

// MNIST as CSV (label, 784 pixels per line) -> [28, 28, 1]
class MNISTDataLoader : public BaseDataLoader {
 public:
  MNISTDataLoader() { sample_shape_ = {28, 28, 1}; classes_ = 10; }
  bool load_data(const std::string& path) override;

 protected:
  void fill_sample(size_t index, float* dst, int32_t* label) override;
  std::vector<uint8_t> pixels_;
  std::vector<uint8_t> labels_;
};

struct DataLoaderPair {
  std::unique_ptr<BaseDataLoader> train, val;
};
class DataLoaderFactory {
 public:
  // "mnist" | "cifar10" | "cifar100" | "synthetic[:H,W,C,classes,n]" ; missing files fall back to synthetic data of the named shape
  static DataLoaderPair create(const std::string& name, const std::string& path);
};

}  // namespace tnn
