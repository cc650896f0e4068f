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

#include "tnn/augmentation.hpp"
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
  // applied to every [B, H, W, C] batch this loader hands out (training loaders)
  void set_augmentation(std::shared_ptr<Augmentation> a) { augmentation_ = std::move(a); }
  const std::shared_ptr<Augmentation>& augmentation() const { return augmentation_; }

 protected:
  virtual void fill_sample(size_t index, float* dst, int32_t* label) = 0;
  size_t n_ = 0, cursor_ = 0, classes_ = 0;
  Shape sample_shape_;
  std::vector<uint32_t> order_;
  std::mt19937_64 rng_{12345};
  bool pinned_ = false;
  std::shared_ptr<Augmentation> augmentation_;
  uint64_t batches_served_ = 0;
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

// uint16 token stream (OpenWebText train.bin / val.bin, reference: open_webtext_data_loader.hpp): x [B, L] INT32 windows,
// y [B, L] INT32 = the next tokens.  size() = number of non-overlapping windows; windows are drawn at random offsets.
class OpenWebTextDataLoader : public BaseDataLoader {
 public:
  explicit OpenWebTextDataLoader(size_t seq_len = 1024) : seq_len_(seq_len) { sample_shape_ = {seq_len}; classes_ = 50257; }
  ~OpenWebTextDataLoader() override;
  bool load_data(const std::string& path) override;      // a .bin file
  bool get_batch(size_t batch_size, Tensor& x, Tensor& y) override;
  size_t num_tokens() const;

 protected:
  void fill_sample(size_t, float*, int32_t*) override {}
  size_t seq_len_;
  struct Impl;
  Impl* impl_ = nullptr;
};

// Numeric CSV with a header line: the last `n_targets` columns are regression targets (reference:
// regression_data_loader.hpp); x [B, F] FP32, y [B, n_targets] FP32
class RegressionCSVDataLoader : public BaseDataLoader {
 public:
  explicit RegressionCSVDataLoader(size_t n_targets = 1, bool has_header = true) : n_targets_(n_targets), header_(has_header) {}
  bool load_data(const std::string& path) override;
  bool get_batch(size_t batch_size, Tensor& x, Tensor& y) override;
  size_t num_features() const { return features_; }

 protected:
  void fill_sample(size_t, float*, int32_t*) override {}
  size_t n_targets_, features_ = 0;
  bool header_;
  std::vector<float> xs_, ys_;
};

// UJIIndoorLoc-style WiFi CSV (reference: wifi_data_loader.hpp): a block of feature columns and a block of target
// columns; regression (float targets) or classification (first target column as the class); normalize_data() z-scores
// features (and regression targets) and keeps the statistics for de-normalisation
class WiFiDataLoader : public BaseDataLoader {
 public:
  explicit WiFiDataLoader(bool is_regression = true) : regression_(is_regression) {}
  bool load_data(const std::string& path) override { return load_data(path, 0, 0, 0, 0, true); }
  bool load_data(const std::string& path, size_t feature_start, size_t feature_end, size_t target_start, size_t target_end, bool has_header = true);
  bool get_batch(size_t batch_size, Tensor& x, Tensor& y) override;
  void normalize_data();
  bool is_normalized() const { return normalized_; }
  const std::vector<float>& feature_means() const { return fmean_; }
  const std::vector<float>& feature_stds() const { return fstd_; }
  const std::vector<float>& target_means() const { return tmean_; }
  const std::vector<float>& target_stds() const { return tstd_; }
  size_t num_features() const { return features_; }
  size_t num_targets() const { return targets_; }

 protected:
  void fill_sample(size_t, float*, int32_t*) override {}
  bool regression_, normalized_ = false;
  size_t features_ = 0, targets_ = 0;
  std::vector<float> xs_, ys_, fmean_, fstd_, tmean_, tstd_;
};

// Directory trees of JPEG images or their decoded caches (reference: image_data_loader.hpp + stb_image, tiny_imagenet_data_loader.hpp,
// imagenet100_data_loader.hpp; Python host: data/loaders.py::ImageFolderDataLoader -- same layouts, same cache files).
//   * `<path>/<split>_x.npy` (uint8 [N, S, S, 3]) + `<split>_y.npy` (int32 [N]): memory mapped, no decoding;
//   * otherwise the tree: TinyImageNet (`train/<wnid>/images/*.JPEG`, `val/images` + `val/val_annotations.txt`, `wnids.txt`) or
//     class-per-directory (`train/<class>/*`, `val/<class>/*`: ImageNet-100), decoded per batch on the thread pool by the
//     baseline JPEG decoder of tnn/jpeg.hpp and resized (bilinear) to size x size.
class ImageFolderDataLoader : public BaseDataLoader {
 public:
  ImageFolderDataLoader(size_t size, size_t classes, bool train) : train_(train) { sample_shape_ = {size, size, 3}; classes_ = classes; }
  ~ImageFolderDataLoader() override;
  bool load_data(const std::string& path) override;           // the dataset root
  bool get_batch(size_t batch_size, Tensor& x, Tensor& y) override;
  const std::vector<std::string>& class_names() const { return names_; }
  bool uses_cache() const { return cache_x_ != nullptr; }

 protected:
  void fill_sample(size_t index, float* dst, int32_t* label) override;
  bool train_;
  std::vector<std::string> files_, names_;
  std::vector<int32_t> labels_;
  const uint8_t* cache_x_ = nullptr;       // mapped <split>_x.npy payload
  const int32_t* cache_y_ = nullptr;
  void* map_x_ = nullptr;
  void* map_y_ = nullptr;
  size_t map_x_bytes_ = 0, map_y_bytes_ = 0;
};
class TinyImageNetDataLoader : public ImageFolderDataLoader {
 public:
  explicit TinyImageNetDataLoader(bool train = true) : ImageFolderDataLoader(64, 200, train) {}
};
class ImageNet100DataLoader : public ImageFolderDataLoader {
 public:
  explicit ImageNet100DataLoader(bool train = true, size_t size = 224) : ImageFolderDataLoader(size, 100, train) {}
};

struct DataLoaderPair {
  std::unique_ptr<BaseDataLoader> train, val;
};
class DataLoaderFactory {
 public:
  // "mnist" | "cifar10" | "cifar100" | "openwebtext" | "regression" | "wifi" | "synthetic[:H,W,C,classes,n]"; missing image
  // files fall back to synthetic data of the named shape
  static DataLoaderPair create(const std::string& name, const std::string& path);
};

}  // namespace tnn
