// Data augmentation (reference: include/data_augmentation/*.hpp, augmentation.hpp AugmentationBuilder): nine strategies
// applied in order to NHWC float batches on the host, each with its own probability, deterministic per (seed, batch,
// sample).  BaseDataLoader::set_augmentation() applies one to every training batch.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "tnn/tensor.hpp"

namespace tnn {

class AugmentationStrategy {
 public:
  virtual ~AugmentationStrategy() = default;
  virtual std::string name() const = 0;
  // one image [H, W, C] in place
  virtual void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const = 0;

 protected:
  static bool hit(std::mt19937_64& rng, double p) { return p >= 1.0 || std::uniform_real_distribution<double>(0.0, 1.0)(rng) < p; }
};

class HorizontalFlip : public AugmentationStrategy {
 public:
  explicit HorizontalFlip(double p = 0.5) : p_(p) {}
  std::string name() const override { return "horizontal_flip"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_)) return;
    for (size_t h = 0; h < H; ++h)
      for (size_t w = 0; w < W / 2; ++w)
        for (size_t c = 0; c < C; ++c) std::swap(img[(h * W + w) * C + c], img[(h * W + (W - 1 - w)) * C + c]);
  }
 private:
  double p_;
};
class VerticalFlip : public AugmentationStrategy {
 public:
  explicit VerticalFlip(double p = 0.5) : p_(p) {}
  std::string name() const override { return "vertical_flip"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_)) return;
    for (size_t h = 0; h < H / 2; ++h)
      for (size_t i = 0; i < W * C; ++i) std::swap(img[h * W * C + i], img[(H - 1 - h) * W * C + i]);
  }
 private:
  double p_;
};
// rotation about the image centre by a uniform angle in [-max_degrees, max_degrees], bilinear, zero outside
class Rotation : public AugmentationStrategy {
 public:
  explicit Rotation(double p = 0.5, double max_degrees = 15.0) : p_(p), deg_(max_degrees) {}
  std::string name() const override { return "rotation"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_)) return;
    const double a = std::uniform_real_distribution<double>(-deg_, deg_)(rng) * 3.14159265358979323846 / 180.0;
    const double ca = std::cos(a), sa = std::sin(a), cy = (H - 1) / 2.0, cx = (W - 1) / 2.0;
    std::vector<float> src(img, img + H * W * C);
    for (size_t h = 0; h < H; ++h)
      for (size_t w = 0; w < W; ++w) {
        const double y = h - cy, x = w - cx;
        const double sx = ca * x + sa * y + cx, sy = -sa * x + ca * y + cy;
        const long x0 = (long)std::floor(sx), y0 = (long)std::floor(sy);
        const double fx = sx - x0, fy = sy - y0;
        for (size_t c = 0; c < C; ++c) {
          auto at = [&](long yy, long xx) -> double { return (yy < 0 || xx < 0 || yy >= (long)H || xx >= (long)W) ? 0.0 : src[((size_t)yy * W + (size_t)xx) * C + c]; };
          img[(h * W + w) * C + c] = (float)((1 - fy) * ((1 - fx) * at(y0, x0) + fx * at(y0, x0 + 1)) + fy * ((1 - fx) * at(y0 + 1, x0) + fx * at(y0 + 1, x0 + 1)));
        }
      }
  }
 private:
  double p_, deg_;
};
class Brightness : public AugmentationStrategy {
 public:
  explicit Brightness(double p = 0.5, double delta = 0.2) : p_(p), d_(delta) {}
  std::string name() const override { return "brightness"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_)) return;
    const float b = (float)std::uniform_real_distribution<double>(-d_, d_)(rng);
    for (size_t i = 0; i < H * W * C; ++i) img[i] = std::min(1.f, std::max(0.f, img[i] + b));
  }
 private:
  double p_, d_;
};
class Contrast : public AugmentationStrategy {
 public:
  explicit Contrast(double p = 0.5, double factor = 0.2) : p_(p), f_(factor) {}
  std::string name() const override { return "contrast"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_)) return;
    const float k = (float)(1.0 + std::uniform_real_distribution<double>(-f_, f_)(rng));
    for (size_t i = 0; i < H * W * C; ++i) img[i] = std::min(1.f, std::max(0.f, (img[i] - 0.5f) * k + 0.5f));
  }
 private:
  double p_, f_;
};
class GaussianNoise : public AugmentationStrategy {
 public:
  explicit GaussianNoise(double p = 0.5, double std_dev = 0.05) : p_(p), s_(std_dev) {}
  std::string name() const override { return "gaussian_noise"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_)) return;
    std::normal_distribution<float> n(0.f, (float)s_);
    for (size_t i = 0; i < H * W * C; ++i) img[i] = std::min(1.f, std::max(0.f, img[i] + n(rng)));
  }
 private:
  double p_, s_;
};
// zero-pad by `padding` on every side, then crop a random H x W window
class RandomCrop : public AugmentationStrategy {
 public:
  explicit RandomCrop(double p = 1.0, size_t padding = 4) : p_(p), pad_(padding) {}
  std::string name() const override { return "random_crop"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_) || pad_ == 0) return;
    const long dy = (long)(rng() % (2 * pad_ + 1)) - (long)pad_, dx = (long)(rng() % (2 * pad_ + 1)) - (long)pad_;
    std::vector<float> src(img, img + H * W * C);
    for (size_t h = 0; h < H; ++h)
      for (size_t w = 0; w < W; ++w) {
        const long sh = (long)h + dy, sw = (long)w + dx;
        const bool in = sh >= 0 && sw >= 0 && sh < (long)H && sw < (long)W;
        for (size_t c = 0; c < C; ++c) img[(h * W + w) * C + c] = in ? src[((size_t)sh * W + (size_t)sw) * C + c] : 0.f;
      }
  }
 private:
  double p_;
  size_t pad_;
};
class Cutout : public AugmentationStrategy {
 public:
  explicit Cutout(double p = 0.5, size_t size = 8) : p_(p), size_(size) {}
  std::string name() const override { return "cutout"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64& rng) const override {
    if (!hit(rng, p_) || size_ == 0) return;
    const long cy = (long)(rng() % H), cx = (long)(rng() % W), half = (long)size_ / 2;
    for (long h = std::max(0L, cy - half); h <= std::min((long)H - 1, cy + half); ++h)
      for (long w = std::max(0L, cx - half); w <= std::min((long)W - 1, cx + half); ++w)
        for (size_t c = 0; c < C; ++c) img[((size_t)h * W + (size_t)w) * C + c] = 0.f;
  }
 private:
  double p_;
  size_t size_;
};
class Normalize : public AugmentationStrategy {
 public:
  Normalize(std::vector<float> mean, std::vector<float> stdv) : mean_(std::move(mean)), std_(std::move(stdv)) {}
  std::string name() const override { return "normalize"; }
  void apply(float* img, size_t H, size_t W, size_t C, std::mt19937_64&) const override {
    for (size_t i = 0; i < H * W; ++i)
      for (size_t c = 0; c < C; ++c) img[i * C + c] = (img[i * C + c] - mean_[c % mean_.size()]) / std_[c % std_.size()];
  }
 private:
  std::vector<float> mean_, std_;
};

class Augmentation {
 public:
  explicit Augmentation(uint64_t seed = 0) : seed_(seed) {}
  void add(std::unique_ptr<AugmentationStrategy> s) { strategies_.push_back(std::move(s)); }
  // strategy management of the reference (augmentation.hpp:67-97): add_augmentation / remove_augmentation(index | name) / clear_augmentations
  void add_augmentation(std::unique_ptr<AugmentationStrategy> s) { add(std::move(s)); }
  void remove_augmentation(size_t index) {
    if (index >= strategies_.size()) throw std::out_of_range("Augmentation: no strategy at index " + std::to_string(index));
    strategies_.erase(strategies_.begin() + (std::ptrdiff_t)index);
  }
  // removes every strategy with that name; returns how many were removed
  size_t remove_augmentation(const std::string& name) {
    const size_t before = strategies_.size();
    strategies_.erase(std::remove_if(strategies_.begin(), strategies_.end(), [&](const std::unique_ptr<AugmentationStrategy>& s) { return s->name() == name; }), strategies_.end());
    return before - strategies_.size();
  }
  void clear_augmentations() { strategies_.clear(); }
  size_t size() const { return strategies_.size(); }
  std::vector<std::string> names() const {
    std::vector<std::string> n;
    for (auto& s : strategies_) n.push_back(s->name());
    return n;
  }
  // x: host FP32 [B, H, W, C], in place; `batch_index` decorrelates batches of one epoch
  void apply(TensorImpl& x, uint64_t batch_index = 0) const {
    if (x.data_type() != DType_t::FP32 || x.is_gpu() || x.shape().size() != 4) throw std::invalid_argument("Augmentation: host FP32 [B, H, W, C] expected");
    const size_t B = x.dimension(0), H = x.dimension(1), W = x.dimension(2), C = x.dimension(3);
    for (size_t b = 0; b < B; ++b) {
      std::mt19937_64 rng(seed_ * 0x9E3779B97F4A7C15ull + batch_index * 1000003ull + b);
      for (auto& s : strategies_) s->apply(x.data_as<float>() + b * H * W * C, H, W, C, rng);
    }
  }

 private:
  uint64_t seed_;
  std::vector<std::unique_ptr<AugmentationStrategy>> strategies_;
};

// AugmentationBuilder().horizontal_flip(0.5).random_crop(1.0, 4).normalize(mean, std).build()
class AugmentationBuilder {
 public:
  explicit AugmentationBuilder(uint64_t seed = 0) : aug_(std::make_shared<Augmentation>(seed)) {}
  AugmentationBuilder& horizontal_flip(double p = 0.5) { aug_->add(std::make_unique<HorizontalFlip>(p)); return *this; }
  AugmentationBuilder& vertical_flip(double p = 0.5) { aug_->add(std::make_unique<VerticalFlip>(p)); return *this; }
  AugmentationBuilder& rotation(double p = 0.5, double max_degrees = 15.0) { aug_->add(std::make_unique<Rotation>(p, max_degrees)); return *this; }
  AugmentationBuilder& brightness(double p = 0.5, double delta = 0.2) { aug_->add(std::make_unique<Brightness>(p, delta)); return *this; }
  AugmentationBuilder& contrast(double p = 0.5, double factor = 0.2) { aug_->add(std::make_unique<Contrast>(p, factor)); return *this; }
  AugmentationBuilder& gaussian_noise(double p = 0.5, double std_dev = 0.05) { aug_->add(std::make_unique<GaussianNoise>(p, std_dev)); return *this; }
  AugmentationBuilder& random_crop(double p = 1.0, size_t padding = 4) { aug_->add(std::make_unique<RandomCrop>(p, padding)); return *this; }
  AugmentationBuilder& cutout(double p = 0.5, size_t size = 8) { aug_->add(std::make_unique<Cutout>(p, size)); return *this; }
  AugmentationBuilder& normalize(std::vector<float> mean, std::vector<float> stdv) { aug_->add(std::make_unique<Normalize>(std::move(mean), std::move(stdv))); return *this; }
  // any user strategy (docs/USAGE.md "custom augmentation"): applied in sequence with the built-in ones
  AugmentationBuilder& custom_augmentation(std::unique_ptr<AugmentationStrategy> s) { aug_->add(std::move(s)); return *this; }
  std::shared_ptr<Augmentation> build() { return aug_; }

 private:
  std::shared_ptr<Augmentation> aug_;
};

}  // namespace tnn
