// Legacy (NCHW) model zoo and loaders (reference: include/nn/legacy/example_models.hpp:14-81,
// src/nn/legacy/example_models.cpp:21-510, include/data_loading/legacy/*.hpp).  Every legacy model is DERIVED from the
// current NHWC definition: the zoo model's config tree is walked and each spatial layer is replaced by its legacy_*
// twin, so the two zoos cannot drift apart.  legacy::DataLoaderFactory hands out NCHW batches.
#pragma once
#include "tnn/data_loader.hpp"
#include "tnn/example_models.hpp"

namespace tnn {
namespace legacy {

// NHWC Sequential config -> the same network on NCHW tensors built from legacy_* layers
Json to_legacy_config(const Json& nhwc_config);
std::unique_ptr<Layer> to_legacy(Layer& nhwc_model);

class ExampleModels {
 public:
  static void register_defaults() { tnn::ExampleModels::register_defaults(); }
  static std::unique_ptr<Layer> create(const std::string& name, DType_t io = DType_t::FP32);
  static Vec<std::string> available_models() { return tnn::ExampleModels::available_models(); }
};
Graph load_or_create_model(const std::string& name, const std::string& path, const Device& dev, uint64_t seed = 0);

// Presents an NHWC loader's image batches as NCHW
class NCHWLoader : public BaseDataLoader {
 public:
  explicit NCHWLoader(std::unique_ptr<BaseDataLoader> inner) : inner_(std::move(inner)) {}
  bool load_data(const std::string& p) override { return inner_->load_data(p); }
  bool get_batch(size_t batch_size, Tensor& x, Tensor& y) override;
  void reset() override { inner_->reset(); }
  void shuffle() override { inner_->shuffle(); }
  size_t size() const override { return inner_->size(); }
  Shape get_data_shape() const override;
  size_t num_classes() const override { return inner_->num_classes(); }

 protected:
  void fill_sample(size_t, float*, int32_t*) override {}
  std::unique_ptr<BaseDataLoader> inner_;
};
struct DataLoaderFactory {
  static DataLoaderPair create(const std::string& name, const std::string& path);
};

}  // namespace legacy
}  // namespace tnn
