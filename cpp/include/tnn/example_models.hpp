// Model zoo (reference: src/nn/example_models.cpp:21-559, include/nn/example_models.hpp:19-84): 17 named models
// built with the LayerBuilder DSL + load_or_create_model.
#pragma once
#include <functional>
#include <map>
#include <string>

#include "tnn/graph.hpp"
#include "tnn/layer_builder.hpp"

namespace tnn {

class ExampleModels {
 public:
  using Creator = std::function<std::unique_ptr<Sequential>(DType_t io)>;
  static void register_defaults();
  static void register_model(const std::string& name, Creator c);
  static std::unique_ptr<Sequential> create(const std::string& name, DType_t io = DType_t::FP32);
  static Vec<std::string> available_models();
  // sample input shape (batchless) and class count of a zoo model: used by the synthetic loaders of the CLIs
  static Shape input_shape(const std::string& name);
  static size_t num_classes(const std::string& name);

 private:
  static std::map<std::string, Creator>& table();
};

// Loads `path` when it exists, else builds the named zoo model as input -> model -> output.
Graph load_or_create_model(const std::string& name, const std::string& path, const Device& dev, uint64_t seed = 0);

}  // namespace tnn
