// inferencer: load (or create) a model and report validation loss / accuracy (reference: examples/inferencer.cpp:13-49).
// Configuration from the environment / ./.env: MODEL_NAME, MODEL_PATH, DATASET_NAME, DATASET_PATH, DEVICE_TYPE, BATCH_SIZE.
//   inferencer [--synthetic] [--samples N]
#include <cstring>
#include <iostream>

#include "tnn/example_models.hpp"
#include "tnn/train.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  try {
    bool synthetic = false;
    size_t samples = 1024;
    for (int i = 1; i < argc; ++i) {
      if (!strcmp(argv[i], "--synthetic")) synthetic = true;
      else if (!strcmp(argv[i], "--samples") && i + 1 < argc) samples = (size_t)atol(argv[++i]);
    }
    ExampleModels::register_defaults();
    TrainingConfig cfg = TrainingConfig::load_from_env();
    std::cout << "[inferencer] config: " << cfg.to_json().dump() << std::endl;
    const Device dev = cfg.device();
    std::unique_ptr<BaseDataLoader> val;
    if (synthetic || cfg.dataset_name.empty()) {
      if (!synthetic) std::cerr << "[inferencer] DATASET_NAME is not set: using the synthetic loader" << std::endl;
      val = std::make_unique<SyntheticDataLoader>(ExampleModels::input_shape(cfg.model_name), ExampleModels::num_classes(cfg.model_name), samples, 2);
    } else {
      DataLoaderPair data = DataLoaderFactory::create(cfg.dataset_name, cfg.dataset_path);
      val = std::move(data.val);
    }
    Graph graph = load_or_create_model(cfg.model_name, cfg.model_path, dev, cfg.seed ? cfg.seed : 1234);
    auto criterion = LossFactory::create_logsoftmax_crossentropy();
    auto [loss, acc] = validate_model(graph, *val, *criterion, cfg);
    std::cout << "Validation Loss: " << loss << ", Accuracy: " << acc * 100.0 << "%" << std::endl;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "Inference failed: " << e.what() << std::endl;
    return 1;
  }
}
