// legacy_trainer: single-process training of the legacy (NCHW) zoo (reference: examples/legacy_trainer.cpp:13-60).
// Environment / ./.env: MODEL_NAME (default cifar10_resnet9), MODEL_PATH, DATASET_NAME, DATASET_PATH, DEVICE_TYPE, LR_INITIAL...
//   legacy_trainer [--synthetic] [--samples N]
#include <cstring>
#include <iostream>

#include "tnn/legacy.hpp"
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
    legacy::ExampleModels::register_defaults();
    TrainingConfig cfg = TrainingConfig::load_from_env();
    if (!getenv("MODEL_NAME")) cfg.model_name = "cifar10_resnet9";
    std::cout << "[legacy_trainer] config: " << cfg.to_json().dump() << std::endl;
    const Device dev = cfg.device();
    DataLoaderPair data;
    if (synthetic || cfg.dataset_name.empty()) {
      const Shape s = tnn::ExampleModels::input_shape(cfg.model_name);
      const size_t classes = tnn::ExampleModels::num_classes(cfg.model_name);
      data.train = std::make_unique<legacy::NCHWLoader>(std::make_unique<SyntheticDataLoader>(s, classes, samples, 1));
      data.val = std::make_unique<legacy::NCHWLoader>(std::make_unique<SyntheticDataLoader>(s, classes, std::max<size_t>(samples / 8, cfg.batch_size), 2));
    } else {
      data = legacy::DataLoaderFactory::create(cfg.dataset_name, cfg.dataset_path);
    }
    Graph graph = legacy::load_or_create_model(cfg.model_name, cfg.model_path, dev, cfg.seed ? cfg.seed : 1234);
    std::cout << "Training model on device: " << dev.name() << " (" << graph.context().num_parameters() << " parameters)" << std::endl;
    auto criterion = LossFactory::create_logsoftmax_crossentropy();
    // examples/legacy_trainer.cpp:46-49: Adam(lr, 0.9, 0.999, eps 1e-5, wd 1e-4), StepLR over 5 epochs of batches, gamma 0.1
    auto optimizer = OptimizerFactory::create_adam(cfg.lr_initial, 0.9, 0.999, 1e-5, 1e-4, false);
    auto scheduler = SchedulerFactory::create_step_lr(optimizer.get(), 5 * std::max<size_t>(1, data.train->size() / cfg.batch_size), 0.1);
    TrainHistory h = train_model(graph, *data.train, data.val.get(), *optimizer, *criterion, scheduler.get(), cfg);
    std::cout << "[legacy_trainer] done: final train loss " << (h.train_loss.empty() ? 0.0 : h.train_loss.back()) << std::endl;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "Training failed: " << e.what() << std::endl;
    return 1;
  }
}
