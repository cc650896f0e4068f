// trainer: single-process training CLI (reference: examples/trainer.cpp:16-80).
//   trainer [--config cfg.json] [--synthetic] [--samples N]
// Configuration comes from a JSON file or from the environment / ./.env (same keys as the reference's .env.example):
// MODEL_NAME, DATASET_NAME, DATASET_PATH, DEVICE_TYPE, BATCH_SIZE, EPOCHS, MAX_STEPS, LR_INITIAL, ...
#include <cstring>
#include <iostream>

#include "tnn/example_models.hpp"
#include "tnn/train.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  try {
    std::string config_path;
    bool synthetic = false;
    size_t samples = 2048;
    for (int i = 1; i < argc; ++i) {
      if (!strcmp(argv[i], "--config") && i + 1 < argc) config_path = argv[++i];
      else if (!strcmp(argv[i], "--synthetic")) synthetic = true;
      else if (!strcmp(argv[i], "--samples") && i + 1 < argc) samples = (size_t)atol(argv[++i]);
      else if (!strcmp(argv[i], "--help")) {
        std::cout << "usage: trainer [--config cfg.json] [--synthetic] [--samples N]\n";
        return 0;
      }
    }
    ExampleModels::register_defaults();
    TrainingConfig cfg = config_path.empty() ? TrainingConfig::load_from_env() : TrainingConfig::load_from_json(config_path);
    std::cout << "[trainer] config: " << cfg.to_json().dump() << std::endl;
    const Device dev = cfg.device();
    std::cout << "[trainer] device " << dev.name() << " (" << DeviceManager::getInstance().num_gpus() << " GPU(s) visible)" << std::endl;
    DataLoaderPair data;
    if (synthetic || cfg.dataset_name.empty()) {
      const Shape s = ExampleModels::input_shape(cfg.model_name);
      const size_t classes = ExampleModels::num_classes(cfg.model_name);
      data.train = std::make_unique<SyntheticDataLoader>(s, classes, samples, 1);
      data.val = std::make_unique<SyntheticDataLoader>(s, classes, std::max<size_t>(samples / 8, cfg.batch_size), 2);
    } else {
      data = DataLoaderFactory::create(cfg.dataset_name, cfg.dataset_path);
    }
    Graph graph = load_or_create_model(cfg.model_name, cfg.model_path, dev, cfg.seed ? cfg.seed : 1234);
    std::cout << "[trainer] model " << graph.name() << ": " << graph.context().num_parameters() << " parameters" << std::endl;
    auto criterion = LossFactory::create_logsoftmax_crossentropy();
    // examples/trainer.cpp:67-71: Adam(lr, 0.9, 0.999, eps 1e-3, wd 3e-4), StepLR over 5 epochs of batches, gamma 0.1
    auto optimizer = OptimizerFactory::create_adam(cfg.lr_initial, 0.9, 0.999, 1e-3, 3e-4, false);
    const size_t steps_per_epoch = std::max<size_t>(1, data.train->size() / cfg.batch_size);
    auto scheduler = SchedulerFactory::create_step_lr(optimizer.get(), 5 * steps_per_epoch, 0.1);
    TrainHistory h = train_model(graph, *data.train, data.val.get(), *optimizer, *criterion, scheduler.get(), cfg);
    std::cout << "[trainer] done: final train loss " << (h.train_loss.empty() ? 0.0 : h.train_loss.back());
    if (!h.val_acc.empty()) std::cout << ", val acc " << 100.0 * h.val_acc.back() << "%";
    std::cout << std::endl;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[trainer] error: " << e.what() << std::endl;
    return 1;
  }
}
