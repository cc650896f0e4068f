// dp_trainer: data-parallel training on the GPUs of one node inside ONE process (a host thread + replica per GPU, the fused
// reduce + optimizer + broadcast kernel per step).  The reference has no data-parallel trainer; this is the
// C++ counterpart of examples/dp_train.py / bench.py --gpus N.
//   dp_trainer [--model cifar100_wrn16_8] [--gpus N] [--batch 256 (per GPU)] [--steps 40] [--warmup 8] [--no-graph]
//              [--resume FILE] [--save FILE]     checkpoint with the sharded optimizer state gathered (any world size resumes it)
#include <chrono>
#include <cstring>
#include <iostream>

#include "tnn/data_parallel.hpp"
#include "tnn/example_models.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  std::string model = "cifar100_wrn16_8";
  int gpus = 0, steps = 40, warmup = 8;
  size_t batch = 256;
  bool use_graph = true;
  std::string resume, save;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--model") && i + 1 < argc) model = argv[++i];
    else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch = (size_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--no-graph")) use_graph = false;
    else if (!strcmp(argv[i], "--resume") && i + 1 < argc) resume = argv[++i];
    else if (!strcmp(argv[i], "--save") && i + 1 < argc) save = argv[++i];
  }
  try {
    ExampleModels::register_defaults();
    auto& dm = DeviceManager::getInstance();
    if (gpus <= 0) gpus = dm.num_gpus();
    if (gpus < 1) throw std::runtime_error("no GPU visible");
    Adam proto(1e-3, 0.9, 0.999, 1e-3, 3e-4, false);
    DataParallel dp([&]() -> std::unique_ptr<Layer> { return ExampleModels::create(model); }, gpus, 1234, proto, "logsoftmax_crossentropy", use_graph);
    std::cout << "[dp_trainer] " << model << ": " << dp.num_parameters() << " parameters x " << gpus << " replicas" << std::endl;
    if (!resume.empty()) {
      dp.load_checkpoint(resume);
      std::cout << "[dp_trainer] resumed " << resume << " at optimizer step " << dp.step_count() << std::endl;
    }
    const Shape in = ExampleModels::input_shape(model);
    const size_t classes = ExampleModels::num_classes(model);
    Vec<std::unique_ptr<SyntheticDataLoader>> loaders;
    Vec<Vec<Tensor>> xs(4), ys(4);                         // four rotating batches per replica, pinned
    for (int r = 0; r < gpus; ++r) {
      loaders.push_back(std::make_unique<SyntheticDataLoader>(in, classes, batch * 4, 100 + (uint64_t)r));
      loaders.back()->set_pinned(true);
    }
    for (int k = 0; k < 4; ++k)
      for (int r = 0; r < gpus; ++r) {
        Tensor x, y;
        loaders[(size_t)r]->get_batch(batch, x, y);
        xs[(size_t)k].push_back(x);
        ys[(size_t)k].push_back(y);
      }
    double loss = 0;
    for (int i = 0; i < warmup; ++i) loss = dp.step(xs[(size_t)i % 4], ys[(size_t)i % 4]).first;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) loss = dp.step(xs[(size_t)i % 4], ys[(size_t)i % 4]).first;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;
    std::cout << "{\"bench\": \"dp_trainer\", \"model\": \"" << model << "\", \"n_gpus\": " << gpus << ", \"per_gpu_batch\": " << batch << ", \"ms_per_step\": " << ms
              << ", \"img_s\": " << (double)(batch * (size_t)gpus) / ms * 1e3 << ", \"final_loss\": " << loss << ", \"cuda_graph\": " << (dp.captured() ? "true" : "false")
              << ", \"replica_divergence\": " << dp.max_replica_divergence() << ", \"timing\": \"wall clock incl. H2D of every batch and the loss read-back\"}" << std::endl;
    if (!save.empty()) {
      dp.save_checkpoint(save);
      std::cout << "[dp_trainer] saved " << save << " at optimizer step " << dp.step_count() << std::endl;
    }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[dp_trainer] error: " << e.what() << std::endl;
    return 1;
  }
}
