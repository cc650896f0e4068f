// bench_cpp: the flagship training step (WRN-16-8 / CIFAR-100, bf16, batch 256, Adam) driven ONLY by the C++ host --
// no Python, no libtorch.  Prints one JSON line in bench.py's schema (value = device-timed steps with device-resident
// inputs; e2e = every step copies its batch from pinned host memory and the loss of a step is read back).
//   bench_cpp [--steps K] [--warmup W] [--batch B] [--model NAME] [--no-graph] [--gpu I]
#include <chrono>
#include <cstring>
#include <iostream>

#include "kernels/simt_ops.cuh"
#include "tnn/example_models.hpp"
#include "tnn/stepstate.hpp"
#include "tnn/bench.hpp"
#include "tnn/train.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  int steps = 30, warmup = 5, gpu = 0;
  size_t B = 256;
  std::string model = "cifar100_wrn16_8";
  bool use_graph = true;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--batch") && i + 1 < argc) B = (size_t)atol(argv[++i]);
    else if (!strcmp(argv[i], "--model") && i + 1 < argc) model = argv[++i];
    else if (!strcmp(argv[i], "--gpu") && i + 1 < argc) gpu = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--no-graph")) use_graph = false;
  }
  try {
    std::cout << bench_train(model, B, steps, warmup, gpu, use_graph).dump() << std::endl;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "[bench_cpp] error: " << e.what() << std::endl;
    return 1;
  }
}
