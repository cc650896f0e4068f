// bench_train: the flagship training-step benchmark (device-timed and end-to-end numbers in bench.py's JSON schema).
#pragma once
#include <string>

#include "tnn/json.hpp"

namespace tnn {
Json bench_train(const std::string& model = "cifar100_wrn16_8", size_t batch = 256, int steps = 30, int warmup = 5, int gpu = 0, bool use_graph = true);
}
