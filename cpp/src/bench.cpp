// Flagship-step benchmark shared by cpp/bin/bench_cpp and the Python binding (tnn_b200._host.bench_train).
#include <chrono>
#include <iostream>

#include "kernels/simt_ops.cuh"
#include "tnn/bench.hpp"
#include "tnn/example_models.hpp"
#include "tnn/stepstate.hpp"
#include "tnn/train.hpp"

namespace tnn {

Json bench_train(const std::string& model, size_t B, int steps, int warmup, int gpu, bool use_graph) {
  Json out = Json::object();
  if (!DeviceManager::getInstance().has_gpu()) {
    out["impl"] = "tnn_b200_cpp";
    out["unavailable"] = "no CUDA device";
    return out;
  }
  const Device dev = getGPU(gpu);
  dev.set_current();
  Graph graph = load_or_create_model(model, "", dev, 1234);
  auto crit = LossFactory::create_logsoftmax_crossentropy();
  auto opt = OptimizerFactory::create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4, false);
  TrainStep step(graph, *crit, *opt, nullptr, 1, use_graph);
  const Shape s = ExampleModels::input_shape(model);
  const size_t classes = ExampleModels::num_classes(model);
  SyntheticDataLoader loader(s, classes, 4 * B, 100);
  loader.set_pinned(true);
  const int NB = 4;
  Tensor hx[NB], hy[NB], dx[NB], dy[NB];
  for (int i = 0; i < NB; ++i) {
    loader.get_batch(B, hx[i], hy[i]);
    dx[i] = hx[i]->to_device(dev);
    dy[i] = hy[i]->to_device(dev);
  }
  Flow& flow = current_flow(dev);
  for (int i = 0; i < std::max(warmup, 3); ++i) step(dx[i % NB], dy[i % NB]);
  flow.synchronize();
  Event e0(true), e1(true);
  const unsigned long long l0 = g_launch_count;
  const uint64_t r0 = step.replays();
  flow.record(e0);
  Tensor stats;
  for (int i = 0; i < steps; ++i) stats = step(dx[i % NB], dy[i % NB]);
  flow.record(e1);
  flow.synchronize();
  const double ms = Event::elapsed_ms(e0, e1);
  const unsigned long long launches = (g_launch_count - l0) + (step.replays() - r0) * step.captured_launches();
  float st[2];
  stats->copy_to_host(st, 8);
  // ---- end to end: pinned host -> device every step, the previous step's loss read back every step
  Tensor loss_host = make_tensor(DeviceAllocator::pinned(), DType_t::FP32, {2});
  Event copied(false);
  auto run = [&](int n) {
    float last = 0;
    for (int i = 0; i < n; ++i) {
      Tensor s2 = step(hx[i % NB], hy[i % NB]);
      if (i > 0) {
        copied.synchronize();
        last = loss_host->data_as<float>()[0];
      }
      TNN_CUDA(cudaMemcpyAsync(loss_host->data(), s2->data(), 8, cudaMemcpyDeviceToHost, flow.stream()));
      flow.record(copied);
    }
    copied.synchronize();
    return last + loss_host->data_as<float>()[0] * 0.f;
  };
  run(std::max(3, warmup));
  flow.synchronize();
  const auto t0 = std::chrono::steady_clock::now();
  flow.record(e0);
  run(steps);
  flow.record(e1);
  flow.synchronize();
  const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  const double ms2 = std::max((double)Event::elapsed_ms(e0, e1), wall);
  out["metric"] = model == "cifar100_wrn16_8" ? "WRN-16-8 CIFAR-100 training throughput" : model + " training throughput";
  out["value"] = (double)B * steps / (ms / 1e3);
  out["unit"] = "img/s";
  out["n_gpus"] = 1;
  out["steps"] = steps;
  out["warmup"] = warmup;
  out["ms_per_step"] = ms / steps;
  out["higher_is_better"] = true;
  out["scaling"] = "weak";
  out["vs_baseline"] = (double)B * steps / (ms / 1e3) / 2480.0;
  out["dtype"] = "bf16";
  out["data"] = "synthetic";
  out["impl"] = "tnn_b200_cpp";
  Json cfg = Json::object();
  cfg["model"] = model;
  cfg["global_batch"] = B;
  cfg["per_gpu_batch"] = B;
  cfg["optimizer"] = "adam(lr1e-3,eps1e-3,wd3e-4)";
  cfg["parallelism"] = "dp1";
  cfg["host"] = "C++ (libtnn_b200: no libtorch in the step); step captured with cudaStreamBeginCapture";
  cfg["l2"] = "inputs rotate over 4 batches; per-step activation working set >> 126 MB L2";
  out["config"] = cfg;
  out["gpu_launches"] = (long long)launches;
  out["cuda_graph"] = step.captured();
  out["launches_per_step"] = (double)launches / std::max(1, steps);
  out["final_loss"] = st[0];
  Json e2e = Json::object();
  e2e["value"] = (double)B * steps / (ms2 / 1e3);
  e2e["unit"] = "img/s";
  e2e["ms_per_step"] = ms2 / steps;
  e2e["h2d_bytes_per_step"] = (long long)(hx[0]->nbytes() + hy[0]->nbytes());
  e2e["d2h_bytes_per_step"] = 8;
  out["e2e"] = e2e;
  out["workspace_bytes"] = (long long)step.executor().workspace().bytes_reserved();
  return out;
}

}  // namespace tnn
