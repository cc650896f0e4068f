// Training driver (reference: include/nn/train.hpp:45-86, src/nn/train.cpp:50-428).
//   TrainingConfig  env / .env / JSON configuration (same keys)
//   TrainStep       one optimisation step: forward, fused loss + gradient, backward, optimizer.  On a GPU the whole step
//                   is captured with cudaStreamBeginCapture into ONE CUDA graph after two eager warm-up steps and
//                   replayed afterwards; per-step scalars travel through the device step state (stepstate.hpp).
//   train_model / validate_model   epoch + validation loop with best-validation checkpointing
#pragma once
#include <functional>
#include <map>
#include <string>

#include "tnn/data_loader.hpp"
#include "tnn/optimizers.hpp"

namespace tnn {

struct TrainingConfig {
  size_t epochs = 10;
  size_t batch_size = 32;
  long max_steps = -1;
  double lr_initial = 1e-3;
  size_t gradient_accumulation_steps = 1;
  size_t progress_print_interval = 100;
  size_t num_threads = 4;
  std::string profiler_type = "NONE";
  bool print_layer_profiling = false;
  bool print_layer_memory_usage = false;
  std::string model_name = "cifar100_wrn16_8";
  std::string model_path;
  std::string dataset_name;
  std::string dataset_path = "data";
  std::string device_type = "CPU";
  std::string io_dtype = "FP32", param_dtype = "FP32", compute_dtype = "FP32";
  size_t num_microbatches = 2;
  std::string snapshot_dir = "model_snapshots";
  uint64_t seed = 0;
  static TrainingConfig load_from_env(const std::string& dotenv_path = ".env");   // src/nn/train.cpp:50-82
  static TrainingConfig load_from_json(const std::string& path);                  // src/nn/train.cpp:84-127
  Json to_json() const;
  Device device() const;
};

class TrainStep {
 public:
  TrainStep(Graph& g, Loss& criterion, Optimizer& opt, Scheduler* sched = nullptr, size_t accum = 1, bool use_graph = true);
  ~TrainStep();
  // x: FP32 (or BF16) batch, y: INT32 labels; device tensors or PINNED host tensors.  Returns the device-resident
  // stats tensor {mean loss, #correct} of this step (nothing here synchronises the host).
  Tensor operator()(const Tensor& x, const Tensor& y, size_t mb_id = 0);
  bool captured() const { return graph_exec_ != nullptr; }
  uint64_t replays() const { return replays_; }
  uint64_t captured_launches() const { return captured_launches_; }
  GraphExecutor& executor() { return exec_; }
  // replaces the optimizer update of this step (eager and captured): the data-parallel trainer launches the fused
  // cross-GPU reduce + optimizer + broadcast kernel here (data_parallel.hpp)
  std::function<void()> update_override;

 private:
  Tensor eager(const Tensor& x, const Tensor& y, size_t mb_id);
  void refresh_conv_flips();               // every Conv2D's flipped dgrad weights in one kernel launch
  Tensor flip_descs_;
  size_t flip_count_ = 0;
  int flip_blocks_ = 0;
  Tensor core(const Tensor& x, const Tensor& y, size_t mb_id);
  void capture(const Tensor& x, const Tensor& y);
  Graph& g_;
  Loss& crit_;
  Optimizer& opt_;
  Scheduler* sched_;
  size_t accum_, micro_ = 0;
  bool use_graph_;
  GraphExecutor exec_;
  int warm_ = 0;
  cudaGraph_t graph_ = nullptr;
  cudaGraphExec_t graph_exec_ = nullptr;
  Tensor sx_, sy_, sstats_;
  uint64_t replays_ = 0, captured_launches_ = 0;
};

struct TrainHistory {
  Vec<double> train_loss, val_acc, step_ms;
};
// (mean loss, accuracy) over the loader
std::pair<double, double> validate_model(Graph& g, BaseDataLoader& loader, Loss& criterion, const TrainingConfig& cfg);
// the table TrainingConfig::print_layer_memory_usage prints before and after training
std::string layer_memory_table(Graph& g, GraphExecutor& exec);
TrainHistory train_model(Graph& g, BaseDataLoader& train, BaseDataLoader* val, Optimizer& opt, Loss& criterion, Scheduler* sched,
                         const TrainingConfig& cfg, const std::function<void(const std::string&)>& log = nullptr);
// model file (reference format) + trailer with optimizer moments / step / scheduler position
void save_checkpoint(Graph& g, const std::string& path, Optimizer* opt = nullptr, Scheduler* sched = nullptr, uint64_t step = 0);
Graph load_checkpoint(const std::string& path, const Device& dev, Optimizer* opt = nullptr, Scheduler* sched = nullptr, uint64_t* step = nullptr);

}  // namespace tnn
