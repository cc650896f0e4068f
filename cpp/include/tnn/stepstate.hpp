// Per-device "step state": 16 floats in device memory that every graph-replayable kernel reads its per-step scalars
// from (learning rate, Adam bias corrections, dropout seed counter).  The host refreshes it with ONE small
// pinned->device copy per step, issued on the compute flow BEFORE the (captured) step, so a captured training step can
// be replayed while lr schedules, bias corrections and dropout masks keep changing.  (The reference passes these
// scalars by value to one kernel per parameter tensor -- src/nn/optimizers_impl/cuda/adam_kernels.cu:22,59 -- which is
// exactly what cannot be baked into a CUDA graph.)
// layout: [0] lr [1] beta1 [2] beta2 [3] eps [4] weight_decay [5] decoupled [6] grad_scale [7] momentum
//         [8] 1-beta1^t [9] 1-beta2^t [12..13] uint64 seed counter
#pragma once
#include <cstdint>

#include "tnn/tensor.hpp"

namespace tnn {

class Optimizer;

class StepState {
 public:
  static constexpr size_t RING = 2048;
  explicit StepState(const Device& d);
  static StepState& get(const Device& d);
  // bump the seed counter, refresh optimizer scalars (when given), push to the device (async on the current flow)
  void advance(const Optimizer* opt = nullptr);
  const float* dev() const { return dev_ ? dev_->data_as<float>() : host_; }
  const uint64_t* dev_seed_ptr() const { return reinterpret_cast<const uint64_t*>(dev() + 12); }
  const float* host() const { return host_; }
  uint64_t counter() const { return counter_; }

 private:
  Device device_;
  Tensor ring_;          // pinned [RING, 16]: every step has its own source row (the host may run ahead of the device)
  Tensor dev_;
  float* host_ = nullptr;
  uint64_t counter_ = 0;
};

// Work that hangs off the backward critical path (convolution weight gradients) runs on the device's SIDE flow and is
// joined before the optimizer (reference: one Flow per device, dgrad and wgrad back to back -- conv2d_layer.cpp backward).
namespace side_flow {
constexpr flowHandle_t kHandle = 1;
bool enabled();
void set_enabled(bool on);
// run fn() on the side flow after everything queued so far on the current flow; `keep` stays referenced until join()
void run(const Device& d, const std::function<void()>& fn, std::vector<Tensor> keep);
bool has_pending(const Device& d);
void join(const Device& d);      // current flow waits for the side flow; kept references are dropped
}  // namespace side_flow

}  // namespace tnn
