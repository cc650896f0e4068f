// Loss functions, optimizers, LR schedulers (reference: include/nn/loss.hpp:24-523, optimizers.hpp:28-280,
// schedulers.hpp:27-751).  Optimizers step the WHOLE parameter slab in one kernel launch (reference: one launch per
// parameter tensor), read their hyper-parameters from the device step state (graph replayable), refresh the bf16
// compute copies and zero the gradient slab in the same pass.
#pragma once
#include <cmath>
#include <memory>
#include <string>

#include "tnn/graph.hpp"

namespace tnn {

// ------------------------------------------------------------------------------------- loss
class Loss {
 public:
  virtual ~Loss() = default;
  virtual std::string name() const = 0;
  // stats (device fp32 [2]) += {sum of per-row losses * scale / rows, #correct};  grad (same shape as pred) = d(mean loss)*scale
  virtual void loss_and_grad(const Tensor& pred, const Tensor& target, Tensor& stats, Tensor* grad, float scale = 1.f) = 0;
  // reference API (loss.hpp:28-38)
  void compute_loss(const Tensor& pred, const Tensor& target, float& loss);
  void compute_gradient(const Tensor& pred, const Tensor& target, Tensor& grad);
  virtual TConfig get_config() const { return TConfig(name(), name()); }
};
class LossFactory {
 public:
  static std::unique_ptr<Loss> create(const std::string& name);     // crossentropy|ce|logsoftmax_crossentropy|logsoftmax_ce|mse|mae|huber ...
  static std::unique_ptr<Loss> create_crossentropy(bool use_logits = true, double eps = 1e-7);
  static std::unique_ptr<Loss> create_logsoftmax_crossentropy();
  static std::unique_ptr<Loss> create_mse();
  static std::unique_ptr<Loss> create_mae();
  static std::unique_ptr<Loss> create_huber(double delta = 1.0);
  static std::unique_ptr<Loss> create_from_config(const TConfig& c);
};
// #rows whose argmax equals the int label (accuracy.hpp:14-40); synchronises
size_t compute_class_corrects(const Tensor& pred, const Tensor& labels);

// ------------------------------------------------------------------------------------- optimizers
class Optimizer {
 public:
  explicit Optimizer(double lr) : lr_(lr) {}
  virtual ~Optimizer() = default;
  virtual std::string name() const = 0;
  virtual void attach(GraphContext& ctx);
  bool attached() const { return ctx_ != nullptr; }
  GraphContext* context() { return ctx_; }
  // one optimisation step over the slab.  `external_state`: the caller (TrainStep) already pushed the step state
  virtual void update(bool external_state = false) = 0;
  void zero_grads() { if (ctx_) ctx_->zero_grads(); }
  void set_learning_rate(double lr) { lr_ = lr; }
  double get_learning_rate() const { return lr_; }
  virtual TConfig get_config() const = 0;
  std::unique_ptr<Optimizer> clone() const;
  // fills the optimizer part of a step-state row (see stepstate.hpp) for step_count + 1
  virtual void write_step_state(float* row) const = 0;
  virtual std::map<std::string, Tensor> state_tensors() { return {}; }
  uint64_t step_count = 0;
  double grad_scale = 1.0;            // 1 / accumulation steps (and 1 / world in data parallel)
  bool fused_zero_grad = true;

 protected:
  double lr_;
  GraphContext* ctx_ = nullptr;
};

class SGD : public Optimizer {
 public:
  SGD(double lr, double momentum = 0.0, double weight_decay = 0.0) : Optimizer(lr), momentum_(momentum), wd_(weight_decay) {}
  std::string name() const override { return "sgd"; }
  void attach(GraphContext& ctx) override;
  void update(bool external_state = false) override;
  TConfig get_config() const override;
  void write_step_state(float* row) const override;
  std::map<std::string, Tensor> state_tensors() override;

 private:
  double momentum_, wd_;
  Tensor velocity_;
};

class Adam : public Optimizer {
 public:
  Adam(double lr = 1e-3, double beta1 = 0.9, double beta2 = 0.999, double eps = 1e-8, double weight_decay = 0.0, bool decouple = false)
      : Optimizer(lr), b1_(beta1), b2_(beta2), eps_(eps), wd_(weight_decay), decoupled_(decouple) {}
  std::string name() const override { return "adam"; }
  void attach(GraphContext& ctx) override;
  void update(bool external_state = false) override;
  TConfig get_config() const override;
  void write_step_state(float* row) const override;
  std::map<std::string, Tensor> state_tensors() override;

 private:
  double b1_, b2_, eps_, wd_;
  bool decoupled_;
  Tensor m_, v_;
};

class OptimizerFactory {
 public:
  static std::unique_ptr<Optimizer> create_sgd(double lr = 0.01, double momentum = 0.0, double weight_decay = 0.0) { return std::make_unique<SGD>(lr, momentum, weight_decay); }
  static std::unique_ptr<Optimizer> create_adam(double lr = 1e-3, double b1 = 0.9, double b2 = 0.999, double eps = 1e-8, double wd = 0.0, bool decouple = false) {
    return std::make_unique<Adam>(lr, b1, b2, eps, wd, decouple);
  }
  static std::unique_ptr<Optimizer> create_from_config(const TConfig& c);
};

// ------------------------------------------------------------------------------------- schedulers (step() once per optimizer step)
class Scheduler {
 public:
  explicit Scheduler(Optimizer* opt) : opt_(opt), base_lr_(opt ? opt->get_learning_rate() : 0.0) {}
  virtual ~Scheduler() = default;
  virtual std::string name() const = 0;
  void step(double metric = NAN) {
    ++t_;
    const double lr = lr_at(t_, metric);
    if (opt_) opt_->set_learning_rate(lr);
    last_lr_ = lr;
  }
  double get_last_lr() const { return last_lr_ < 0 ? base_lr_ : last_lr_; }
  uint64_t step_index() const { return t_; }
  void set_step_index(uint64_t t) { t_ = t; }
  virtual TConfig get_config() const = 0;

 protected:
  virtual double lr_at(uint64_t t, double metric) = 0;
  Optimizer* opt_;
  double base_lr_;
  double last_lr_ = -1;
  uint64_t t_ = 0;
};

class SchedulerFactory {
 public:
  static std::unique_ptr<Scheduler> create_no_op(Optimizer* o);
  static std::unique_ptr<Scheduler> create_step_lr(Optimizer* o, size_t step_size, double gamma = 0.1);
  static std::unique_ptr<Scheduler> create_multi_step_lr(Optimizer* o, Vec<size_t> milestones, double gamma = 0.1);
  static std::unique_ptr<Scheduler> create_exponential_lr(Optimizer* o, double gamma);
  static std::unique_ptr<Scheduler> create_cosine_annealing_lr(Optimizer* o, size_t t_max, double eta_min = 0.0);
  static std::unique_ptr<Scheduler> create_cosine_annealing_warm_restarts(Optimizer* o, size_t t0, size_t t_mult = 1, double eta_min = 0.0);
  static std::unique_ptr<Scheduler> create_linear_warmup(Optimizer* o, size_t warmup_steps, double start_factor = 0.0);
  static std::unique_ptr<Scheduler> create_warmup_cosine_annealing(Optimizer* o, size_t warmup_steps, size_t total_steps, double eta_min = 0.0);
  static std::unique_ptr<Scheduler> create_reduce_lr_on_plateau(Optimizer* o, double factor = 0.1, size_t patience = 10, double threshold = 1e-4, double min_lr = 0.0, bool mode_max = false);
  static std::unique_ptr<Scheduler> create_polynomial_lr(Optimizer* o, size_t total_steps, double power = 1.0, double end_lr = 0.0);
  static std::unique_ptr<Scheduler> create_one_cycle_lr(Optimizer* o, double max_lr, size_t total_steps, double pct_start = 0.3, double div_factor = 25.0, double final_div_factor = 1e4);
  static std::unique_ptr<Scheduler> create_from_config(const TConfig& c, Optimizer* o);
};

}  // namespace tnn
