// Losses, optimizers, LR schedulers.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "kernels/simt_ops.cuh"
#include "tnn/optimizers.hpp"
#include "tnn/stepstate.hpp"

namespace tnn {

// ===================================================================================== losses
namespace {
// labels as device/host int32 [rows] (class ids); accepts INT32, INT64, float ids and dense one-hot rows
Tensor labels_int32(const Tensor& target, size_t rows, size_t V, const Device& dev) {
  Tensor h = target->is_gpu() && target->data_type() != DType_t::INT32_T ? target->to_host() : target;
  if (h->data_type() == DType_t::INT32_T) return h->device() == dev ? h : h->to_device(dev);
  Tensor o = make_tensor(DType_t::INT32_T, {rows}, getHost());
  const auto v = to_float_vector(*h);
  if (v.size() == rows) {
    for (size_t i = 0; i < rows; ++i) o->data_as<int32_t>()[i] = (int32_t)v[i];
  } else if (v.size() == rows * V) {             // one-hot (open_webtext_data_loader.hpp:38-44 style)
    for (size_t i = 0; i < rows; ++i) o->data_as<int32_t>()[i] = (int32_t)(std::max_element(v.begin() + (long)(i * V), v.begin() + (long)((i + 1) * V)) - (v.begin() + (long)(i * V)));
  } else {
    throw std::invalid_argument("loss: target has " + std::to_string(v.size()) + " elements for " + std::to_string(rows) + " rows");
  }
  return dev.is_gpu() ? o->to_device(dev) : o;
}

class CrossEntropyLoss : public Loss {
 public:
  explicit CrossEntropyLoss(bool use_logits, double eps, const char* n) : logits_(use_logits), eps_(eps), name_(n) {}
  std::string name() const override { return name_; }
  void loss_and_grad(const Tensor& pred, const Tensor& target, Tensor& stats, Tensor* grad, float scale) override {
    const size_t V = pred->shape().back(), rows = pred->size() / V;
    const Device& dev = pred->device();
    Tensor labels = labels_int32(target, rows, V, dev);
    if (!stats) {
      stats = make_tensor(DType_t::FP32, {2}, dev);
      stats->zero();
    }
    if (grad && !*grad) *grad = make_tensor(pred->data_type(), pred->shape(), dev);
    if (dev.is_gpu() && logits_) {
      dev.set_current();
      softmax_ce(pred->data(), pred->data_type() == DType_t::FP32, (long long)V, labels->data_as<int>(), (long long)rows, (int)V, scale / (float)rows,
                 scale / (float)rows, grad ? (*grad)->data() : nullptr, stats->data_as<float>(), current_flow(dev).stream());
      return;
    }
    Tensor ph = dev.is_gpu() ? pred->to_host() : pred;
    const auto x = to_float_vector(*ph);
    std::vector<float> g(grad ? x.size() : 0);
    double loss = 0, correct = 0;
    for (size_t r = 0; r < rows; ++r) {
      const float* row = x.data() + r * V;
      const int y = labels->is_gpu() ? 0 : labels->data_as<int32_t>()[r];
      const size_t arg = (size_t)(std::max_element(row, row + V) - row);
      correct += (int)arg == y;
      if (logits_) {
        float mx = row[arg];
        double sum = 0;
        for (size_t c = 0; c < V; ++c) sum += std::exp((double)row[c] - mx);
        const double lse = mx + std::log(sum);
        loss += lse - row[y];
        if (grad)
          for (size_t c = 0; c < V; ++c) g[r * V + c] = (float)((std::exp((double)row[c] - lse) - (c == (size_t)y ? 1.0 : 0.0)) * scale / (double)rows);
      } else {
        const double p = std::min(std::max((double)row[y], eps_), 1.0 - eps_);
        loss += -std::log(p);
        if (grad)
          for (size_t c = 0; c < V; ++c) g[r * V + c] = c == (size_t)y ? (float)(-1.0 / p * scale / (double)rows) : 0.f;
      }
    }
    float st[2];
    Tensor sh = dev.is_gpu() ? stats->to_host() : stats;
    st[0] = sh->data_as<float>()[0] + (float)(loss * scale / (double)rows);
    st[1] = sh->data_as<float>()[1] + (float)correct;
    stats->copy_from_host(st, 8);
    if (grad) {
      Tensor gh = make_tensor_from(g, pred->shape());
      if ((*grad)->data_type() == DType_t::FP32) (*grad)->copy_from(*gh);
      else (*grad)->copy_from(*dtype_cast(*gh, (*grad)->data_type()));
    }
  }
  TConfig get_config() const override {
    TConfig c(name_, name_);
    c.set("use_logits", logits_).set("epsilon", eps_);
    return c;
  }

 private:
  bool logits_;
  double eps_;
  std::string name_;
};

class RegressionLoss : public Loss {
 public:
  RegressionLoss(int kind, double delta) : kind_(kind), delta_(delta) {}
  std::string name() const override { return kind_ == 0 ? "mse" : kind_ == 1 ? "mae" : "huber"; }
  void loss_and_grad(const Tensor& pred, const Tensor& target, Tensor& stats, Tensor* grad, float scale) override {
    const Device& dev = pred->device();
    const size_t n = pred->size();
    if (!stats) {
      stats = make_tensor(DType_t::FP32, {2}, dev);
      stats->zero();
    }
    if (grad && !*grad) *grad = make_tensor(pred->data_type(), pred->shape(), dev);
    Tensor t = target->device() == dev ? target : target->to_device(dev);
    if (t->data_type() != pred->data_type()) t = dtype_cast(*t, pred->data_type());
    if (dev.is_gpu() && (pred->data_type() == DType_t::FP32 || pred->data_type() == DType_t::BF16)) {
      dev.set_current();
      regression_loss(pred->data(), t->data(), pred->data_type() == DType_t::FP32, (long long)n, kind_, (float)delta_, scale / (float)n,
                      grad ? (*grad)->data() : nullptr, stats->data_as<float>(), current_flow(dev).stream());
      return;
    }
    const auto p = to_float_vector(*pred), y = to_float_vector(*t);
    std::vector<float> g(grad ? n : 0);
    double loss = 0;
    for (size_t i = 0; i < n; ++i) {
      const double d = (double)p[i] - y[i];
      double l, gd;
      if (kind_ == 0) { l = d * d; gd = 2 * d; }
      else if (kind_ == 1) { l = std::fabs(d); gd = d > 0 ? 1 : d < 0 ? -1 : 0; }
      else if (std::fabs(d) <= delta_) { l = 0.5 * d * d; gd = d; }
      else { l = delta_ * (std::fabs(d) - 0.5 * delta_); gd = d > 0 ? delta_ : -delta_; }
      loss += l;
      if (grad) g[i] = (float)(gd * scale / (double)n);
    }
    float st[2] = {stats->data_as<float>()[0] + (float)(loss * scale / (double)n), stats->data_as<float>()[1]};
    stats->copy_from_host(st, 8);
    if (grad) (*grad)->copy_from(*make_tensor_from(g, pred->shape()));
  }
  TConfig get_config() const override {
    TConfig c(name(), name());
    c.set("delta", delta_);
    return c;
  }

 private:
  int kind_;
  double delta_;
};
}  // namespace

void Loss::compute_loss(const Tensor& pred, const Tensor& target, float& loss) {
  Tensor stats;
  loss_and_grad(pred, target, stats, nullptr, 1.f);
  float st[2];
  stats->copy_to_host(st, 8);
  loss = st[0];
}
void Loss::compute_gradient(const Tensor& pred, const Tensor& target, Tensor& grad) {
  Tensor stats;
  loss_and_grad(pred, target, stats, &grad, 1.f);
}
std::unique_ptr<Loss> LossFactory::create_crossentropy(bool use_logits, double eps) { return std::make_unique<CrossEntropyLoss>(use_logits, eps, "crossentropy"); }
std::unique_ptr<Loss> LossFactory::create_logsoftmax_crossentropy() { return std::make_unique<CrossEntropyLoss>(true, 1e-7, "logsoftmax_crossentropy"); }
std::unique_ptr<Loss> LossFactory::create_mse() { return std::make_unique<RegressionLoss>(0, 1.0); }
std::unique_ptr<Loss> LossFactory::create_mae() { return std::make_unique<RegressionLoss>(1, 1.0); }
std::unique_ptr<Loss> LossFactory::create_huber(double delta) { return std::make_unique<RegressionLoss>(2, delta); }
std::unique_ptr<Loss> LossFactory::create(const std::string& name) {
  std::string n = name;
  for (auto& c : n) c = (char)tolower((unsigned char)c);
  if (n == "crossentropy" || n == "ce") return create_crossentropy();
  if (n == "logsoftmax_crossentropy" || n == "logsoftmax_ce") return create_logsoftmax_crossentropy();
  if (n == "mse" || n == "mean_squared_error") return create_mse();
  if (n == "mae" || n == "mean_absolute_error") return create_mae();
  if (n == "huber") return create_huber();
  throw std::invalid_argument("LossFactory: unknown loss '" + name + "'");
}
std::unique_ptr<Loss> LossFactory::create_from_config(const TConfig& c) {
  const std::string n = c.type.empty() ? c.name : c.type;
  if (n == "crossentropy" || n == "ce") return create_crossentropy(c.get<bool>("use_logits", true), c.get<double>("epsilon", 1e-7));
  if (n == "huber") return create_huber(c.get<double>("delta", 1.0));
  return create(n);
}
size_t compute_class_corrects(const Tensor& pred, const Tensor& labels) {
  const size_t V = pred->shape().back(), rows = pred->size() / V;
  Tensor lab = labels_int32(labels, rows, V, getHost());
  const auto x = to_float_vector(*pred);
  size_t ok = 0;
  for (size_t r = 0; r < rows; ++r) ok += (int)(std::max_element(x.begin() + (long)(r * V), x.begin() + (long)((r + 1) * V)) - (x.begin() + (long)(r * V))) == lab->data_as<int32_t>()[r];
  return ok;
}

// ===================================================================================== optimizers
void Optimizer::attach(GraphContext& ctx) { ctx_ = &ctx; }
std::unique_ptr<Optimizer> Optimizer::clone() const { return OptimizerFactory::create_from_config(get_config()); }
std::unique_ptr<Optimizer> OptimizerFactory::create_from_config(const TConfig& c) {
  std::string t = c.type.empty() ? c.name : c.type;
  for (auto& ch : t) ch = (char)tolower((unsigned char)ch);
  if (t == "sgd") return create_sgd(c.get<double>("learning_rate", 0.01), c.get<double>("momentum", 0.0), c.get<double>("weight_decay", 0.0));
  if (t == "adam" || t == "adamw")
    return create_adam(c.get<double>("learning_rate", 1e-3), c.get<double>("beta1", 0.9), c.get<double>("beta2", 0.999), c.get<double>("epsilon", 1e-8),
                       c.get<double>("weight_decay", 0.0), c.get<bool>("decouple_weight_decay", t == "adamw"));
  throw std::invalid_argument("OptimizerFactory: unknown optimizer '" + t + "'");
}

static void fill_common(float* h, double lr, double b1, double b2, double eps, double wd, bool dec, double gs, double mom, uint64_t t) {
  h[0] = (float)lr; h[1] = (float)b1; h[2] = (float)b2; h[3] = (float)eps; h[4] = (float)wd; h[5] = dec ? 1.f : 0.f;
  h[6] = (float)gs; h[7] = (float)mom;
  h[8] = (float)(1.0 - std::pow(b1, (double)t));
  h[9] = (float)(1.0 - std::pow(b2, (double)t));
}
void SGD::attach(GraphContext& ctx) {
  Optimizer::attach(ctx);
  if (momentum_ > 0) {
    velocity_ = make_tensor(DType_t::FP32, ctx.param_slab()->shape(), ctx.device());
    velocity_->zero();
  }
}
TConfig SGD::get_config() const {
  TConfig c("sgd", "sgd");
  c.set("learning_rate", lr_).set("momentum", momentum_).set("weight_decay", wd_);
  return c;
}
void SGD::write_step_state(float* h) const { fill_common(h, lr_, 0, 0, 0, wd_, false, grad_scale, momentum_, step_count + 1); }
std::map<std::string, Tensor> SGD::state_tensors() { return velocity_ ? std::map<std::string, Tensor>{{"velocity", velocity_}} : std::map<std::string, Tensor>{}; }
void SGD::update(bool external_state) {
  if (!ctx_) throw std::runtime_error("SGD::update before attach");
  Tensor& p = ctx_->param_slab();
  Tensor& g = ctx_->grad_slab();
  const long long n = (long long)p->size();
  if (ctx_->device().is_gpu()) {
    ctx_->device().set_current();
    StepState& st = StepState::get(ctx_->device());
    if (!external_state) st.advance(this);
    sgd_step_flat_dev(p->data_as<float>(), g->data_as<float>(), velocity_ ? velocity_->data_as<float>() : nullptr, ctx_->lp_slab()->data(), n, st.dev(),
                      fused_zero_grad ? 1 : 0, current_flow(ctx_->device()).stream());
  } else {
    float* pp = p->data_as<float>();
    float* gp = g->data_as<float>();
    float* vp = velocity_ ? velocity_->data_as<float>() : nullptr;
    for (long long i = 0; i < n; ++i) {
      const float gg = gp[i] * (float)grad_scale + (float)wd_ * pp[i];
      if (vp) {
        vp[i] = (float)momentum_ * vp[i] - (float)lr_ * gg;
        pp[i] += vp[i];
      } else {
        pp[i] -= (float)lr_ * gg;
      }
      if (fused_zero_grad) gp[i] = 0.f;
    }
  }
  ++step_count;
  ctx_->mark_lp_fresh();
}

void Adam::attach(GraphContext& ctx) {
  Optimizer::attach(ctx);
  m_ = make_tensor(DType_t::FP32, ctx.param_slab()->shape(), ctx.device());
  v_ = make_tensor(DType_t::FP32, ctx.param_slab()->shape(), ctx.device());
  m_->zero();
  v_->zero();
}
TConfig Adam::get_config() const {
  TConfig c("adam", "adam");
  c.set("learning_rate", lr_).set("beta1", b1_).set("beta2", b2_).set("epsilon", eps_).set("weight_decay", wd_).set("decouple_weight_decay", decoupled_);
  return c;
}
void Adam::write_step_state(float* h) const { fill_common(h, lr_, b1_, b2_, eps_, wd_, decoupled_, grad_scale, 0, step_count + 1); }
std::map<std::string, Tensor> Adam::state_tensors() { return {{"m", m_}, {"v", v_}}; }
void Adam::update(bool external_state) {
  if (!ctx_) throw std::runtime_error("Adam::update before attach");
  Tensor& p = ctx_->param_slab();
  Tensor& g = ctx_->grad_slab();
  const long long n = (long long)p->size();
  if (ctx_->device().is_gpu()) {
    ctx_->device().set_current();
    StepState& st = StepState::get(ctx_->device());
    if (!external_state) st.advance(this);
    adam_step_flat_dev(p->data_as<float>(), g->data_as<float>(), m_->data_as<float>(), v_->data_as<float>(), ctx_->lp_slab()->data(), n, st.dev(),
                       fused_zero_grad ? 1 : 0, current_flow(ctx_->device()).stream());
  } else {
    const uint64_t t = step_count + 1;
    const float bc1 = (float)(1.0 - std::pow(b1_, (double)t)), bc2 = (float)(1.0 - std::pow(b2_, (double)t));
    float* pp = p->data_as<float>();
    float* gp = g->data_as<float>();
    float* mp = m_->data_as<float>();
    float* vp = v_->data_as<float>();
    const float decay = (wd_ != 0 && decoupled_) ? (float)(1.0 - wd_ * lr_) : 1.f, l2 = (wd_ != 0 && !decoupled_) ? (float)(wd_ * lr_) : 0.f;
    for (long long i = 0; i < n; ++i) {
      const float gg = gp[i] * (float)grad_scale, pi = pp[i] * decay;
      mp[i] = (float)b1_ * mp[i] + (1.f - (float)b1_) * gg;
      vp[i] = (float)b2_ * vp[i] + (1.f - (float)b2_) * gg * gg;
      pp[i] = pi - ((float)lr_ * (mp[i] / bc1) / (std::sqrt(vp[i] / bc2) + (float)eps_) + l2 * pi);
      if (fused_zero_grad) gp[i] = 0.f;
    }
  }
  ++step_count;
  ctx_->mark_lp_fresh();
}

// ===================================================================================== schedulers
namespace {
constexpr double kPi = 3.14159265358979323846;
// closed-form schedules: lr = f(step index, base lr)
class FnScheduler : public Scheduler {
 public:
  FnScheduler(Optimizer* o, std::string type, TConfig cfg, std::function<double(uint64_t, double)> fn)
      : Scheduler(o), type_(std::move(type)), cfg_(std::move(cfg)), fn_(std::move(fn)) {}
  std::string name() const override { return type_; }
  TConfig get_config() const override { return cfg_; }

 protected:
  double lr_at(uint64_t t, double) override { return fn_(t, base_lr_); }
  std::string type_;
  TConfig cfg_;
  std::function<double(uint64_t, double)> fn_;
};
TConfig sched_cfg(const char* type) { return TConfig(type, type); }

class ReduceOnPlateau : public Scheduler {
 public:
  ReduceOnPlateau(Optimizer* o, double f, size_t pat, double thr, double mn, bool mx) : Scheduler(o), factor_(f), patience_(pat), thr_(thr), min_(mn), mode_max_(mx), cur_(base_lr_) {}
  std::string name() const override { return "reduce_lr_on_plateau"; }
  TConfig get_config() const override {
    TConfig c("reduce_lr_on_plateau", "reduce_lr_on_plateau");
    c.set("factor", factor_).set("patience", patience_).set("threshold", thr_).set("min_lr", min_).set("mode_max", mode_max_);
    return c;
  }

 protected:
  double lr_at(uint64_t, double metric) override {
    if (std::isnan(metric)) return cur_;
    const bool better = !have_ || (mode_max_ ? metric > best_ * (1 + thr_) : metric < best_ * (1 - thr_));
    if (better) { best_ = metric; have_ = true; bad_ = 0; }
    else if (++bad_ > patience_) { cur_ = std::max(min_, cur_ * factor_); bad_ = 0; }
    return cur_;
  }
  double factor_;
  size_t patience_;
  double thr_, min_;
  bool mode_max_;
  double cur_, best_ = 0;
  bool have_ = false;
  size_t bad_ = 0;
};
}  // namespace

std::unique_ptr<Scheduler> SchedulerFactory::create_no_op(Optimizer* o) {
  return std::make_unique<FnScheduler>(o, "no_op", sched_cfg("no_op"), [](uint64_t, double base) { return base; });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_step_lr(Optimizer* o, size_t step, double gamma) {
  step = std::max<size_t>(step, 1);
  TConfig c = sched_cfg("step_lr");
  c.set("step_size", step).set("gamma", gamma);
  return std::make_unique<FnScheduler>(o, "step_lr", c, [=](uint64_t t, double base) { return base * std::pow(gamma, (double)(t / step)); });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_multi_step_lr(Optimizer* o, Vec<size_t> ms, double gamma) {
  std::sort(ms.begin(), ms.end());
  TConfig c = sched_cfg("multi_step_lr");
  c.set("milestones", ms).set("gamma", gamma);
  return std::make_unique<FnScheduler>(o, "multi_step_lr", c, [=](uint64_t t, double base) {
    size_t k = 0;
    for (size_t m : ms) k += t >= m;
    return base * std::pow(gamma, (double)k);
  });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_exponential_lr(Optimizer* o, double gamma) {
  TConfig c = sched_cfg("exponential_lr");
  c.set("gamma", gamma);
  return std::make_unique<FnScheduler>(o, "exponential_lr", c, [=](uint64_t t, double base) { return base * std::pow(gamma, (double)t); });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_cosine_annealing_lr(Optimizer* o, size_t tmax, double eta) {
  tmax = std::max<size_t>(tmax, 1);
  TConfig c = sched_cfg("cosine_annealing_lr");
  c.set("T_max", tmax).set("eta_min", eta);
  return std::make_unique<FnScheduler>(o, "cosine_annealing_lr", c, [=](uint64_t t, double base) {
    return eta + (base - eta) * (1 + std::cos(kPi * (double)(t % (2 * tmax)) / (double)tmax)) / 2;
  });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_cosine_annealing_warm_restarts(Optimizer* o, size_t t0, size_t mult, double eta) {
  t0 = std::max<size_t>(t0, 1);
  mult = std::max<size_t>(mult, 1);
  TConfig c = sched_cfg("cosine_annealing_warm_restarts");
  c.set("T_0", t0).set("T_mult", mult).set("eta_min", eta);
  return std::make_unique<FnScheduler>(o, "cosine_annealing_warm_restarts", c, [=](uint64_t t, double base) {
    uint64_t cur = t;
    size_t ti = t0;
    while (cur >= ti) {
      cur -= ti;
      ti *= mult;
    }
    return eta + (base - eta) * (1 + std::cos(kPi * (double)cur / (double)ti)) / 2;
  });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_linear_warmup(Optimizer* o, size_t warm, double start) {
  warm = std::max<size_t>(warm, 1);
  TConfig c = sched_cfg("linear_warmup");
  c.set("warmup_steps", warm).set("start_factor", start);
  return std::make_unique<FnScheduler>(o, "linear_warmup", c, [=](uint64_t t, double base) {
    return t >= warm ? base : base * (start + (1 - start) * (double)t / (double)warm);
  });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_warmup_cosine_annealing(Optimizer* o, size_t warm, size_t total, double eta) {
  total = std::max(total, warm + 1);
  TConfig c = sched_cfg("warmup_cosine_annealing");
  c.set("warmup_steps", warm).set("total_steps", total).set("eta_min", eta);
  return std::make_unique<FnScheduler>(o, "warmup_cosine_annealing", c, [=](uint64_t t, double base) {
    if (t < warm) return base * (double)(t + 1) / (double)warm;
    const double p = std::min(1.0, (double)(t - warm) / (double)(total - warm));
    return eta + (base - eta) * (1 + std::cos(kPi * p)) / 2;
  });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_reduce_lr_on_plateau(Optimizer* o, double f, size_t p, double thr, double mn, bool mx) {
  return std::make_unique<ReduceOnPlateau>(o, f, p, thr, mn, mx);
}
std::unique_ptr<Scheduler> SchedulerFactory::create_polynomial_lr(Optimizer* o, size_t total, double power, double end) {
  total = std::max<size_t>(total, 1);
  TConfig c = sched_cfg("polynomial_lr");
  c.set("total_steps", total).set("power", power).set("end_lr", end);
  return std::make_unique<FnScheduler>(o, "polynomial_lr", c, [=](uint64_t t, double base) {
    return end + (base - end) * std::pow(1.0 - std::min(1.0, (double)t / (double)total), power);
  });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_one_cycle_lr(Optimizer* o, double mx, size_t total, double pct, double div, double fdiv) {
  total = std::max<size_t>(total, 2);
  TConfig c = sched_cfg("one_cycle_lr");
  c.set("max_lr", mx).set("total_steps", total).set("pct_start", pct).set("div_factor", div).set("final_div_factor", fdiv);
  return std::make_unique<FnScheduler>(o, "one_cycle_lr", c, [=](uint64_t t, double) {
    const double init = mx / div, fin = init / fdiv, up = std::max(1.0, pct * (double)total - 1);
    const double tt = (double)std::min<uint64_t>(t, total - 1);
    if (tt <= up) return init + (mx - init) * (1 - std::cos(kPi * tt / up)) / 2;
    const double down = std::max(1.0, (double)total - 1 - up);
    return fin + (mx - fin) * (1 + std::cos(kPi * (tt - up) / down)) / 2;
  });
}
std::unique_ptr<Scheduler> SchedulerFactory::create_from_config(const TConfig& c, Optimizer* o) {
  const std::string t = c.type.empty() ? c.name : c.type;
  if (t == "no_op" || t == "noop" || t.empty()) return create_no_op(o);
  if (t == "step_lr") return create_step_lr(o, c.get<size_t>("step_size", 1), c.get<double>("gamma", 0.1));
  if (t == "multi_step_lr") return create_multi_step_lr(o, c.get<Vec<size_t>>("milestones"), c.get<double>("gamma", 0.1));
  if (t == "exponential_lr") return create_exponential_lr(o, c.get<double>("gamma", 0.9));
  if (t == "cosine_annealing_lr") return create_cosine_annealing_lr(o, c.get<size_t>("T_max", 1), c.get<double>("eta_min", 0.0));
  if (t == "cosine_annealing_warm_restarts") return create_cosine_annealing_warm_restarts(o, c.get<size_t>("T_0", 1), c.get<size_t>("T_mult", 1), c.get<double>("eta_min", 0.0));
  if (t == "linear_warmup") return create_linear_warmup(o, c.get<size_t>("warmup_steps", 1), c.get<double>("start_factor", 0.0));
  if (t == "warmup_cosine_annealing") return create_warmup_cosine_annealing(o, c.get<size_t>("warmup_steps", 0), c.get<size_t>("total_steps", 1), c.get<double>("eta_min", 0.0));
  if (t == "reduce_lr_on_plateau") return create_reduce_lr_on_plateau(o, c.get<double>("factor", 0.1), c.get<size_t>("patience", 10), c.get<double>("threshold", 1e-4), c.get<double>("min_lr", 0.0), c.get<bool>("mode_max", false));
  if (t == "polynomial_lr") return create_polynomial_lr(o, c.get<size_t>("total_steps", 1), c.get<double>("power", 1.0), c.get<double>("end_lr", 0.0));
  if (t == "one_cycle_lr") return create_one_cycle_lr(o, c.get<double>("max_lr", 0.1), c.get<size_t>("total_steps", 2), c.get<double>("pct_start", 0.3), c.get<double>("div_factor", 25.0), c.get<double>("final_div_factor", 1e4));
  throw std::invalid_argument("SchedulerFactory: unknown scheduler '" + t + "'");
}

}  // namespace tnn
