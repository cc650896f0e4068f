// Layer base class, factory, step state, side flow.
#include <cstring>
#include <random>

#include "kernels/simt_ops.cuh"
#include "tnn/layers.hpp"
#include "tnn/optimizers.hpp"
#include "tnn/stepstate.hpp"

namespace tnn {

std::function<void(TensorImpl&, uint64_t)> uniform_init(double b) {
  return [b](TensorImpl& t, uint64_t seed) { t.fill_random_uniform(-b, b, seed); };
}
std::function<void(TensorImpl&, uint64_t)> const_init(double v) {
  return [v](TensorImpl& t, uint64_t) { t.fill(v); };
}
std::function<void(TensorImpl&, uint64_t)> normal_init(double s) {
  return [s](TensorImpl& t, uint64_t seed) { t.fill_random_normal(0.0, s, seed); };
}

void Layer::init(const Device& dev, uint64_t seed) {
  device_ = dev;
  if (seed) seed_ = seed;
  size_t idx = 0;
  for (const auto& d : param_descriptors()) {
    if (!params_.count(d.name)) {
      Tensor p = make_tensor(DType_t::FP32, d.shape, dev);
      // the reference reuses one seed for every tensor of a layer (dense_layer.cpp:46-61); here each tensor gets its
      // own stream derived from (seed, layer name, index) so that weights and biases are not correlated
      uint64_t h = 1469598103934665603ull;
      for (char c : name_) h = (h ^ (unsigned char)c) * 1099511628211ull;
      if (d.init) d.init(*p, seed_ * 0x9E3779B97F4A7C15ull + h + idx * 7919);
      params_[d.name] = p;
      if (d.trainable) {
        Tensor g = make_tensor(DType_t::FP32, d.shape, dev);
        g->zero();
        grads_[d.name] = g;
      }
    }
    ++idx;
  }
  for (Layer* c : children()) c->init(dev, seed_);
  lp_stale_ = true;
  initialized_ = true;
}
void Block::init(const Device& dev, uint64_t seed) { Layer::init(dev, seed); }

void Layer::bind(const std::string& n, Tensor data, Tensor grad, Tensor lp) {
  params_[n] = std::move(data);
  if (grad) grads_[n] = std::move(grad);
  if (lp) lp_[n] = std::move(lp);
  lp_stale_ = true;
  ++param_version_;
}
const Tensor& Layer::lp(const std::string& n) {
  Tensor& p = params_.at(n);
  if (!device_.is_gpu()) return p;
  auto it = lp_.find(n);
  if (it == lp_.end()) {
    lp_[n] = make_tensor(DType_t::BF16, p->shape(), device_);
    lp_stale_ = true;
    it = lp_.find(n);
  }
  if (lp_stale_) {
    device_.set_current();
    for (auto& kv : lp_) {
      Tensor& src = params_.at(kv.first);
      cast_f32_to_bf16(src->data_as<float>(), kv.second->data(), (long long)src->size(), stream());
    }
    lp_stale_ = false;
  }
  return it->second;
}
TConfig Layer::get_config() const {
  TConfig c(name_, type());
  fill_config(c);
  return c;
}
Vec<Layer*> Layer::all_layers() {
  Vec<Layer*> out{this};
  for (Layer* c : children()) {
    auto sub = c->all_layers();
    out.insert(out.end(), sub.begin(), sub.end());
  }
  return out;
}
Tensor Layer::new_tensor(DType_t dt, const Shape& s) const {
  if (workspace_ && workspace_->device() == device_) return make_tensor(*workspace_, dt, s);
  return make_tensor(dt, s, device_);
}
Tensor Layer::new_persistent(DType_t dt, const Shape& s, bool zero) const {
  Tensor t = make_tensor(dt, s, device_);
  if (zero) t->zero();
  return t;
}
Tensor Layer::as_compute(const Tensor& x) const {
  if (!device_.is_gpu()) {
    if (x->is_gpu()) return x->to_host();
    return x->data_type() == DType_t::FP32 ? x : dtype_cast(*x, DType_t::FP32);
  }
  Tensor src = x;
  if (!x->is_gpu()) src = x->to_device(device_, workspace_);       // auto-migration of host inputs (layer.cpp:36-53)
  if (src->data_type() == DType_t::BF16) return src;
  if (src->data_type() != DType_t::FP32) return src;                // integer ids etc. pass through
  Tensor y = new_tensor(DType_t::BF16, src->shape());
  device_.set_current();
  cast_f32_to_bf16(src->data_as<float>(), y->data(), (long long)src->size(), stream());
  return y;
}

// ------------------------------------------------------------------------------------- factory
std::map<std::string, LayerFactory::Creator>& LayerFactory::table() {
  static std::map<std::string, Creator> t;
  return t;
}
void LayerFactory::register_type(const std::string& type, Creator c) { table()[type] = std::move(c); }
bool LayerFactory::has(const std::string& type) {
  register_defaults();
  return table().count(type) != 0;
}
Vec<std::string> LayerFactory::registered_types() {
  register_defaults();
  Vec<std::string> v;
  for (auto& kv : table()) v.push_back(kv.first);
  return v;
}
std::unique_ptr<Layer> LayerFactory::create(const TConfig& cfg) {
  register_defaults();
  auto it = table().find(cfg.type);
  if (it == table().end()) throw std::invalid_argument("LayerFactory: unknown layer type '" + cfg.type + "'");
  return it->second(cfg);
}

// ------------------------------------------------------------------------------------- step state
StepState::StepState(const Device& d) : device_(d) {
  IAllocator& ha = d.is_gpu() ? static_cast<IAllocator&>(DeviceAllocator::pinned()) : HostAllocator();
  ring_ = make_tensor(ha, DType_t::FP32, {RING, 16});
  std::memset(ring_->data(), 0, ring_->nbytes());
  host_ = ring_->data_as<float>();
  if (d.is_gpu()) {
    dev_ = make_tensor(DType_t::FP32, {16}, d);
    dev_->zero();
  }
}
StepState& StepState::get(const Device& d) {
  static std::mutex mu;
  static auto* all = new std::map<int, StepState*>();      // leaked: see DeviceManager::getInstance
  std::lock_guard<std::mutex> g(mu);
  auto& s = (*all)[d.is_gpu() ? d.id() : -1];
  if (!s) s = new StepState(d);
  return *s;
}
void StepState::advance(const Optimizer* opt) {
  ++counter_;
  float* h = ring_->data_as<float>() + (counter_ % RING) * 16;
  std::memcpy(h, host_, 64);                   // scalars persist from step to step unless refreshed
  host_ = h;
  std::memcpy(h + 12, &counter_, 8);
  if (opt) opt->write_step_state(h);
  if (dev_) {
    device_.set_current();
    TNN_CUDA(cudaMemcpyAsync(dev_->data(), h, 64, cudaMemcpyHostToDevice, current_flow(device_).stream()));
  }
}

// ------------------------------------------------------------------------------------- side flow
namespace side_flow {
static bool g_enabled = true;
static std::mutex g_mu;
static std::map<int, std::vector<Tensor>>& g_keep = *new std::map<int, std::vector<Tensor>>();
static std::map<int, bool>& g_pending = *new std::map<int, bool>();
bool enabled() { return g_enabled; }
void set_enabled(bool on) { g_enabled = on; }
void run(const Device& d, const std::function<void()>& fn, std::vector<Tensor> keep) {
  Flow& main = current_flow(d);
  Flow& side = DeviceManager::getInstance().getFlow(d, kHandle);
  side.wait(main);
  {
    FlowGuard g(d, kHandle);
    fn();
  }
  std::lock_guard<std::mutex> g(g_mu);
  auto& k = g_keep[d.id()];
  k.insert(k.end(), keep.begin(), keep.end());
  g_pending[d.id()] = true;
}
bool has_pending(const Device& d) {
  std::lock_guard<std::mutex> g(g_mu);
  return g_pending[d.id()];
}
void join(const Device& d) {
  {
    std::lock_guard<std::mutex> g(g_mu);
    if (!g_pending[d.id()]) return;
    g_pending[d.id()] = false;
  }
  current_flow(d).wait(DeviceManager::getInstance().getFlow(d, kHandle));
  std::vector<Tensor> drop;
  {
    std::lock_guard<std::mutex> g(g_mu);
    drop.swap(g_keep[d.id()]);
  }
}
}  // namespace side_flow

}  // namespace tnn
