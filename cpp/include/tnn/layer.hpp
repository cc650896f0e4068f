// Layer base, ParamDescriptor, LayerFactory (reference: include/nn/layer.hpp:23-144, src/nn/layer.cpp:25-230,
// include/nn/layers.hpp:95-167, block.hpp:6-71).
//
// Differences by design: parameters are fp32 masters that GraphContext re-homes into ONE slab (with a
// gradient slab and a bf16 compute-copy slab beside it); activations on the GPU are bf16; there is no stream
// synchronisation inside a pass; cross-layer fusion is an explicit plan (FwdAux / BwdAux + the flags a Block sets on
// its children at plan time), not something layers discover from tensor attributes or environment variables.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "tnn/config.hpp"
#include "tnn/profiler.hpp"
#include "tnn/tensor.hpp"

namespace tnn {

template <class T>
using Vec = std::vector<T>;

// What a producer may fold into its epilogue for its consumer (set by the enclosing Block's fusion plan).
struct FwdAux {
  Tensor residual;        // added to the output before the activation (ResidualBlock shortcut)
  int act = 0;            // 0 none, 1 relu, 2 gelu -- applied after the residual add
};
struct BwdAux {
  Tensor extra;           // in : gradient to ADD to this layer's input gradient (shortcut branch)
  Tensor colsum_in;       // in : per-feature column sums of dy, when the consumer's backward already produced them
  Tensor colsum_out;      // out: per-feature column sums of the returned dx (= bias gradient of the producer)
  bool need_dx = true;    // false: first layer of the model, skip the data gradient
};

struct ParamDescriptor {
  std::string name;
  DType_t dtype = DType_t::FP32;
  Shape shape;
  bool trainable = true;                                   // false: buffers (BatchNorm running statistics)
  std::function<void(TensorImpl&, uint64_t)> init;         // fills the fp32 tensor (host or device)
};

// Per-microbatch state a layer keeps between forward and backward (layer.cpp:169-188): tensors + scalars by name.
struct LayerCache {
  std::map<std::string, Tensor> t;
  std::map<std::string, double> s;
  std::map<std::string, uint64_t> u;
  std::map<std::string, Shape> sh;
  Tensor get(const std::string& k) const { auto it = t.find(k); return it == t.end() ? nullptr : it->second; }
};

class Layer {
 public:
  explicit Layer(std::string name) : name_(std::move(name)) {}
  virtual ~Layer() = default;

  virtual std::string type() const = 0;
  const std::string& name() const { return name_; }
  void set_name(const std::string& n) { name_ = n; }

  // ---- configuration (blocks forward every setter to their children)
  virtual void set_training(bool f) { training_ = f; for (Layer* c : children()) c->set_training(f); }
  bool is_training() const { return training_; }
  virtual void set_seed(uint64_t s) { seed_ = s; for (Layer* c : children()) c->set_seed(s); }
  uint64_t seed() const { return seed_; }
  // workspace allocator for activations (the executor's DELAllocatorV2); persistent state uses the pool
  virtual void set_allocator(IAllocator* a) { workspace_ = a; for (Layer* c : children()) c->set_allocator(a); }
  IAllocator* allocator() const { return workspace_; }
  virtual void set_io_dtype(DType_t d) { io_dtype_ = d; for (Layer* c : children()) c->set_io_dtype(d); }
  virtual void set_param_dtype(DType_t d) { param_dtype_ = d; for (Layer* c : children()) c->set_param_dtype(d); }
  virtual void set_compute_dtype(DType_t d) { compute_dtype_ = d; for (Layer* c : children()) c->set_compute_dtype(d); }
  DType_t io_dtype() const { return io_dtype_; }
  DType_t param_dtype() const { return param_dtype_; }
  DType_t compute_dtype() const { return compute_dtype_; }

  // ---- parameters
  virtual Vec<ParamDescriptor> param_descriptors() const { return {}; }
  // Creates and initialises this layer's own parameters on `dev` (and its children's).  GraphContext later
  // re-homes them with bind().
  virtual void init(const Device& dev, uint64_t seed = 0);
  bool initialized() const { return initialized_; }
  const Device& device() const { return device_; }
  Tensor& param(const std::string& n) { return params_.at(n); }
  Tensor& grad(const std::string& n) { return grads_.at(n); }
  bool has_param(const std::string& n) const { return params_.count(n) != 0; }
  // bf16 compute copy of a parameter (GPU); refreshed by the optimizer kernel or lazily here
  const Tensor& lp(const std::string& n);
  void bind(const std::string& n, Tensor data, Tensor grad, Tensor lp);
  void mark_lp_fresh() { lp_stale_ = false; ++param_version_; }
  void mark_params_dirty() { lp_stale_ = true; ++param_version_; }
  uint64_t param_version() const { return param_version_; }
  std::map<std::string, Tensor>& params() { return params_; }
  std::map<std::string, Tensor>& grads() { return grads_; }

  // ---- execution
  Tensor forward(const Tensor& x, size_t mb_id = 0) { return forward_impl(x, mb_id, FwdAux{}); }
  Tensor backward(const Tensor& dy, size_t mb_id = 0) {
    BwdAux aux;
    return backward_impl(dy, mb_id, aux);
  }
  // multi-input / multi-output form used by the Graph executor (op_node.hpp:24-29); SISO layers map to the above
  virtual Vec<Tensor> forward(const Vec<Tensor>& xs, size_t mb_id) { return {forward(xs.at(0), mb_id)}; }
  virtual Vec<Tensor> backward(const Vec<Tensor>& dys, size_t mb_id) { return {backward(dys.at(0), mb_id)}; }
  virtual Tensor forward_impl(const Tensor& x, size_t mb_id, const FwdAux& aux) = 0;
  virtual Tensor backward_impl(const Tensor& dy, size_t mb_id, BwdAux& aux) = 0;
  virtual bool fuses_fwd_residual() const { return false; }     // honours FwdAux::residual / act in its kernel
  virtual bool fuses_bwd_extra() const { return false; }        // honours BwdAux::extra in its kernel

  virtual Shape output_shape(const Shape& in) const { return in; }
  virtual uint64_t forward_flops(const Shape&) const { return 0; }
  virtual TConfig get_config() const;                           // {name, type, parameters: config_params()}
  virtual Vec<Layer*> children() { return {}; }
  Vec<Layer*> all_layers();                                     // depth-first, leaves and blocks
  void clear_cache(size_t mb_id) { cache_.erase(mb_id); }
  void clear_caches() { cache_.clear(); }

 protected:
  virtual void fill_config(TConfig&) const {}
  LayerCache& cache(size_t mb_id) { return cache_[mb_id]; }
  // activation buffer from the workspace allocator (falls back to the device pool outside an executor)
  Tensor new_tensor(DType_t dt, const Shape& s) const;
  // persistent per-layer device state (statistics workspaces, prepared weights): never from the step workspace
  Tensor new_persistent(DType_t dt, const Shape& s, bool zero = false) const;
  DType_t act_dtype() const { return device_.is_gpu() ? DType_t::BF16 : DType_t::FP32; }
  // bf16 view of an incoming activation on the GPU (fp32 inputs are cast by the hand-written kernel)
  Tensor as_compute(const Tensor& x) const;
  cudaStream_t stream() const { return current_flow(device_).stream(); }

  std::string name_;
  bool training_ = true;
  uint64_t seed_ = 0;
  Device device_ = Device(DeviceType::CPU, 0);
  bool initialized_ = false;
  IAllocator* workspace_ = nullptr;
  DType_t io_dtype_ = DType_t::FP32, param_dtype_ = DType_t::FP32, compute_dtype_ = DType_t::FP32;
  std::map<std::string, Tensor> params_, grads_, lp_;
  bool lp_stale_ = true;
  uint64_t param_version_ = 0;
  std::unordered_map<size_t, LayerCache> cache_;
};

// A layer that owns children (block.hpp:6-71)
class Block : public Layer {
 public:
  using Layer::Layer;
  void init(const Device& dev, uint64_t seed = 0) override;
};

class LayerFactory {
 public:
  using Creator = std::function<std::unique_ptr<Layer>(const TConfig&)>;
  static void register_type(const std::string& type, Creator c);
  static void register_defaults();                              // idempotent
  static std::unique_ptr<Layer> create(const TConfig& cfg);
  static std::unique_ptr<Layer> create(const std::string& json) { return create(TConfig::from_string(json)); }
  static Vec<std::string> registered_types();
  static bool has(const std::string& type);

 private:
  static std::map<std::string, Creator>& table();
};

// fills `t` (fp32, host or device) uniformly in [-b, b] / with a constant / N(0, std)
std::function<void(TensorImpl&, uint64_t)> uniform_init(double bound);
std::function<void(TensorImpl&, uint64_t)> const_init(double v);
std::function<void(TensorImpl&, uint64_t)> normal_init(double stddev);

// Scope that charges its wall time to a layer in the LayerProfiler (one relaxed load when profiling is off).
struct LayerTimer {
  LayerTimer(const std::string& layer, const Device& dev, bool backward) : active_(LayerProfiler::instance().enabled()) {
    if (!active_) return;
    name_ = &layer;
    dev_ = dev;
    backward_ = backward;
    if (dev_.is_gpu()) current_flow(dev_).synchronize();             // work queued by the previous layer is not this layer's time
    t0_ = std::chrono::steady_clock::now();
  }
  ~LayerTimer() {
    if (!active_) return;
    try {
      if (dev_.is_gpu()) current_flow(dev_).synchronize();
    } catch (...) {                                                   // a failed launch surfaces at the caller's next CUDA call
    }
    LayerProfiler::instance().add(*name_, backward_, std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0_).count());
  }
  LayerTimer(const LayerTimer&) = delete;
  LayerTimer& operator=(const LayerTimer&) = delete;

 private:
  bool active_;
  bool backward_ = false;
  const std::string* name_ = nullptr;
  Device dev_;
  std::chrono::steady_clock::time_point t0_;
};

}  // namespace tnn
