// Concrete layers and blocks (reference: include/nn/layers_impl/*.hpp, include/nn/blocks_impl/*.hpp,
// include/nn/activations.hpp:27-58).  TYPE_NAMEs and config keys are the reference's (include/nn/layers.hpp:125-158);
// NHWC / feature-last layout; `legacy_*` names are registered as aliases that present NCHW tensors.
//
// CPU path = plain fp32 loops + the packed SGEMM (the oracle);  GPU path = the hand-written sm_100a kernels in
// tnn_b200/csrc/kernels (tcgen05 GEMM / implicit-GEMM conv / attention, SIMT norms, pools, activations).
#pragma once
#include "tnn/layer.hpp"

namespace tnn {

// =============================================================================== dense
class DenseLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "dense";
  DenseLayer(size_t in, size_t out, bool use_bias = true, const std::string& name = "dense")
      : Layer(name), in_(in), out_(out), use_bias_(use_bias) {}
  std::string type() const override { return TYPE_NAME; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  bool fuses_fwd_residual() const override { return true; }
  Shape output_shape(const Shape& s) const override { Shape o = s; o.back() = out_; return o; }
  uint64_t forward_flops(const Shape& s) const override { return 2ull * (shape_numel(s) / s.back()) * in_ * out_; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  size_t input_features() const { return in_; }
  size_t output_features() const { return out_; }
  bool use_bias() const { return use_bias_; }

 protected:
  void fill_config(TConfig& c) const override;
  size_t in_, out_;
  bool use_bias_;
};

// =============================================================================== conv2d (NHWC, weights [K,R,S,C])
struct ConvPlan;   // geometry + tap lists of one input shape (layers_conv.cpp)
class Conv2DLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "conv2d";
  Conv2DLayer(size_t in_ch, size_t out_ch, size_t kh, size_t kw, size_t sh = 1, size_t sw = 1, size_t ph = 0, size_t pw = 0,
              bool use_bias = true, const std::string& name = "conv2d");
  ~Conv2DLayer() override;
  std::string type() const override { return TYPE_NAME; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  bool fuses_fwd_residual() const override { return true; }
  Shape output_shape(const Shape& s) const override;
  uint64_t forward_flops(const Shape& s) const override;
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  size_t in_channels() const { return C_; }
  size_t out_channels() const { return K_; }
  // the flipped dgrad weights, for the trainer's batched refresh (one launch for every convolution after the optimizer
  // step instead of one launch per layer inside backward): null until the first backward allocated them
  Tensor& flipped_weights() { return wt_; }
  void mark_flipped_fresh() { wt_version_ = param_version_; }
  size_t kernel_h() const { return R_; }
  size_t kernel_w() const { return S_; }

 protected:
  void fill_config(TConfig& c) const override;
  const ConvPlan& plan_for(const Shape& x);
  size_t C_, K_, R_, S_, sh_, sw_, ph_, pw_;
  bool use_bias_;
  std::map<Shape, std::unique_ptr<ConvPlan>> plans_;
  Tensor wt_;                    // bf16 [C,R,S,K], (r,s) flipped: dgrad operand, refreshed when the weights change
  uint64_t wt_version_ = ~0ull;
};

// =============================================================================== batchnorm (channel-last)
class BatchNormLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "batchnorm";
  BatchNormLayer(size_t features, double eps = 1e-5, double momentum = 0.1, bool affine = true, bool use_relu = false,
                 const std::string& name = "batchnorm")
      : Layer(name), C_(features), eps_(eps), momentum_(momentum), affine_(affine), relu_(use_relu) {}
  std::string type() const override { return TYPE_NAME; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  bool fuses_bwd_extra() const override { return true; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  // fusion plan (set by Sequential): a Dropout that directly follows is applied inside this layer's apply kernel
  double fused_dropout_p = 0.0;
  size_t num_features() const { return C_; }
  bool use_relu() const { return relu_; }
  bool affine() const { return affine_; }
  double epsilon() const { return eps_; }
  double momentum() const { return momentum_; }

 protected:
  void fill_config(TConfig& c) const override;
  void ensure_buffers();
  size_t C_;
  double eps_, momentum_;
  bool affine_, relu_;
  Tensor ws_, cws_, ss_, coef_;      // persistent fp32 workspaces of the GPU kernels
  uint64_t cpu_step_ = 0;
};

// =============================================================================== pooling (NHWC)
class Pool2DLayer : public Layer {
 public:
  Pool2DLayer(bool is_max, size_t ph, size_t pw, size_t sh, size_t sw, size_t padh, size_t padw, const std::string& name)
      : Layer(name), max_(is_max), kh_(ph), kw_(pw), sh_(sh ? sh : ph), sw_(sw ? sw : pw), ph_(padh), pw_(padw) {}
  std::string type() const override { return max_ ? "maxpool2d" : "avgpool2d"; }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override;
  static std::unique_ptr<Layer> create_max(const TConfig& c);
  static std::unique_ptr<Layer> create_avg(const TConfig& c);

 protected:
  void fill_config(TConfig& c) const override;
  bool max_;
  size_t kh_, kw_, sh_, sw_, ph_, pw_;
};
class MaxPool2DLayer : public Pool2DLayer {
 public:
  static constexpr const char* TYPE_NAME = "maxpool2d";
  MaxPool2DLayer(size_t ph, size_t pw, size_t sh = 0, size_t sw = 0, size_t padh = 0, size_t padw = 0, const std::string& name = "maxpool2d")
      : Pool2DLayer(true, ph, pw, sh, sw, padh, padw, name) {}
};
class AvgPool2DLayer : public Pool2DLayer {
 public:
  static constexpr const char* TYPE_NAME = "avgpool2d";
  AvgPool2DLayer(size_t ph, size_t pw, size_t sh = 0, size_t sw = 0, size_t padh = 0, size_t padw = 0, const std::string& name = "avgpool2d")
      : Pool2DLayer(false, ph, pw, sh, sw, padh, padw, name) {}
};

// =============================================================================== dropout / flatten / activation / identity
class DropoutLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "dropout";
  explicit DropoutLayer(double rate, const std::string& name = "dropout") : Layer(name), p_(rate) {}
  std::string type() const override { return TYPE_NAME; }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  bool fuses_fwd_residual() const override { return !fused_into_prev; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  double rate() const { return p_; }
  bool fused_into_prev = false;     // fusion plan: the preceding BatchNorm applies this dropout

 protected:
  void fill_config(TConfig& c) const override { c.set("dropout_rate", p_); }
  uint64_t layer_seed(size_t mb) const;
  double p_;
  uint64_t cpu_step_ = 0;
};

class FlattenLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "flatten";
  FlattenLayer(int start_dim = 1, int end_dim = -1, const std::string& name = "flatten") : Layer(name), start_(start_dim), end_(end_dim) {}
  std::string type() const override { return TYPE_NAME; }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override;
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);

 protected:
  void fill_config(TConfig& c) const override { c.set("start_dim", start_).set("end_dim", end_); }
  int start_, end_;
};

// activation ∈ none | linear | relu | leaky_relu | sigmoid | tanh | elu | gelu | softmax  (activations.hpp:47-54)
class ActivationLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "activation";
  explicit ActivationLayer(const std::string& activation, const std::string& name = "activation");
  std::string type() const override { return TYPE_NAME; }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  const std::string& activation() const { return act_; }

 protected:
  void fill_config(TConfig& c) const override { c.set("activation", act_); }
  std::string act_;
  int kind_;        // ActKind of simt_ops.cuh; -1 = softmax over the last dim
  float slope_ = 0.01f;
};

class IdentityLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "identity";
  explicit IdentityLayer(const std::string& name = "identity") : Layer(name) {}
  std::string type() const override { return TYPE_NAME; }
  Tensor forward_impl(const Tensor& x, size_t, const FwdAux&) override { return x; }
  Tensor backward_impl(const Tensor& dy, size_t, BwdAux&) override { return dy; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c) { return std::make_unique<IdentityLayer>(c.name.empty() ? "identity" : c.name); }
};

// =============================================================================== normalisation over the last dim / groups
class LayerNormLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "layer_norm";
  LayerNormLayer(size_t normalized_shape, double eps = 1e-5, bool affine = true, const std::string& name = "layer_norm")
      : Layer(name), C_(normalized_shape), eps_(eps), affine_(affine) {}
  std::string type() const override { return TYPE_NAME; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  bool fuses_bwd_extra() const override { return true; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  size_t normalized_shape() const { return C_; }
  double epsilon() const { return eps_; }
  bool affine() const { return affine_; }

 protected:
  void fill_config(TConfig& c) const override { c.set("normalized_shape", C_).set("epsilon", eps_).set("affine", affine_); }
  size_t C_;
  double eps_;
  bool affine_;
};

class GroupNormLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "groupnorm";
  GroupNormLayer(size_t groups, size_t channels, double eps = 1e-5, bool affine = true, const std::string& name = "groupnorm")
      : Layer(name), G_(groups), C_(channels), eps_(eps), affine_(affine) {}
  std::string type() const override { return TYPE_NAME; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);

 protected:
  void fill_config(TConfig& c) const override { c.set("num_groups", G_).set("num_channels", C_).set("epsilon", eps_).set("affine", affine_); }
  size_t G_, C_;
  double eps_;
  bool affine_;
};

// =============================================================================== embedding family
class EmbeddingLayer : public Layer {       // ids: INT32 (or float ids as the reference stores them) [B, L] -> [B, L, D]
 public:
  static constexpr const char* TYPE_NAME = "embedding";
  EmbeddingLayer(size_t vocab, size_t dim, long padding_idx = -1, const std::string& name = "embedding")
      : Layer(name), V_(vocab), D_(dim), pad_(padding_idx) {}
  std::string type() const override { return TYPE_NAME; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override { Shape o = s; o.push_back(D_); return o; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  size_t vocab_size() const { return V_; }
  size_t embed_dim() const { return D_; }

 protected:
  void fill_config(TConfig& c) const override { c.set("vocab_size", V_).set("embed_dim", D_).set("padding_idx", (long)pad_); }
  size_t V_, D_;
  long pad_;
};

class PositionalEmbeddingLayer : public Layer {   // x [B, L, D] + learned table [seq_len, D]
 public:
  static constexpr const char* TYPE_NAME = "pos_embedding";
  PositionalEmbeddingLayer(size_t dim, size_t seq_len, const std::string& name = "pos_embedding") : Layer(name), D_(dim), L_(seq_len) {}
  std::string type() const override { return TYPE_NAME; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  size_t embed_dim() const { return D_; }
  size_t seq_len() const { return L_; }

 protected:
  void fill_config(TConfig& c) const override { c.set("embed_dim", D_).set("seq_len", L_); }
  size_t D_, L_;
};

class ClassTokenLayer : public Layer {            // [B, L, D] -> [B, L+1, D] (token first)
 public:
  static constexpr const char* TYPE_NAME = "class_token";
  explicit ClassTokenLayer(size_t dim, const std::string& name = "class_token") : Layer(name), D_(dim) {}
  std::string type() const override { return TYPE_NAME; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override { Shape o = s; o[1] += 1; return o; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);

 protected:
  void fill_config(TConfig& c) const override { c.set("embed_dim", D_); }
  size_t D_;
};

// =============================================================================== shape ops
class SliceLayer : public Layer {                 // x.narrow(axis, start, length)
 public:
  static constexpr const char* TYPE_NAME = "slice";
  SliceLayer(size_t axis, size_t start, size_t length, const std::string& name = "slice") : Layer(name), axis_(axis), start_(start), len_(length) {}
  std::string type() const override { return TYPE_NAME; }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override { Shape o = s; o[axis_] = len_; return o; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);

 protected:
  void fill_config(TConfig& c) const override { c.set("axis", axis_).set("start", start_).set("length", len_); }
  size_t axis_, start_, len_;
};

class TransposeLayer : public Layer {             // swaps the last two dims
 public:
  static constexpr const char* TYPE_NAME = "transpose";
  explicit TransposeLayer(const std::string& name = "transpose") : Layer(name) {}
  std::string type() const override { return TYPE_NAME; }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override { Shape o = s; std::swap(o[o.size() - 1], o[o.size() - 2]); return o; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c) { return std::make_unique<TransposeLayer>(c.name.empty() ? "transpose" : c.name); }
};

// N-ary joins (n_ary_layer): add / sub / mul over N same-shaped inputs
class NAryLayer : public Layer {
 public:
  NAryLayer(int op, const std::string& name) : Layer(name), op_(op) {}
  std::string type() const override { return op_ == 0 ? "add" : op_ == 1 ? "sub" : "mul"; }
  Vec<Tensor> forward(const Vec<Tensor>& xs, size_t mb) override;
  Vec<Tensor> backward(const Vec<Tensor>& dys, size_t mb) override;
  Tensor forward_impl(const Tensor& x, size_t, const FwdAux&) override { return x; }
  Tensor backward_impl(const Tensor& dy, size_t, BwdAux&) override { return dy; }
  static std::unique_ptr<Layer> create_add(const TConfig& c) { return std::make_unique<NAryLayer>(0, c.name.empty() ? "add" : c.name); }
  static std::unique_ptr<Layer> create_sub(const TConfig& c) { return std::make_unique<NAryLayer>(1, c.name.empty() ? "sub" : c.name); }
  static std::unique_ptr<Layer> create_mul(const TConfig& c) { return std::make_unique<NAryLayer>(2, c.name.empty() ? "mul" : c.name); }

 protected:
  int op_;
};

// 1 input -> M references to it (fan-out feeding MSequential branches); backward sums the M gradients
class MBroadcastLayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "mbroadcast";
  explicit MBroadcastLayer(size_t m, const std::string& name = "mbroadcast") : Layer(name), m_(m) {}
  std::string type() const override { return TYPE_NAME; }
  Vec<Tensor> forward(const Vec<Tensor>& xs, size_t mb) override;
  Vec<Tensor> backward(const Vec<Tensor>& dys, size_t mb) override;
  Tensor forward_impl(const Tensor& x, size_t, const FwdAux&) override { return x; }
  Tensor backward_impl(const Tensor& dy, size_t, BwdAux&) override { return dy; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c) { return std::make_unique<MBroadcastLayer>(c.get<size_t>("m", 2), c.name.empty() ? "mbroadcast" : c.name); }

 protected:
  void fill_config(TConfig& c) const override { c.set("m", m_); }
  size_t m_;
};

// =============================================================================== blocks
class Sequential : public Block {
 public:
  static constexpr const char* TYPE_NAME = "sequential";
  explicit Sequential(Vec<std::unique_ptr<Layer>> layers = {}, const std::string& name = "sequential");
  std::string type() const override { return TYPE_NAME; }
  void add(std::unique_ptr<Layer> l);
  Vec<Layer*> children() override;
  Vec<Layer*> layers() { return children(); }
  size_t size() const { return layers_.size(); }
  Layer* at(size_t i) { return layers_.at(i).get(); }
  Vec<std::unique_ptr<Layer>> release_layers() { return std::move(layers_); }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  bool fuses_fwd_residual() const override { return !layers_.empty() && layers_.back()->fuses_fwd_residual(); }
  bool fuses_bwd_extra() const override { return !layers_.empty() && layers_.front()->fuses_bwd_extra(); }
  Shape output_shape(const Shape& s) const override;
  uint64_t forward_flops(const Shape& s) const override;
  TConfig get_config() const override;
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  // optional per-layer backward-completion callback (data-parallel bucket overlap)
  std::function<void(size_t)> backward_hook;
  void print_summary(const Shape& input_shape) const;

 private:
  void plan_fusion();           // the explicit fusion plan: BatchNorm+Dropout pairs
  Vec<std::unique_ptr<Layer>> layers_;
};

// M parallel Sequential branches joined by an N-ary layer; branches run in descending (peak - retained) workspace
// order measured on the first call (msequential.cpp:40-163)
class MSequential : public Block {
 public:
  static constexpr const char* TYPE_NAME = "msequential";
  MSequential(Vec<std::unique_ptr<Layer>> sequences, std::unique_ptr<Layer> join, const std::string& name = "msequential");
  std::string type() const override { return TYPE_NAME; }
  Vec<Layer*> children() override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override { return seqs_.front()->output_shape(s); }
  TConfig get_config() const override;
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  const Vec<size_t>& order() const { return order_; }

 private:
  Vec<std::unique_ptr<Layer>> seqs_;
  std::unique_ptr<Layer> join_;
  Vec<size_t> order_;
};

// act(F(x) + shortcut(x)) (residual_block.cpp:63-142)
class ResidualBlock : public Block {
 public:
  static constexpr const char* TYPE_NAME = "residual_block";
  ResidualBlock(std::unique_ptr<Layer> main_path, std::unique_ptr<Layer> shortcut_path, const std::string& activation = "relu",
                const std::string& name = "residual_block");
  std::string type() const override { return TYPE_NAME; }
  Vec<Layer*> children() override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override { return main_->output_shape(s); }
  uint64_t forward_flops(const Shape& s) const override { return main_->forward_flops(s) + (shortcut_ ? shortcut_->forward_flops(s) : 0); }
  TConfig get_config() const override;
  static std::unique_ptr<Layer> create_from_config(const TConfig& c);
  Layer* main_path() { return main_.get(); }
  Layer* shortcut_path() { return shortcut_.get(); }
  const std::string& activation() const { return act_; }

 private:
  std::unique_ptr<Layer> main_, shortcut_;
  std::string act_;
  std::unique_ptr<ActivationLayer> act_layer_;
};

// Multi-head self-attention [B, L, E] -> [B, L, E]: one fused QKV GEMM, fused flash attention, output projection
// (attention_block.cpp:42-297, flash_attention_block.cpp:59-260).  head_dim 64 on the GPU.
class AttentionBlock : public Block {
 public:
  AttentionBlock(size_t embed_dim, size_t num_heads, bool causal = true, const std::string& name = "attention_block", bool flash = false)
      : Block(name), E_(embed_dim), H_(num_heads), causal_(causal), flash_(flash) {}
  std::string type() const override { return flash_ ? "flash_attention_block" : "attention_block"; }
  Vec<ParamDescriptor> param_descriptors() const override;
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  bool fuses_fwd_residual() const override { return true; }
  static std::unique_ptr<Layer> create_attention(const TConfig& c);
  static std::unique_ptr<Layer> create_flash(const TConfig& c);
  size_t embed_dim() const { return E_; }
  size_t num_heads() const { return H_; }
  bool is_causal() const { return causal_; }

 protected:
  void fill_config(TConfig& c) const override { c.set("embed_dim", E_).set("num_heads", H_).set("is_causal", causal_); }
  size_t E_, H_;
  bool causal_, flash_;
};

// scaled dot-product attention over pre-projected [B, L, 3, H, D] (sdpa_layer.cpp)
class SDPALayer : public Layer {
 public:
  static constexpr const char* TYPE_NAME = "sdpa";
  explicit SDPALayer(bool causal = false, const std::string& name = "sdpa") : Layer(name), causal_(causal) {}
  std::string type() const override { return TYPE_NAME; }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override { return {s[0], s[1], s[3] * s[4]}; }
  static std::unique_ptr<Layer> create_from_config(const TConfig& c) { return std::make_unique<SDPALayer>(c.get<bool>("is_causal", false), c.name.empty() ? "sdpa" : c.name); }

 protected:
  void fill_config(TConfig& c) const override { c.set("is_causal", causal_); }
  bool causal_;
};

// NCHW facade: presents an NHWC layer to NCHW tensors (legacy_conv2d / legacy_batchnorm / legacy_maxpool2d /
// legacy_avgpool2d; legacy_dense is DenseLayer itself).  Layout conversion runs in the tiled transpose kernels --
// except legacy_batchnorm on the GPU, which is normalised in its own NCHW layout by kernels/norm_nchw.cu.
class LegacyNCHWLayer : public Block {
 public:
  LegacyNCHWLayer(std::unique_ptr<Layer> inner, const std::string& legacy_type);
  std::string type() const override { return legacy_type_; }
  Vec<Layer*> children() override { return {inner_.get()}; }
  Tensor forward_impl(const Tensor& x, size_t mb, const FwdAux& aux) override;
  Tensor backward_impl(const Tensor& dy, size_t mb, BwdAux& aux) override;
  Shape output_shape(const Shape& s) const override;
  TConfig get_config() const override;

 private:
  Tensor bn_forward_nchw(BatchNormLayer& bn, const Tensor& x, size_t mb);
  Tensor bn_backward_nchw(BatchNormLayer& bn, const Tensor& dy, size_t mb);
  std::unique_ptr<Layer> inner_;
  std::string legacy_type_;
  Tensor nchw_ws_, nchw_ss_, nchw_coef_;   // persistent workspaces of the native NCHW BatchNorm kernels
};

// ---- helpers shared by layer implementations and tests
Tensor nchw_to_nhwc_tensor(const Tensor& x, IAllocator* alloc = nullptr);
Tensor nhwc_to_nchw_tensor(const Tensor& x, IAllocator* alloc = nullptr);

}  // namespace tnn
