// LayerBuilder: fluent model-construction DSL with shape inference (reference: include/nn/layer_builder.hpp:11-624).
// Shapes are batchless ({H, W, C} / {L} / {L, E}); a batch dimension of 1 is prepended internally, as the reference does.
#pragma once
#include <string>

#include "tnn/layers.hpp"

namespace tnn {

class LayerBuilder {
 public:
  explicit LayerBuilder(const Shape& batchless_input_shape) : shape_{1} { shape_.insert(shape_.end(), batchless_input_shape.begin(), batchless_input_shape.end()); }
  LayerBuilder& dtype(DType_t io) { io_ = io; return *this; }
  Shape get_current_shape() const { return shape_; }
  Shape get_batchless_current_shape() const { return Shape(shape_.begin() + 1, shape_.end()); }
  Vec<std::unique_ptr<Layer>> build() { return std::move(layers_); }
  std::unique_ptr<Sequential> build_sequential(const std::string& name) { return std::make_unique<Sequential>(std::move(layers_), name); }
  LayerBuilder& add_layer(std::unique_ptr<Layer> l) { return push(std::move(l)); }

  LayerBuilder& dense(size_t out, bool bias = true, const std::string& name = "") { return push(std::make_unique<DenseLayer>(shape_.back(), out, bias, nm(name, "dense"))); }
  LayerBuilder& conv2d(size_t out_ch, size_t kh, size_t kw, size_t sh = 1, size_t sw = 1, size_t ph = 0, size_t pw = 0, bool bias = true, const std::string& name = "") {
    return push(std::make_unique<Conv2DLayer>(shape_.back(), out_ch, kh, kw, sh, sw, ph, pw, bias, nm(name, "conv2d")));
  }
  // epsilon < 0: the io dtype's epsilon, as the reference model zoo passes (example_models.cpp:25)
  LayerBuilder& batchnorm(double eps = -1, double momentum = 0.1, bool affine = true, bool use_relu = false, const std::string& name = "") {
    return push(std::make_unique<BatchNormLayer>(shape_.back(), eps < 0 ? dtype_eps(io_) : eps, momentum, affine, use_relu, nm(name, "batchnorm")));
  }
  LayerBuilder& groupnorm(size_t groups, double eps = 1e-5, bool affine = true, const std::string& name = "") { return push(std::make_unique<GroupNormLayer>(groups, shape_.back(), eps, affine, nm(name, "groupnorm"))); }
  LayerBuilder& layernorm(double eps = 1e-5, bool affine = true, const std::string& name = "") { return push(std::make_unique<LayerNormLayer>(shape_.back(), eps, affine, nm(name, "layer_norm"))); }
  LayerBuilder& maxpool2d(size_t ph, size_t pw, size_t sh = 0, size_t sw = 0, size_t padh = 0, size_t padw = 0, const std::string& name = "") { return push(std::make_unique<MaxPool2DLayer>(ph, pw, sh, sw, padh, padw, nm(name, "maxpool2d"))); }
  LayerBuilder& avgpool2d(size_t ph, size_t pw, size_t sh = 0, size_t sw = 0, size_t padh = 0, size_t padw = 0, const std::string& name = "") { return push(std::make_unique<AvgPool2DLayer>(ph, pw, sh, sw, padh, padw, nm(name, "avgpool2d"))); }
  LayerBuilder& activation(const std::string& act, const std::string& name = "") { return push(std::make_unique<ActivationLayer>(act, nm(name, act.c_str()))); }
  LayerBuilder& dropout(double rate, const std::string& name = "") { return push(std::make_unique<DropoutLayer>(rate, nm(name, "dropout"))); }
  LayerBuilder& flatten(int start = 1, int end = -1, const std::string& name = "") { return push(std::make_unique<FlattenLayer>(start, end, nm(name, "flatten"))); }
  LayerBuilder& identity(const std::string& name = "") { return push(std::make_unique<IdentityLayer>(nm(name, "identity"))); }
  LayerBuilder& class_token(size_t dim = 0, const std::string& name = "") { return push(std::make_unique<ClassTokenLayer>(dim ? dim : shape_.back(), nm(name, "class_token"))); }
  LayerBuilder& positional_embedding(size_t dim = 0, size_t seq_len = 0, const std::string& name = "") {
    return push(std::make_unique<PositionalEmbeddingLayer>(dim ? dim : shape_.back(), seq_len ? seq_len : shape_[shape_.size() - 2], nm(name, "pos_embedding")));
  }
  LayerBuilder& embedding(size_t vocab, size_t dim, long padding_idx = -1, const std::string& name = "") { return push(std::make_unique<EmbeddingLayer>(vocab, dim, padding_idx, nm(name, "embedding"))); }
  LayerBuilder& slice(size_t axis, size_t start, size_t length, const std::string& name = "") { return push(std::make_unique<SliceLayer>(axis, start, length, nm(name, "slice"))); }
  LayerBuilder& transpose(const std::string& name = "") { return push(std::make_unique<TransposeLayer>(nm(name, "transpose"))); }
  LayerBuilder& attention(size_t embed, size_t heads, bool causal = true, const std::string& name = "") { return push(std::make_unique<AttentionBlock>(embed, heads, causal, nm(name, "attention"), false)); }
  LayerBuilder& flash_attention(size_t embed, size_t heads, bool causal = true, const std::string& name = "") { return push(std::make_unique<AttentionBlock>(embed, heads, causal, nm(name, "flash_attention"), true)); }
  LayerBuilder& broadcast_m_seq(Vec<std::unique_ptr<Layer>> sequences, std::unique_ptr<Layer> join, const std::string& name = "") {
    return push(std::make_unique<MSequential>(std::move(sequences), std::move(join), nm(name, "msequential")));
  }
  // legacy NCHW twins: the current shape is {1, C, H, W}
  LayerBuilder& legacy_dense(size_t out, bool bias = true, const std::string& name = "") { return dense(out, bias, name.empty() ? nm(name, "legacy_dense") : name); }
  LayerBuilder& legacy_conv2d(size_t out_ch, size_t kh, size_t kw, size_t sh = 1, size_t sw = 1, size_t ph = 0, size_t pw = 0, bool bias = true, const std::string& name = "") {
    return push(std::make_unique<LegacyNCHWLayer>(std::make_unique<Conv2DLayer>(shape_[1], out_ch, kh, kw, sh, sw, ph, pw, bias, nm(name, "legacy_conv2d")), "legacy_conv2d"));
  }
  LayerBuilder& legacy_batchnorm(double eps = 1e-5, double momentum = 0.1, bool affine = true, bool use_relu = false, const std::string& name = "") {
    return push(std::make_unique<LegacyNCHWLayer>(std::make_unique<BatchNormLayer>(shape_[1], eps, momentum, affine, use_relu, nm(name, "legacy_batchnorm")), "legacy_batchnorm"));
  }
  LayerBuilder& legacy_maxpool2d(size_t ph, size_t pw, size_t sh = 0, size_t sw = 0, size_t padh = 0, size_t padw = 0, const std::string& name = "") {
    return push(std::make_unique<LegacyNCHWLayer>(std::make_unique<MaxPool2DLayer>(ph, pw, sh, sw, padh, padw, nm(name, "legacy_maxpool2d")), "legacy_maxpool2d"));
  }
  LayerBuilder& legacy_avgpool2d(size_t ph, size_t pw, size_t sh = 0, size_t sw = 0, size_t padh = 0, size_t padw = 0, const std::string& name = "") {
    return push(std::make_unique<LegacyNCHWLayer>(std::make_unique<AvgPool2DLayer>(ph, pw, sh, sw, padh, padw, nm(name, "legacy_avgpool2d")), "legacy_avgpool2d"));
  }

  // ---- composite blocks (layer_builder.hpp:339-403, 408, 531-575)
  LayerBuilder& residual_block(Vec<std::unique_ptr<Layer>> main_layers, Vec<std::unique_ptr<Layer>> shortcut_layers, const std::string& act = "relu", const std::string& name = "") {
    const std::string n = nm(name, "residual_block");
    std::unique_ptr<Layer> main = std::make_unique<Sequential>(std::move(main_layers), n + "_main_path");
    std::unique_ptr<Layer> sc;
    if (!shortcut_layers.empty()) sc = std::make_unique<Sequential>(std::move(shortcut_layers), n + "_shortcut_path");
    return push(std::make_unique<ResidualBlock>(std::move(main), std::move(sc), act, n));
  }
  LayerBuilder& basic_residual_block(size_t in, size_t out, size_t stride = 1, const std::string& name = "basic_residual_block") {
    LayerBuilder m = sub();
    m.conv2d(out, 3, 3, stride, stride, 1, 1, false, name + "_conv1").batchnorm(-1, 0.1, true, true, name + "_bn0")
        .conv2d(out, 3, 3, 1, 1, 1, 1, false, name + "_conv2").batchnorm(-1, 0.1, true, false, name + "_bn1");
    Vec<std::unique_ptr<Layer>> sc;
    if (stride != 1 || in != out) {
      LayerBuilder s = sub();
      s.conv2d(out, 1, 1, stride, stride, 0, 0, false, name + "_conv0").batchnorm(-1, 0.1, true, false, name + "_bn0s");
      sc = s.build();
    }
    return residual_block(m.build(), std::move(sc), "relu", name);
  }
  // pre-activation WRN block: BN-ReLU-conv3x3(stride)-BN-ReLU-dropout-conv3x3, 1x1 shortcut when the shape changes, linear after the add
  LayerBuilder& wide_residual_block(size_t in, size_t out, size_t stride = 1, double dropout_rate = 0.0, const std::string& name = "wide_residual_block") {
    LayerBuilder m = sub();
    m.batchnorm(-1, 0.1, true, true, name + "_bn1").conv2d(out, 3, 3, stride, stride, 1, 1, true, name + "_conv1").batchnorm(-1, 0.1, true, true, name + "_bn2");
    if (dropout_rate > 0) m.dropout(dropout_rate, name + "_dropout");
    m.conv2d(out, 3, 3, 1, 1, 1, 1, true, name + "_conv2");
    Vec<std::unique_ptr<Layer>> sc;
    if (stride != 1 || in != out) {
      LayerBuilder s = sub();
      s.conv2d(out, 1, 1, stride, stride, 0, 0, false, name + "_shortcut");
      sc = s.build();
    }
    return residual_block(m.build(), std::move(sc), "linear", name);
  }
  LayerBuilder& bottleneck_residual_block(size_t in, size_t mid, size_t out, size_t stride = 1, const std::string& name = "bottleneck_residual_block") {
    LayerBuilder m = sub();
    m.conv2d(mid, 1, 1, 1, 1, 0, 0, false, name + "_conv0").batchnorm(-1, 0.1, true, true, name + "_bn0")
        .conv2d(mid, 3, 3, stride, stride, 1, 1, false, name + "_conv1").batchnorm(-1, 0.1, true, true, name + "_bn1")
        .conv2d(out, 1, 1, 1, 1, 0, 0, false, name + "_conv2").batchnorm(-1, 0.1, true, true, name + "_bn2");
    Vec<std::unique_ptr<Layer>> sc;
    if (stride != 1 || in != out) {
      LayerBuilder s = sub();
      s.conv2d(out, 1, 1, stride, stride, 0, 0, false, name + "_conv3").batchnorm(-1, 0.1, true, false, name + "_bn3");
      sc = s.build();
    }
    return residual_block(m.build(), std::move(sc), "none", name);
  }
  LayerBuilder& gpt_block(size_t embed, size_t heads, size_t ffn, double dropout_rate = 0.1, bool causal = false, const std::string& act = "gelu",
                          const std::string& name = "", bool flash = false) {
    const std::string n = name.empty() ? "gpt_block_" + std::to_string(layers_.size()) : name;
    LayerBuilder a = sub();
    a.layernorm(1e-5, true, n + "_ln_1");
    if (flash) a.flash_attention(embed, heads, causal, n + "_attn");
    else a.attention(embed, heads, causal, n + "_attn");
    if (dropout_rate > 0) a.dropout(dropout_rate, n + "_attn_dropout");
    residual_block(a.build(), {}, "linear", n + "_attn_res");
    LayerBuilder f = sub();
    f.layernorm(1e-5, true, n + "_ln_2").dense(ffn, true, n + "_mlp_fc1").activation(act, n + "_mlp_act").dense(embed, true, n + "_mlp_fc2");
    if (dropout_rate > 0) f.dropout(dropout_rate, n + "_mlp_dropout");
    return residual_block(f.build(), {}, "linear", n + "_ffn_res");
  }
  LayerBuilder& flash_gpt_block(size_t embed, size_t heads, size_t ffn, double dropout_rate = 0.1, bool causal = false, const std::string& act = "gelu", const std::string& name = "") {
    return gpt_block(embed, heads, ffn, dropout_rate, causal, act, name, true);
  }

 private:
  LayerBuilder sub() const {
    LayerBuilder b(get_batchless_current_shape());
    b.io_ = io_;
    return b;
  }
  LayerBuilder& push(std::unique_ptr<Layer> l) {
    l->set_io_dtype(io_);
    shape_ = l->output_shape(shape_);
    layers_.push_back(std::move(l));
    return *this;
  }
  std::string nm(const std::string& name, const char* def) const { return name.empty() ? std::string(def) + "_" + std::to_string(layers_.size()) : name; }
  Shape shape_;
  Vec<std::unique_ptr<Layer>> layers_;
  DType_t io_ = DType_t::FP32;
};

}  // namespace tnn
