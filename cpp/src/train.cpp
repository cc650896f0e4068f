// Model zoo, data loaders, TrainStep (CUDA-graph captured), train / validate loops, checkpoints.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <sys/stat.h>

#include "kernels/simt_ops.cuh"
#include "runtime/data_loader.hpp"
#include "tnn/common.hpp"
#include "tnn/example_models.hpp"
#include "tnn/hardware_info.hpp"
#include "tnn/profiler.hpp"
#include "tnn/stepstate.hpp"
#include "tnn/train.hpp"

namespace tnn {

// ===================================================================================== model zoo
namespace {
std::unique_ptr<Sequential> mnist_cnn(DType_t io) {
  LayerBuilder b({28, 28, 1});
  b.dtype(io);
  b.conv2d(8, 5, 5, 1, 1, 0, 0, false, "conv1").batchnorm(-1, 0.1, true, true, "bn1").maxpool2d(3, 3, 3, 3, 0, 0, "pool1")
      .conv2d(16, 1, 1, 1, 1, 0, 0, false, "conv2_1x1").batchnorm(-1, 0.1, true, true, "bn2_1x1").activation("relu", "relu2")
      .conv2d(48, 5, 5, 1, 1, 0, 0, false, "conv3").batchnorm(-1, 0.1, true, true, "bn3").maxpool2d(2, 2, 2, 2, 0, 0, "pool2")
      .flatten(1, -1, "flatten").dense(10, false, "output");
  return b.build_sequential("mnist_cnn");
}
std::unique_ptr<Sequential> cifar10_vgg(DType_t io) {
  LayerBuilder b({32, 32, 3});
  b.dtype(io);
  struct It { int ch; const char* id; bool bn; };
  const Vec<Vec<It>> groups = {{{64, "0", true}, {64, "1", true}}, {{128, "2", true}, {128, "3", true}}, {{256, "4", true}, {256, "5", false}, {256, "6", true}},
                               {{512, "7", true}, {512, "8", true}, {512, "9", true}}};
  for (size_t g = 0; g < groups.size(); ++g) {
    for (const It& it : groups[g]) {
      b.conv2d((size_t)it.ch, 3, 3, 1, 1, 1, 1, false, std::string("conv") + it.id);
      if (it.bn) b.batchnorm(-1, 0.1, true, true, std::string("bn") + it.id);
      else b.activation("relu", std::string("relu") + it.id);
    }
    b.maxpool2d(2, 2, 2, 2, 0, 0, "pool" + std::to_string(g));
  }
  b.flatten(1, -1, "flatten").dense(512, true, "fc0").activation("relu", "relu10").dense(10, true, "fc1");
  return b.build_sequential("cifar10_vgg");
}
std::unique_ptr<Sequential> cifar10_resnet9(DType_t io) {
  LayerBuilder b({32, 32, 3});
  b.dtype(io);
  b.conv2d(64, 3, 3, 1, 1, 1, 1, false, "conv1").batchnorm(-1, 0.1, true, true, "bn1").conv2d(128, 3, 3, 1, 1, 1, 1, false, "conv2")
      .batchnorm(-1, 0.1, true, true, "bn2").maxpool2d(2, 2, 2, 2, 0, 0, "pool1").basic_residual_block(128, 128, 1, "res_block1")
      .basic_residual_block(128, 128, 1, "res_block2").conv2d(256, 3, 3, 1, 1, 1, 1, false, "conv3").batchnorm(-1, 0.1, true, true, "bn3")
      .maxpool2d(2, 2, 2, 2, 0, 0, "pool2").basic_residual_block(256, 256, 1, "res_block3").basic_residual_block(256, 256, 1, "res_block4")
      .conv2d(512, 3, 3, 1, 1, 1, 1, false, "conv4").batchnorm(-1, 0.1, true, true, "bn4").maxpool2d(2, 2, 2, 2, 0, 0, "pool3")
      .basic_residual_block(512, 512, 1, "res_block5").avgpool2d(4, 4, 1, 1, 0, 0, "avgpool").flatten(1, -1, "flatten").dense(10, true, "output");
  return b.build_sequential("cifar10_resnet9");
}
std::unique_ptr<Sequential> resnet18(const Shape& shape, size_t classes, size_t pool, const std::string& name, DType_t io) {
  LayerBuilder b(shape);
  b.dtype(io);
  b.conv2d(32, 3, 3, 1, 1, 1, 1, false, "conv1").batchnorm(-1, 0.1, true, true, "bn1").maxpool2d(2, 2, 2, 2, 0, 0, "maxpool");
  const size_t spec[8][3] = {{32, 64, 1}, {64, 64, 1}, {64, 128, 2}, {128, 128, 1}, {128, 256, 2}, {256, 256, 1}, {256, 512, 2}, {512, 512, 1}};
  for (int i = 0; i < 8; ++i) b.basic_residual_block(spec[i][0], spec[i][1], spec[i][2], "layer" + std::to_string(i / 2 + 1) + "_block" + std::to_string(i % 2 + 1));
  b.avgpool2d(pool, pool, 1, 1, 0, 0, "avgpool").flatten(1, -1, "flatten").dense(classes, true, "fc");
  return b.build_sequential(name);
}
std::unique_ptr<Sequential> wrn16_8(const Shape& shape, size_t classes, const std::string& name, DType_t io, double dropout = 0.3) {
  LayerBuilder b(shape);
  b.dtype(io);
  b.conv2d(16, 3, 3, 1, 1, 1, 1, true, "conv1");
  const size_t spec[6][3] = {{16, 128, 1}, {128, 128, 1}, {128, 256, 2}, {256, 256, 1}, {256, 512, 2}, {512, 512, 1}};
  for (int i = 0; i < 6; ++i) b.wide_residual_block(spec[i][0], spec[i][1], spec[i][2], dropout, "group" + std::to_string(i / 2 + 1) + "_block" + std::to_string(i % 2 + 1));
  const size_t pool = b.get_current_shape()[1];
  b.batchnorm(-1, 0.1, true, true, "bn_final").avgpool2d(pool >= 8 ? 8 : pool, pool >= 8 ? 8 : pool, 1, 1, 0, 0, "avgpool").flatten(1, -1, "flatten").dense(classes, true, "fc");
  return b.build_sequential(name);
}
std::unique_ptr<Sequential> resnet50(const Shape& shape, size_t classes, bool stem7, size_t pool, const std::string& name, DType_t io) {
  LayerBuilder b(shape);
  b.dtype(io);
  if (stem7) b.conv2d(64, 7, 7, 2, 2, 3, 3, true, "conv1");
  else b.conv2d(64, 3, 3, 1, 1, 1, 1, true, "conv1");
  b.batchnorm(-1, 0.1, true, true, "bn1").maxpool2d(3, 3, 2, 2, 1, 1, "maxpool");
  const size_t cfg[4][5] = {{64, 64, 256, 1, 3}, {256, 128, 512, 2, 4}, {512, 256, 1024, 2, 6}, {1024, 512, 2048, 2, 3}};
  for (int li = 0; li < 4; ++li)
    for (size_t j = 0; j < cfg[li][4]; ++j)
      b.bottleneck_residual_block(j == 0 ? cfg[li][0] : cfg[li][2], cfg[li][1], cfg[li][2], j == 0 ? cfg[li][3] : 1, "layer" + std::to_string(li + 1) + "_block" + std::to_string(j + 1));
  b.avgpool2d(pool, pool, 1, 1, 0, 0, "avgpool").flatten(1, -1, "flatten").dense(classes, true, "fc");
  return b.build_sequential(name);
}
std::unique_ptr<Sequential> vit(bool flash, const std::string& name, DType_t io) {
  const size_t patch = 4, E = 256, H = 4, ratio = 4, depth = 4, classes = 200, seq = (64 / patch) * (64 / patch) + 1;
  LayerBuilder b({64, 64, 3});
  b.dtype(io);
  b.conv2d(E, patch, patch, patch, patch, 0, 0, true, "patch_embed").flatten(1, 2, "flatten_patches").class_token(E, "cls")
      .positional_embedding(E, seq, "pos_embed").dropout(0.1, "embed_dropout");
  for (size_t i = 0; i < depth; ++i) {
    const std::string p = "enc" + std::to_string(i);
    LayerBuilder a({seq, E});
    a.dtype(io).layernorm(1e-5, true, p + "_ln_attn");
    if (flash) a.flash_attention(E, H, false, p + "_attn");
    else a.attention(E, H, false, p + "_attn");
    a.dropout(0.1, p + "_attn_dropout");
    b.residual_block(a.build(), {}, "linear", "encoder_" + std::to_string(i) + "_attn");
    LayerBuilder m({seq, E});
    m.dtype(io).layernorm(1e-5, true, p + "_ln_mlp").dense(E * ratio, false, p + "_fc1").activation("gelu", p + "_gelu").dropout(0.1, p + "_d1")
        .dense(E, false, p + "_fc2").dropout(0.1, p + "_d2");
    b.residual_block(m.build(), {}, "linear", "encoder_" + std::to_string(i) + "_mlp");
  }
  b.layernorm(1e-5, true, "ln_final").slice(1, 0, 1, "extract_cls").flatten(1, -1, "flatten_cls").dense(classes, true, "head");
  return b.build_sequential(name);
}
std::unique_ptr<Sequential> gpt2(size_t E, size_t H, size_t layers, bool flash, const std::string& name, DType_t io, size_t seq_len = 1024, size_t vocab = 50257, double dropout = 0.1) {
  LayerBuilder b({seq_len});
  b.dtype(io);
  b.embedding(vocab, E, -1, "token_embed").positional_embedding(E, seq_len, "pos_embed");
  if (dropout > 0) b.dropout(dropout, "embed_dropout");
  for (size_t i = 0; i < layers; ++i) b.gpt_block(E, H, 4 * E, dropout, true, "gelu", "h" + std::to_string(i), flash);
  b.layernorm(1e-5, true, "ln_f").dense(vocab, true, "head");
  return b.build_sequential(name);
}
}  // namespace

std::map<std::string, ExampleModels::Creator>& ExampleModels::table() {
  static std::map<std::string, Creator> t;
  return t;
}
void ExampleModels::register_model(const std::string& name, Creator c) { table()[name] = std::move(c); }
void ExampleModels::register_defaults() {
  if (!table().empty()) return;
  register_model("mnist_cnn", mnist_cnn);
  register_model("cifar10_vgg", cifar10_vgg);
  register_model("cifar10_resnet9", cifar10_resnet9);
  register_model("cifar100_resnet18", [](DType_t io) { return resnet18({32, 32, 3}, 100, 2, "cifar100_resnet18", io); });
  register_model("tiny_imagenet_resnet18", [](DType_t io) { return resnet18({64, 64, 3}, 200, 4, "tiny_imagenet_resnet18", io); });
  register_model("cifar100_wrn16_8", [](DType_t io) { return wrn16_8({32, 32, 3}, 100, "cifar100_wrn16_8", io); });
  register_model("tiny_imagenet_wrn16_8", [](DType_t io) { return wrn16_8({64, 64, 3}, 200, "tiny_imagenet_wrn16_8", io); });
  register_model("tiny_imagenet_resnet50", [](DType_t io) { return resnet50({64, 64, 3}, 200, false, 4, "tiny_imagenet_resnet50", io); });
  register_model("imagenet_resnet50", [](DType_t io) { return resnet50({224, 224, 3}, 1000, true, 7, "imagenet_resnet50", io); });
  register_model("tiny_imagenet_vit", [](DType_t io) { return vit(false, "tiny_imagenet_vit", io); });
  register_model("tiny_imagenet_flash_vit", [](DType_t io) { return vit(true, "tiny_imagenet_flash_vit", io); });
  register_model("gpt2_small", [](DType_t io) { return gpt2(768, 12, 12, false, "gpt2_small", io); });
  register_model("gpt2", [](DType_t io) { return gpt2(768, 12, 12, false, "gpt2_small", io); });
  register_model("flash_gpt2_small", [](DType_t io) { return gpt2(768, 12, 12, true, "flash_gpt2_small", io); });
  register_model("gpt2_medium", [](DType_t io) { return gpt2(1024, 16, 24, false, "gpt2_medium", io); });
  register_model("flash_gpt2_medium", [](DType_t io) { return gpt2(1024, 16, 24, true, "flash_gpt2_medium", io); });
  register_model("gpt2_large", [](DType_t io) { return gpt2(1280, 20, 36, false, "gpt2_large", io); });
  register_model("flash_gpt2_large", [](DType_t io) { return gpt2(1280, 20, 36, true, "flash_gpt2_large", io); });
}
std::unique_ptr<Sequential> ExampleModels::create(const std::string& name, DType_t io) {
  register_defaults();
  auto it = table().find(name);
  if (it == table().end()) throw std::invalid_argument("ExampleModels: unknown model '" + name + "'");
  return it->second(io);
}
Vec<std::string> ExampleModels::available_models() {
  register_defaults();
  Vec<std::string> v;
  for (auto& kv : table()) v.push_back(kv.first);
  return v;
}
Shape ExampleModels::input_shape(const std::string& n) {
  if (n.find("mnist") != std::string::npos) return {28, 28, 1};
  if (n.find("tiny_imagenet") != std::string::npos) return {64, 64, 3};
  if (n.find("imagenet") != std::string::npos) return {224, 224, 3};
  if (n.find("gpt2") != std::string::npos) return {1024};
  return {32, 32, 3};
}
size_t ExampleModels::num_classes(const std::string& n) {
  if (n.find("mnist") != std::string::npos || n.find("cifar10_") != std::string::npos) return 10;
  if (n.find("cifar100") != std::string::npos) return 100;
  if (n.find("tiny_imagenet") != std::string::npos) return 200;
  if (n.find("imagenet") != std::string::npos) return 1000;
  if (n.find("gpt2") != std::string::npos) return 50257;
  return 10;
}
Graph load_or_create_model(const std::string& name, const std::string& path, const Device& dev, uint64_t seed) {
  struct stat st;
  if (!path.empty() && stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode)) return Graph::load_state(path, dev);
  return graph_from_sequential(ExampleModels::create(name), dev, seed, name);
}

// ===================================================================================== data loaders
bool BaseDataLoader::get_batch(size_t B, Tensor& x, Tensor& y) {
  if (cursor_ >= n_) return false;
  const size_t b = std::min(B, n_ - cursor_);
  Shape xs{b};
  xs.insert(xs.end(), sample_shape_.begin(), sample_shape_.end());
  IAllocator& a = pinned_ ? static_cast<IAllocator&>(DeviceAllocator::pinned()) : HostAllocator();
  if (!x || x->shape() != xs || x->is_gpu()) x = make_tensor(a, DType_t::FP32, xs);
  if (!y || y->size() != b || y->is_gpu()) y = make_tensor(a, DType_t::INT32_T, {b});
  const size_t per = shape_numel(sample_shape_);
  if (order_.size() != n_) {
    order_.resize(n_);
    std::iota(order_.begin(), order_.end(), 0u);
  }
  for (size_t i = 0; i < b; ++i) fill_sample(order_[cursor_ + i], x->data_as<float>() + i * per, y->data_as<int32_t>() + i);
  cursor_ += b;
  if (augmentation_ && sample_shape_.size() == 3) augmentation_->apply(*x, batches_served_);
  ++batches_served_;
  return true;
}
void BaseDataLoader::shuffle() {
  if (order_.size() != n_) {
    order_.resize(n_);
    std::iota(order_.begin(), order_.end(), 0u);
  }
  std::shuffle(order_.begin(), order_.end(), rng_);
}
SyntheticDataLoader::SyntheticDataLoader(const Shape& s, size_t classes, size_t n, uint64_t seed) : seed_(seed) {
  sample_shape_ = s;
  classes_ = classes;
  n_ = n;
}
void SyntheticDataLoader::fill_sample(size_t index, float* dst, int32_t* label) {
  std::mt19937_64 g(seed_ * 1000003ull + index);
  const int32_t cls = (int32_t)(g() % classes_);
  *label = cls;
  const size_t per = shape_numel(sample_shape_);
  if (sample_shape_.size() == 1) {                        // token sequences: ids in [0, classes)
    for (size_t i = 0; i < per; ++i) dst[i] = (float)(g() % classes_);
    return;
  }
  std::normal_distribution<float> nd(0.f, 0.5f);
  std::mt19937_64 cg(7919ull * (uint64_t)(cls + 1) + 12345ull);   // class prototype: shared by every loader (train / val draw different noise)
  std::normal_distribution<float> pd(0.5f, 0.25f);
  for (size_t i = 0; i < per; ++i) dst[i] = pd(cg) + nd(g) * 0.3f;
}
bool CIFARDataLoader::load_data(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t rec = c100_ ? 3074 : 3073;
  if (buf.size() % rec != 0) return false;
  records_.insert(records_.end(), buf.begin(), buf.end());
  n_ = records_.size() / rec;
  return true;
}
void CIFARDataLoader::fill_sample(size_t index, float* dst, int32_t* label) {
  const size_t rec = c100_ ? 3074 : 3073, hdr = c100_ ? 2 : 1;
  const uint8_t* r = records_.data() + index * rec;
  *label = c100_ ? r[1] : r[0];
  const uint8_t* px = r + hdr;                            // CHW planes -> HWC, / 255 (cifar10_data_loader.hpp:152-165)
  for (size_t p = 0; p < 1024; ++p)
    for (size_t c = 0; c < 3; ++c) dst[p * 3 + c] = px[c * 1024 + p] * (1.f / 255.f);
}
bool MNISTDataLoader::load_data(const std::string& path) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  while (std::getline(f, line)) {
    if (line.empty() || !isdigit((unsigned char)line[0])) continue;          // header row
    std::stringstream ss(line);
    std::string tok;
    std::getline(ss, tok, ',');
    labels_.push_back((uint8_t)atoi(tok.c_str()));
    size_t k = 0;
    while (std::getline(ss, tok, ',') && k < 784) {
      pixels_.push_back((uint8_t)atoi(tok.c_str()));
      ++k;
    }
    pixels_.resize(labels_.size() * 784, 0);
  }
  n_ = labels_.size();
  return n_ > 0;
}
void MNISTDataLoader::fill_sample(size_t index, float* dst, int32_t* label) {
  *label = labels_[index];
  for (size_t i = 0; i < 784; ++i) dst[i] = pixels_[index * 784 + i] * (1.f / 255.f);
}
// ---- OpenWebText token stream
struct OpenWebTextDataLoader::Impl {
  std::unique_ptr<TokenDataset> ds;
};
OpenWebTextDataLoader::~OpenWebTextDataLoader() { delete impl_; }
bool OpenWebTextDataLoader::load_data(const std::string& path) {
  struct stat st;
  std::string f = path;
  if (stat(f.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) f = path + "/train.bin";
  if (stat(f.c_str(), &st) != 0) return false;
  if (!impl_) impl_ = new Impl();
  impl_->ds = std::make_unique<TokenDataset>(f);
  n_ = impl_->ds->num_tokens() / seq_len_;
  cursor_ = 0;
  return n_ > 0;
}
size_t OpenWebTextDataLoader::num_tokens() const { return impl_ && impl_->ds ? impl_->ds->num_tokens() : 0; }
bool OpenWebTextDataLoader::get_batch(size_t B, Tensor& x, Tensor& y) {
  if (!impl_ || !impl_->ds || cursor_ >= n_) return false;
  const size_t b = std::min(B, n_ - cursor_);
  IAllocator& a = pinned_ ? static_cast<IAllocator&>(DeviceAllocator::pinned()) : HostAllocator();
  x = make_tensor(a, DType_t::INT32_T, {b, seq_len_});
  y = make_tensor(a, DType_t::INT32_T, {b, seq_len_});
  impl_->ds->sample((int)b, (int)seq_len_, rng_(), x->data_as<int32_t>(), y->data_as<int32_t>());
  cursor_ += b;
  return true;
}

// ---- numeric CSV helpers
static bool read_csv(const std::string& path, bool header, std::vector<std::vector<float>>& rows) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  if (header) std::getline(f, line);
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    std::vector<float> r;
    std::stringstream ss(line);
    std::string cell;
    while (std::getline(ss, cell, ',')) r.push_back(cell.empty() ? 0.f : (float)atof(cell.c_str()));
    if (!r.empty()) rows.push_back(std::move(r));
  }
  return !rows.empty();
}
static bool csv_batch(size_t B, size_t& cursor, size_t n, const std::vector<uint32_t>& order, const std::vector<float>& xs, size_t F, Tensor& x,
                      bool pinned) {
  if (cursor >= n) return false;
  const size_t b = std::min(B, n - cursor);
  IAllocator& a = pinned ? static_cast<IAllocator&>(DeviceAllocator::pinned()) : HostAllocator();
  x = make_tensor(a, DType_t::FP32, {b, F});
  for (size_t i = 0; i < b; ++i) std::memcpy(x->data_as<float>() + i * F, xs.data() + (size_t)order[cursor + i] * F, F * sizeof(float));
  return true;
}
bool RegressionCSVDataLoader::load_data(const std::string& path) {
  std::vector<std::vector<float>> rows;
  if (!read_csv(path, header_, rows)) return false;
  const size_t cols = rows[0].size();
  if (cols <= n_targets_) return false;
  features_ = cols - n_targets_;
  n_ = rows.size();
  xs_.resize(n_ * features_);
  ys_.resize(n_ * n_targets_);
  for (size_t i = 0; i < n_; ++i) {
    if (rows[i].size() != cols) return false;
    std::copy(rows[i].begin(), rows[i].begin() + (long)features_, xs_.begin() + (long)(i * features_));
    std::copy(rows[i].begin() + (long)features_, rows[i].end(), ys_.begin() + (long)(i * n_targets_));
  }
  sample_shape_ = {features_};
  classes_ = n_targets_;
  cursor_ = 0;
  order_.clear();
  return true;
}
bool RegressionCSVDataLoader::get_batch(size_t B, Tensor& x, Tensor& y) {
  if (order_.size() != n_) {
    order_.resize(n_);
    std::iota(order_.begin(), order_.end(), 0u);
  }
  if (!csv_batch(B, cursor_, n_, order_, xs_, features_, x, pinned_)) return false;
  size_t c2 = cursor_;
  csv_batch(B, c2, n_, order_, ys_, n_targets_, y, pinned_);
  cursor_ += x->dimension(0);
  return true;
}
bool WiFiDataLoader::load_data(const std::string& path, size_t fs, size_t fe, size_t ts, size_t te, bool has_header) {
  std::vector<std::vector<float>> rows;
  if (!read_csv(path, has_header, rows)) return false;
  const size_t cols = rows[0].size();
  if (fe <= fs) {                                   // reference default: all but the last column
    fs = 0;
    fe = cols - 1;
  }
  if (te <= ts) {
    ts = fe;
    te = cols;
  }
  fe = std::min(fe, cols);
  te = std::min(te, cols);
  features_ = fe - fs;
  targets_ = regression_ ? te - ts : 1;
  n_ = rows.size();
  xs_.resize(n_ * features_);
  ys_.resize(n_ * targets_);
  for (size_t i = 0; i < n_; ++i) {
    if (rows[i].size() != cols) return false;
    for (size_t c = 0; c < features_; ++c) xs_[i * features_ + c] = rows[i][fs + c];
    for (size_t c = 0; c < targets_; ++c) ys_[i * targets_ + c] = rows[i][ts + c];
  }
  sample_shape_ = {features_};
  classes_ = targets_;
  normalized_ = false;
  cursor_ = 0;
  order_.clear();
  return true;
}
void WiFiDataLoader::normalize_data() {
  auto zscore = [&](std::vector<float>& v, size_t F, std::vector<float>& mean, std::vector<float>& sd) {
    mean.assign(F, 0.f);
    sd.assign(F, 0.f);
    for (size_t i = 0; i < n_; ++i)
      for (size_t c = 0; c < F; ++c) mean[c] += v[i * F + c];
    for (size_t c = 0; c < F; ++c) mean[c] /= (float)n_;
    for (size_t i = 0; i < n_; ++i)
      for (size_t c = 0; c < F; ++c) sd[c] += (v[i * F + c] - mean[c]) * (v[i * F + c] - mean[c]);
    for (size_t c = 0; c < F; ++c) sd[c] = std::max(1e-8f, std::sqrt(sd[c] / (float)n_));
    for (size_t i = 0; i < n_; ++i)
      for (size_t c = 0; c < F; ++c) v[i * F + c] = (v[i * F + c] - mean[c]) / sd[c];
  };
  if (normalized_ || n_ == 0) return;
  zscore(xs_, features_, fmean_, fstd_);
  if (regression_) zscore(ys_, targets_, tmean_, tstd_);
  normalized_ = true;
}
bool WiFiDataLoader::get_batch(size_t B, Tensor& x, Tensor& y) {
  if (order_.size() != n_) {
    order_.resize(n_);
    std::iota(order_.begin(), order_.end(), 0u);
  }
  if (!csv_batch(B, cursor_, n_, order_, xs_, features_, x, pinned_)) return false;
  const size_t b = x->dimension(0);
  IAllocator& a = pinned_ ? static_cast<IAllocator&>(DeviceAllocator::pinned()) : HostAllocator();
  if (regression_) {
    y = make_tensor(a, DType_t::FP32, {b, targets_});
    for (size_t i = 0; i < b; ++i) std::memcpy(y->data_as<float>() + i * targets_, ys_.data() + (size_t)order_[cursor_ + i] * targets_, targets_ * sizeof(float));
  } else {
    y = make_tensor(a, DType_t::INT32_T, {b});
    for (size_t i = 0; i < b; ++i) y->data_as<int32_t>()[i] = (int32_t)ys_[(size_t)order_[cursor_ + i]];
  }
  cursor_ += b;
  return true;
}

DataLoaderPair DataLoaderFactory::create(const std::string& name, const std::string& path) {
  DataLoaderPair p;
  auto exists = [](const std::string& f) { struct stat st; return stat(f.c_str(), &st) == 0; };
  if (name == "cifar10" || name == "cifar100") {
    const bool c100 = name == "cifar100";
    auto tr = std::make_unique<CIFARDataLoader>(c100), va = std::make_unique<CIFARDataLoader>(c100);
    bool ok = false;
    if (c100) ok = tr->load_data(path + "/train.bin") && va->load_data(path + "/test.bin");
    else {
      ok = true;
      for (int i = 1; i <= 5; ++i) ok &= tr->load_data(path + "/data_batch_" + std::to_string(i) + ".bin");
      ok &= va->load_data(path + "/test_batch.bin");
    }
    if (ok) {
      p.train = std::move(tr);
      p.val = std::move(va);
      return p;
    }
    p.train = std::make_unique<SyntheticDataLoader>(Shape{32, 32, 3}, c100 ? 100 : 10, 8192, 1);
    p.val = std::make_unique<SyntheticDataLoader>(Shape{32, 32, 3}, c100 ? 100 : 10, 1024, 1);
    return p;
  }
  if (name == "mnist") {
    auto tr = std::make_unique<MNISTDataLoader>(), va = std::make_unique<MNISTDataLoader>();
    if (exists(path + "/mnist_train.csv") && tr->load_data(path + "/mnist_train.csv") && va->load_data(path + "/mnist_test.csv")) {
      p.train = std::move(tr);
      p.val = std::move(va);
      return p;
    }
    p.train = std::make_unique<SyntheticDataLoader>(Shape{28, 28, 1}, 10, 4096, 1);
    p.val = std::make_unique<SyntheticDataLoader>(Shape{28, 28, 1}, 10, 512, 1);
    return p;
  }
  if (name == "openwebtext" || name == "open_webtext") {
    auto tr = std::make_unique<OpenWebTextDataLoader>(1024), va = std::make_unique<OpenWebTextDataLoader>(1024);
    if (!tr->load_data(path + "/train.bin")) throw std::runtime_error("DataLoaderFactory: no token file " + path + "/train.bin");
    p.train = std::move(tr);
    if (va->load_data(path + "/val.bin")) p.val = std::move(va);
    return p;
  }
  if (name == "tiny_imagenet" || name == "imagenet100") {
    const bool tiny = name == "tiny_imagenet";
    auto mk = [&](bool train) -> std::unique_ptr<ImageFolderDataLoader> {
      if (tiny) return std::make_unique<TinyImageNetDataLoader>(train);
      return std::make_unique<ImageNet100DataLoader>(train);
    };
    auto tr = mk(true), va = mk(false);
    if (!tr->load_data(path)) throw std::runtime_error("DataLoaderFactory: no image tree or .npy cache under " + path);
    p.train = std::move(tr);
    if (va->load_data(path)) p.val = std::move(va);
    return p;
  }
  if (name == "regression" || name == "wifi" || name == "uji") {
    const bool wifi = name != "regression";
    auto mk = [&](const std::string& f) -> std::unique_ptr<BaseDataLoader> {
      if (wifi) {
        auto l = std::make_unique<WiFiDataLoader>(true);
        if (!l->load_data(f)) return nullptr;
        l->normalize_data();
        return l;
      }
      auto l = std::make_unique<RegressionCSVDataLoader>(1);
      if (!l->load_data(f)) return nullptr;
      return l;
    };
    p.train = mk(exists(path + "/train.csv") ? path + "/train.csv" : path);
    if (!p.train) throw std::runtime_error("DataLoaderFactory: cannot read " + path);
    if (exists(path + "/val.csv")) p.val = mk(path + "/val.csv");
    return p;
  }
  // "synthetic[:H,W,C,classes,n]"
  size_t H = 32, W = 32, C = 3, classes = 10, n = 4096;
  const size_t colon = name.find(':');
  if (colon != std::string::npos) sscanf(name.c_str() + colon + 1, "%zu,%zu,%zu,%zu,%zu", &H, &W, &C, &classes, &n);
  p.train = std::make_unique<SyntheticDataLoader>(Shape{H, W, C}, classes, n, 1);
  p.val = std::make_unique<SyntheticDataLoader>(Shape{H, W, C}, classes, std::max<size_t>(n / 8, 1), 1);
  return p;
}

// ===================================================================================== training config
static std::map<std::string, std::string> read_dotenv(const std::string& path) {
  std::map<std::string, std::string> kv;
  std::ifstream f(path);
  std::string line;
  while (f && std::getline(f, line)) {
    const size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    const size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    auto trim = [](std::string s) {
      while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
      size_t i = 0;
      while (i < s.size() && isspace((unsigned char)s[i])) ++i;
      s = s.substr(i);
      if (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front()) s = s.substr(1, s.size() - 2);
      return s;
    };
    kv[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
  }
  return kv;
}
TrainingConfig TrainingConfig::load_from_env(const std::string& dotenv_path) {
  const auto file = read_dotenv(dotenv_path);
  auto get = [&](const char* k, const std::string& def) {
    const char* e = getenv(k);
    if (e) return std::string(e);
    auto it = file.find(k);
    return it == file.end() ? def : it->second;
  };
  TrainingConfig c;
  c.epochs = (size_t)atol(get("EPOCHS", std::to_string(c.epochs)).c_str());
  c.batch_size = (size_t)atol(get("BATCH_SIZE", std::to_string(c.batch_size)).c_str());
  c.max_steps = atol(get("MAX_STEPS", "-1").c_str());
  c.lr_initial = atof(get("LR_INITIAL", std::to_string(c.lr_initial)).c_str());
  c.gradient_accumulation_steps = (size_t)std::max(1l, atol(get("GRADIENT_ACCUMULATION_STEPS", "1").c_str()));
  c.progress_print_interval = (size_t)atol(get("PROGRESS_PRINT_INTERVAL", std::to_string(c.progress_print_interval)).c_str());
  c.num_threads = (size_t)atol(get("NUM_THREADS", "4").c_str());
  c.profiler_type = get("PROFILER_TYPE", "NONE");
  c.print_layer_profiling = get("PRINT_LAYER_PROFILING", "0") == "1";
  c.print_layer_memory_usage = get("PRINT_LAYER_MEMORY_USAGE", "0") == "1";
  c.model_name = get("MODEL_NAME", c.model_name);
  c.model_path = get("MODEL_PATH", "");
  c.dataset_name = get("DATASET_NAME", "");
  c.dataset_path = get("DATASET_PATH", "data");
  c.device_type = get("DEVICE_TYPE", "CPU");
  c.io_dtype = get("IO_DTYPE", "FP32");
  c.param_dtype = get("PARAM_DTYPE", "FP32");
  c.compute_dtype = get("COMPUTE_DTYPE", "FP32");
  c.num_microbatches = (size_t)std::max(1l, atol(get("NUM_MICROBATCHES", "2").c_str()));
  c.snapshot_dir = get("SNAPSHOT_DIR", "model_snapshots");
  c.seed = (uint64_t)atoll(get("SEED", "0").c_str());
  return c;
}
TrainingConfig TrainingConfig::load_from_json(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("TrainingConfig: cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  const Json j = Json::parse(ss.str());
  TrainingConfig c;
  c.epochs = (size_t)j.value("epochs", (int64_t)c.epochs);
  c.batch_size = (size_t)j.value("batch_size", (int64_t)c.batch_size);
  c.max_steps = (long)j.value("max_steps", (int64_t)-1);
  c.lr_initial = j.value("lr_initial", c.lr_initial);
  c.gradient_accumulation_steps = (size_t)j.value("gradient_accumulation_steps", (int64_t)1);
  c.progress_print_interval = (size_t)j.value("progress_print_interval", (int64_t)c.progress_print_interval);
  c.num_threads = (size_t)j.value("num_threads", (int64_t)4);
  c.profiler_type = j.value("profiler_type", "NONE");
  c.print_layer_profiling = j.value("print_layer_profiling", false);
  c.print_layer_memory_usage = j.value("print_layer_memory_usage", false);
  c.model_name = j.value("model_name", c.model_name);
  c.model_path = j.value("model_path", "");
  c.dataset_name = j.value("dataset_name", "");
  c.dataset_path = j.value("dataset_path", "data");
  c.device_type = j.value("device_type", "CPU");
  c.io_dtype = j.value("io_dtype", "FP32");
  c.param_dtype = j.value("param_dtype", "FP32");
  c.compute_dtype = j.value("compute_dtype", "FP32");
  c.num_microbatches = (size_t)j.value("num_microbatches", (int64_t)2);
  c.snapshot_dir = j.value("snapshot_dir", "model_snapshots");
  c.seed = (uint64_t)j.value("seed", (int64_t)0);
  return c;
}
Json TrainingConfig::to_json() const {
  Json j = Json::object();
  j["epochs"] = epochs; j["batch_size"] = batch_size; j["max_steps"] = (long long)max_steps; j["lr_initial"] = lr_initial;
  j["gradient_accumulation_steps"] = gradient_accumulation_steps; j["progress_print_interval"] = progress_print_interval;
  j["num_threads"] = num_threads; j["profiler_type"] = profiler_type; j["print_layer_profiling"] = print_layer_profiling;
  j["print_layer_memory_usage"] = print_layer_memory_usage; j["model_name"] = model_name; j["model_path"] = model_path;
  j["dataset_name"] = dataset_name; j["dataset_path"] = dataset_path; j["device_type"] = device_type; j["io_dtype"] = io_dtype;
  j["param_dtype"] = param_dtype; j["compute_dtype"] = compute_dtype; j["num_microbatches"] = num_microbatches; j["snapshot_dir"] = snapshot_dir;
  j["seed"] = (long long)seed;
  return j;
}
Device TrainingConfig::device() const {
  std::string t = device_type;
  for (auto& c : t) c = (char)toupper((unsigned char)c);
  if (t.rfind("GPU", 0) == 0 || t.rfind("CUDA", 0) == 0) return DeviceManager::getInstance().getDevice(t);
  return getHost();
}

// ===================================================================================== TrainStep
TrainStep::TrainStep(Graph& g, Loss& criterion, Optimizer& opt, Scheduler* sched, size_t accum, bool use_graph)
    : g_(g), crit_(criterion), opt_(opt), sched_(sched), accum_(std::max<size_t>(accum, 1)), exec_(g) {
  if (!opt_.attached()) opt_.attach(g.context());
  opt_.grad_scale = 1.0 / (double)accum_;
  use_graph_ = use_graph && g.device().is_gpu() && accum_ == 1;
}
TrainStep::~TrainStep() {
  if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
  if (graph_) cudaGraphDestroy(graph_);
}
Tensor TrainStep::core(const Tensor& x, const Tensor& y, size_t mb) {
  Tensor out;
  {
    NvtxRange r("step:forward");
    out = exec_.forward(x, mb);
  }
  Tensor stats = make_tensor(exec_.workspace(), DType_t::FP32, {2});
  stats->zero();
  Tensor grad = make_tensor(exec_.workspace(), out->data_type(), out->shape());
  {
    NvtxRange r("step:loss");
    crit_.loss_and_grad(out, y, stats, &grad, 1.f);
  }
  NvtxRange r("step:backward");
  exec_.backward(grad, mb);
  return stats;
}
void TrainStep::refresh_conv_flips() {
  const Device& dev = g_.device();
  if (!dev.is_gpu()) return;
  Vec<Conv2DLayer*> convs;
  for (Layer* top : g_.get_layers())
    for (Layer* l : top->all_layers())
      if (auto* c = dynamic_cast<Conv2DLayer*>(l))
        if (c->flipped_weights()) convs.push_back(c);
  if (convs.empty()) return;
  if (convs.size() != flip_count_) {                       // (re)build the device descriptor table: pointers are stable
    std::vector<ConvFlipDesc> host(convs.size());
    int blocks = 0;
    for (size_t i = 0; i < convs.size(); ++i) {
      Conv2DLayer* c = convs[i];
      ConvFlipDesc& d = host[i];
      d.w = c->param("weight")->data_as<float>();
      d.wt = c->flipped_weights()->data();
      d.K = (int)c->out_channels();
      d.R = (int)c->kernel_h();
      d.S = (int)c->kernel_w();
      d.C = (int)c->in_channels();
      d.block_begin = blocks;
      d.pad_ = 0;
      blocks += ((d.C + 31) / 32) * ((d.K + 31) / 32) * d.R * d.S;
    }
    flip_descs_ = make_tensor(DType_t::UINT8_T, {host.size() * sizeof(ConvFlipDesc)}, dev);
    flip_descs_->copy_from_host(host.data(), host.size() * sizeof(ConvFlipDesc));
    flip_count_ = convs.size();
    flip_blocks_ = blocks;
  }
  conv_weight_flip_batch(static_cast<const ConvFlipDesc*>(flip_descs_->data()), (int)flip_count_, flip_blocks_, current_flow(dev).stream());
  for (Conv2DLayer* c : convs) c->mark_flipped_fresh();
}

Tensor TrainStep::eager(const Tensor& x, const Tensor& y, size_t mb) {
  const Device& dev = g_.device();
  Tensor dx = x->device() == dev ? x : x->to_device(dev, &exec_.workspace());
  Tensor dy = y->device() == dev ? y : y->to_device(dev, &exec_.workspace());
  ++micro_;
  const bool do_update = micro_ % accum_ == 0;
  if (dev.is_gpu()) StepState::get(dev).advance(do_update ? &opt_ : nullptr);
  Tensor stats = core(dx, dy, mb);
  if (do_update) {
    if (update_override) {
      update_override();
      ++opt_.step_count;
    } else {
      opt_.update(/*external_state=*/dev.is_gpu());
    }
    refresh_conv_flips();
    if (sched_) sched_->step();
  }
  // the stats tensor must outlive the workspace reuse of the next step: hand back a pool-owned copy
  Tensor keep = make_tensor(DType_t::FP32, {2}, dev);
  keep->copy_from(*stats);
  return keep;
}
void TrainStep::capture(const Tensor& x, const Tensor& y) {
  const Device& dev = g_.device();
  dev.set_current();
  Flow& flow = current_flow(dev);
  sx_ = make_tensor(x->data_type(), x->shape(), dev);
  sy_ = make_tensor(y->data_type(), y->shape(), dev);
  sstats_ = make_tensor(DType_t::FP32, {2}, dev);
  flow.synchronize();
  const uint64_t count = opt_.step_count, l0 = g_launch_count;
  TNN_CUDA(cudaStreamBeginCapture(flow.stream(), cudaStreamCaptureModeRelaxed));
  try {
    Tensor stats = core(sx_, sy_, 0);
    if (update_override) update_override();
    else opt_.update(/*external_state=*/true);
    refresh_conv_flips();                    // the two eager warm-up steps allocated every flipped copy: table is final here
    TNN_CUDA(cudaMemcpyAsync(sstats_->data(), stats->data(), 8, cudaMemcpyDeviceToDevice, flow.stream()));
  } catch (...) {
    cudaGraph_t dead = nullptr;
    cudaStreamEndCapture(flow.stream(), &dead);
    if (dead) cudaGraphDestroy(dead);
    throw;
  }
  TNN_CUDA(cudaStreamEndCapture(flow.stream(), &graph_));
  TNN_CUDA(cudaGraphInstantiate(&graph_exec_, graph_, 0));
  captured_launches_ = g_launch_count - l0;
  opt_.step_count = count;                 // capture does not execute; replays advance it
}
Tensor TrainStep::operator()(const Tensor& x, const Tensor& y, size_t mb) {
  if (!use_graph_) return eager(x, y, mb);
  if (!graph_exec_) {
    if (warm_ < 2) {
      ++warm_;
      return eager(x, y, mb);
    }
    capture(x, y);
  } else if (x->shape() != sx_->shape()) {
    return eager(x, y, mb);                // ragged last batch
  }
  const Device& dev = g_.device();
  dev.set_current();
  Flow& flow = current_flow(dev);
  // inputs straight into the graph's static buffers (async from pinned host memory or device to device)
  const cudaMemcpyKind kx = x->is_gpu() ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  TNN_CUDA(cudaMemcpyAsync(sx_->data(), x->data(), sx_->nbytes(), kx, flow.stream()));
  TNN_CUDA(cudaMemcpyAsync(sy_->data(), y->data(), sy_->nbytes(), y->is_gpu() ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, flow.stream()));
  ++micro_;
  StepState::get(dev).advance(&opt_);
  TNN_CUDA(cudaGraphLaunch(graph_exec_, flow.stream()));
  ++replays_;
  ++opt_.step_count;
  if (sched_) sched_->step();
  return sstats_;
}

// ===================================================================================== loops
std::pair<double, double> validate_model(Graph& g, BaseDataLoader& loader, Loss& criterion, const TrainingConfig& cfg) {
  g.set_training(false);
  GraphExecutor ex(g);
  loader.reset();
  Tensor x, y;
  double loss = 0, correct = 0;
  size_t n = 0, batches = 0;
  while (loader.get_batch(cfg.batch_size, x, y)) {
    Tensor dx = g.device().is_gpu() ? x->to_device(g.device(), &ex.workspace()) : x;
    Tensor dy = g.device().is_gpu() ? y->to_device(g.device(), &ex.workspace()) : y;
    Tensor out = ex.forward(dx, 0);
    Tensor stats;
    criterion.loss_and_grad(out, dy, stats, nullptr, 1.f);
    float st[2];
    stats->copy_to_host(st, 8);
    loss += st[0];
    correct += st[1];
    n += y->size();
    ++batches;
  }
  g.set_training(true);
  return {loss / (double)std::max<size_t>(batches, 1), correct / (double)std::max<size_t>(n, 1)};
}
void save_checkpoint(Graph& g, const std::string& path, Optimizer* opt, Scheduler* sched, uint64_t step) {
  std::map<std::string, Tensor> tensors;
  Json meta = Json::object();
  meta["step"] = (long long)step;
  if (opt) {
    for (auto& kv : opt->state_tensors()) tensors["opt." + kv.first] = kv.second;
    meta["opt.config"] = opt->get_config().to_json();
    meta["opt.step"] = (long long)opt->step_count;
    meta["opt.lr"] = opt->get_learning_rate();
  }
  if (sched) {
    meta["sched.config"] = sched->get_config().to_json();
    meta["sched.step"] = (long long)sched->step_index();
  }
  g.save_state(path, &tensors, &meta);
}
Graph load_checkpoint(const std::string& path, const Device& dev, Optimizer* opt, Scheduler* sched, uint64_t* step) {
  std::map<std::string, Tensor> tensors;
  Json meta;
  Graph g = Graph::load_state(path, dev, &tensors, &meta);
  if (step && meta.is_object()) *step = (uint64_t)meta.value("step", (int64_t)0);
  if (opt) {
    opt->attach(g.context());
    for (auto& kv : opt->state_tensors()) {
      auto it = tensors.find("opt." + kv.first);
      if (it != tensors.end() && it->second->size() == kv.second->size()) kv.second->copy_from(*it->second);
    }
    if (meta.is_object()) {
      opt->step_count = (uint64_t)meta.value("opt.step", (int64_t)0);
      opt->set_learning_rate(meta.value("opt.lr", opt->get_learning_rate()));
    }
  }
  if (sched && meta.is_object()) sched->set_step_index((uint64_t)meta.value("sched.step", (int64_t)0));
  return g;
}
std::string layer_memory_table(Graph& g, GraphExecutor& exec) {
  // TrainingConfig::print_layer_memory_usage: parameter bytes per leaf layer (fp32 master + gradient, + the bf16 copy on the
  // GPU) and the executor's activation workspace (in use now / peak / reserved slabs)
  std::string out = "layer memory usage:\n";
  char b[512];
  const bool gpu = g.device().is_gpu();
  size_t total = 0, w = 5;
  Vec<Layer*> leaves;
  for (Layer* top : g.get_layers())
    for (Layer* l : top->all_layers())
      if (!l->param_descriptors().empty()) {
        leaves.push_back(l);
        w = std::max(w, l->name().size());
      }
  snprintf(b, sizeof b, "  %-*s %-18s %14s %14s\n", (int)w, "layer", "type", "parameters", "bytes");
  out += b;
  for (Layer* l : leaves) {
    size_t n = 0, bytes = 0;
    for (const auto& d : l->param_descriptors()) {
      const size_t k = shape_numel(d.shape);
      n += k;
      bytes += k * 4 + (d.trainable ? k * 4 + (gpu ? k * 2 : 0) : 0);
    }
    total += bytes;
    snprintf(b, sizeof b, "  %-*s %-18s %14zu %14zu\n", (int)w, l->name().c_str(), l->type().c_str(), n, bytes);
    out += b;
  }
  DELAllocatorV2& ws = exec.workspace();
  snprintf(b, sizeof b, "  parameters + gradients%s: %.2f MB; activation workspace: %.2f MB in use, %.2f MB peak, %.2f MB reserved in %zu slab(s); process RSS %.1f MB",
           gpu ? " + bf16 copies" : "", total / 1048576.0, ws.bytes_in_use() / 1048576.0, ws.peak_in_use() / 1048576.0, ws.bytes_reserved() / 1048576.0, ws.num_slabs(),
           get_memory_usage_kb() / 1024.0);
  return out + b;
}
TrainHistory train_model(Graph& g, BaseDataLoader& train, BaseDataLoader* val, Optimizer& opt, Loss& criterion, Scheduler* sched,
                         const TrainingConfig& cfg, const std::function<void(const std::string&)>& log_in) {
  auto log = log_in ? log_in : [](const std::string& s) { std::cout << s << std::endl; };
  const Device& dev = g.device();
  g.set_training(true);
  train.set_pinned(dev.is_gpu());
  if (val) val->set_pinned(dev.is_gpu());
  // per-layer timing synchronises inside the step, so a profiled run is eager (no CUDA-graph capture)
  const ProfilerType ptype = string_to_profiler_type(cfg.profiler_type);
  const bool layer_prof = cfg.print_layer_profiling && ptype != ProfilerType::NONE;
  LayerProfiler& lprof = LayerProfiler::instance();
  if (layer_prof) {
    lprof.reset();
    lprof.enable(true);
  }
  struct ProfOff {
    bool on;
    ~ProfOff() {
      if (on) LayerProfiler::instance().enable(false);
    }
  } prof_off{layer_prof};
  TrainStep step(g, criterion, opt, sched, cfg.gradient_accumulation_steps, /*use_graph=*/!layer_prof);
  if (cfg.print_layer_memory_usage) log(layer_memory_table(g, step.executor()));
  TrainHistory hist;
  double best_acc = -1;
  uint64_t global_step = 0;
  mkdir(cfg.snapshot_dir.c_str(), 0755);
  const size_t epochs = cfg.max_steps > 0 ? (size_t)1 << 30 : cfg.epochs;
  for (size_t epoch = 0; epoch < epochs; ++epoch) {
    train.shuffle();
    train.reset();
    double loss_sum = 0, correct = 0;
    size_t seen = 0, batches = 0;
    const auto t_epoch = std::chrono::steady_clock::now();
    auto t_last = t_epoch;
    // two host batches in flight: batch k+1 is produced (and its H2D copy queued) while the device runs batch k
    Tensor xs[2], ys[2];
    Tensor pending_stats;
    size_t pending_n = 0;
    int slot = 0;
    auto drain = [&]() {
      if (!pending_stats) return;
      float st[2];
      pending_stats->copy_to_host(st, 8);
      loss_sum += st[0];
      correct += st[1];
      seen += pending_n;
      ++batches;
      pending_stats = nullptr;
    };
    while (train.get_batch(cfg.batch_size, xs[slot], ys[slot])) {
      Tensor stats = step(xs[slot], ys[slot]);
      drain();                                          // reads the PREVIOUS step's loss: no bubble on the current one
      pending_stats = stats;
      if (step.captured()) {
        // the captured graph writes one static stats buffer: read it before the next replay
        drain();
      }
      pending_n = ys[slot]->size();
      slot ^= 1;
      ++global_step;
      if (cfg.progress_print_interval > 0 && global_step % cfg.progress_print_interval == 0) {
        drain();
        const auto now = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(now - t_last).count() / (double)cfg.progress_print_interval;
        t_last = now;
        hist.step_ms.push_back(ms);
        char b[256];
        snprintf(b, sizeof b, "epoch %zu step %llu: loss %.4f acc %.2f%% %.2f ms/step lr %.2e", epoch + 1, (unsigned long long)global_step,
                 loss_sum / (double)std::max<size_t>(batches, 1), 100.0 * correct / (double)std::max<size_t>(seen, 1), ms, opt.get_learning_rate());
        log(b);
        if (layer_prof && ptype == ProfilerType::NORMAL) {
          log(lprof.table("layer profile of the last " + std::to_string(cfg.progress_print_interval) + " step(s):"));
          lprof.reset();
        }
      }
      if (cfg.max_steps > 0 && (long)global_step >= cfg.max_steps) break;
    }
    drain();
    hist.train_loss.push_back(loss_sum / (double)std::max<size_t>(batches, 1));
    {
      char b[128];
      snprintf(b, sizeof b, "epoch %zu done in %.2fs", epoch + 1, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_epoch).count());
      log(b);
    }
    if (val) {
      const auto va = validate_model(g, *val, criterion, cfg);
      hist.val_acc.push_back(va.second);
      char b[128];
      snprintf(b, sizeof b, "validation: loss %.4f acc %.2f%%", va.first, 100.0 * va.second);
      log(b);
      if (va.second > best_acc) {
        best_acc = va.second;
        save_checkpoint(g, cfg.snapshot_dir + "/" + g.name(), &opt, sched, global_step);
      }
    }
    if (cfg.max_steps > 0 && (long)global_step >= cfg.max_steps) {
      save_checkpoint(g, cfg.snapshot_dir + "/" + g.name(), &opt, sched, global_step);
      break;
    }
  }
  if (dev.is_gpu()) current_flow(dev).synchronize();
  if (layer_prof && ptype == ProfilerType::CUMULATIVE) log(lprof.table("cumulative layer profile of " + std::to_string(global_step) + " step(s):"));
  if (cfg.print_layer_memory_usage) log(layer_memory_table(g, step.executor()));
  return hist;
}

}  // namespace tnn
