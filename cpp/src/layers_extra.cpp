// LayerNorm / GroupNorm / embeddings / shape ops / joins / MSequential / attention / legacy NCHW facades + the
// LayerFactory registration table.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "kernels/attention.cuh"
#include "kernels/simt_ops.cuh"
#include "kernels/umma_gemm.cuh"
#include "runtime/cpu_math.hpp"
#include "tnn/layers.hpp"
#include "tnn/stepstate.hpp"

namespace tnn {

static std::string cfg_name(const TConfig& c, const char* def) { return c.name.empty() ? def : c.name; }

// ===================================================================================== LayerNorm
Vec<ParamDescriptor> LayerNormLayer::param_descriptors() const {
  if (!affine_) return {};
  return {{"gamma", DType_t::FP32, {C_}, true, const_init(1.0)}, {"beta", DType_t::FP32, {C_}, true, const_init(0.0)}};
}
std::unique_ptr<Layer> LayerNormLayer::create_from_config(const TConfig& c) {
  return std::make_unique<LayerNormLayer>(c.get<size_t>("normalized_shape"), c.get<double>("epsilon", 1e-5), c.get<bool>("affine", true), cfg_name(c, "layer_norm"));
}
Tensor LayerNormLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  Tensor x = as_compute(xin);
  const size_t rows = x->size() / C_;
  Tensor y = new_tensor(act_dtype(), x->shape());
  Tensor mean = new_tensor(DType_t::FP32, {rows}), rstd = new_tensor(DType_t::FP32, {rows});
  const float* g = affine_ ? param("gamma")->data_as<float>() : nullptr;
  const float* b = affine_ ? param("beta")->data_as<float>() : nullptr;
  if (device_.is_gpu()) {
    device_.set_current();
    layernorm_fwd(x->data(), nullptr, y->data(), nullptr, (long long)rows, (int)C_, g, b, (float)eps_, mean->data_as<float>(), rstd->data_as<float>(), stream());
  } else {
    const float* xp = x->data_as<float>();
    float* yp = y->data_as<float>();
    for (size_t r = 0; r < rows; ++r) {
      double s = 0, q = 0;
      for (size_t c = 0; c < C_; ++c) s += xp[r * C_ + c];
      const double m = s / (double)C_;
      for (size_t c = 0; c < C_; ++c) q += (xp[r * C_ + c] - m) * (xp[r * C_ + c] - m);
      const double rs = 1.0 / std::sqrt(q / (double)C_ + eps_);
      mean->data_as<float>()[r] = (float)m;
      rstd->data_as<float>()[r] = (float)rs;
      for (size_t c = 0; c < C_; ++c) {
        const float v = (float)((xp[r * C_ + c] - m) * rs);
        yp[r * C_ + c] = g ? v * g[c] + b[c] : v;
      }
    }
  }
  if (training_) {
    auto& c = cache(mb);
    c.t["x"] = x;
    c.t["mean"] = mean;
    c.t["rstd"] = rstd;
  }
  return y;
}
Tensor LayerNormLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux& aux) {
  auto& c = cache(mb);
  Tensor x = c.get("x"), mean = c.get("mean"), rstd = c.get("rstd");
  Tensor dy = as_compute(dyin);
  const size_t rows = x->size() / C_;
  Tensor dx = new_tensor(act_dtype(), x->shape());
  const float* g = affine_ ? param("gamma")->data_as<float>() : nullptr;
  float* dg = affine_ ? grad("gamma")->data_as<float>() : nullptr;
  float* db = affine_ ? grad("beta")->data_as<float>() : nullptr;
  if (device_.is_gpu()) {
    device_.set_current();
    layernorm_bwd(dy->data(), x->data(), (long long)rows, (int)C_, g, mean->data_as<float>(), rstd->data_as<float>(), dx->data(), dg, db,
                  aux.extra ? aux.extra->data() : nullptr, stream());
  } else {
    const float* xp = x->data_as<float>();
    const float* dyp = dy->data_as<float>();
    float* dxp = dx->data_as<float>();
    const float* ep = aux.extra ? aux.extra->data_as<float>() : nullptr;
    for (size_t r = 0; r < rows; ++r) {
      const float m = mean->data_as<float>()[r], rs = rstd->data_as<float>()[r];
      double s1 = 0, s2 = 0;
      for (size_t k = 0; k < C_; ++k) {
        const float xh = (xp[r * C_ + k] - m) * rs, gy = dyp[r * C_ + k] * (g ? g[k] : 1.f);
        s1 += gy;
        s2 += gy * xh;
        if (dg) {
          dg[k] += dyp[r * C_ + k] * xh;
          db[k] += dyp[r * C_ + k];
        }
      }
      for (size_t k = 0; k < C_; ++k) {
        const float xh = (xp[r * C_ + k] - m) * rs, gy = dyp[r * C_ + k] * (g ? g[k] : 1.f);
        dxp[r * C_ + k] = rs * (gy - (float)(s1 / (double)C_) - xh * (float)(s2 / (double)C_)) + (ep ? ep[r * C_ + k] : 0.f);
      }
    }
  }
  clear_cache(mb);
  return dx;
}

// ===================================================================================== GroupNorm (NHWC: [N, H, W, C] or [N, HW, C])
Vec<ParamDescriptor> GroupNormLayer::param_descriptors() const {
  if (!affine_) return {};
  return {{"gamma", DType_t::FP32, {C_}, true, const_init(1.0)}, {"beta", DType_t::FP32, {C_}, true, const_init(0.0)}};
}
std::unique_ptr<Layer> GroupNormLayer::create_from_config(const TConfig& c) {
  return std::make_unique<GroupNormLayer>(c.get<size_t>("num_groups"), c.get<size_t>("num_channels"), c.get<double>("epsilon", 1e-5), c.get<bool>("affine", true), cfg_name(c, "groupnorm"));
}
Tensor GroupNormLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  Tensor x = as_compute(xin);
  const size_t N = x->dimension(0), HW = x->size() / (N * C_), cg = C_ / G_;
  Tensor y = new_tensor(act_dtype(), x->shape());
  Tensor mean = new_tensor(DType_t::FP32, {N * G_}), rstd = new_tensor(DType_t::FP32, {N * G_});
  const float* g = affine_ ? param("gamma")->data_as<float>() : nullptr;
  const float* b = affine_ ? param("beta")->data_as<float>() : nullptr;
  if (device_.is_gpu()) {
    device_.set_current();
    groupnorm_fwd(x->data(), y->data(), (int)N, (int)HW, (int)C_, (int)G_, g, b, (float)eps_, mean->data_as<float>(), rstd->data_as<float>(), stream());
  } else {
    const float* xp = x->data_as<float>();
    float* yp = y->data_as<float>();
    for (size_t n = 0; n < N; ++n)
      for (size_t gi = 0; gi < G_; ++gi) {
        double s = 0, q = 0;
        for (size_t p = 0; p < HW; ++p)
          for (size_t k = 0; k < cg; ++k) s += xp[(n * HW + p) * C_ + gi * cg + k];
        const double m = s / (double)(HW * cg);
        for (size_t p = 0; p < HW; ++p)
          for (size_t k = 0; k < cg; ++k) { const double d = xp[(n * HW + p) * C_ + gi * cg + k] - m; q += d * d; }
        const double rs = 1.0 / std::sqrt(q / (double)(HW * cg) + eps_);
        mean->data_as<float>()[n * G_ + gi] = (float)m;
        rstd->data_as<float>()[n * G_ + gi] = (float)rs;
        for (size_t p = 0; p < HW; ++p)
          for (size_t k = 0; k < cg; ++k) {
            const size_t ch = gi * cg + k, i = (n * HW + p) * C_ + ch;
            const float v = (float)((xp[i] - m) * rs);
            yp[i] = g ? v * g[ch] + b[ch] : v;
          }
      }
  }
  if (training_) {
    auto& c = cache(mb);
    c.t["x"] = x;
    c.t["mean"] = mean;
    c.t["rstd"] = rstd;
  }
  return y;
}
Tensor GroupNormLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux&) {
  auto& c = cache(mb);
  Tensor x = c.get("x"), mean = c.get("mean"), rstd = c.get("rstd");
  Tensor dy = as_compute(dyin);
  const size_t N = x->dimension(0), HW = x->size() / (N * C_), cg = C_ / G_;
  Tensor dx = new_tensor(act_dtype(), x->shape());
  const float* g = affine_ ? param("gamma")->data_as<float>() : nullptr;
  float* dg = affine_ ? grad("gamma")->data_as<float>() : nullptr;
  float* db = affine_ ? grad("beta")->data_as<float>() : nullptr;
  if (device_.is_gpu()) {
    device_.set_current();
    groupnorm_bwd(dy->data(), x->data(), (int)N, (int)HW, (int)C_, (int)G_, g, mean->data_as<float>(), rstd->data_as<float>(), dx->data(), dg, db, stream());
  } else {
    const float* xp = x->data_as<float>();
    const float* dyp = dy->data_as<float>();
    float* dxp = dx->data_as<float>();
    const double cnt = (double)(HW * cg);
    for (size_t n = 0; n < N; ++n)
      for (size_t gi = 0; gi < G_; ++gi) {
        const float m = mean->data_as<float>()[n * G_ + gi], rs = rstd->data_as<float>()[n * G_ + gi];
        double s1 = 0, s2 = 0;
        for (size_t p = 0; p < HW; ++p)
          for (size_t k = 0; k < cg; ++k) {
            const size_t ch = gi * cg + k, i = (n * HW + p) * C_ + ch;
            const float xh = (xp[i] - m) * rs, gy = dyp[i] * (g ? g[ch] : 1.f);
            s1 += gy;
            s2 += gy * xh;
            if (dg) {
              dg[ch] += dyp[i] * xh;
              db[ch] += dyp[i];
            }
          }
        for (size_t p = 0; p < HW; ++p)
          for (size_t k = 0; k < cg; ++k) {
            const size_t ch = gi * cg + k, i = (n * HW + p) * C_ + ch;
            const float xh = (xp[i] - m) * rs, gy = dyp[i] * (g ? g[ch] : 1.f);
            dxp[i] = rs * (gy - (float)(s1 / cnt) - xh * (float)(s2 / cnt));
          }
      }
  }
  clear_cache(mb);
  return dx;
}

// ===================================================================================== embeddings
static Tensor ids_as_int32(const Tensor& ids, const Device& dev) {
  if (ids->data_type() == DType_t::INT32_T) return ids->device() == dev ? ids : ids->to_device(dev);
  // the reference carries token ids as floats (embedding_ops.cu:17-75): convert on the host
  Tensor h = ids->is_gpu() ? ids->to_host() : ids;
  Tensor o = make_tensor(DType_t::INT32_T, h->shape(), getHost());
  const auto v = to_float_vector(*h);
  for (size_t i = 0; i < v.size(); ++i) o->data_as<int32_t>()[i] = (int32_t)v[i];
  return dev.is_gpu() ? o->to_device(dev) : o;
}
Vec<ParamDescriptor> EmbeddingLayer::param_descriptors() const { return {{"weight", DType_t::FP32, {V_, D_}, true, normal_init(0.02)}}; }
std::unique_ptr<Layer> EmbeddingLayer::create_from_config(const TConfig& c) {
  return std::make_unique<EmbeddingLayer>(c.get<size_t>("vocab_size"), c.get<size_t>("embed_dim"), (long)c.get<long>("padding_idx", -1), cfg_name(c, "embedding"));
}
Tensor EmbeddingLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  Tensor ids = ids_as_int32(xin, device_);
  const size_t T = ids->size();
  Shape os = ids->shape();
  os.push_back(D_);
  Tensor y = new_tensor(act_dtype(), os);
  if (device_.is_gpu()) {
    device_.set_current();
    embedding_fwd(ids->data_as<int>(), param("weight")->data_as<float>(), nullptr, y->data(), (long long)T, (int)(os.size() >= 2 ? os[os.size() - 2] : 1), (int)D_, (int)V_, stream());
  } else {
    const float* w = param("weight")->data_as<float>();
    for (size_t t = 0; t < T; ++t) {
      const int id = ids->data_as<int32_t>()[t];
      if (id < 0 || (size_t)id >= V_ || id == pad_) std::memset(y->data_as<float>() + t * D_, 0, D_ * 4);
      else std::memcpy(y->data_as<float>() + t * D_, w + (size_t)id * D_, D_ * 4);
    }
  }
  if (training_) cache(mb).t["ids"] = ids;
  return y;
}
Tensor EmbeddingLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux&) {
  Tensor ids = cache(mb).get("ids");
  Tensor dy = as_compute(dyin);
  const size_t T = ids->size();
  if (device_.is_gpu()) {
    device_.set_current();
    embedding_bwd(ids->data_as<int>(), dy->data(), grad("weight")->data_as<float>(), nullptr, (long long)T, 1, (int)D_, (int)V_, stream());
  } else {
    float* dw = grad("weight")->data_as<float>();
    const float* dyp = dy->data_as<float>();
    for (size_t t = 0; t < T; ++t) {
      const int id = ids->data_as<int32_t>()[t];
      if (id < 0 || (size_t)id >= V_ || id == pad_) continue;
      for (size_t d = 0; d < D_; ++d) dw[(size_t)id * D_ + d] += dyp[t * D_ + d];
    }
  }
  clear_cache(mb);
  return nullptr;       // ids carry no gradient
}

Vec<ParamDescriptor> PositionalEmbeddingLayer::param_descriptors() const { return {{"weight", DType_t::FP32, {L_, D_}, true, normal_init(0.02)}}; }
std::unique_ptr<Layer> PositionalEmbeddingLayer::create_from_config(const TConfig& c) {
  return std::make_unique<PositionalEmbeddingLayer>(c.get<size_t>("embed_dim"), c.get<size_t>("seq_len"), cfg_name(c, "pos_embedding"));
}
Tensor PositionalEmbeddingLayer::forward_impl(const Tensor& xin, size_t, const FwdAux&) {
  Tensor x = as_compute(xin);
  const size_t B = x->dimension(0), L = x->dimension(1);
  if (L > L_) throw std::invalid_argument(name_ + ": sequence longer than the positional table");
  Tensor y = new_tensor(act_dtype(), x->shape());
  if (device_.is_gpu()) {
    device_.set_current();
    // y[b] = x[b] + table (bf16 view of the fp32 table): one binary kernel per batch row block
    const Tensor& t = lp("weight");
    for (size_t b = 0; b < B; ++b)
      add_bf16(static_cast<const char*>(x->data()) + b * L * D_ * 2, t->data(), static_cast<char*>(y->data()) + b * L * D_ * 2, (long long)(L * D_), 0, stream());
  } else {
    const float* w = param("weight")->data_as<float>();
    const float* xp = x->data_as<float>();
    float* yp = y->data_as<float>();
    for (size_t b = 0; b < B; ++b)
      for (size_t i = 0; i < L * D_; ++i) yp[b * L * D_ + i] = xp[b * L * D_ + i] + w[i];
  }
  return y;
}
Tensor PositionalEmbeddingLayer::backward_impl(const Tensor& dyin, size_t, BwdAux&) {
  Tensor dy = as_compute(dyin);
  const size_t B = dy->dimension(0), L = dy->dimension(1);
  float* dw = grad("weight")->data_as<float>();
  if (device_.is_gpu()) {
    device_.set_current();
    // dtable[l, :] += sum_b dy[b, l, :]: column sums of the [B, L*D] view
    colsum_bf16(dy->data(), (long long)B, (int)(L * D_), (long long)(L * D_), 1.f, dw, stream());
  } else {
    const float* dyp = dy->data_as<float>();
    for (size_t b = 0; b < B; ++b)
      for (size_t i = 0; i < L * D_; ++i) dw[i] += dyp[b * L * D_ + i];
  }
  return dy;
}

Vec<ParamDescriptor> ClassTokenLayer::param_descriptors() const { return {{"token", DType_t::FP32, {D_}, true, normal_init(0.02)}}; }
std::unique_ptr<Layer> ClassTokenLayer::create_from_config(const TConfig& c) { return std::make_unique<ClassTokenLayer>(c.get<size_t>("embed_dim"), cfg_name(c, "class_token")); }
Tensor ClassTokenLayer::forward_impl(const Tensor& xin, size_t, const FwdAux&) {
  Tensor x = as_compute(xin);
  const size_t B = x->dimension(0), L = x->dimension(1), es = get_dtype_size(x->data_type());
  Tensor y = new_tensor(act_dtype(), {B, L + 1, D_});
  if (device_.is_gpu()) {
    device_.set_current();
    const Tensor& t = lp("token");
    TNN_CUDA(cudaMemcpy2DAsync(y->data(), (L + 1) * D_ * es, t->data(), 0, D_ * es, B, cudaMemcpyDeviceToDevice, stream()));
    TNN_CUDA(cudaMemcpy2DAsync(static_cast<char*>(y->data()) + D_ * es, (L + 1) * D_ * es, x->data(), L * D_ * es, L * D_ * es, B, cudaMemcpyDeviceToDevice, stream()));
  } else {
    for (size_t b = 0; b < B; ++b) {
      std::memcpy(y->data_as<float>() + b * (L + 1) * D_, param("token")->data(), D_ * 4);
      std::memcpy(y->data_as<float>() + b * (L + 1) * D_ + D_, x->data_as<float>() + b * L * D_, L * D_ * 4);
    }
  }
  return y;
}
Tensor ClassTokenLayer::backward_impl(const Tensor& dyin, size_t, BwdAux&) {
  Tensor dy = as_compute(dyin);
  const size_t B = dy->dimension(0), L = dy->dimension(1) - 1, es = get_dtype_size(dy->data_type());
  Tensor dx = new_tensor(act_dtype(), {B, L, D_});
  float* dt = grad("token")->data_as<float>();
  if (device_.is_gpu()) {
    device_.set_current();
    colsum_bf16(dy->data(), (long long)B, (int)D_, (long long)((L + 1) * D_), 1.f, dt, stream());
    TNN_CUDA(cudaMemcpy2DAsync(dx->data(), L * D_ * es, static_cast<const char*>(dy->data()) + D_ * es, (L + 1) * D_ * es, L * D_ * es, B, cudaMemcpyDeviceToDevice, stream()));
  } else {
    for (size_t b = 0; b < B; ++b) {
      for (size_t d = 0; d < D_; ++d) dt[d] += dy->data_as<float>()[b * (L + 1) * D_ + d];
      std::memcpy(dx->data_as<float>() + b * L * D_, dy->data_as<float>() + b * (L + 1) * D_ + D_, L * D_ * 4);
    }
  }
  return dx;
}

// ===================================================================================== slice / transpose
std::unique_ptr<Layer> SliceLayer::create_from_config(const TConfig& c) {
  return std::make_unique<SliceLayer>(c.get<size_t>("axis"), c.get<size_t>("start"), c.get<size_t>("length"), cfg_name(c, "slice"));
}
// copies between x[outer, A, inner] and y[outer, len, inner] (start offset inside A)
static void slice_copy(const Tensor& big, const Tensor& small, size_t axis, size_t start, size_t len, bool big_to_small, cudaStream_t s) {
  const Shape& bs = big->shape();
  size_t outer = 1, inner = 1;
  for (size_t i = 0; i < axis; ++i) outer *= bs[i];
  for (size_t i = axis + 1; i < bs.size(); ++i) inner *= bs[i];
  const size_t es = get_dtype_size(big->data_type()), A = bs[axis];
  const size_t bpitch = A * inner * es, spitch = len * inner * es, width = len * inner * es;
  char* bp = static_cast<char*>(big->data()) + start * inner * es;
  char* sp = static_cast<char*>(small->data());
  if (big->is_gpu()) {
    if (big_to_small) TNN_CUDA(cudaMemcpy2DAsync(sp, spitch, bp, bpitch, width, outer, cudaMemcpyDeviceToDevice, s));
    else TNN_CUDA(cudaMemcpy2DAsync(bp, bpitch, sp, spitch, width, outer, cudaMemcpyDeviceToDevice, s));
  } else {
    for (size_t o = 0; o < outer; ++o) {
      if (big_to_small) std::memcpy(sp + o * spitch, bp + o * bpitch, width);
      else std::memcpy(bp + o * bpitch, sp + o * spitch, width);
    }
  }
}
Tensor SliceLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  Tensor x = as_compute(xin);
  Tensor y = new_tensor(x->data_type(), output_shape(x->shape()));
  if (device_.is_gpu()) device_.set_current();
  slice_copy(x, y, axis_, start_, len_, true, device_.is_gpu() ? stream() : nullptr);
  if (training_) cache(mb).sh["x"] = x->shape();
  return y;
}
Tensor SliceLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux&) {
  Tensor dy = as_compute(dyin);
  Tensor dx = new_tensor(dy->data_type(), cache(mb).sh.at("x"));
  dx->zero();
  if (device_.is_gpu()) device_.set_current();
  slice_copy(dx, dy, axis_, start_, len_, false, device_.is_gpu() ? stream() : nullptr);
  clear_cache(mb);
  return dx;
}
static Tensor transpose_last2(const Tensor& x, const std::function<Tensor(DType_t, const Shape&)>& alloc, const Device& dev, cudaStream_t s) {
  Shape os = x->shape();
  const size_t R = os[os.size() - 2], C = os[os.size() - 1], batch = x->size() / (R * C);
  std::swap(os[os.size() - 1], os[os.size() - 2]);
  Tensor y = alloc(x->data_type(), os);
  if (dev.is_gpu()) {
    for (size_t b = 0; b < batch; ++b)
      transpose_bf16(static_cast<const char*>(x->data()) + b * R * C * 2, static_cast<char*>(y->data()) + b * R * C * 2, (int)R, (int)C, s);
  } else {
    const float* xp = x->data_as<float>();
    float* yp = y->data_as<float>();
    for (size_t b = 0; b < batch; ++b)
      for (size_t r = 0; r < R; ++r)
        for (size_t c = 0; c < C; ++c) yp[b * R * C + c * R + r] = xp[b * R * C + r * C + c];
  }
  return y;
}
Tensor TransposeLayer::forward_impl(const Tensor& xin, size_t, const FwdAux&) {
  Tensor x = as_compute(xin);
  if (device_.is_gpu()) device_.set_current();
  return transpose_last2(x, [this](DType_t d, const Shape& s) { return new_tensor(d, s); }, device_, device_.is_gpu() ? stream() : nullptr);
}
Tensor TransposeLayer::backward_impl(const Tensor& dyin, size_t, BwdAux&) {
  Tensor dy = as_compute(dyin);
  if (device_.is_gpu()) device_.set_current();
  return transpose_last2(dy, [this](DType_t d, const Shape& s) { return new_tensor(d, s); }, device_, device_.is_gpu() ? stream() : nullptr);
}

// ===================================================================================== N-ary joins / broadcast
static void ew(const Tensor& a, const Tensor& b, Tensor& o, int op, const Device& dev, cudaStream_t s) {   // op 0 add 1 sub 2 mul
  if (dev.is_gpu()) binary_bf16(a->data(), b->data(), o->data(), (long long)a->size(), op, s);
  else {
    const float* ap = a->data_as<float>();
    const float* bp = b->data_as<float>();
    float* op_ = o->data_as<float>();
    for (size_t i = 0, n = a->size(); i < n; ++i) op_[i] = op == 0 ? ap[i] + bp[i] : op == 1 ? ap[i] - bp[i] : ap[i] * bp[i];
  }
}
Vec<Tensor> NAryLayer::forward(const Vec<Tensor>& xs, size_t mb) {
  if (xs.empty()) throw std::invalid_argument(name_ + ": no inputs");
  Vec<Tensor> in;
  for (auto& x : xs) in.push_back(as_compute(x));
  if (device_.is_gpu()) device_.set_current();
  Tensor acc = in[0];
  for (size_t i = 1; i < in.size(); ++i) {
    Tensor o = new_tensor(acc->data_type(), acc->shape());
    ew(acc, in[i], o, op_, device_, device_.is_gpu() ? stream() : nullptr);
    acc = o;
  }
  if (training_ && op_ == 2) {
    auto& c = cache(mb);
    for (size_t i = 0; i < in.size(); ++i) c.t["x" + std::to_string(i)] = in[i];
  }
  if (training_) cache(mb).s["n"] = (double)in.size();
  return {acc};
}
Vec<Tensor> NAryLayer::backward(const Vec<Tensor>& dys, size_t mb) {
  Tensor dy = as_compute(dys.at(0));
  auto& c = cache(mb);
  const size_t n = (size_t)c.s["n"];
  Vec<Tensor> out(n);
  if (device_.is_gpu()) device_.set_current();
  cudaStream_t s = device_.is_gpu() ? stream() : nullptr;
  for (size_t i = 0; i < n; ++i) {
    if (op_ == 0 || (op_ == 1 && i == 0)) out[i] = dy;
    else if (op_ == 1) {
      Tensor neg = new_tensor(dy->data_type(), dy->shape());
      if (device_.is_gpu()) tensor_unary(dy->data(), neg->data(), (long long)dy->size(), UOP_NEG, 0.f, 0.f, 0, s);
      else for (size_t k = 0; k < dy->size(); ++k) neg->data_as<float>()[k] = -dy->data_as<float>()[k];
      out[i] = neg;
    } else {
      Tensor g = dy;
      for (size_t j = 0; j < n; ++j) {
        if (j == i) continue;
        Tensor o = new_tensor(dy->data_type(), dy->shape());
        ew(g, c.get("x" + std::to_string(j)), o, 2, device_, s);
        g = o;
      }
      out[i] = g;
    }
  }
  clear_cache(mb);
  return out;
}
Vec<Tensor> MBroadcastLayer::forward(const Vec<Tensor>& xs, size_t) { return Vec<Tensor>(m_, xs.at(0)); }
Vec<Tensor> MBroadcastLayer::backward(const Vec<Tensor>& dys, size_t) {
  Tensor acc;
  if (device_.is_gpu()) device_.set_current();
  for (auto& d : dys) {
    if (!d) continue;
    Tensor g = as_compute(d);
    if (!acc) { acc = g; continue; }
    Tensor o = new_tensor(g->data_type(), g->shape());
    ew(acc, g, o, 0, device_, device_.is_gpu() ? stream() : nullptr);
    acc = o;
  }
  return {acc};
}

// ===================================================================================== MSequential
MSequential::MSequential(Vec<std::unique_ptr<Layer>> sequences, std::unique_ptr<Layer> join, const std::string& name)
    : Block(name), seqs_(std::move(sequences)), join_(std::move(join)) {}
Vec<Layer*> MSequential::children() {
  Vec<Layer*> v;
  for (auto& s : seqs_) v.push_back(s.get());
  v.push_back(join_.get());
  return v;
}
Tensor MSequential::forward_impl(const Tensor& x, size_t mb, const FwdAux&) {
  Vec<Tensor> outs(seqs_.size());
  auto* del = dynamic_cast<DELAllocatorV2*>(workspace_);
  if (order_.empty()) {
    // first call: measure each branch's peak and retained workspace and run the hungriest branch first from then on
    Vec<std::pair<long long, size_t>> score;
    for (size_t i = 0; i < seqs_.size(); ++i) {
      const size_t base = del ? del->bytes_in_use() : 0;
      const size_t peak0 = del ? del->peak_in_use() : 0;
      outs[i] = seqs_[i]->forward(x, mb);
      const size_t retained = del ? del->bytes_in_use() - base : outs[i]->nbytes();
      const size_t peak = del ? std::max(del->peak_in_use(), peak0) - base : retained;
      score.emplace_back(-(long long)(peak - std::min(peak, retained)), i);
    }
    std::stable_sort(score.begin(), score.end());
    for (auto& s : score) order_.push_back(s.second);
  } else {
    for (size_t i : order_) outs[i] = seqs_[i]->forward(x, mb);
  }
  return join_->forward(outs, mb).at(0);
}
Tensor MSequential::backward_impl(const Tensor& dy, size_t mb, BwdAux&) {
  Vec<Tensor> gs = join_->backward(Vec<Tensor>{dy}, mb);
  if (gs.size() == 1 && seqs_.size() > 1) gs = Vec<Tensor>(seqs_.size(), gs[0]);
  Tensor total;
  if (device_.is_gpu()) device_.set_current();
  for (size_t i = 0; i < seqs_.size(); ++i) {
    Tensor d = seqs_[i]->backward(gs[i], mb);
    if (!d) continue;
    if (!total) { total = d; continue; }
    Tensor o = new_tensor(d->data_type(), d->shape());
    ew(total, d, o, 0, device_, device_.is_gpu() ? stream() : nullptr);
    total = o;
  }
  return total;
}
TConfig MSequential::get_config() const {
  TConfig c(name_, TYPE_NAME);
  Vec<TConfig> ss;
  for (auto& s : seqs_) ss.push_back(s->get_config());
  c.set("sequences", ss);
  c.set("join_layer", join_->get_config());
  return c;
}
std::unique_ptr<Layer> MSequential::create_from_config(const TConfig& c) {
  Vec<std::unique_ptr<Layer>> ss;
  for (const auto& sc : c.get_config_array("sequences")) ss.push_back(LayerFactory::create(sc));
  return std::make_unique<MSequential>(std::move(ss), LayerFactory::create(c.get_config("join_layer")), cfg_name(c, "msequential"));
}

// ===================================================================================== attention
// CPU oracle: qkv [B, L, 3, H, D] fp32 -> o [B, L, H*D], lse [B, H, L]
static void cpu_attention_fwd(const float* qkv, float* o, float* lse, size_t B, size_t L, size_t H, size_t D, bool causal) {
  const float scale = 1.f / std::sqrt((float)D);
  const size_t tok = 3 * H * D;
  std::vector<float> sc(L);
  for (size_t b = 0; b < B; ++b)
    for (size_t h = 0; h < H; ++h)
      for (size_t i = 0; i < L; ++i) {
        const float* q = qkv + (b * L + i) * tok + h * D;
        const size_t lim = causal ? i + 1 : L;
        float mx = -INFINITY;
        for (size_t j = 0; j < lim; ++j) {
          const float* k = qkv + (b * L + j) * tok + H * D + h * D;
          float s = 0;
          for (size_t d = 0; d < D; ++d) s += q[d] * k[d];
          sc[j] = s * scale;
          mx = std::max(mx, sc[j]);
        }
        float sum = 0;
        for (size_t j = 0; j < lim; ++j) sum += (sc[j] = std::exp(sc[j] - mx));
        float* out = o + (b * L + i) * H * D + h * D;
        std::fill(out, out + D, 0.f);
        for (size_t j = 0; j < lim; ++j) {
          const float* v = qkv + (b * L + j) * tok + 2 * H * D + h * D;
          const float p = sc[j] / sum;
          for (size_t d = 0; d < D; ++d) out[d] += p * v[d];
        }
        lse[(b * H + h) * L + i] = mx + std::log(sum);
      }
}
static void cpu_attention_bwd(const float* qkv, const float* dO, const float* lse, float* dqkv, size_t B, size_t L, size_t H, size_t D, bool causal) {
  const float scale = 1.f / std::sqrt((float)D);
  const size_t tok = 3 * H * D;
  std::memset(dqkv, 0, B * L * tok * 4);
  std::vector<float> p(L), dp(L);
  for (size_t b = 0; b < B; ++b)
    for (size_t h = 0; h < H; ++h)
      for (size_t i = 0; i < L; ++i) {
        const float* q = qkv + (b * L + i) * tok + h * D;
        const float* go = dO + (b * L + i) * H * D + h * D;
        const size_t lim = causal ? i + 1 : L;
        float delta = 0;
        for (size_t j = 0; j < lim; ++j) {
          const float* k = qkv + (b * L + j) * tok + H * D + h * D;
          const float* v = qkv + (b * L + j) * tok + 2 * H * D + h * D;
          float s = 0, d_ = 0;
          for (size_t d = 0; d < D; ++d) {
            s += q[d] * k[d];
            d_ += go[d] * v[d];
          }
          p[j] = std::exp(s * scale - lse[(b * H + h) * L + i]);
          dp[j] = d_;
          delta += p[j] * d_;
        }
        float* dq = dqkv + (b * L + i) * tok + h * D;
        for (size_t j = 0; j < lim; ++j) {
          const float ds = p[j] * (dp[j] - delta) * scale;
          const float* k = qkv + (b * L + j) * tok + H * D + h * D;
          float* dk = dqkv + (b * L + j) * tok + H * D + h * D;
          float* dv = dqkv + (b * L + j) * tok + 2 * H * D + h * D;
          for (size_t d = 0; d < D; ++d) {
            dq[d] += ds * k[d];
            dk[d] += ds * q[d];
            dv[d] += p[j] * go[d];
          }
        }
      }
}

Vec<ParamDescriptor> AttentionBlock::param_descriptors() const {
  const double b = 1.0 / std::sqrt((double)E_);
  return {{"qkv_weight", DType_t::FP32, {3 * E_, E_}, true, uniform_init(b)}, {"qkv_bias", DType_t::FP32, {3 * E_}, true, uniform_init(b)},
          {"out_weight", DType_t::FP32, {E_, E_}, true, uniform_init(b)}, {"out_bias", DType_t::FP32, {E_}, true, uniform_init(b)}};
}
std::unique_ptr<Layer> AttentionBlock::create_attention(const TConfig& c) {
  return std::make_unique<AttentionBlock>(c.get<size_t>("embed_dim"), c.get<size_t>("num_heads"), c.get<bool>("is_causal", true), cfg_name(c, "attention_block"), false);
}
std::unique_ptr<Layer> AttentionBlock::create_flash(const TConfig& c) {
  return std::make_unique<AttentionBlock>(c.get<size_t>("embed_dim"), c.get<size_t>("num_heads"), c.get<bool>("is_causal", true), cfg_name(c, "flash_attention_block"), true);
}
Tensor AttentionBlock::forward_impl(const Tensor& xin, size_t mb, const FwdAux& aux) {
  Tensor x = as_compute(xin);
  const size_t B = x->dimension(0), L = x->dimension(1), D = E_ / H_, M = B * L;
  Tensor qkv = new_tensor(act_dtype(), {B, L, 3, H_, D});
  Tensor o = new_tensor(act_dtype(), {B, L, E_});
  Tensor lse = new_tensor(DType_t::FP32, {B, H_, L});
  Tensor y = new_tensor(act_dtype(), {B, L, E_});
  if (device_.is_gpu()) {
    if (D != 64) throw std::runtime_error(name_ + ": the fused attention kernels implement head_dim 64");
    device_.set_current();
    cudaStream_t s = stream();
    Epilogue e;
    e.out = qkv->data();
    e.ldc = (long long)(3 * E_);
    e.bias = param("qkv_bias")->data_as<float>();
    gemm_bf16(x->data(), lp("qkv_weight")->data(), (int)M, (int)(3 * E_), (int)E_, (long long)E_, (long long)E_, false, false, e, 1, s);
    const char* base = static_cast<const char*>(qkv->data());
    const long long tok = (long long)(3 * E_), bat = (long long)(L * 3 * E_);
    attention_fwd_bf16(base, tok, bat, base + E_ * 2, tok, bat, base + 2 * E_ * 2, tok, bat, o->data(), lse->data_as<float>(), (int)B, (int)H_, (int)L,
                       (int)L, causal_ ? 1 : 0, 0, s);
    Epilogue e2;
    e2.out = y->data();
    e2.ldc = (long long)E_;
    e2.bias = param("out_bias")->data_as<float>();
    e2.residual = aux.residual ? aux.residual->data() : nullptr;
    gemm_bf16(o->data(), lp("out_weight")->data(), (int)M, (int)E_, (int)E_, (long long)E_, (long long)E_, false, false, e2, 1, s);
  } else {
    float* qp = qkv->data_as<float>();
    cpu_sgemm(x->data_as<float>(), param("qkv_weight")->data_as<float>(), qp, (int64_t)M, (int64_t)(3 * E_), (int64_t)E_, false, true, 1.f, 0.f,
              (int64_t)E_, (int64_t)E_, (int64_t)(3 * E_));
    const float* qb = param("qkv_bias")->data_as<float>();
    for (size_t i = 0; i < M; ++i)
      for (size_t j = 0; j < 3 * E_; ++j) qp[i * 3 * E_ + j] += qb[j];
    cpu_attention_fwd(qp, o->data_as<float>(), lse->data_as<float>(), B, L, H_, D, causal_);
    float* yp = y->data_as<float>();
    cpu_sgemm(o->data_as<float>(), param("out_weight")->data_as<float>(), yp, (int64_t)M, (int64_t)E_, (int64_t)E_, false, true, 1.f, 0.f, (int64_t)E_,
              (int64_t)E_, (int64_t)E_);
    const float* ob = param("out_bias")->data_as<float>();
    const float* rp = aux.residual ? aux.residual->data_as<float>() : nullptr;
    for (size_t i = 0; i < M; ++i)
      for (size_t j = 0; j < E_; ++j) yp[i * E_ + j] += ob[j] + (rp ? rp[i * E_ + j] : 0.f);
  }
  if (training_) {
    auto& c = cache(mb);
    c.t["x"] = x;
    c.t["qkv"] = qkv;
    c.t["o"] = o;
    c.t["lse"] = lse;
  }
  return y;
}
Tensor AttentionBlock::backward_impl(const Tensor& dyin, size_t mb, BwdAux& aux) {
  auto& c = cache(mb);
  Tensor x = c.get("x"), qkv = c.get("qkv"), o = c.get("o"), lse = c.get("lse");
  Tensor dy = as_compute(dyin);
  const size_t B = x->dimension(0), L = x->dimension(1), D = E_ / H_, M = B * L;
  Tensor dO = new_tensor(act_dtype(), {B, L, E_});
  Tensor dqkv = new_tensor(act_dtype(), {B, L, 3, H_, D});
  Tensor dx = new_tensor(act_dtype(), {B, L, E_});
  if (device_.is_gpu()) {
    device_.set_current();
    cudaStream_t s = stream();
    auto wgrad = [&](const Tensor& g, size_t N, const Tensor& in, size_t K, float* dw) {
      Epilogue e;
      e.out = dw;
      e.out_mode = OUT_F32_ATOMIC;
      e.ldc = (long long)K;
      const int tiles = (int)(((N + 127) / 128) * ((K + 127) / 128));
      const int splits = std::max(1, std::min((int)((M + 63) / 64), (2 * 148) / std::max(1, tiles)));
      gemm_bf16(g->data(), in->data(), (int)N, (int)K, (int)M, (long long)N, (long long)K, true, true, e, splits, s);
    };
    wgrad(dy, E_, o, E_, grad("out_weight")->data_as<float>());
    colsum_bf16(dy->data(), (long long)M, (int)E_, (long long)E_, 1.f, grad("out_bias")->data_as<float>(), s);
    Epilogue e;
    e.out = dO->data();
    e.ldc = (long long)E_;
    gemm_bf16(dy->data(), lp("out_weight")->data(), (int)M, (int)E_, (int)E_, (long long)E_, (long long)E_, false, true, e, 1, s);
    Tensor delta = new_tensor(DType_t::FP32, {B, H_, L}), dq_acc = new_tensor(DType_t::FP32, {B, L, E_});
    const char* base = static_cast<const char*>(qkv->data());
    char* dbase = static_cast<char*>(dqkv->data());
    const long long tok = (long long)(3 * E_), bat = (long long)(L * 3 * E_);
    attention_bwd_bf16(base, base + E_ * 2, base + 2 * E_ * 2, tok, bat, dO->data(), o->data(), lse->data_as<float>(), delta->data_as<float>(),
                       dq_acc->data_as<float>(), dbase, dbase + E_ * 2, dbase + 2 * E_ * 2, tok, bat, (int)B, (int)H_, (int)L, causal_ ? 1 : 0, s);
    wgrad(dqkv, 3 * E_, x, E_, grad("qkv_weight")->data_as<float>());
    colsum_bf16(dqkv->data(), (long long)M, (int)(3 * E_), (long long)(3 * E_), 1.f, grad("qkv_bias")->data_as<float>(), s);
    Epilogue e3;
    e3.out = dx->data();
    e3.ldc = (long long)E_;
    e3.residual = aux.extra ? aux.extra->data() : nullptr;
    gemm_bf16(dqkv->data(), lp("qkv_weight")->data(), (int)M, (int)E_, (int)(3 * E_), (long long)(3 * E_), (long long)E_, false, true, e3, 1, s);
  } else {
    const float* dyp = dy->data_as<float>();
    cpu_sgemm(dyp, o->data_as<float>(), grad("out_weight")->data_as<float>(), (int64_t)E_, (int64_t)E_, (int64_t)M, true, false, 1.f, 1.f, (int64_t)E_, (int64_t)E_, (int64_t)E_);
    float* ob = grad("out_bias")->data_as<float>();
    for (size_t i = 0; i < M; ++i)
      for (size_t j = 0; j < E_; ++j) ob[j] += dyp[i * E_ + j];
    cpu_sgemm(dyp, param("out_weight")->data_as<float>(), dO->data_as<float>(), (int64_t)M, (int64_t)E_, (int64_t)E_, false, false, 1.f, 0.f, (int64_t)E_, (int64_t)E_, (int64_t)E_);
    cpu_attention_bwd(qkv->data_as<float>(), dO->data_as<float>(), lse->data_as<float>(), dqkv->data_as<float>(), B, L, H_, D, causal_);
    const float* dq = dqkv->data_as<float>();
    cpu_sgemm(dq, x->data_as<float>(), grad("qkv_weight")->data_as<float>(), (int64_t)(3 * E_), (int64_t)E_, (int64_t)M, true, false, 1.f, 1.f, (int64_t)(3 * E_), (int64_t)E_, (int64_t)E_);
    float* qb = grad("qkv_bias")->data_as<float>();
    for (size_t i = 0; i < M; ++i)
      for (size_t j = 0; j < 3 * E_; ++j) qb[j] += dq[i * 3 * E_ + j];
    cpu_sgemm(dq, param("qkv_weight")->data_as<float>(), dx->data_as<float>(), (int64_t)M, (int64_t)E_, (int64_t)(3 * E_), false, false, 1.f, 0.f, (int64_t)(3 * E_), (int64_t)E_, (int64_t)E_);
    if (aux.extra) dx->add(*aux.extra);
  }
  clear_cache(mb);
  return dx;
}

Tensor SDPALayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  Tensor qkv = as_compute(xin);
  const size_t B = qkv->dimension(0), L = qkv->dimension(1), H = qkv->dimension(3), D = qkv->dimension(4), E = H * D;
  Tensor o = new_tensor(act_dtype(), {B, L, E});
  Tensor lse = new_tensor(DType_t::FP32, {B, H, L});
  if (device_.is_gpu()) {
    if (D != 64) throw std::runtime_error(name_ + ": head_dim 64 only on the GPU");
    device_.set_current();
    const char* base = static_cast<const char*>(qkv->data());
    const long long tok = (long long)(3 * E), bat = (long long)(L * 3 * E);
    attention_fwd_bf16(base, tok, bat, base + E * 2, tok, bat, base + 2 * E * 2, tok, bat, o->data(), lse->data_as<float>(), (int)B, (int)H, (int)L, (int)L,
                       causal_ ? 1 : 0, 0, stream());
  } else {
    cpu_attention_fwd(qkv->data_as<float>(), o->data_as<float>(), lse->data_as<float>(), B, L, H, D, causal_);
  }
  if (training_) {
    auto& c = cache(mb);
    c.t["qkv"] = qkv;
    c.t["o"] = o;
    c.t["lse"] = lse;
  }
  return o;
}
Tensor SDPALayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux&) {
  auto& c = cache(mb);
  Tensor qkv = c.get("qkv"), o = c.get("o"), lse = c.get("lse");
  Tensor dO = as_compute(dyin);
  const size_t B = qkv->dimension(0), L = qkv->dimension(1), H = qkv->dimension(3), D = qkv->dimension(4), E = H * D;
  Tensor dqkv = new_tensor(act_dtype(), qkv->shape());
  if (device_.is_gpu()) {
    device_.set_current();
    Tensor delta = new_tensor(DType_t::FP32, {B, H, L}), dq_acc = new_tensor(DType_t::FP32, {B, L, E});
    const char* base = static_cast<const char*>(qkv->data());
    char* dbase = static_cast<char*>(dqkv->data());
    const long long tok = (long long)(3 * E), bat = (long long)(L * 3 * E);
    attention_bwd_bf16(base, base + E * 2, base + 2 * E * 2, tok, bat, dO->data(), o->data(), lse->data_as<float>(), delta->data_as<float>(),
                       dq_acc->data_as<float>(), dbase, dbase + E * 2, dbase + 2 * E * 2, tok, bat, (int)B, (int)H, (int)L, causal_ ? 1 : 0, stream());
  } else {
    cpu_attention_bwd(qkv->data_as<float>(), dO->data_as<float>(), lse->data_as<float>(), dqkv->data_as<float>(), B, L, H, D, causal_);
  }
  clear_cache(mb);
  return dqkv;
}

// ===================================================================================== legacy NCHW facades
Tensor nchw_to_nhwc_tensor(const Tensor& x, IAllocator* alloc) {
  const size_t N = x->dimension(0), C = x->dimension(1), H = x->dimension(2), W = x->dimension(3);
  Tensor y = alloc ? make_tensor(*alloc, x->data_type(), {N, H, W, C}) : make_tensor(x->data_type(), {N, H, W, C}, x->device());
  if (x->is_gpu()) {
    x->device().set_current();
    nchw_to_nhwc(x->data(), y->data(), (int)N, (int)C, (int)H, (int)W, x->data_type() == DType_t::FP32, current_flow(x->device()).stream());
  } else {
    const float* xp = x->data_as<float>();
    float* yp = y->data_as<float>();
    for (size_t n = 0; n < N; ++n)
      for (size_t c = 0; c < C; ++c)
        for (size_t p = 0; p < H * W; ++p) yp[(n * H * W + p) * C + c] = xp[(n * C + c) * H * W + p];
  }
  return y;
}
Tensor nhwc_to_nchw_tensor(const Tensor& x, IAllocator* alloc) {
  const size_t N = x->dimension(0), H = x->dimension(1), W = x->dimension(2), C = x->dimension(3);
  Tensor y = alloc ? make_tensor(*alloc, x->data_type(), {N, C, H, W}) : make_tensor(x->data_type(), {N, C, H, W}, x->device());
  if (x->is_gpu()) {
    x->device().set_current();
    nhwc_to_nchw(x->data(), y->data(), (int)N, (int)C, (int)H, (int)W, x->data_type() == DType_t::FP32, current_flow(x->device()).stream());
  } else {
    const float* xp = x->data_as<float>();
    float* yp = y->data_as<float>();
    for (size_t n = 0; n < N; ++n)
      for (size_t c = 0; c < C; ++c)
        for (size_t p = 0; p < H * W; ++p) yp[(n * C + c) * H * W + p] = xp[(n * H * W + p) * C + c];
  }
  return y;
}
LegacyNCHWLayer::LegacyNCHWLayer(std::unique_ptr<Layer> inner, const std::string& legacy_type)
    : Block(inner->name()), inner_(std::move(inner)), legacy_type_(legacy_type) {}
// legacy_batchnorm on the GPU: normalise the [N, C, H, W] tensor in place of its layout (norm_nchw.cu): statistics,
// finalize, apply = 3 launches; reduce, coefficients, apply = 3 launches backward; no transposes
Tensor LegacyNCHWLayer::bn_forward_nchw(BatchNormLayer& bn, const Tensor& x, size_t mb) {
  device_.set_current();
  cudaStream_t s = stream();
  const size_t N = x->shape()[0], C = x->shape()[1], HW = x->shape()[2] * x->shape()[3];
  if (!nchw_ws_) {
    nchw_ws_ = new_persistent(DType_t::FP64, {2 * C}, true);
    nchw_ss_ = new_persistent(DType_t::FP32, {2 * C}, true);
    nchw_coef_ = new_persistent(DType_t::FP32, {3 * C}, true);
  }
  auto& P = bn.params();
  const float* gamma = bn.affine() ? P.at("gamma")->data_as<float>() : nullptr;
  const float* beta = bn.affine() ? P.at("beta")->data_as<float>() : nullptr;
  float* rmean = P.at("running_mean")->data_as<float>();
  float* rvar = P.at("running_var")->data_as<float>();
  const int f32 = x->data_type() == DType_t::FP32;
  Tensor y = new_tensor(x->data_type(), x->shape());
  const float* ss = nchw_ss_->data_as<float>();
  if (training_) {
    Tensor rec = new_tensor(DType_t::FP32, {4 * C});                       // mean | invstd | (scale, shift) pairs
    float* r = rec->data_as<float>();
    bn_nchw_train_stats(x->data(), f32, (int)N, (int)C, (int)HW, gamma, beta, (float)bn.epsilon(), (float)bn.momentum(), rmean, rvar, r, r + C,
                        r + 2 * C, nchw_ws_->data(), s);
    ss = r + 2 * C;
    auto& c = cache(mb);
    c.t["x"] = x;
    c.t["rec"] = rec;
  } else {
    bn_nchw_eval_scale_shift((int)C, gamma, beta, rmean, rvar, (float)bn.epsilon(), nchw_ss_->data_as<float>(), s);
  }
  bn_nchw_apply(x->data(), y->data(), f32, (int)N, (int)C, (int)HW, ss, bn.use_relu() ? 1 : 0, s);
  return y;
}
Tensor LegacyNCHWLayer::bn_backward_nchw(BatchNormLayer& bn, const Tensor& dy, size_t mb) {
  device_.set_current();
  auto& c = cache(mb);
  Tensor x = c.t.at("x"), rec = c.t.at("rec");
  const size_t N = x->shape()[0], C = x->shape()[1], HW = x->shape()[2] * x->shape()[3];
  const float* r = rec->data_as<float>();
  Tensor dx = new_tensor(x->data_type(), x->shape());
  auto& G = bn.grads();
  bn_nchw_backward(dy->data(), x->data(), dx->data(), x->data_type() == DType_t::FP32, (int)N, (int)C, (int)HW,
                   bn.affine() ? bn.params().at("gamma")->data_as<float>() : nullptr, r, r + C, r + 2 * C, bn.use_relu() ? 1 : 0,
                   bn.affine() ? G.at("gamma")->data_as<float>() : nullptr, bn.affine() ? G.at("beta")->data_as<float>() : nullptr,
                   nchw_coef_->data_as<float>(), nchw_ws_->data(), stream());
  clear_cache(mb);
  return dx;
}
Tensor LegacyNCHWLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  Tensor x = as_compute(xin);
  if (x->dims() != 4) return inner_->forward(x, mb);
  if (device_.is_gpu()) {
    auto* bn = dynamic_cast<BatchNormLayer*>(inner_.get());
    if (bn && bn->fused_dropout_p == 0.0 && x->shape()[1] == bn->num_features()) return bn_forward_nchw(*bn, x, mb);
  }
  Tensor y = inner_->forward(nchw_to_nhwc_tensor(x, workspace_), mb);
  return y->dims() == 4 ? nhwc_to_nchw_tensor(y, workspace_) : y;
}
Tensor LegacyNCHWLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux&) {
  Tensor dy = as_compute(dyin);
  if (device_.is_gpu() && dy->dims() == 4) {
    auto* bn = dynamic_cast<BatchNormLayer*>(inner_.get());
    if (bn && cache(mb).get("rec")) return bn_backward_nchw(*bn, dy, mb);
  }
  Tensor dx = inner_->backward(dy->dims() == 4 ? nchw_to_nhwc_tensor(dy, workspace_) : dy, mb);
  if (!dx) return dx;
  return dx->dims() == 4 ? nhwc_to_nchw_tensor(dx, workspace_) : dx;
}
Shape LegacyNCHWLayer::output_shape(const Shape& s) const {
  if (s.size() != 4) return inner_->output_shape(s);
  const Shape o = inner_->output_shape({s[0], s[2], s[3], s[1]});
  return {o[0], o[3], o[1], o[2]};
}
TConfig LegacyNCHWLayer::get_config() const {
  TConfig c = inner_->get_config();
  c.type = legacy_type_;
  return c;
}

// ===================================================================================== factory table
void LayerFactory::register_defaults() {
  static bool done = false;
  if (done) return;
  done = true;
  register_type("dense", DenseLayer::create_from_config);
  register_type("legacy_dense", DenseLayer::create_from_config);
  register_type("conv2d", Conv2DLayer::create_from_config);
  register_type("batchnorm", BatchNormLayer::create_from_config);
  register_type("maxpool2d", Pool2DLayer::create_max);
  register_type("avgpool2d", Pool2DLayer::create_avg);
  register_type("dropout", DropoutLayer::create_from_config);
  register_type("flatten", FlattenLayer::create_from_config);
  register_type("activation", ActivationLayer::create_from_config);
  register_type("identity", IdentityLayer::create_from_config);
  register_type("layer_norm", LayerNormLayer::create_from_config);
  register_type("groupnorm", GroupNormLayer::create_from_config);
  register_type("embedding", EmbeddingLayer::create_from_config);
  register_type("pos_embedding", PositionalEmbeddingLayer::create_from_config);
  register_type("class_token", ClassTokenLayer::create_from_config);
  register_type("slice", SliceLayer::create_from_config);
  register_type("transpose", TransposeLayer::create_from_config);
  register_type("add", NAryLayer::create_add);
  register_type("sub", NAryLayer::create_sub);
  register_type("mul", NAryLayer::create_mul);
  register_type("mbroadcast", MBroadcastLayer::create_from_config);
  register_type("sdpa", SDPALayer::create_from_config);
  register_type("sequential", Sequential::create_from_config);
  register_type("msequential", MSequential::create_from_config);
  register_type("residual_block", ResidualBlock::create_from_config);
  register_type("attention_block", AttentionBlock::create_attention);
  register_type("flash_attention_block", AttentionBlock::create_flash);
  auto legacy = [](const char* inner_type, const char* legacy_type) {
    return [inner_type, legacy_type](const TConfig& c) -> std::unique_ptr<Layer> {
      TConfig ic = c;
      ic.type = inner_type;
      return std::make_unique<LegacyNCHWLayer>(LayerFactory::create(ic), legacy_type);
    };
  };
  register_type("legacy_conv2d", legacy("conv2d", "legacy_conv2d"));
  register_type("legacy_batchnorm", legacy("batchnorm", "legacy_batchnorm"));
  register_type("legacy_maxpool2d", legacy("maxpool2d", "legacy_maxpool2d"));
  register_type("legacy_avgpool2d", legacy("avgpool2d", "legacy_avgpool2d"));
}

}  // namespace tnn
