// Dense / Conv2D / BatchNorm / pooling / dropout / flatten / activation + Sequential / ResidualBlock.
// GPU: hand-written sm_100a kernels (tnn_b200/csrc/kernels); CPU: fp32 loops + packed SGEMM (the oracle).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <random>

#include "kernels/simt_ops.cuh"
#include "kernels/umma_gemm.cuh"
#include "tnn/profiler.hpp"
#include "runtime/cpu_math.hpp"
#include "tnn/layers.hpp"
#include "tnn/stepstate.hpp"

namespace tnn {

static inline size_t round8(size_t n) { return (n + 7) / 8 * 8; }
static uint64_t name_hash(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (char c : s) h = (h ^ (unsigned char)c) * 1099511628211ull;
  return h;
}
static std::string cfg_name(const TConfig& c, const char* def) { return c.name.empty() ? def : c.name; }

// [rows, cols] bf16 matrix whose row pitch must be a multiple of 8 elements for TMA: returns the pointer + pitch to
// use, copying into a padded workspace buffer only when cols % 8 != 0 (classifier heads: 10 / 100 / 200 classes).
struct PitchedView {
  const void* ptr;
  long long ld;
  Tensor keep;
};
static PitchedView pitched_bf16(const Tensor& t, size_t rows, size_t cols, const std::function<Tensor(DType_t, const Shape&)>& alloc,
                                cudaStream_t s) {
  if (cols % 8 == 0 && (reinterpret_cast<uintptr_t>(t->data()) & 15) == 0) return {t->data(), (long long)cols, nullptr};
  const size_t ld = round8(cols);
  Tensor buf = alloc(DType_t::BF16, {rows, ld});
  TNN_CUDA(cudaMemsetAsync(buf->data(), 0, buf->nbytes(), s));
  TNN_CUDA(cudaMemcpy2DAsync(buf->data(), ld * 2, t->data(), cols * 2, cols * 2, rows, cudaMemcpyDeviceToDevice, s));
  return {buf->data(), (long long)ld, buf};
}

// ===================================================================================== Dense
Vec<ParamDescriptor> DenseLayer::param_descriptors() const {
  const double b = 1.0 / std::sqrt((double)in_);
  Vec<ParamDescriptor> d{{"weight", DType_t::FP32, {out_, in_}, true, uniform_init(b)}};
  if (use_bias_) d.push_back({"bias", DType_t::FP32, {out_}, true, uniform_init(b)});
  return d;
}
void DenseLayer::fill_config(TConfig& c) const { c.set("input_features", in_).set("output_features", out_).set("use_bias", use_bias_); }
std::unique_ptr<Layer> DenseLayer::create_from_config(const TConfig& c) {
  return std::make_unique<DenseLayer>(c.get<size_t>("input_features"), c.get<size_t>("output_features"), c.get<bool>("use_bias", true), cfg_name(c, "dense"));
}
Tensor DenseLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux& aux) {
  Tensor x = as_compute(xin);
  const size_t M = x->size() / in_;
  Shape oshape = x->shape();
  oshape.back() = out_;
  Tensor y;
  if (device_.is_gpu()) {
    device_.set_current();
    cudaStream_t s = stream();
    auto alloc = [this](DType_t dt, const Shape& sh) { return new_tensor(dt, sh); };
    PitchedView xa = pitched_bf16(x, M, in_, alloc, s);
    y = new_tensor(DType_t::BF16, oshape);
    Epilogue e;
    e.out = y->data();
    e.out_mode = OUT_BF16;
    e.act = aux.act;
    e.bias = use_bias_ ? param("bias")->data_as<float>() : nullptr;
    e.residual = aux.residual ? aux.residual->data() : nullptr;
    e.ldc = (long long)out_;
    const Tensor& w = lp("weight");
    if (in_ % 8 != 0) throw std::runtime_error(name_ + ": GPU dense needs input_features % 8 == 0");
    if (M <= 32 && out_ % 8 == 0 && xa.ld % 8 == 0)
      skinny_gemm_bf16(xa.ptr, xa.ld, w->data(), (long long)in_, e.bias, e.residual, (long long)out_, y->data(), (long long)out_, (int)M,
                       (int)out_, (int)in_, aux.act, 0, s);
    else
      gemm_bf16(xa.ptr, w->data(), (int)M, (int)out_, (int)in_, xa.ld, (long long)in_, false, false, e, 1, s);
    if (training_) {
      auto& c = cache(mb);
      c.t["x"] = x;
      if (xa.keep) c.t["x_pad"] = xa.keep;
    }
    return y;
  }
  y = new_tensor(DType_t::FP32, oshape);
  const float* W = param("weight")->data_as<float>();
  cpu_sgemm(x->data_as<float>(), W, y->data_as<float>(), (int64_t)M, (int64_t)out_, (int64_t)in_, false, true, 1.f, 0.f, (int64_t)in_,
            (int64_t)in_, (int64_t)out_);
  float* yp = y->data_as<float>();
  const float* bp = use_bias_ ? param("bias")->data_as<float>() : nullptr;
  const float* rp = aux.residual ? aux.residual->data_as<float>() : nullptr;
  for (size_t i = 0; i < M; ++i)
    for (size_t j = 0; j < out_; ++j) {
      float v = yp[i * out_ + j] + (bp ? bp[j] : 0.f) + (rp ? rp[i * out_ + j] : 0.f);
      if (aux.act == ACT_RELU) v = std::max(v, 0.f);
      else if (aux.act == ACT_GELU) v = 0.5f * v * (1.f + std::tanh(0.7978845608028654f * (v + 0.044715f * v * v * v)));
      yp[i * out_ + j] = v;
    }
  if (training_) cache(mb).t["x"] = x;
  return y;
}
Tensor DenseLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux& aux) {
  auto& c = cache(mb);
  Tensor x = c.get("x");
  if (!x) throw std::runtime_error(name_ + ": backward without a cached forward (mb " + std::to_string(mb) + ")");
  Tensor dy = as_compute(dyin);
  const size_t M = x->size() / in_;
  Shape dshape = dy->shape();
  dshape.back() = in_;
  Tensor dx;
  if (device_.is_gpu()) {
    device_.set_current();
    cudaStream_t s = stream();
    auto alloc = [this](DType_t dt, const Shape& sh) { return new_tensor(dt, sh); };
    PitchedView dya = pitched_bf16(dy, M, out_, alloc, s);
    const void* xp = c.get("x_pad") ? c.get("x_pad")->data() : x->data();
    // dW[N,K] += dy^T x : both operands MN-major, split-K over the batch rows, fp32 atomic accumulate (beta = 1)
    {
      Epilogue e;
      e.out = grad("weight")->data();
      e.out_mode = OUT_F32_ATOMIC;
      e.ldc = (long long)in_;
      const int tiles = (int)(((out_ + 127) / 128) * ((in_ + 127) / 128));
      const int splits = std::max(1, std::min((int)((M + 63) / 64), (2 * 148) / std::max(1, tiles)));
      gemm_bf16(dya.ptr, xp, (int)out_, (int)in_, (int)M, dya.ld, (long long)round8(in_), true, true, e, splits, s);
    }
    if (use_bias_) {
      if (aux.colsum_in) axpy_f32(grad("bias")->data_as<float>(), aux.colsum_in->data_as<float>(), 1.f, (long long)out_, s);
      else colsum_bf16(dya.ptr, (long long)M, (int)out_, dya.ld, 1.f, grad("bias")->data_as<float>(), s);
    }
    if (aux.need_dx) {
      dx = new_tensor(DType_t::BF16, dshape);
      Epilogue e;
      e.out = dx->data();
      e.out_mode = OUT_BF16;
      e.ldc = (long long)in_;
      e.residual = aux.extra ? aux.extra->data() : nullptr;
      // dx[M,K] = dy[M,N] W[N,K]: W consumed MN-major (no transpose copy); needs out_features % 8 == 0 rows of 16 B
      const Tensor& w = lp("weight");
      gemm_bf16(dya.ptr, w->data(), (int)M, (int)in_, (int)out_, dya.ld, (long long)in_, false, true, e, 1, s);
    }
    clear_cache(mb);
    return dx;
  }
  const float* dyp = dy->data_as<float>();
  cpu_sgemm(dyp, x->data_as<float>(), grad("weight")->data_as<float>(), (int64_t)out_, (int64_t)in_, (int64_t)M, true, false, 1.f, 1.f,
            (int64_t)out_, (int64_t)in_, (int64_t)in_);
  if (use_bias_) {
    float* db = grad("bias")->data_as<float>();
    for (size_t i = 0; i < M; ++i)
      for (size_t j = 0; j < out_; ++j) db[j] += dyp[i * out_ + j];
  }
  if (aux.need_dx) {
    dx = new_tensor(DType_t::FP32, dshape);
    cpu_sgemm(dyp, param("weight")->data_as<float>(), dx->data_as<float>(), (int64_t)M, (int64_t)in_, (int64_t)out_, false, false, 1.f, 0.f,
              (int64_t)out_, (int64_t)in_, (int64_t)in_);
    if (aux.extra) dx->add(*aux.extra);
  }
  clear_cache(mb);
  return dx;
}

// ===================================================================================== Conv2D
static inline size_t conv_out(size_t h, size_t k, size_t s, size_t p) { return (h + 2 * p - k) / s + 1; }

struct ConvPlan {
  enum Path { REF, DIRECT, TMA, IM2COL } path = REF;
  int N, H, W, C, K, R, S, stride, pad, OH, OW;
  ConvGeom f_geom{};                        // fprop / wgrad geometry + taps
  struct DLaunch {
    ConvGeom geom;
    bool grid = false;
    int gOH, gOW;
    long long sN, sH, sW, base;
  };
  Vec<DLaunch> d_launches;                  // dgrad: one launch (stride 1) or one per output phase (stride 2)
  bool d_needs_zero = false;
};
static ConvGeom make_geom(int IN, int IH, int IW, int C, int ON, int OH, int OW, const Vec<ConvTap>& taps) {
  ConvGeom g{};
  g.IN = IN; g.IH = IH; g.IW = IW; g.C = C; g.ON = ON; g.OH = OH; g.OW = OW;
  g.ntaps = (int)taps.size();
  for (size_t i = 0; i < taps.size() && i < (size_t)kMaxTaps; ++i) g.taps[i] = taps[i];
  return g;
}
static std::unique_ptr<ConvPlan> make_conv_plan(int N, int H, int W, int C, int K, int R, int S, int stride, int pad, bool gpu) {
  auto p = std::make_unique<ConvPlan>();
  p->N = N; p->H = H; p->W = W; p->C = C; p->K = K; p->R = R; p->S = S; p->stride = stride; p->pad = pad;
  p->OH = (int)conv_out(H, R, stride, pad);
  p->OW = (int)conv_out(W, S, stride, pad);
  if (!gpu) return p;
  if (C < 8 && K % 8 == 0 && K * R * S * C <= 2048) {
    p->path = ConvPlan::DIRECT;
    return p;
  }
  p->path = ConvPlan::IM2COL;
  if (!(R * S <= 9 && (stride == 1 || stride == 2) && C % 8 == 0 && K % 8 == 0 && (R * S == 1 || (R * S) % 3 == 0))) return p;
  if (stride == 2 && (H % 2 || W % 2)) return p;
  Vec<ConvTap> taps;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      if (stride == 1) taps.push_back(ConvTap{r - pad, s - pad, 0, (r * S + s) * C});
      else {
        const int eh = r - pad, ew = s - pad;
        const int ph = ((eh % 2) + 2) % 2, pw = ((ew % 2) + 2) % 2;
        taps.push_back(ConvTap{(eh - ph) / 2, (ew - pw) / 2, (ph * 2 + pw) * N, (r * S + s) * C});
      }
    }
  ConvGeom geom = stride == 1 ? make_geom(N, H, W, C, N, p->OH, p->OW, taps) : make_geom(4 * N, H / 2, W / 2, C, N, p->OH, p->OW, taps);
  if (!conv_geom_supported(geom)) return p;
  Vec<ConvPlan::DLaunch> launches;
  bool needs_zero = false;
  if (stride == 1) {
    Vec<ConvTap> dt;
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) dt.push_back(ConvTap{r - (R - 1 - pad), s - (S - 1 - pad), 0, (r * S + s) * K});
    ConvPlan::DLaunch L;
    L.geom = make_geom(N, p->OH, p->OW, K, N, H, W, dt);
    if (!conv_geom_supported(L.geom)) return p;
    launches.push_back(L);
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        Vec<ConvTap> dt;
        for (int r = 0; r < R; ++r) {
          if (((ph + pad - r) % 2 + 2) % 2) continue;
          for (int s = 0; s < S; ++s) {
            if (((pw + pad - s) % 2 + 2) % 2) continue;
            const int rf = R - 1 - r, sf = S - 1 - s;
            // floor division (the shifts may be negative)
            auto fdiv2 = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };
            dt.push_back(ConvTap{fdiv2(ph + pad - r), fdiv2(pw + pad - s), 0, (rf * S + sf) * K});
          }
        }
        if (dt.empty()) {
          needs_zero = true;
          continue;
        }
        ConvPlan::DLaunch L;
        L.geom = make_geom(N, p->OH, p->OW, K, N, H / 2, W / 2, dt);
        if (!conv_geom_supported(L.geom)) return p;
        L.grid = true;
        L.gOH = H / 2;
        L.gOW = W / 2;
        L.sN = (long long)H * W * C;
        L.sH = 2ll * W * C;
        L.sW = 2ll * C;
        L.base = ((long long)ph * W + pw) * C;
        launches.push_back(L);
      }
  }
  p->path = ConvPlan::TMA;
  p->f_geom = geom;
  p->d_launches = std::move(launches);
  p->d_needs_zero = needs_zero;
  return p;
}

Conv2DLayer::Conv2DLayer(size_t in_ch, size_t out_ch, size_t kh, size_t kw, size_t sh, size_t sw, size_t ph, size_t pw, bool use_bias,
                         const std::string& name)
    : Layer(name), C_(in_ch), K_(out_ch), R_(kh), S_(kw), sh_(sh), sw_(sw), ph_(ph), pw_(pw), use_bias_(use_bias) {}
Conv2DLayer::~Conv2DLayer() = default;
Vec<ParamDescriptor> Conv2DLayer::param_descriptors() const {
  const double b = 1.0 / std::sqrt((double)(C_ * R_ * S_));
  Vec<ParamDescriptor> d{{"weight", DType_t::FP32, {K_, R_, S_, C_}, true, uniform_init(b)}};
  if (use_bias_) d.push_back({"bias", DType_t::FP32, {K_}, true, uniform_init(b)});
  return d;
}
void Conv2DLayer::fill_config(TConfig& c) const {
  c.set("in_channels", C_).set("out_channels", K_).set("kernel_h", R_).set("kernel_w", S_).set("stride_h", sh_).set("stride_w", sw_)
      .set("pad_h", ph_).set("pad_w", pw_).set("use_bias", use_bias_);
}
std::unique_ptr<Layer> Conv2DLayer::create_from_config(const TConfig& c) {
  return std::make_unique<Conv2DLayer>(c.get<size_t>("in_channels"), c.get<size_t>("out_channels"), c.get<size_t>("kernel_h"),
                                       c.get<size_t>("kernel_w"), c.get<size_t>("stride_h", 1), c.get<size_t>("stride_w", 1),
                                       c.get<size_t>("pad_h", 0), c.get<size_t>("pad_w", 0), c.get<bool>("use_bias", true), cfg_name(c, "conv2d"));
}
Shape Conv2DLayer::output_shape(const Shape& s) const { return {s[0], conv_out(s[1], R_, sh_, ph_), conv_out(s[2], S_, sw_, pw_), K_}; }
uint64_t Conv2DLayer::forward_flops(const Shape& s) const {
  const Shape o = output_shape(s);
  return 2ull * o[0] * o[1] * o[2] * o[3] * R_ * S_ * C_;
}
const ConvPlan& Conv2DLayer::plan_for(const Shape& x) {
  auto it = plans_.find(x);
  if (it == plans_.end()) {
    if (x.size() != 4 || x[3] != C_) throw std::invalid_argument(name_ + ": expected NHWC input with " + std::to_string(C_) + " channels, got " + shape_to_string(x));
    const bool iso = sh_ == sw_ && ph_ == pw_;
    // anisotropic stride / padding: the generic im2col path of the same kernels (no library fallback)
    auto p = make_conv_plan((int)x[0], (int)x[1], (int)x[2], (int)C_, (int)K_, (int)R_, (int)S_, (int)sh_, (int)ph_, device_.is_gpu() && iso);
    if (device_.is_gpu() && !iso) p->path = ConvPlan::IM2COL;
    p->OH = (int)conv_out(x[1], R_, sh_, ph_);
    p->OW = (int)conv_out(x[2], S_, sw_, pw_);
    it = plans_.emplace(x, std::move(p)).first;
  }
  return *it->second;
}

// ---- CPU reference: im2col + SGEMM
static void cpu_im2col(const float* x, float* col, int N, int H, int W, int C, int R, int S, int sh, int sw, int ph, int pw, int OH, int OW) {
  const size_t rsc = (size_t)R * S * C;
  for (int n = 0; n < N; ++n)
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow) {
        float* dst = col + (((size_t)n * OH + oh) * OW + ow) * rsc;
        for (int r = 0; r < R; ++r)
          for (int s = 0; s < S; ++s) {
            const int ih = oh * sh - ph + r, iw = ow * sw - pw + s;
            float* d = dst + ((size_t)r * S + s) * C;
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) std::memset(d, 0, (size_t)C * 4);
            else std::memcpy(d, x + (((size_t)n * H + ih) * W + iw) * C, (size_t)C * 4);
          }
      }
}
static void cpu_col2im(const float* col, float* dx, int N, int H, int W, int C, int R, int S, int sh, int sw, int ph, int pw, int OH, int OW) {
  const size_t rsc = (size_t)R * S * C;
  std::memset(dx, 0, (size_t)N * H * W * C * 4);
  for (int n = 0; n < N; ++n)
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow) {
        const float* src = col + (((size_t)n * OH + oh) * OW + ow) * rsc;
        for (int r = 0; r < R; ++r)
          for (int s = 0; s < S; ++s) {
            const int ih = oh * sh - ph + r, iw = ow * sw - pw + s;
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
            float* d = dx + (((size_t)n * H + ih) * W + iw) * C;
            const float* q = src + ((size_t)r * S + s) * C;
            for (int c = 0; c < C; ++c) d[c] += q[c];
          }
      }
}

Tensor Conv2DLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux& aux) {
  Tensor x = as_compute(xin);
  const ConvPlan& p = plan_for(x->shape());
  const size_t M = (size_t)p.N * p.OH * p.OW, rsc = R_ * S_ * C_;
  const Shape oshape{(size_t)p.N, (size_t)p.OH, (size_t)p.OW, K_};
  if (!device_.is_gpu()) {
    Tensor col = new_tensor(DType_t::FP32, {M, rsc});
    cpu_im2col(x->data_as<float>(), col->data_as<float>(), p.N, p.H, p.W, p.C, p.R, p.S, (int)sh_, (int)sw_, (int)ph_, (int)pw_, p.OH, p.OW);
    Tensor y = new_tensor(DType_t::FP32, oshape);
    cpu_sgemm(col->data_as<float>(), param("weight")->data_as<float>(), y->data_as<float>(), (int64_t)M, (int64_t)K_, (int64_t)rsc, false, true,
              1.f, 0.f, (int64_t)rsc, (int64_t)rsc, (int64_t)K_);
    float* yp = y->data_as<float>();
    const float* bp = use_bias_ ? param("bias")->data_as<float>() : nullptr;
    const float* rp = aux.residual ? aux.residual->data_as<float>() : nullptr;
    for (size_t i = 0; i < M; ++i)
      for (size_t k = 0; k < K_; ++k) {
        float v = yp[i * K_ + k] + (bp ? bp[k] : 0.f) + (rp ? rp[i * K_ + k] : 0.f);
        yp[i * K_ + k] = aux.act == ACT_RELU ? std::max(v, 0.f) : v;
      }
    if (training_) {
      cache(mb).t["col"] = col;
      cache(mb).sh["x"] = x->shape();
    }
    return y;
  }
  device_.set_current();
  cudaStream_t s = stream();
  Tensor y = new_tensor(DType_t::BF16, oshape);
  const float* bias = use_bias_ ? param("bias")->data_as<float>() : nullptr;
  auto& c = cache(mb);
  if (p.path == ConvPlan::DIRECT) {
    if (aux.residual || aux.act) throw std::runtime_error(name_ + ": the direct first-layer kernel has no epilogue fusion");
    conv_direct_fwd(x->data(), param("weight")->data_as<float>(), bias, y->data(), p.N, p.H, p.W, p.C, p.K, p.R, p.S, p.stride, p.pad, p.OH, p.OW, s);
  } else if (p.path == ConvPlan::TMA) {
    Tensor xin2 = x;
    if (p.stride == 2) {
      xin2 = new_tensor(DType_t::BF16, {(size_t)4 * p.N, (size_t)p.H / 2, (size_t)p.W / 2, C_});
      phase_split(x->data(), xin2->data(), p.N, p.H, p.W, p.C, s);
      if (training_) c.t["x_ps"] = xin2;
    }
    Epilogue e;
    e.out = y->data();
    e.out_mode = OUT_BF16;
    e.act = aux.act;
    e.bias = bias;
    e.residual = aux.residual ? aux.residual->data() : nullptr;
    e.ldc = (long long)K_;
    conv_fprop_bf16(xin2->data(), lp("weight")->data(), (long long)rsc, (int)K_, p.f_geom, e, s);
  } else {
    const size_t ld = round8(rsc);
    Tensor col = new_tensor(DType_t::BF16, {M, ld});
    if (ld != rsc) TNN_CUDA(cudaMemsetAsync(col->data(), 0, col->nbytes(), s));
    im2col_nhwc2(x->data(), col->data(), (long long)ld, p.N, p.H, p.W, p.C, p.R, p.S, (int)sh_, (int)sw_, (int)ph_, (int)pw_, p.OH, p.OW, s);
    Tensor wm = lp("weight");
    Tensor wpad;
    const void* wptr = wm->data();
    if (ld != rsc) {
      wpad = new_tensor(DType_t::BF16, {K_, ld});
      TNN_CUDA(cudaMemsetAsync(wpad->data(), 0, wpad->nbytes(), s));
      TNN_CUDA(cudaMemcpy2DAsync(wpad->data(), ld * 2, wm->data(), rsc * 2, rsc * 2, K_, cudaMemcpyDeviceToDevice, s));
      wptr = wpad->data();
    }
    Epilogue e;
    e.out = y->data();
    e.out_mode = OUT_BF16;
    e.act = aux.act;
    e.bias = bias;
    e.residual = aux.residual ? aux.residual->data() : nullptr;
    e.ldc = (long long)K_;
    gemm_bf16(col->data(), wptr, (int)M, (int)K_, (int)ld, (long long)ld, (long long)ld, false, false, e, 1, s);
    if (training_) c.t["col"] = col;
  }
  if (training_) c.t["x"] = x;
  else clear_cache(mb);
  return y;
}

Tensor Conv2DLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux& aux) {
  auto& c = cache(mb);
  Tensor dy = as_compute(dyin);
  const size_t rsc = R_ * S_ * C_;
  if (!device_.is_gpu()) {
    Tensor col = c.get("col");
    if (!col) throw std::runtime_error(name_ + ": backward without a cached forward");
    const size_t M = col->dimension(0);
    const ConvPlan* p = &plan_for(c.sh.at("x"));
    const float* dyp = dy->data_as<float>();
    cpu_sgemm(dyp, col->data_as<float>(), grad("weight")->data_as<float>(), (int64_t)K_, (int64_t)rsc, (int64_t)M, true, false, 1.f, 1.f,
              (int64_t)K_, (int64_t)rsc, (int64_t)rsc);
    if (use_bias_) {
      float* db = grad("bias")->data_as<float>();
      for (size_t i = 0; i < M; ++i)
        for (size_t k = 0; k < K_; ++k) db[k] += dyp[i * K_ + k];
    }
    Tensor dx;
    if (aux.need_dx) {
      Tensor dcol = new_tensor(DType_t::FP32, {M, rsc});
      cpu_sgemm(dyp, param("weight")->data_as<float>(), dcol->data_as<float>(), (int64_t)M, (int64_t)rsc, (int64_t)K_, false, false, 1.f, 0.f,
                (int64_t)K_, (int64_t)rsc, (int64_t)rsc);
      dx = new_tensor(DType_t::FP32, {(size_t)p->N, (size_t)p->H, (size_t)p->W, C_});
      cpu_col2im(dcol->data_as<float>(), dx->data_as<float>(), p->N, p->H, p->W, p->C, p->R, p->S, (int)sh_, (int)sw_, (int)ph_, (int)pw_, p->OH, p->OW);
      if (aux.extra) dx->add(*aux.extra);
    }
    clear_cache(mb);
    return dx;
  }
  Tensor x = c.get("x");
  if (!x) throw std::runtime_error(name_ + ": backward without a cached forward");
  const ConvPlan& p = plan_for(x->shape());
  const size_t M = (size_t)p.N * p.OH * p.OW;
  device_.set_current();
  cudaStream_t s = stream();
  float* db = use_bias_ ? grad("bias")->data_as<float>() : nullptr;
  if (db && aux.colsum_in && p.path != ConvPlan::DIRECT) {
    axpy_f32(db, aux.colsum_in->data_as<float>(), 1.f, (long long)K_, s);       // the bias gradient came fused out of the consumer's backward
    db = nullptr;
  }
  float* dw = grad("weight")->data_as<float>();
  if (p.path == ConvPlan::DIRECT) {
    conv_direct_wgrad(x->data(), dy->data(), dw, db, p.N, p.H, p.W, p.C, p.K, p.R, p.S, p.stride, p.pad, p.OH, p.OW, s);
    clear_cache(mb);
    return nullptr;
  }
  if (p.path == ConvPlan::TMA) {
    Tensor xps = p.stride == 2 ? c.get("x_ps") : x;
    const ConvGeom geom = p.f_geom;
    const int Kout = (int)K_;
    const long long ldw = (long long)rsc;
    auto wgrad = [=]() {
      cudaStream_t ws = current_flow(device_).stream();
      conv_wgrad_bf16(dy->data(), (long long)Kout, Kout, xps->data(), geom, dw, ldw, 1.f, ws);
      if (db) colsum_bf16(dy->data(), (long long)M, Kout, (long long)Kout, 1.f, db, ws);
    };
    // weight gradient on the side flow: overlaps the dgrad / BatchNorm backward that follow
    if (side_flow::enabled()) side_flow::run(device_, wgrad, {dy, xps});
    else wgrad();
  } else {
    Tensor col = c.get("col");
    const size_t ld = round8(rsc);
    const int tiles = (int)(((K_ + 127) / 128) * ((rsc + 127) / 128));
    const int splits = std::max(1, std::min((int)((M + 63) / 64), (2 * 148 + tiles - 1) / std::max(1, tiles)));
    Epilogue e;
    e.out = dw;
    e.out_mode = OUT_F32_ATOMIC;
    e.ldc = (long long)rsc;
    gemm_bf16(dy->data(), col->data(), (int)K_, (int)rsc, (int)M, (long long)K_, (long long)ld, true, true, e, splits, s);
    if (db) colsum_bf16(dy->data(), (long long)M, (int)K_, (long long)K_, 1.f, db, s);
  }
  Tensor dx;
  if (aux.need_dx) {
    dx = new_tensor(DType_t::BF16, x->shape());
    if (p.path == ConvPlan::TMA) {
      if (!wt_ || wt_version_ != param_version_) {
        if (!wt_) wt_ = new_persistent(DType_t::BF16, {C_, R_, S_, K_});
        conv_weight_prep(param("weight")->data_as<float>(), nullptr, wt_->data(), (int)K_, (int)R_, (int)S_, (int)C_, s);
        wt_version_ = param_version_;
      }
      if (p.d_needs_zero) {
        if (aux.extra) dx->copy_from(*aux.extra);     // phases without taps keep the shortcut gradient, the others overwrite
        else TNN_CUDA(cudaMemsetAsync(dx->data(), 0, dx->nbytes(), s));
      }
      for (const auto& L : p.d_launches) {
        Epilogue e;
        e.out = dx->data();
        e.out_mode = OUT_BF16;
        e.ldc = (long long)C_;
        e.residual = aux.extra ? aux.extra->data() : nullptr;
        if (L.grid) {
          e.grid_mode = 1;
          e.OH = L.gOH; e.OW = L.gOW; e.sN = L.sN; e.sH = L.sH; e.sW = L.sW; e.base = L.base;
        }
        conv_fprop_bf16(dy->data(), wt_->data(), (long long)(R_ * S_ * K_), (int)C_, L.geom, e, s);
      }
    } else {
      const size_t ld = round8(rsc);
      Tensor wm = lp("weight");
      Tensor wpad;
      const void* wptr = wm->data();
      if (ld != rsc) {
        wpad = new_tensor(DType_t::BF16, {K_, ld});
        TNN_CUDA(cudaMemsetAsync(wpad->data(), 0, wpad->nbytes(), s));
        TNN_CUDA(cudaMemcpy2DAsync(wpad->data(), ld * 2, wm->data(), rsc * 2, rsc * 2, K_, cudaMemcpyDeviceToDevice, s));
        wptr = wpad->data();
      }
      Tensor dcol = new_tensor(DType_t::BF16, {M, ld});
      Epilogue e;
      e.out = dcol->data();
      e.out_mode = OUT_BF16;
      e.ldc = (long long)ld;
      gemm_bf16(dy->data(), wptr, (int)M, (int)ld, (int)K_, (long long)K_, (long long)ld, false, true, e, 1, s);
      col2im_nhwc2(dcol->data(), (long long)ld, dx->data(), p.N, p.H, p.W, p.C, p.R, p.S, (int)sh_, (int)sw_, (int)ph_, (int)pw_, p.OH, p.OW, s);
      if (aux.extra) add_bf16(dx->data(), aux.extra->data(), dx->data(), (long long)dx->size(), 0, s);
    }
  }
  clear_cache(mb);
  return dx;
}

// ===================================================================================== BatchNorm
Vec<ParamDescriptor> BatchNormLayer::param_descriptors() const {
  Vec<ParamDescriptor> d;
  if (affine_) {
    d.push_back({"gamma", DType_t::FP32, {C_}, true, const_init(1.0)});
    d.push_back({"beta", DType_t::FP32, {C_}, true, const_init(0.0)});
  }
  d.push_back({"running_mean", DType_t::FP32, {C_}, false, const_init(0.0)});
  d.push_back({"running_var", DType_t::FP32, {C_}, false, const_init(1.0)});
  return d;
}
void BatchNormLayer::fill_config(TConfig& c) const {
  c.set("num_features", C_).set("epsilon", eps_).set("momentum", momentum_).set("affine", affine_).set("use_relu", relu_);
}
std::unique_ptr<Layer> BatchNormLayer::create_from_config(const TConfig& c) {
  return std::make_unique<BatchNormLayer>(c.get<size_t>("num_features"), c.get<double>("epsilon", 1e-5), c.get<double>("momentum", 0.1),
                                          c.get<bool>("affine", true), c.get<bool>("use_relu", false), cfg_name(c, "batchnorm"));
}
void BatchNormLayer::ensure_buffers() {
  if (ws_) return;
  const size_t n = (size_t)bn_ws_floats((int)C_);
  ws_ = new_persistent(DType_t::FP32, {n}, true);
  cws_ = new_persistent(DType_t::FP32, {n}, true);
  ss_ = new_persistent(DType_t::FP32, {2 * C_}, true);
  coef_ = new_persistent(DType_t::FP32, {3 * C_}, true);
}
Tensor BatchNormLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  Tensor x = as_compute(xin);
  if (x->shape().back() != C_) throw std::invalid_argument(name_ + ": expected " + std::to_string(C_) + " channels, got " + shape_to_string(x->shape()));
  const size_t M = x->size() / C_;
  const double p_drop = training_ ? fused_dropout_p : 0.0;
  const float* gamma = affine_ ? param("gamma")->data_as<float>() : nullptr;
  const float* beta = affine_ ? param("beta")->data_as<float>() : nullptr;
  float* rmean = param("running_mean")->data_as<float>();
  float* rvar = param("running_var")->data_as<float>();
  if (device_.is_gpu()) {
    device_.set_current();
    cudaStream_t s = stream();
    ensure_buffers();
    Tensor y = new_tensor(DType_t::BF16, x->shape());
    const uint64_t seed = (seed_ * 1000003ull + (name_hash(name_) & 0xFFFFFF) * 7919ull + mb) & 0x7FFFFFFFFFFFFFFFull;
    Tensor rec;
    if (training_) {
      // this microbatch's statistics record, written by the statistics kernel itself:
      // [0,4C) (invstd, -mean*invstd, scale, shift) per channel, [4C,5C) mean, [5C,6C) invstd
      rec = new_tensor(DType_t::FP32, {6 * C_});
      float* r = rec->data_as<float>();
      bn_train_stats(x->data(), (long long)M, (int)C_, gamma, beta, (float)eps_, (float)momentum_, rmean, rvar, r + 4 * C_, r + 5 * C_,
                     ss_->data_as<float>(), ws_->data_as<float>(), r, s);
    } else {
      bn_eval_scale_shift((int)C_, gamma, beta, rmean, rvar, (float)eps_, ss_->data_as<float>(), s);
    }
    Tensor bits;
    if (training_ && (relu_ || p_drop > 0) && C_ % 8 == 0) bits = new_tensor(DType_t::UINT8_T, {M * (C_ / 8)});   // keep bits: all backward needs of y
    bn_apply(x->data(), y->data(), (long long)M, (int)C_, ss_->data_as<float>(), relu_ ? 1 : 0, (float)p_drop, seed,
             p_drop > 0 ? StepState::get(device_).dev_seed_ptr() : nullptr, bits ? bits->data_as<uint8_t>() : nullptr, s);
    if (training_) {
      auto& c = cache(mb);
      c.t["x"] = x;
      c.t["rec"] = rec;
      if (bits) c.t["bits"] = bits;
      else if (relu_ || p_drop > 0) c.t["y"] = y;
      c.s["p"] = p_drop;
      c.u["seed"] = seed;
    }
    return y;
  }
  // ---- CPU oracle
  const float* xp = x->data_as<float>();
  Tensor mean = new_tensor(DType_t::FP32, {C_}), invstd = new_tensor(DType_t::FP32, {C_});
  float* mp = mean->data_as<float>();
  float* ip = invstd->data_as<float>();
  if (training_) {
    std::vector<double> sum(C_, 0.0), sq(C_, 0.0);
    for (size_t i = 0; i < M; ++i)
      for (size_t c = 0; c < C_; ++c) {
        const double v = xp[i * C_ + c];
        sum[c] += v;
        sq[c] += v * v;
      }
    for (size_t c = 0; c < C_; ++c) {
      const double m = sum[c] / (double)M;
      const double var = std::max(sq[c] / (double)M - m * m, 0.0);
      mp[c] = (float)m;
      ip[c] = (float)(1.0 / std::sqrt(var + eps_));
      rmean[c] = (float)((1 - momentum_) * rmean[c] + momentum_ * m);
      rvar[c] = (float)((1 - momentum_) * rvar[c] + momentum_ * var * ((double)M / (double)std::max<size_t>(M - 1, 1)));
    }
  } else {
    for (size_t c = 0; c < C_; ++c) {
      mp[c] = rmean[c];
      ip[c] = (float)(1.0 / std::sqrt((double)rvar[c] + eps_));
    }
  }
  Tensor y = new_tensor(DType_t::FP32, x->shape());
  float* yp = y->data_as<float>();
  std::mt19937_64 gen((seed_ * 1000003ull + cpu_step_ * 7919ull + mb) & 0x7FFFFFFFull);
  if (p_drop > 0) ++cpu_step_;
  std::uniform_real_distribution<float> uni(0.f, 1.f);
  const float dscale = p_drop > 0 ? (float)(1.0 / (1.0 - p_drop)) : 1.f;
  for (size_t i = 0; i < M; ++i)
    for (size_t c = 0; c < C_; ++c) {
      float v = (xp[i * C_ + c] - mp[c]) * ip[c];
      if (affine_) v = v * gamma[c] + beta[c];
      if (relu_) v = std::max(v, 0.f);
      if (p_drop > 0) v = uni(gen) >= (float)p_drop ? v * dscale : 0.f;
      yp[i * C_ + c] = v;
    }
  if (training_) {
    auto& c = cache(mb);
    c.t["x"] = x;
    c.t["y"] = y;
    c.t["mean"] = mean;
    c.t["invstd"] = invstd;
    c.s["p"] = p_drop;
  }
  return y;
}
Tensor BatchNormLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux& aux) {
  auto& c = cache(mb);
  Tensor x = c.get("x");
  if (!x) throw std::runtime_error(name_ + ": backward without a cached forward");
  Tensor dy = as_compute(dyin);
  const size_t M = x->size() / C_;
  const double p_drop = c.s["p"];
  const bool masked = relu_ || p_drop > 0;
  float* dgamma = affine_ ? grad("gamma")->data_as<float>() : nullptr;
  float* dbeta = affine_ ? grad("beta")->data_as<float>() : nullptr;
  const float* gamma = affine_ ? param("gamma")->data_as<float>() : nullptr;
  const float* beta = affine_ ? param("beta")->data_as<float>() : nullptr;
  if (device_.is_gpu()) {
    device_.set_current();
    cudaStream_t s = stream();
    ensure_buffers();
    Tensor rec = c.get("rec"), y = c.get("y"), bits = c.get("bits");
    const float* r = rec->data_as<float>();
    const uint64_t seed = c.u["seed"];
    const uint64_t* sp = p_drop > 0 ? StepState::get(device_).dev_seed_ptr() : nullptr;
    bn_bwd_reduce(dy->data(), x->data(), y ? y->data() : nullptr, (long long)M, (int)C_, gamma, r + 4 * C_, r + 5 * C_, masked ? 1 : 0, (float)p_drop,
                  dgamma, dbeta, coef_->data_as<float>(), ws_->data_as<float>(), beta, relu_ ? 1 : 0, seed, sp,
                  bits ? bits->data_as<uint8_t>() : nullptr, s);
    Tensor dx = new_tensor(DType_t::BF16, x->shape());
    Tensor colsum = new_tensor(DType_t::FP32, {C_});
    bn_bwd_apply(dy->data(), x->data(), y ? y->data() : nullptr, aux.extra ? aux.extra->data() : nullptr, dx->data(), (long long)M, (int)C_,
                 r + 4 * C_, r + 5 * C_, coef_->data_as<float>(), masked ? 1 : 0, (float)p_drop, colsum->data_as<float>(), cws_->data_as<float>(),
                 gamma, beta, relu_ ? 1 : 0, seed, sp, bits ? bits->data_as<uint8_t>() : nullptr, s);
    aux.colsum_out = colsum;          // column sums of dx = bias gradient of the producing conv / dense
    clear_cache(mb);
    return dx;
  }
  const float* xp = x->data_as<float>();
  const float* yp = c.get("y")->data_as<float>();
  const float* mp = c.get("mean")->data_as<float>();
  const float* ip = c.get("invstd")->data_as<float>();
  const float* dyp = dy->data_as<float>();
  const float dscale = p_drop > 0 ? (float)(1.0 / (1.0 - p_drop)) : 1.f;
  std::vector<double> sg(C_, 0.0), sgx(C_, 0.0);
  Tensor g = new_tensor(DType_t::FP32, x->shape());
  float* gp = g->data_as<float>();
  for (size_t i = 0; i < M; ++i)
    for (size_t ch = 0; ch < C_; ++ch) {
      float v = dyp[i * C_ + ch];
      if (masked) v = yp[i * C_ + ch] != 0.f ? v * dscale : 0.f;
      gp[i * C_ + ch] = v;
      sg[ch] += v;
      sgx[ch] += (double)v * (xp[i * C_ + ch] - mp[ch]) * ip[ch];
    }
  Tensor dx = new_tensor(DType_t::FP32, x->shape());
  float* dxp = dx->data_as<float>();
  for (size_t ch = 0; ch < C_; ++ch) {
    if (affine_) {
      dgamma[ch] += (float)sgx[ch];
      dbeta[ch] += (float)sg[ch];
    }
  }
  const float* ep = aux.extra ? aux.extra->data_as<float>() : nullptr;
  for (size_t i = 0; i < M; ++i)
    for (size_t ch = 0; ch < C_; ++ch) {
      const float xh = (xp[i * C_ + ch] - mp[ch]) * ip[ch];
      const float coef = (affine_ ? gamma[ch] : 1.f) * ip[ch];
      float v = coef * (gp[i * C_ + ch] - (float)(sg[ch] / (double)M) - xh * (float)(sgx[ch] / (double)M));
      if (ep) v += ep[i * C_ + ch];
      dxp[i * C_ + ch] = v;
    }
  clear_cache(mb);
  return dx;
}

// ===================================================================================== pooling
void Pool2DLayer::fill_config(TConfig& c) const {
  c.set("pool_h", kh_).set("pool_w", kw_).set("stride_h", sh_).set("stride_w", sw_).set("pad_h", ph_).set("pad_w", pw_);
}
Shape Pool2DLayer::output_shape(const Shape& s) const {
  return {s[0], (s[1] + 2 * ph_ - kh_) / sh_ + 1, (s[2] + 2 * pw_ - kw_) / sw_ + 1, s[3]};
}
std::unique_ptr<Layer> Pool2DLayer::create_max(const TConfig& c) {
  return std::make_unique<MaxPool2DLayer>(c.get<size_t>("pool_h"), c.get<size_t>("pool_w"), c.get<size_t>("stride_h", 0), c.get<size_t>("stride_w", 0),
                                          c.get<size_t>("pad_h", 0), c.get<size_t>("pad_w", 0), cfg_name(c, "maxpool2d"));
}
std::unique_ptr<Layer> Pool2DLayer::create_avg(const TConfig& c) {
  return std::make_unique<AvgPool2DLayer>(c.get<size_t>("pool_h"), c.get<size_t>("pool_w"), c.get<size_t>("stride_h", 0), c.get<size_t>("stride_w", 0),
                                          c.get<size_t>("pad_h", 0), c.get<size_t>("pad_w", 0), cfg_name(c, "avgpool2d"));
}
Tensor Pool2DLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  Tensor x = as_compute(xin);
  const Shape os = output_shape(x->shape());
  const int N = (int)x->dimension(0), H = (int)x->dimension(1), W = (int)x->dimension(2), C = (int)x->dimension(3), OH = (int)os[1], OW = (int)os[2];
  Tensor y = new_tensor(act_dtype(), os);
  if (device_.is_gpu()) {
    if (C % 8 != 0) throw std::runtime_error(name_ + ": GPU pooling needs channels % 8 == 0");
    device_.set_current();
    cudaStream_t s = stream();
    if (max_) {
      Tensor mask = training_ ? new_tensor(DType_t::UINT8_T, os) : nullptr;
      maxpool_fwd(x->data(), y->data(), mask ? mask->data_as<uint8_t>() : nullptr, N, H, W, C, (int)kh_, (int)kw_, (int)sh_, (int)sw_, (int)ph_,
                  (int)pw_, OH, OW, s);
      if (training_) cache(mb).t["mask"] = mask;
    } else {
      avgpool_fwd(x->data(), y->data(), N, H, W, C, (int)kh_, (int)kw_, (int)sh_, (int)sw_, (int)ph_, (int)pw_, OH, OW, s);
    }
    if (training_) cache(mb).sh["x"] = x->shape();
    return y;
  }
  const float* xp = x->data_as<float>();
  float* yp = y->data_as<float>();
  Tensor idx = max_ && training_ ? new_tensor(DType_t::INT32_T, os) : nullptr;
  for (int n = 0; n < N; ++n)
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow)
        for (int c = 0; c < C; ++c) {
          float best = -INFINITY, sum = 0.f;
          int arg = -1;
          for (int r = 0; r < (int)kh_; ++r)
            for (int q = 0; q < (int)kw_; ++q) {
              const int ih = oh * (int)sh_ - (int)ph_ + r, iw = ow * (int)sw_ - (int)pw_ + q;
              if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
              const float v = xp[(((size_t)n * H + ih) * W + iw) * C + c];
              sum += v;
              if (v > best) {
                best = v;
                arg = ih * W + iw;
              }
            }
          const size_t o = (((size_t)n * OH + oh) * OW + ow) * C + c;
          yp[o] = max_ ? best : sum / (float)(kh_ * kw_);
          if (idx) idx->data_as<int32_t>()[o] = arg;
        }
  if (training_) {
    if (idx) cache(mb).t["idx"] = idx;
    cache(mb).sh["x"] = x->shape();
  }
  return y;
}
Tensor Pool2DLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux&) {
  auto& c = cache(mb);
  Tensor dy = as_compute(dyin);
  const Shape& xs = c.sh.at("x");
  const int N = (int)xs[0], H = (int)xs[1], W = (int)xs[2], C = (int)xs[3];
  const int OH = (int)dy->dimension(1), OW = (int)dy->dimension(2);
  Tensor dx = new_tensor(act_dtype(), {(size_t)N, (size_t)H, (size_t)W, (size_t)C});
  if (device_.is_gpu()) {
    device_.set_current();
    cudaStream_t s = stream();
    if (max_) maxpool_bwd(dy->data(), c.get("mask")->data_as<uint8_t>(), dx->data(), N, H, W, C, (int)kh_, (int)kw_, (int)sh_, (int)sw_, (int)ph_, (int)pw_, OH, OW, s);
    else avgpool_bwd(dy->data(), dx->data(), N, H, W, C, (int)kh_, (int)kw_, (int)sh_, (int)sw_, (int)ph_, (int)pw_, OH, OW, s);
    clear_cache(mb);
    return dx;
  }
  const float* dyp = dy->data_as<float>();
  float* dxp = dx->data_as<float>();
  std::memset(dxp, 0, dx->nbytes());
  for (int n = 0; n < N; ++n)
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow)
        for (int ch = 0; ch < C; ++ch) {
          const size_t o = (((size_t)n * OH + oh) * OW + ow) * C + ch;
          if (max_) {
            const int arg = c.get("idx")->data_as<int32_t>()[o];
            if (arg >= 0) dxp[((size_t)n * H * W + arg) * C + ch] += dyp[o];
          } else {
            const float g = dyp[o] / (float)(kh_ * kw_);
            for (int r = 0; r < (int)kh_; ++r)
              for (int q = 0; q < (int)kw_; ++q) {
                const int ih = oh * (int)sh_ - (int)ph_ + r, iw = ow * (int)sw_ - (int)pw_ + q;
                if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                dxp[(((size_t)n * H + ih) * W + iw) * C + ch] += g;
              }
          }
        }
  clear_cache(mb);
  return dx;
}

// ===================================================================================== dropout
std::unique_ptr<Layer> DropoutLayer::create_from_config(const TConfig& c) { return std::make_unique<DropoutLayer>(c.get<double>("dropout_rate", 0.5), cfg_name(c, "dropout")); }
uint64_t DropoutLayer::layer_seed(size_t mb) const { return (seed_ * 1000003ull + (name_hash(name_) & 0xFFFFFF) * 7919ull + mb) & 0x7FFFFFFFFFFFFFFFull; }
Tensor DropoutLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux& aux) {
  if (fused_into_prev || !training_ || p_ <= 0) {
    if (!aux.residual) return xin;
    Tensor x = as_compute(xin);
    Tensor y = new_tensor(act_dtype(), x->shape());
    if (device_.is_gpu()) add_bf16(x->data(), aux.residual->data(), y->data(), (long long)x->size(), 0, stream());
    else {
      y->copy_from(*x);
      y->add(*aux.residual);
    }
    return y;
  }
  Tensor x = as_compute(xin);
  Tensor y = new_tensor(act_dtype(), x->shape());
  if (device_.is_gpu()) {
    device_.set_current();
    dropout_fwd(x->data(), y->data(), (long long)x->size(), (float)p_, layer_seed(mb), StepState::get(device_).dev_seed_ptr(),
                aux.residual ? aux.residual->data() : nullptr, stream());
    return y;
  }
  std::mt19937_64 gen((seed_ * 1000003ull + cpu_step_++ * 7919ull + mb) & 0x7FFFFFFFull);
  std::uniform_real_distribution<float> uni(0.f, 1.f);
  Tensor mask = new_tensor(DType_t::FP32, x->shape());
  const float sc = (float)(1.0 / (1.0 - p_));
  const float* xp = x->data_as<float>();
  float* yp = y->data_as<float>();
  float* mp = mask->data_as<float>();
  const float* rp = aux.residual ? aux.residual->data_as<float>() : nullptr;
  for (size_t i = 0, n = x->size(); i < n; ++i) {
    mp[i] = uni(gen) >= (float)p_ ? sc : 0.f;
    yp[i] = xp[i] * mp[i] + (rp ? rp[i] : 0.f);
  }
  cache(mb).t["mask"] = mask;
  return y;
}
Tensor DropoutLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux&) {
  if (fused_into_prev || !training_ || p_ <= 0) return dyin;
  Tensor dy = as_compute(dyin);
  Tensor dx = new_tensor(act_dtype(), dy->shape());
  if (device_.is_gpu()) {
    device_.set_current();
    // the mask is a pure function of (seed, step counter, index): recomputed, never stored
    dropout_fwd(dy->data(), dx->data(), (long long)dy->size(), (float)p_, layer_seed(mb), StepState::get(device_).dev_seed_ptr(), nullptr, stream());
    return dx;
  }
  const float* mp = cache(mb).get("mask")->data_as<float>();
  const float* dyp = dy->data_as<float>();
  float* dxp = dx->data_as<float>();
  for (size_t i = 0, n = dy->size(); i < n; ++i) dxp[i] = dyp[i] * mp[i];
  clear_cache(mb);
  return dx;
}

// ===================================================================================== flatten / activation
std::unique_ptr<Layer> FlattenLayer::create_from_config(const TConfig& c) {
  return std::make_unique<FlattenLayer>(c.get<int>("start_dim", 1), c.get<int>("end_dim", -1), cfg_name(c, "flatten"));
}
Shape FlattenLayer::output_shape(const Shape& s) const {
  const int nd = (int)s.size();
  const int a = start_ < 0 ? start_ + nd : start_, b = end_ < 0 ? end_ + nd : end_;
  Shape o(s.begin(), s.begin() + a);
  size_t prod = 1;
  for (int i = a; i <= b; ++i) prod *= s[(size_t)i];
  o.push_back(prod);
  o.insert(o.end(), s.begin() + b + 1, s.end());
  return o;
}
Tensor FlattenLayer::forward_impl(const Tensor& x, size_t mb, const FwdAux&) {
  if (training_) cache(mb).sh["x"] = x->shape();
  return x->reshape(output_shape(x->shape()));
}
Tensor FlattenLayer::backward_impl(const Tensor& dy, size_t mb, BwdAux&) {
  const Shape s = cache(mb).sh.at("x");
  clear_cache(mb);
  return dy->reshape(s);
}

ActivationLayer::ActivationLayer(const std::string& activation, const std::string& name) : Layer(name), act_(activation) {
  std::string a = activation;
  for (auto& ch : a) ch = (char)tolower((unsigned char)ch);
  act_ = a;
  if (a == "none" || a == "linear" || a.empty()) kind_ = AK_NONE;
  else if (a == "relu") kind_ = AK_RELU;
  else if (a == "leaky_relu") kind_ = AK_LEAKY_RELU;
  else if (a == "sigmoid") kind_ = AK_SIGMOID;
  else if (a == "tanh") kind_ = AK_TANH;
  else if (a == "elu") kind_ = AK_ELU;
  else if (a == "gelu") kind_ = AK_GELU;
  else if (a == "softmax") kind_ = -1;
  else throw std::invalid_argument("ActivationLayer: unknown activation '" + activation + "'");
  if (kind_ == AK_ELU) slope_ = 1.0f;
}
std::unique_ptr<Layer> ActivationLayer::create_from_config(const TConfig& c) {
  return std::make_unique<ActivationLayer>(c.get<std::string>("activation", "relu"), cfg_name(c, "activation"));
}
static float act_f(int k, float x, float slope) {
  switch (k) {
    case AK_RELU: return std::max(x, 0.f);
    case AK_LEAKY_RELU: return x > 0 ? x : slope * x;
    case AK_SIGMOID: return 1.f / (1.f + std::exp(-x));
    case AK_TANH: return std::tanh(x);
    case AK_ELU: return x > 0 ? x : slope * (std::exp(x) - 1.f);
    case AK_GELU: return 0.5f * x * (1.f + std::tanh(0.7978845608028654f * (x + 0.044715f * x * x * x)));
    default: return x;
  }
}
static float act_df(int k, float x, float slope) {
  switch (k) {
    case AK_RELU: return x > 0 ? 1.f : 0.f;
    case AK_LEAKY_RELU: return x > 0 ? 1.f : slope;
    case AK_SIGMOID: { const float s = 1.f / (1.f + std::exp(-x)); return s * (1.f - s); }
    case AK_TANH: { const float t = std::tanh(x); return 1.f - t * t; }
    case AK_ELU: return x > 0 ? 1.f : slope * std::exp(x);
    case AK_GELU: {
      const float k0 = 0.7978845608028654f, k1 = 0.044715f;
      const float u = k0 * (x + k1 * x * x * x), t = std::tanh(u);
      return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k0 * (1.f + 3.f * k1 * x * x);
    }
    default: return 1.f;
  }
}
Tensor ActivationLayer::forward_impl(const Tensor& xin, size_t mb, const FwdAux&) {
  if (kind_ == AK_NONE) return xin;
  Tensor x = as_compute(xin);
  Tensor y = new_tensor(act_dtype(), x->shape());
  const size_t C = x->shape().back(), rows = x->size() / C;
  if (device_.is_gpu()) {
    device_.set_current();
    if (kind_ == -1) softmax_rows(x->data(), y->data(), (long long)rows, (int)C, 0, stream());
    else act_fwd(x->data(), y->data(), (long long)x->size(), kind_, slope_, 0, stream());
  } else {
    const float* xp = x->data_as<float>();
    float* yp = y->data_as<float>();
    if (kind_ == -1) {
      for (size_t r = 0; r < rows; ++r) {
        float mx = -INFINITY, sum = 0.f;
        for (size_t c = 0; c < C; ++c) mx = std::max(mx, xp[r * C + c]);
        for (size_t c = 0; c < C; ++c) sum += (yp[r * C + c] = std::exp(xp[r * C + c] - mx));
        for (size_t c = 0; c < C; ++c) yp[r * C + c] /= sum;
      }
    } else {
      for (size_t i = 0, n = x->size(); i < n; ++i) yp[i] = act_f(kind_, xp[i], slope_);
    }
  }
  if (training_) cache(mb).t[kind_ == -1 ? "y" : "x"] = kind_ == -1 ? y : x;
  return y;
}
Tensor ActivationLayer::backward_impl(const Tensor& dyin, size_t mb, BwdAux&) {
  if (kind_ == AK_NONE) return dyin;
  Tensor dy = as_compute(dyin);
  Tensor dx = new_tensor(act_dtype(), dy->shape());
  auto& c = cache(mb);
  const size_t C = dy->shape().back(), rows = dy->size() / C;
  if (device_.is_gpu()) {
    device_.set_current();
    if (kind_ == -1) softmax_rows_bwd(c.get("y")->data(), dy->data(), dx->data(), (long long)rows, (int)C, 0, stream());
    else act_bwd(c.get("x")->data(), dy->data(), dx->data(), (long long)dy->size(), kind_, slope_, 0, stream());
  } else {
    const float* dyp = dy->data_as<float>();
    float* dxp = dx->data_as<float>();
    if (kind_ == -1) {
      const float* yp = c.get("y")->data_as<float>();
      for (size_t r = 0; r < rows; ++r) {
        float dot = 0.f;
        for (size_t k = 0; k < C; ++k) dot += yp[r * C + k] * dyp[r * C + k];
        for (size_t k = 0; k < C; ++k) dxp[r * C + k] = yp[r * C + k] * (dyp[r * C + k] - dot);
      }
    } else {
      const float* xp = c.get("x")->data_as<float>();
      for (size_t i = 0, n = dy->size(); i < n; ++i) dxp[i] = dyp[i] * act_df(kind_, xp[i], slope_);
    }
  }
  clear_cache(mb);
  return dx;
}

// ===================================================================================== Sequential
Sequential::Sequential(Vec<std::unique_ptr<Layer>> layers, const std::string& name) : Block(name), layers_(std::move(layers)) { plan_fusion(); }
void Sequential::add(std::unique_ptr<Layer> l) {
  layers_.push_back(std::move(l));
  plan_fusion();
}
Vec<Layer*> Sequential::children() {
  Vec<Layer*> v;
  for (auto& l : layers_) v.push_back(l.get());
  return v;
}
void Sequential::plan_fusion() {
  // BatchNorm(+ReLU) directly followed by Dropout -> ONE apply kernel (statistics-normalise-ReLU-dropout-keep-bits)
  for (size_t i = 0; i + 1 < layers_.size(); ++i) {
    auto* bn = dynamic_cast<BatchNormLayer*>(layers_[i].get());
    auto* dr = dynamic_cast<DropoutLayer*>(layers_[i + 1].get());
    if (bn && dr && dr->rate() > 0) {
      bn->fused_dropout_p = dr->rate();
      dr->fused_into_prev = true;
    }
  }
}
Tensor Sequential::forward_impl(const Tensor& xin, size_t mb, const FwdAux& aux) {
  Tensor x = xin;
  auto* del = dynamic_cast<DELAllocatorV2*>(workspace_);
  const size_t n = layers_.size();
  for (size_t i = 0; i < n; ++i) {
    const bool last = i + 1 == n;
    NvtxRange range("fwd:", layers_[i]->name());
    LayerTimer timer(layers_[i]->name(), device_, false);
    if (last && (aux.residual || aux.act) && layers_[i]->fuses_fwd_residual()) x = layers_[i]->forward_impl(x, mb, aux);
    else x = layers_[i]->forward(x, mb);
    if (del) del->flip();            // consecutive layer outputs come from opposite ends of the workspace (sequential.cpp:33-44)
  }
  if ((aux.residual || aux.act) && (n == 0 || !layers_.back()->fuses_fwd_residual()))
    throw std::logic_error(name_ + ": fused residual requested but the last layer cannot take it");
  return x;
}
Tensor Sequential::backward_impl(const Tensor& dyin, size_t mb, BwdAux& aux) {
  Tensor dy = dyin;
  Tensor colsum = aux.colsum_in;
  auto* del = dynamic_cast<DELAllocatorV2*>(workspace_);
  for (size_t k = layers_.size(); k-- > 0;) {
    Layer* layer = layers_[k].get();
    if (!dy) {
      if (backward_hook) backward_hook(k);
      continue;
    }
    NvtxRange range("bwd:", layer->name());
    LayerTimer timer(layer->name(), device_, true);
    BwdAux a;
    a.colsum_in = colsum;
    a.need_dx = k > 0 || aux.need_dx;
    bool extra_pending = k == 0 && aux.extra != nullptr;
    if (extra_pending && layer->fuses_bwd_extra()) {
      a.extra = aux.extra;
      extra_pending = false;
    }
    dy = layer->backward_impl(dy, mb, a);
    colsum = a.colsum_out;
    if (extra_pending && dy) {
      // the first layer cannot fold the shortcut gradient into its kernel: add it here
      Tensor sum = new_tensor(dy->data_type(), dy->shape());
      if (device_.is_gpu()) add_bf16(dy->data(), aux.extra->data(), sum->data(), (long long)dy->size(), 0, stream());
      else {
        sum->copy_from(*dy);
        sum->add(*aux.extra);
      }
      dy = sum;
      colsum = nullptr;
    }
    if (del) del->flip();
    if (backward_hook) backward_hook(k);
  }
  aux.colsum_out = colsum;
  return dy;
}
Shape Sequential::output_shape(const Shape& s) const {
  Shape o = s;
  for (auto& l : layers_) o = l->output_shape(o);
  return o;
}
uint64_t Sequential::forward_flops(const Shape& s) const {
  uint64_t total = 0;
  Shape o = s;
  for (auto& l : layers_) {
    total += l->forward_flops(o);
    o = l->output_shape(o);
  }
  return total;
}
TConfig Sequential::get_config() const {
  TConfig c(name_, TYPE_NAME);
  Vec<TConfig> ls;
  for (auto& l : layers_) ls.push_back(l->get_config());
  c.set("layers", ls);
  return c;
}
std::unique_ptr<Layer> Sequential::create_from_config(const TConfig& c) {
  Vec<std::unique_ptr<Layer>> ls;
  for (const auto& lc : c.get_config_array("layers")) ls.push_back(LayerFactory::create(lc));
  return std::make_unique<Sequential>(std::move(ls), cfg_name(c, "sequential"));
}
void Sequential::print_summary(const Shape& input_shape) const {
  Shape s = input_shape;
  std::cout << "Sequential '" << name_ << "'\n";
  for (auto& l : layers_) {
    const Shape o = l->output_shape(s);
    std::cout << "  " << l->type() << " (" << l->name() << ")  " << shape_to_string(s) << " -> " << shape_to_string(o) << "\n";
    s = o;
  }
}

// ===================================================================================== ResidualBlock
ResidualBlock::ResidualBlock(std::unique_ptr<Layer> main_path, std::unique_ptr<Layer> shortcut_path, const std::string& activation,
                             const std::string& name)
    : Block(name), main_(std::move(main_path)), shortcut_(std::move(shortcut_path)), act_(activation.empty() ? "linear" : activation) {
  for (auto& ch : act_) ch = (char)tolower((unsigned char)ch);
  if (act_ != "linear" && act_ != "none" && act_ != "relu") act_layer_ = std::make_unique<ActivationLayer>(act_, name + "_act");
}
Vec<Layer*> ResidualBlock::children() {
  Vec<Layer*> v{main_.get()};
  if (shortcut_) v.push_back(shortcut_.get());
  if (act_layer_) v.push_back(act_layer_.get());
  return v;
}
Tensor ResidualBlock::forward_impl(const Tensor& x, size_t mb, const FwdAux&) {
  Tensor sc = shortcut_ ? shortcut_->forward(x, mb) : as_compute(x);
  const bool relu = act_ == "relu";
  Tensor y;
  if (main_->fuses_fwd_residual() && !act_layer_) {
    FwdAux a;
    a.residual = sc;
    a.act = relu ? ACT_RELU : ACT_NONE;
    y = main_->forward_impl(x, mb, a);         // the shortcut is added in the last GEMM's epilogue
  } else {
    Tensor m = main_->forward(x, mb);
    y = new_tensor(m->data_type(), m->shape());
    if (device_.is_gpu()) add_bf16(m->data(), sc->data(), y->data(), (long long)m->size(), relu ? 1 : 0, stream());
    else {
      y->copy_from(*m);
      y->add(*sc);
      if (relu) {
        float* p = y->data_as<float>();
        for (size_t i = 0, n = y->size(); i < n; ++i) p[i] = std::max(p[i], 0.f);
      }
    }
    if (act_layer_) y = act_layer_->forward(y, mb);
  }
  if (training_ && relu) cache(mb).t["y"] = y;
  return y;
}
Tensor ResidualBlock::backward_impl(const Tensor& dyin, size_t mb, BwdAux& aux) {
  Tensor dy = as_compute(dyin);          // the identity shortcut hands dy on as `extra`: it must be in the compute dtype
  Tensor colsum = aux.colsum_in;
  if (act_ == "relu") {
    Tensor y = cache(mb).get("y");
    Tensor d = dy;
    Tensor g = new_tensor(d->data_type(), d->shape());
    if (device_.is_gpu()) act_bwd(y->data(), d->data(), g->data(), (long long)d->size(), AK_RELU, 0.f, 0, stream());
    else {
      const float* yp = y->data_as<float>();
      const float* dp = d->data_as<float>();
      float* gp = g->data_as<float>();
      for (size_t i = 0, n = d->size(); i < n; ++i) gp[i] = yp[i] > 0 ? dp[i] : 0.f;
    }
    dy = g;
    colsum = nullptr;
    clear_cache(mb);
  } else if (act_layer_) {
    dy = act_layer_->backward(dy, mb);
    colsum = nullptr;
  }
  Tensor d_short = dy;
  if (shortcut_) {
    BwdAux sa;
    sa.need_dx = aux.need_dx;
    d_short = shortcut_->backward_impl(dy, mb, sa);
  }
  BwdAux ma;
  ma.colsum_in = colsum;
  ma.need_dx = aux.need_dx;
  const bool main_takes_extra = dynamic_cast<Sequential*>(main_.get()) != nullptr || main_->fuses_bwd_extra();
  if (main_takes_extra) ma.extra = d_short;      // folded into the first layer's backward kernel (Sequential adds it itself otherwise)
  Tensor dx = main_->backward_impl(dy, mb, ma);
  aux.colsum_out = ma.colsum_out;
  if (!main_takes_extra && dx && d_short) {
    Tensor sum = new_tensor(dx->data_type(), dx->shape());
    if (device_.is_gpu()) add_bf16(dx->data(), d_short->data(), sum->data(), (long long)dx->size(), 0, stream());
    else {
      sum->copy_from(*dx);
      sum->add(*d_short);
    }
    dx = sum;
    aux.colsum_out = nullptr;
  }
  return dx;
}
TConfig ResidualBlock::get_config() const {
  TConfig c(name_, TYPE_NAME);
  c.set("activation", act_);
  c.set("main_path", main_->get_config());
  if (shortcut_) c.set("shortcut_path", shortcut_->get_config());
  else c.set("shortcut_path", Json::object());
  return c;
}
std::unique_ptr<Layer> ResidualBlock::create_from_config(const TConfig& c) {
  auto main = LayerFactory::create(c.get_config("main_path"));
  std::unique_ptr<Layer> sc;
  const Json* sj = c.raw("shortcut_path");
  if (sj && sj->is_object() && !sj->empty() && sj->contains("type")) sc = LayerFactory::create(TConfig::from_json(*sj));
  return std::make_unique<ResidualBlock>(std::move(main), std::move(sc), c.get<std::string>("activation", "relu"), cfg_name(c, "residual_block"));
}

}  // namespace tnn
