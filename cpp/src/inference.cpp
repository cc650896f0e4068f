// GPT2Engine: see tnn/inference.hpp.  GPU path = hand-written kernels on engine-owned buffers (graph replayable);
// CPU path = the same cached algorithm in fp32.
#include "tnn/inference.hpp"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstring>

#include "kernels/attention.cuh"
#include "kernels/decode_mega.cuh"
#include "kernels/mxfp8.cuh"
#include "kernels/simt_ops.cuh"
#include "kernels/umma_gemm.cuh"
#include "runtime/cpu_math.hpp"

namespace tnn {
namespace {

template <class T>
T* find_first(const Vec<Layer*>& ls) {
  for (Layer* l : ls)
    if (auto* p = dynamic_cast<T*>(l)) return p;
  return nullptr;
}
Vec<Layer*> flat_children(Layer* l) {            // a residual main path is a Sequential; a bare layer stands for itself
  Vec<Layer*> c = l->children();
  return c.empty() ? Vec<Layer*>{l} : c;
}

// out[M, N] = act(x[M, K] . w[N, K]^T + bias + residual) on the GPU; decode-sized M goes to the cluster split-K kernel
void gemm_gpu(const void* x, long long ldx, const void* w, int M, int N, int K, const float* bias, const void* residual, void* out, long long ldo,
              int act, cudaStream_t s) {
  if (M <= 32 && ldx % 8 == 0 && K % 8 == 0) {
    skinny_gemm_bf16(x, ldx, w, (long long)K, bias, residual, ldo, out, ldo, M, N, K, act, /*static_weights=*/1, s);
    return;
  }
  Epilogue e;
  e.out = out;
  e.ldc = ldo;
  e.bias = bias;
  e.residual = residual;
  e.act = act;
  gemm_bf16(x, w, M, N, K, ldx, (long long)K, false, false, e, 1, s);
}

void cpu_layernorm(const float* x, float* y, size_t rows, size_t C, const float* g, const float* b, float eps) {
  for (size_t r = 0; r < rows; ++r) {
    const float* xr = x + r * C;
    double m = 0, v = 0;
    for (size_t c = 0; c < C; ++c) m += xr[c];
    m /= (double)C;
    for (size_t c = 0; c < C; ++c) v += (xr[c] - m) * (xr[c] - m);
    const float rs = 1.f / std::sqrt((float)(v / (double)C) + eps);
    for (size_t c = 0; c < C; ++c) y[r * C + c] = ((xr[c] - (float)m) * rs) * (g ? g[c] : 1.f) + (b ? b[c] : 0.f);
  }
}
// y = act(x W^T + bias) (+ residual)
void cpu_linear(const float* x, const float* W, const float* bias, const float* residual, float* y, size_t M, size_t N, size_t K, bool gelu) {
  cpu_sgemm(x, W, y, (int64_t)M, (int64_t)N, (int64_t)K, false, true, 1.f, 0.f, (int64_t)K, (int64_t)K, (int64_t)N);
  for (size_t i = 0; i < M; ++i)
    for (size_t j = 0; j < N; ++j) {
      float v = y[i * N + j] + (bias ? bias[j] : 0.f);
      if (gelu) v = 0.5f * v * (1.f + std::tanh(0.7978845608028654f * (v + 0.044715f * v * v * v)));
      y[i * N + j] = v + (residual ? residual[i * N + j] : 0.f);
    }
}

}  // namespace

cudaStream_t GPT2Engine::stream() const { return current_flow(dev_).stream(); }

GPT2Engine::GPT2Engine(Layer& model, size_t max_batch, size_t max_len, const std::string& precision)
    : model_(model), dev_(model.device()), gpu_(model.device().is_gpu()), max_batch_(std::max<size_t>(1, max_batch)) {
  model_.set_training(false);
  Vec<Layer*> top = model_.children();
  wte_ = find_first<EmbeddingLayer>(top);
  wpe_ = find_first<PositionalEmbeddingLayer>(top);
  if (!wte_ || !wpe_) throw std::runtime_error("GPT2Engine: the model has no token / positional embedding");
  Vec<ResidualBlock*> res;
  size_t last_res = 0;
  for (size_t i = 0; i < top.size(); ++i)
    if (auto* r = dynamic_cast<ResidualBlock*>(top[i])) {
      res.push_back(r);
      last_res = i;
    }
  if (res.empty() || res.size() % 2) throw std::runtime_error("GPT2Engine: expected (attention, mlp) residual pairs");
  for (size_t i = 0; i < res.size(); i += 2) {
    Block b;
    Vec<Layer*> al = flat_children(res[i]->main_path()), fl = flat_children(res[i + 1]->main_path());
    b.ln1 = find_first<LayerNormLayer>(al);
    b.attn = find_first<AttentionBlock>(al);
    b.ln2 = find_first<LayerNormLayer>(fl);
    Vec<DenseLayer*> dense;
    for (Layer* l : fl)
      if (auto* d = dynamic_cast<DenseLayer*>(l)) dense.push_back(d);
    if (!b.ln1 || !b.attn || !b.ln2 || dense.size() != 2) throw std::runtime_error("GPT2Engine: unexpected transformer block structure");
    if (auto* a = find_first<ActivationLayer>(fl))
      if (a->activation() != "gelu") throw std::runtime_error("GPT2Engine: the MLP epilogue implements GELU, got " + a->activation());
    b.fc1 = dense[0];
    b.fc2 = dense[1];
    blocks_.push_back(std::move(b));
  }
  Vec<Layer*> tail(top.begin() + (long)last_res + 1, top.end());
  ln_f_ = find_first<LayerNormLayer>(tail);
  head_ = find_first<DenseLayer>(tail);
  if (!ln_f_ || !head_) throw std::runtime_error("GPT2Engine: no final LayerNorm / LM head");
  E_ = wte_->embed_dim();
  V_ = wte_->vocab_size();
  H_ = blocks_[0].attn->num_heads();
  F_ = blocks_[0].fc1->output_features();
  max_len_ = max_len ? std::min(max_len, wpe_->seq_len()) : wpe_->seq_len();
  if (gpu_ && E_ / H_ != 64) throw std::runtime_error("GPT2Engine: the fused attention kernels implement head_dim 64");
  const DType_t act = gpu_ ? DType_t::BF16 : DType_t::FP32;
  for (auto& b : blocks_) {
    b.kc = make_tensor(act, {max_batch_, max_len_, E_}, dev_);
    b.vc = make_tensor(act, {max_batch_, max_len_, E_}, dev_);
  }
  if (precision == "fp8") enable_fp8();
  else if (precision != "bf16") throw std::runtime_error("GPT2Engine: precision must be bf16 or fp8, not " + precision);
}

GPT2Engine::~GPT2Engine() {
  if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
  if (graph_) cudaGraphDestroy(graph_);
}

void GPT2Engine::enable_fp8() {
  if (!gpu_) throw std::runtime_error("GPT2Engine: the fp8 path needs a GPU (tcgen05 block-scaled MMA)");
  if (E_ % 128 || F_ % 128) throw std::runtime_error("GPT2Engine: MXFP8 needs feature sizes that are multiples of 128");
  dev_.set_current();
  cudaStream_t s = stream();
  auto quant = [&](Tensor& w, Tensor& q, Tensor& sf) {
    const size_t rows = w->dimension(0), K = w->dimension(1);
    q = make_tensor(DType_t::UINT8_T, {rows, K}, dev_);
    sf = make_tensor(DType_t::UINT8_T, {(size_t)mxfp8_sf_bytes((long long)rows, (long long)K)}, dev_);
    mxfp8_quantize_f32(w->data_as<float>(), (long long)K, (long long)rows, (int)K, q->data(), (long long)K, sf->data(), s);
  };
  for (auto& b : blocks_) {
    quant(b.attn->param("qkv_weight"), b.qkv_q, b.qkv_sf);
    quant(b.attn->param("out_weight"), b.out_q, b.out_sf);
    quant(b.fc1->param("weight"), b.fc1_q, b.fc1_sf);
    quant(b.fc2->param("weight"), b.fc2_q, b.fc2_sf);
  }
  // fp8 shadow KV cache for the persistent decode kernel: OFF by default -- measured slower (B = 32, context ~1000: attention
  // phase 36.8 us vs 25.3 us per layer, 1.057 vs 0.868 ms per token, and +0.3 ms of prefill for the shadow writes): the phase
  // is instruction / latency bound (e4m3 converts, two scale-byte loads per key), not bandwidth bound.  TNN_KV_FP8=1 enables.
  const char* kve = getenv("TNN_KV_FP8");
  kv8_enabled_ = kve && atoi(kve) != 0 && E_ % 64 == 0;
  if (kv8_enabled_)
    for (auto& b : blocks_) {
      b.k8 = make_tensor(DType_t::UINT8_T, {max_batch_, max_len_, E_}, dev_);
      b.v8 = make_tensor(DType_t::UINT8_T, {max_batch_, max_len_, E_}, dev_);
      b.ks8 = make_tensor(DType_t::UINT8_T, {max_batch_, max_len_, E_ / 32}, dev_);
      b.vs8 = make_tensor(DType_t::UINT8_T, {max_batch_, max_len_, E_ / 32}, dev_);
    }
  kv8_len_ = 0;
  current_flow(dev_).synchronize();
  precision_ = "fp8";
  cap_tokens_ = 0;                   // the fp8 activation buffers are allocated with the next ensure_buffers()
}

void GPT2Engine::ensure_buffers(size_t tokens) {
  if (tokens <= cap_tokens_) return;
  const size_t T = std::max(tokens, max_batch_);
  const DType_t act = gpu_ ? DType_t::BF16 : DType_t::FP32;
  ids_ = make_tensor(DType_t::INT32_T, {T}, dev_);
  x_ = make_tensor(act, {T, E_}, dev_);
  x2_ = make_tensor(act, {T, E_}, dev_);
  y_ = make_tensor(act, {T, E_}, dev_);
  qkv_ = make_tensor(act, {T, 3 * E_}, dev_);
  o_ = make_tensor(act, {T, E_}, dev_);
  h_ = make_tensor(act, {T, F_}, dev_);
  const size_t vpad = (V_ + 7) / 8 * 8;
  last_ = make_tensor(act, {max_batch_, E_}, dev_);
  logits_ = make_tensor(act, {max_batch_, vpad}, dev_);
  next_ = make_tensor(DType_t::INT32_T, {max_batch_}, dev_);
  if (precision_ == "fp8") {
    q8_ = make_tensor(DType_t::UINT8_T, {T, E_}, dev_);
    q8_sf_ = make_tensor(DType_t::UINT8_T, {(size_t)mxfp8_sf_bytes((long long)T, (long long)E_)}, dev_);
    h8_ = make_tensor(DType_t::UINT8_T, {T, F_}, dev_);
    h8_sf_ = make_tensor(DType_t::UINT8_T, {(size_t)mxfp8_sf_bytes((long long)T, (long long)F_)}, dev_);
    // scale rows beyond the live tokens stay at 1.0 (ue8m0 byte 127): filled once here instead of a pad launch per LayerNorm
    TNN_CUDA(cudaMemsetAsync(q8_sf_->data(), 127, q8_sf_->nbytes(), stream()));
    TNN_CUDA(cudaMemsetAsync(h8_sf_->data(), 127, h8_sf_->nbytes(), stream()));
  }
  cap_tokens_ = T;
  if (graph_exec_) {                 // the captured graph points into the old buffers
    cudaGraphExecDestroy(graph_exec_);
    cudaGraphDestroy(graph_);
    graph_exec_ = nullptr;
    graph_ = nullptr;
  }
}

// ids_ already holds the B*L token ids (device on the GPU); writes x_ [B*L, E]
void GPT2Engine::embed(const int32_t*, size_t B, size_t L, size_t pos) {
  const float* wte = wte_->param("weight")->data_as<float>();
  const float* wpe = wpe_->param("weight")->data_as<float>() + pos * E_;
  if (gpu_) {
    embedding_fwd(ids_->data_as<int>(), wte, wpe, x_->data(), (long long)(B * L), (int)L, (int)E_, (int)V_, stream());
    return;
  }
  const int32_t* ids = ids_->data_as<int32_t>();
  float* x = x_->data_as<float>();
  for (size_t t = 0; t < B * L; ++t) {
    const size_t id = (size_t)std::min<long>(std::max<long>(ids[t], 0), (long)V_ - 1);
    for (size_t e = 0; e < E_; ++e) x[t * E_ + e] = wte[id * E_ + e] + wpe[(t % L) * E_ + e];
  }
}

void GPT2Engine::layers_gpu(size_t B, size_t L, size_t pos, bool fp8) {
  cudaStream_t s = stream();
  const int T = (int)(B * L), E = (int)E_, F = (int)F_;
  const long long tokE = (long long)E_, batE = (long long)(max_len_ * E_);
  void* x = x_->data();
  void* x2 = x2_->data();
  const bool ln8 = fp8 && E_ % 256 == 0 && E_ <= 1280;  // LayerNorm -> MXFP8 in one kernel (C = 256 .. 1280 in steps of 256)
  for (auto& b : blocks_) {
    const float* g1 = b.ln1->param("gamma")->data_as<float>();
    const float* b1 = b.ln1->param("beta")->data_as<float>();
    const float* g2 = b.ln2->param("gamma")->data_as<float>();
    const float* b2 = b.ln2->param("beta")->data_as<float>();
    const float* qkv_b = b.attn->param("qkv_bias")->data_as<float>();
    const float* out_b = b.attn->param("out_bias")->data_as<float>();
    const float* fc1_b = b.fc1->use_bias() ? b.fc1->param("bias")->data_as<float>() : nullptr;
    const float* fc2_b = b.fc2->use_bias() ? b.fc2->param("bias")->data_as<float>() : nullptr;
    // ---- attention half: x2 = x + out_proj(attn(qkv(LN(x))))
    if (fp8) {
      if (ln8) layernorm_fwd_mxfp8(x, nullptr, nullptr, T, E, g1, b1, (float)b.ln1->epsilon(), q8_->data(), q8_sf_->data(), nullptr, s, /*pad_scales=*/false);
      else {
        layernorm_fwd(x, nullptr, y_->data(), nullptr, T, E, g1, b1, (float)b.ln1->epsilon(), nullptr, nullptr, s);
        mxfp8_quantize_bf16(y_->data(), E, T, E, q8_->data(), E, q8_sf_->data(), s);
      }
      // Q -> qkv_ used as a dense [T, E] buffer; K / V -> straight into the caches at [pos, pos + L) (no cache-store pass)
      Fp8Epilogue e;
      e.bias = qkv_b;
      e.seg_cols = E;
      e.seg_rows = (int)L;
      e.seg_out[0] = qkv_->data();
      e.seg_tok[0] = E;
      e.seg_batch[0] = (long long)L * E;
      e.seg_out[1] = static_cast<char*>(b.kc->data()) + pos * E_ * 2;
      e.seg_out[2] = static_cast<char*>(b.vc->data()) + pos * E_ * 2;
      e.seg_tok[1] = e.seg_tok[2] = tokE;
      e.seg_batch[1] = e.seg_batch[2] = batE;
      if (kv8_enabled_) {                                 // the fp8 shadow cache of the decode kernel, same positions
        e.seg_q[1] = static_cast<char*>(b.k8->data()) + pos * E_;
        e.seg_q[2] = static_cast<char*>(b.v8->data()) + pos * E_;
        e.seg_qs[1] = static_cast<char*>(b.ks8->data()) + pos * (E_ / 32);
        e.seg_qs[2] = static_cast<char*>(b.vs8->data()) + pos * (E_ / 32);
      }
      gemm_mxfp8(q8_->data(), E, q8_sf_->data(), b.qkv_q->data(), E, b.qkv_sf->data(), T, 3 * E, E, e, s);
    } else {
      layernorm_fwd(x, nullptr, y_->data(), nullptr, T, E, g1, b1, (float)b.ln1->epsilon(), nullptr, nullptr, s);
      gemm_gpu(y_->data(), E, b.attn->lp("qkv_weight")->data(), T, 3 * E, E, qkv_b, nullptr, qkv_->data(), 3 * E, ACT_NONE, s);
    }
    const char* base = static_cast<const char*>(qkv_->data());
    if (fp8) {
      // (fp8 prefill has L > 1) queries from the dense Q buffer, keys / values of [0, pos + L) from the cache
      attention_fwd_bf16(base, tokE, (long long)L * E, b.kc->data(), tokE, batE, b.vc->data(), tokE, batE, o_->data(), nullptr, (int)B, (int)H_, (int)L,
                         (int)(pos + L), b.attn->is_causal() ? 1 : 0, (int)pos, s);
    } else if (L == 1) {
      // single token per row: the decode kernel over the cache (host-known position)
      kv_cache_store_bf16(qkv_->data(), b.kc->data(), b.vc->data(), tokE, batE, (int)B, 1, E, (int)pos, s);
      attention_decode_bf16(base, 3 * E, b.kc->data(), b.vc->data(), tokE, batE, o_->data(), (int)B, (int)H_, (int)(pos + 1), s);
    } else {
      kv_cache_store_bf16(qkv_->data(), b.kc->data(), b.vc->data(), tokE, batE, (int)B, (int)L, E, (int)pos, s);
      const long long tok = 3 * E, bat = (long long)L * 3 * E;
      if (pos == 0)
        attention_fwd_bf16(base, tok, bat, base + E_ * 2, tok, bat, base + 2 * E_ * 2, tok, bat, o_->data(), nullptr, (int)B, (int)H_, (int)L, (int)L,
                           b.attn->is_causal() ? 1 : 0, 0, s);
      else   // chunked prefill: keys / values of [0, pos + L) from the cache
        attention_fwd_bf16(base, tok, bat, b.kc->data(), tokE, batE, b.vc->data(), tokE, batE, o_->data(), nullptr, (int)B, (int)H_, (int)L,
                           (int)(pos + L), b.attn->is_causal() ? 1 : 0, (int)pos, s);
    }
    if (fp8) {
      mxfp8_quantize_bf16(o_->data(), E, T, E, q8_->data(), E, q8_sf_->data(), s);
      Fp8Epilogue e;
      e.out = x2;
      e.ldc = E;
      e.bias = out_b;
      e.residual = x;
      gemm_mxfp8(q8_->data(), E, q8_sf_->data(), b.out_q->data(), E, b.out_sf->data(), T, E, E, e, s);
    } else {
      gemm_gpu(o_->data(), E, b.attn->lp("out_weight")->data(), T, E, E, out_b, x, x2, E, ACT_NONE, s);
    }
    // ---- MLP half: x = x2 + fc2(gelu(fc1(LN(x2))))
    if (fp8) {
      if (ln8) layernorm_fwd_mxfp8(x2, nullptr, nullptr, T, E, g2, b2, (float)b.ln2->epsilon(), q8_->data(), q8_sf_->data(), nullptr, s, /*pad_scales=*/false);
      else {
        layernorm_fwd(x2, nullptr, y_->data(), nullptr, T, E, g2, b2, (float)b.ln2->epsilon(), nullptr, nullptr, s);
        mxfp8_quantize_bf16(y_->data(), E, T, E, q8_->data(), E, q8_sf_->data(), s);
      }
      Fp8Epilogue e1;                                   // GELU + re-quantisation for fc2 in the fc1 epilogue
      e1.bias = fc1_b;
      e1.act = ACT_GELU;
      e1.out_q = h8_->data();
      e1.ldq = F;
      e1.out_sf = h8_sf_->data();
      gemm_mxfp8(q8_->data(), E, q8_sf_->data(), b.fc1_q->data(), E, b.fc1_sf->data(), T, F, E, e1, s);
      Fp8Epilogue e2;
      e2.out = x;
      e2.ldc = E;
      e2.bias = fc2_b;
      e2.residual = x2;
      gemm_mxfp8(h8_->data(), F, h8_sf_->data(), b.fc2_q->data(), F, b.fc2_sf->data(), T, E, F, e2, s);
    } else {
      layernorm_fwd(x2, nullptr, y_->data(), nullptr, T, E, g2, b2, (float)b.ln2->epsilon(), nullptr, nullptr, s);
      gemm_gpu(y_->data(), E, b.fc1->lp("weight")->data(), T, F, E, fc1_b, nullptr, h_->data(), F, ACT_GELU, s);
      gemm_gpu(h_->data(), F, b.fc2->lp("weight")->data(), T, E, F, fc2_b, x2, x, E, ACT_NONE, s);
    }
  }
}

void GPT2Engine::layers_cpu(size_t B, size_t L, size_t pos) {
  const size_t T = B * L, D = E_ / H_;
  float* x = x_->data_as<float>();
  float* x2 = x2_->data_as<float>();
  float* y = y_->data_as<float>();
  float* qkv = qkv_->data_as<float>();
  float* o = o_->data_as<float>();
  float* h = h_->data_as<float>();
  const float scale = 1.f / std::sqrt((float)D);
  std::vector<float> p(pos + L);
  for (auto& b : blocks_) {
    cpu_layernorm(x, y, T, E_, b.ln1->param("gamma")->data_as<float>(), b.ln1->param("beta")->data_as<float>(), (float)b.ln1->epsilon());
    cpu_linear(y, b.attn->param("qkv_weight")->data_as<float>(), b.attn->param("qkv_bias")->data_as<float>(), nullptr, qkv, T, 3 * E_, E_, false);
    float* kc = b.kc->data_as<float>();
    float* vc = b.vc->data_as<float>();
    for (size_t bi = 0; bi < B; ++bi)
      for (size_t l = 0; l < L; ++l) {
        const float* row = qkv + (bi * L + l) * 3 * E_;
        std::memcpy(kc + (bi * max_len_ + pos + l) * E_, row + E_, E_ * sizeof(float));
        std::memcpy(vc + (bi * max_len_ + pos + l) * E_, row + 2 * E_, E_ * sizeof(float));
      }
    const bool causal = b.attn->is_causal();
    for (size_t bi = 0; bi < B; ++bi)
      for (size_t hh = 0; hh < H_; ++hh)
        for (size_t l = 0; l < L; ++l) {
          const float* q = qkv + (bi * L + l) * 3 * E_ + hh * D;
          const size_t lim = causal ? pos + l + 1 : pos + L;
          float mx = -1e30f;
          for (size_t j = 0; j < lim; ++j) {
            const float* k = kc + (bi * max_len_ + j) * E_ + hh * D;
            float sdot = 0;
            for (size_t d = 0; d < D; ++d) sdot += q[d] * k[d];
            p[j] = sdot * scale;
            mx = std::max(mx, p[j]);
          }
          float den = 0;
          for (size_t j = 0; j < lim; ++j) {
            p[j] = std::exp(p[j] - mx);
            den += p[j];
          }
          float* out = o + (bi * L + l) * E_ + hh * D;
          std::fill(out, out + D, 0.f);
          for (size_t j = 0; j < lim; ++j) {
            const float w = p[j] / den;
            const float* v = vc + (bi * max_len_ + j) * E_ + hh * D;
            for (size_t d = 0; d < D; ++d) out[d] += w * v[d];
          }
        }
    cpu_linear(o, b.attn->param("out_weight")->data_as<float>(), b.attn->param("out_bias")->data_as<float>(), x, x2, T, E_, E_, false);
    cpu_layernorm(x2, y, T, E_, b.ln2->param("gamma")->data_as<float>(), b.ln2->param("beta")->data_as<float>(), (float)b.ln2->epsilon());
    cpu_linear(y, b.fc1->param("weight")->data_as<float>(), b.fc1->use_bias() ? b.fc1->param("bias")->data_as<float>() : nullptr, nullptr, h, T, F_, E_, true);
    cpu_linear(h, b.fc2->param("weight")->data_as<float>(), b.fc2->use_bias() ? b.fc2->param("bias")->data_as<float>() : nullptr, x2, x, T, E_, F_, false);
  }
}

// last position of every row -> final LayerNorm -> LM head -> arg-max (+ logits on request)
std::vector<int32_t> GPT2Engine::finish(size_t B, size_t L, std::vector<float>* logits) {
  const size_t vpad = (V_ + 7) / 8 * 8;
  const float* g = ln_f_->param("gamma")->data_as<float>();
  const float* be = ln_f_->param("beta")->data_as<float>();
  const float* hb = head_->use_bias() ? head_->param("bias")->data_as<float>() : nullptr;
  std::vector<int32_t> next(B);
  if (gpu_) {
    cudaStream_t s = stream();
    const char* src = static_cast<const char*>(x_->data()) + (L - 1) * E_ * 2;
    TNN_CUDA(cudaMemcpy2DAsync(last_->data(), E_ * 2, src, L * E_ * 2, E_ * 2, B, cudaMemcpyDeviceToDevice, s));
    layernorm_fwd(last_->data(), nullptr, y_->data(), nullptr, (long long)B, (int)E_, g, be, (float)ln_f_->epsilon(), nullptr, nullptr, s);
    gemm_gpu(y_->data(), (long long)E_, head_->lp("weight")->data(), (int)B, (int)V_, (int)E_, hb, nullptr, logits_->data(), (long long)vpad, ACT_NONE, s);
    argmax_rows(logits_->data(), 0, (long long)vpad, (long long)B, (int)V_, next_->data_as<int>(), s);
    TNN_CUDA(cudaMemcpyAsync(next.data(), next_->data(), B * 4, cudaMemcpyDeviceToHost, s));
    if (logits) {
      std::vector<uint16_t> raw(B * vpad);
      TNN_CUDA(cudaMemcpyAsync(raw.data(), logits_->data(), raw.size() * 2, cudaMemcpyDeviceToHost, s));
      current_flow(dev_).synchronize();
      logits->resize(B * V_);
      for (size_t b = 0; b < B; ++b)
        for (size_t v = 0; v < V_; ++v) (*logits)[b * V_ + v] = bf16_to_float(bf16_t{raw[b * vpad + v]});
    } else {
      current_flow(dev_).synchronize();
    }
    return next;
  }
  float* last = last_->data_as<float>();
  const float* x = x_->data_as<float>();
  for (size_t b = 0; b < B; ++b) std::memcpy(last + b * E_, x + (b * L + L - 1) * E_, E_ * sizeof(float));
  float* y = y_->data_as<float>();
  cpu_layernorm(last, y, B, E_, g, be, (float)ln_f_->epsilon());
  std::vector<float> lg(B * V_);
  cpu_linear(y, head_->param("weight")->data_as<float>(), hb, nullptr, lg.data(), B, V_, E_, false);
  for (size_t b = 0; b < B; ++b) next[b] = (int32_t)(std::max_element(lg.begin() + (long)(b * V_), lg.begin() + (long)((b + 1) * V_)) - (lg.begin() + (long)(b * V_)));
  if (logits) *logits = std::move(lg);
  return next;
}

std::vector<int32_t> GPT2Engine::prefill(const std::vector<int32_t>& ids, size_t B, size_t L, std::vector<float>* logits) {
  if (ids.size() != B * L) throw std::runtime_error("GPT2Engine::prefill: ids must hold B*L tokens");
  if (B == 0 || L == 0 || B > max_batch_ || L > max_len_)
    throw std::runtime_error("GPT2Engine::prefill: prompt [" + std::to_string(B) + ", " + std::to_string(L) + "] exceeds engine capacity [" +
                             std::to_string(max_batch_) + ", " + std::to_string(max_len_) + "]");
  if (gpu_) dev_.set_current();
  ensure_buffers(B * L);
  cur_batch_ = B;
  cur_len_ = 0;
  if (gpu_) TNN_CUDA(cudaMemcpyAsync(ids_->data(), ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, stream()));
  else std::memcpy(ids_->data(), ids.data(), ids.size() * 4);
  embed(nullptr, B, L, 0);
  const bool fp8_path = gpu_ && precision_ == "fp8" && B * L >= 128;
  if (gpu_) layers_gpu(B, L, 0, fp8_path);
  else layers_cpu(B, L, 0);
  cur_len_ = L;
  kv8_len_ = (fp8_path && kv8_enabled_) ? L : 0;
  return finish(B, L, logits);
}

std::vector<int32_t> GPT2Engine::decode(const std::vector<int32_t>& tokens, std::vector<float>* logits) {
  const size_t B = cur_batch_;
  if (tokens.size() != B) throw std::runtime_error("GPT2Engine::decode: one token per cached row expected");
  if (cur_len_ >= max_len_) throw std::runtime_error("GPT2Engine::decode: the KV cache is full");
  if (gpu_) dev_.set_current();
  ensure_buffers(B);
  if (gpu_) TNN_CUDA(cudaMemcpyAsync(ids_->data(), tokens.data(), B * 4, cudaMemcpyHostToDevice, stream()));
  else std::memcpy(ids_->data(), tokens.data(), B * 4);
  embed(nullptr, B, 1, cur_len_);
  if (gpu_) layers_gpu(B, 1, cur_len_, false);
  else layers_cpu(B, 1, cur_len_);
  ++cur_len_;
  return finish(B, 1, logits);
}

void GPT2Engine::build_mega() {
  // LayerNorm folding (decode_mega.cuh): W' = bf16(W * gamma), column sums of W', b' = b + W . beta -- once per engine
  cudaStream_t s = stream();
  mega_fold_.clear();
  auto fold = [&](Tensor& W, const float* bias, LayerNormLayer* ln, const void*& wf, const float*& cs, const float*& bf) {
    const size_t N = W->dimension(0), K = W->dimension(1);
    Tensor w = make_tensor(DType_t::BF16, {N, K}, dev_), c = make_tensor(DType_t::FP32, {N}, dev_), b = make_tensor(DType_t::FP32, {N}, dev_);
    decode_fold_ln(W->data_as<float>(), bias, ln->param("gamma")->data_as<float>(), ln->param("beta")->data_as<float>(), (int)N, (int)K, w->data(),
                   c->data_as<float>(), b->data_as<float>(), s);
    wf = w->data();
    cs = c->data_as<float>();
    bf = b->data_as<float>();
    mega_fold_.push_back(w);
    mega_fold_.push_back(c);
    mega_fold_.push_back(b);
  };
  std::vector<DecodeLayerParams> host(blocks_.size());
  for (size_t i = 0; i < blocks_.size(); ++i) {
    Block& b = blocks_[i];
    DecodeLayerParams& L = host[i];
    fold(b.attn->param("qkv_weight"), b.attn->param("qkv_bias")->data_as<float>(), b.ln1, L.w_qkv, L.c_qkv, L.b_qkv);
    L.w_out = b.attn->lp("out_weight")->data();
    L.b_out = b.attn->param("out_bias")->data_as<float>();
    fold(b.fc1->param("weight"), b.fc1->use_bias() ? b.fc1->param("bias")->data_as<float>() : nullptr, b.ln2, L.w_fc1, L.c_fc1, L.b_fc1);
    L.w_fc2 = b.fc2->lp("weight")->data();
    L.b_fc2 = b.fc2->use_bias() ? b.fc2->param("bias")->data_as<float>() : nullptr;
    L.kc = b.kc->data();
    L.vc = b.vc->data();
    L.k8 = b.k8 ? b.k8->data() : nullptr;
    L.v8 = b.v8 ? b.v8->data() : nullptr;
    L.ks8 = b.ks8 ? b.ks8->data() : nullptr;
    L.vs8 = b.vs8 ? b.vs8->data() : nullptr;
    L.eps1 = (float)b.ln1->epsilon();
    L.eps2 = (float)b.ln2->epsilon();
  }
  fold(head_->param("weight"), head_->use_bias() ? head_->param("bias")->data_as<float>() : nullptr, ln_f_, mega_head_w_, mega_head_c_, mega_head_b_);
  mega_layers_ = make_tensor(DType_t::UINT8_T, {host.size() * sizeof(DecodeLayerParams)}, dev_);
  mega_layers_->copy_from_host(host.data(), host.size() * sizeof(DecodeLayerParams));
  mega_part_ = make_tensor(DType_t::UINT8_T, {decode_mega_part_bytes((int)max_batch_, (int)H_)}, dev_);
  mega_cnt_ = make_tensor(DType_t::INT32_T, {max_batch_ * H_}, dev_);
  mega_cnt_->zero();
  mega_bar_ = make_tensor(DType_t::INT32_T, {4}, dev_);
  mega_bar_->zero();
  mega_stats_ = make_tensor(DType_t::FP32, {128}, dev_);
  mega_stats_->zero();
  mega_amax_ = make_tensor(DType_t::INT64_T, {32}, dev_);
  mega_amax_->zero();
  const char* pe = getenv("TNN_DECODE_PROF");
  if (pe && atoi(pe)) {
    mega_prof_ = make_tensor(DType_t::INT64_T, {4 + 5 * blocks_.size()}, dev_);
    mega_prof_->zero();
  }
  current_flow(dev_).synchronize();
}

void GPT2Engine::decode_step_device(size_t B) {
  cudaStream_t s = stream();
  if (use_mega_) {
    DecodeParams p{};
    p.layers = static_cast<const DecodeLayerParams*>(mega_layers_->data());
    p.n_layers = (int)blocks_.size();
    p.B = (int)B;
    p.E = (int)E_;
    p.H = (int)H_;
    p.F = (int)F_;
    p.V = (int)V_;
    p.Lmax = (int)max_len_;
    p.vpad = (long long)((V_ + 7) / 8 * 8);
    p.wte = wte_->param("weight")->data_as<float>();
    p.wpe = wpe_->param("weight")->data_as<float>();
    p.w_head = mega_head_w_;
    p.c_head = mega_head_c_;
    p.b_head = mega_head_b_;
    p.epsf = (float)ln_f_->epsilon();
    p.tok = d_tok_->data_as<int>();
    p.pos = d_pos_->data_as<int>();
    p.step = d_step_->data_as<int>();
    p.hist = d_hist_->data_as<int>();
    p.hist_cap = (int)hist_cap_;
    p.x = x_->data();
    p.x2 = x2_->data();
    p.qkv = qkv_->data();
    p.o = o_->data();
    p.h = h_->data();
    p.stats = mega_stats_->data_as<float>();
    p.amax = static_cast<unsigned long long*>(mega_amax_->data());
    p.attn_part = static_cast<float*>(mega_part_->data());
    p.attn_cnt = mega_cnt_->data_as<int>();
    p.logits = logits_->data();
    p.bar = static_cast<unsigned int*>(mega_bar_->data());
    p.kv_fp8 = graph_kv8_ ? 1 : 0;
    p.prof = mega_prof_ ? static_cast<unsigned long long*>(mega_prof_->data()) : nullptr;
    decode_step_mega(p, s);
    return;
  }
  const int E = (int)E_, F = (int)F_;
  const long long tokE = (long long)E_, batE = (long long)(max_len_ * E_);
  const size_t vpad = (V_ + 7) / 8 * 8;
  void* x = x_->data();
  void* x2 = x2_->data();
  embed_decode(d_tok_->data_as<int>(), wte_->param("weight")->data_as<float>(), wpe_->param("weight")->data_as<float>(), x, (int)B, E, (int)V_,
               d_pos_->data_as<int>(), s);
  for (auto& b : blocks_) {
    layernorm_fwd(x, nullptr, y_->data(), nullptr, (long long)B, E, b.ln1->param("gamma")->data_as<float>(), b.ln1->param("beta")->data_as<float>(),
                  (float)b.ln1->epsilon(), nullptr, nullptr, s);
    gemm_gpu(y_->data(), E, b.attn->lp("qkv_weight")->data(), (int)B, 3 * E, E, b.attn->param("qkv_bias")->data_as<float>(), nullptr, qkv_->data(), 3 * E,
             ACT_NONE, s);
    attention_decode_append_bf16(qkv_->data(), b.kc->data(), b.vc->data(), tokE, batE, o_->data(), (int)B, (int)H_, d_pos_->data_as<int>(), s);
    gemm_gpu(o_->data(), E, b.attn->lp("out_weight")->data(), (int)B, E, E, b.attn->param("out_bias")->data_as<float>(), x, x2, E, ACT_NONE, s);
    layernorm_fwd(x2, nullptr, y_->data(), nullptr, (long long)B, E, b.ln2->param("gamma")->data_as<float>(), b.ln2->param("beta")->data_as<float>(),
                  (float)b.ln2->epsilon(), nullptr, nullptr, s);
    gemm_gpu(y_->data(), E, b.fc1->lp("weight")->data(), (int)B, F, E, b.fc1->use_bias() ? b.fc1->param("bias")->data_as<float>() : nullptr, nullptr,
             h_->data(), F, ACT_GELU, s);
    gemm_gpu(h_->data(), F, b.fc2->lp("weight")->data(), (int)B, E, F, b.fc2->use_bias() ? b.fc2->param("bias")->data_as<float>() : nullptr, x2, x, E,
             ACT_NONE, s);
  }
  layernorm_fwd(x, nullptr, y_->data(), nullptr, (long long)B, E, ln_f_->param("gamma")->data_as<float>(), ln_f_->param("beta")->data_as<float>(),
                (float)ln_f_->epsilon(), nullptr, nullptr, s);
  gemm_gpu(y_->data(), E, head_->lp("weight")->data(), (int)B, (int)V_, E, head_->use_bias() ? head_->param("bias")->data_as<float>() : nullptr, nullptr,
           logits_->data(), (long long)vpad, ACT_NONE, s);
  argmax_rows(logits_->data(), 0, (long long)vpad, (long long)B, (int)V_, next_->data_as<int>(), s);
  decode_epilogue(next_->data_as<int>(), d_tok_->data_as<int>(), d_hist_->data_as<int>(), (int)B, (int)hist_cap_, d_pos_->data_as<int>(),
                  d_step_->data_as<int>(), s);
}

std::vector<int32_t> GPT2Engine::decode_graph(const std::vector<int32_t>& tokens, size_t steps) {
  const size_t B = cur_batch_;
  if (tokens.size() != B) throw std::runtime_error("GPT2Engine::decode_graph: one token per cached row expected");
  steps = std::min(steps, max_len_ - cur_len_);
  std::vector<int32_t> out(B * steps);
  if (steps == 0) return out;
  if (!gpu_) {
    std::vector<int32_t> tok = tokens;
    for (size_t s = 0; s < steps; ++s) {
      tok = decode(tok);
      for (size_t b = 0; b < B; ++b) out[b * steps + s] = tok[b];
    }
    return out;
  }
  dev_.set_current();
  ensure_buffers(B);
  Flow& flow = current_flow(dev_);
  const bool want_kv8 = kv8_enabled_ && kv8_len_ == cur_len_ && cur_len_ > 0;
  if (graph_exec_ && want_kv8 != graph_kv8_) {            // the captured kernel variant attends over the other cache
    cudaGraphExecDestroy(graph_exec_);
    cudaGraphDestroy(graph_);
    graph_exec_ = nullptr;
    graph_ = nullptr;
  }
  graph_kv8_ = want_kv8;
  if (!d_tok_ || graph_batch_ != B || hist_cap_ < steps) {
    hist_cap_ = std::max<size_t>(steps, 64);
    d_tok_ = make_tensor(DType_t::INT32_T, {B}, dev_);
    d_pos_ = make_tensor(DType_t::INT32_T, {1}, dev_);
    d_step_ = make_tensor(DType_t::INT32_T, {1}, dev_);
    d_hist_ = make_tensor(DType_t::INT32_T, {hist_cap_, B}, dev_);
    graph_batch_ = B;
    if (graph_exec_) {
      cudaGraphExecDestroy(graph_exec_);
      cudaGraphDestroy(graph_);
      graph_exec_ = nullptr;
      graph_ = nullptr;
    }
  }
  const int32_t pos0 = (int32_t)cur_len_, zero = 0;
  auto rewind = [&] {
    TNN_CUDA(cudaMemcpyAsync(d_tok_->data(), tokens.data(), B * 4, cudaMemcpyHostToDevice, flow.stream()));
    TNN_CUDA(cudaMemcpyAsync(d_pos_->data(), &pos0, 4, cudaMemcpyHostToDevice, flow.stream()));
    TNN_CUDA(cudaMemcpyAsync(d_step_->data(), &zero, 4, cudaMemcpyHostToDevice, flow.stream()));
    flow.synchronize();
  };
  rewind();
  if (!graph_exec_) {
    for (auto& b : blocks_) {          // make sure the bf16 weight copies exist before the capture
      b.attn->lp("qkv_weight");
      b.attn->lp("out_weight");
      b.fc1->lp("weight");
      b.fc2->lp("weight");
    }
    head_->lp("weight");
    const char* env = getenv("TNN_DECODE_MEGA");
    use_mega_ = decode_mega_supported((int)B, (int)E_, (int)H_, (int)F_) && !(env && atoi(env) == 0);
    if (use_mega_) build_mega();
    decode_step_device(B);             // warm-up (kernel attributes, lazy state) ...
    rewind();                          // ... then undo its side effects
    TNN_CUDA(cudaStreamBeginCapture(flow.stream(), cudaStreamCaptureModeRelaxed));
    try {
      decode_step_device(B);
    } catch (...) {
      cudaGraph_t dead = nullptr;
      cudaStreamEndCapture(flow.stream(), &dead);
      if (dead) cudaGraphDestroy(dead);
      throw;
    }
    TNN_CUDA(cudaStreamEndCapture(flow.stream(), &graph_));
    TNN_CUDA(cudaGraphInstantiate(&graph_exec_, graph_, 0));
  }
  for (size_t s = 0; s < steps; ++s) TNN_CUDA(cudaGraphLaunch(graph_exec_, flow.stream()));
  std::vector<int32_t> hist(steps * B);
  TNN_CUDA(cudaMemcpyAsync(hist.data(), d_hist_->data(), hist.size() * 4, cudaMemcpyDeviceToHost, flow.stream()));
  flow.synchronize();
  cur_len_ += steps;
  if (use_mega_ && graph_kv8_) kv8_len_ = cur_len_;       // the persistent kernel appended both caches
  for (size_t s = 0; s < steps; ++s)
    for (size_t b = 0; b < B; ++b) out[b * steps + s] = hist[s * B + b];
  return out;
}

std::vector<int32_t> GPT2Engine::generate(const std::vector<int32_t>& prompt, size_t B, size_t L, size_t max_new_tokens, long eos_token, size_t* n_out) {
  std::vector<std::vector<int32_t>> cols;
  std::vector<int32_t> tok = prefill(prompt, B, L);
  cols.push_back(tok);
  std::vector<char> done(B, 0);
  for (size_t i = 1; i < max_new_tokens && cur_len_ < max_len_; ++i) {
    if (eos_token >= 0) {
      bool all = true;
      for (size_t b = 0; b < B; ++b) {
        done[b] = done[b] || tok[b] == (int32_t)eos_token;
        all = all && done[b];
      }
      if (all) break;
    }
    tok = decode(tok);
    cols.push_back(tok);
  }
  const size_t n = cols.size();
  std::vector<int32_t> out(B * n);
  for (size_t s = 0; s < n; ++s)
    for (size_t b = 0; b < B; ++b) out[b * n + s] = cols[s][b];
  if (n_out) *n_out = n;
  return out;
}

std::vector<float> GPT2Engine::full_forward_logits(const std::vector<int32_t>& ids, size_t B, size_t L) {
  model_.set_training(false);
  Tensor t = make_tensor(DType_t::INT32_T, {B, L}, getHost());
  std::memcpy(t->data(), ids.data(), ids.size() * 4);
  if (gpu_) t = t->to_device(dev_);
  Tensor y = model_.forward(t, 0);
  if (gpu_) current_flow(dev_).synchronize();
  return to_float_vector(*y);
}

std::vector<float> GPT2Engine::last_logits() {
  const size_t B = cur_batch_, vpad = (V_ + 7) / 8 * 8;
  std::vector<float> out(B * V_);
  if (!gpu_ || !logits_) throw std::runtime_error("GPT2Engine::last_logits: GPU engine after a step only");
  std::vector<uint16_t> raw(B * vpad);
  current_flow(dev_).synchronize();
  logits_->copy_to_host(raw.data(), raw.size() * 2);
  for (size_t b = 0; b < B; ++b)
    for (size_t v = 0; v < V_; ++v) out[b * V_ + v] = bf16_to_float(bf16_t{raw[b * vpad + v]});
  return out;
}

std::vector<double> GPT2Engine::persistent_decode_profile() {
  std::vector<double> out;
  if (!mega_prof_) return out;
  current_flow(dev_).synchronize();
  std::vector<unsigned long long> t(mega_prof_->size());
  mega_prof_->copy_to_host(t.data(), t.size() * 8);
  for (size_t i = 1; i < t.size() && t[i]; ++i) out.push_back((double)(t[i] - t[i - 1]));
  return out;
}

double GPT2Engine::time_prefill_ms(size_t B, size_t L, int iters, int warmup) {
  std::vector<int32_t> ids(B * L);
  for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int32_t)((i * 2654435761ull) % V_);
  for (int i = 0; i < warmup; ++i) prefill(ids, B, L);
  if (!gpu_) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) prefill(ids, B, L);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
  }
  Flow& flow = current_flow(dev_);
  Event e0(true), e1(true);
  float total = 0;
  for (int i = 0; i < iters; ++i) {          // device time of embed + layers (+ last-row head) per prefill
    flow.synchronize();
    flow.record(e0);
    prefill(ids, B, L);
    flow.record(e1);
    e1.synchronize();
    total += Event::elapsed_ms(e0, e1);
  }
  return total / iters;
}

double GPT2Engine::time_decode_ms(size_t B, size_t context, size_t steps, bool graph) {
  std::vector<int32_t> ids(B * context);
  for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int32_t)((i * 2246822519ull) % V_);
  std::vector<int32_t> tok = prefill(ids, B, context);
  if (graph) decode_graph(tok, 2);            // capture + settle
  else decode(tok);
  tok = prefill(ids, B, context);
  if (!gpu_) {
    const auto t0 = std::chrono::steady_clock::now();
    decode_graph(tok, steps);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;
  }
  Flow& flow = current_flow(dev_);
  Event e0(true), e1(true);
  flow.synchronize();
  flow.record(e0);
  if (graph) decode_graph(tok, steps);
  else
    for (size_t s = 0; s < steps; ++s) tok = decode(tok);
  flow.record(e1);
  e1.synchronize();
  return Event::elapsed_ms(e0, e1) / (double)steps;
}

}  // namespace tnn
