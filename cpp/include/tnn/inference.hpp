// GPT-2 inference engine of the C++ host: batched prefill + incremental decode over per-layer KV caches, greedy
// sampling on the device, CUDA-graph decode, optional block-scaled fp8 (MXFP8) prefill.
//
// The reference's gpt2_inference (examples/gpt2_inference.cpp:44-121) has no KV cache: every generated token
// re-runs the whole model on a zero-padded [1, 512] float tensor, copies [1, 512, 50257] logits to the host and
// arg-maxes there.  This engine walks a zoo GPT model (Sequential of embedding, positional embedding, (attention,
// mlp) residual pairs, final LayerNorm, LM head), binds the layers' parameters, and drives the hand-written kernels
// (csrc/kernels: tcgen05 GEMM, flash attention, split-lane decode attention, LayerNorm, MXFP8) on its own buffers.
// On the CPU the same cached algorithm runs in fp32 (the oracle the tests compare with a full recompute).
#pragma once
#include <string>
#include <vector>

#include "tnn/graph.hpp"
#include "tnn/layers.hpp"

namespace tnn {

class GPT2Engine {
 public:
  // `model`: a zoo gpt2_* / flash_gpt2_* Sequential (already initialised on its device).  The engine keeps pointers
  // into it; the model must outlive the engine.  precision: "bf16" | "fp8" (GPU only).
  GPT2Engine(Layer& model, size_t max_batch = 1, size_t max_len = 0, const std::string& precision = "bf16");
  ~GPT2Engine();
  GPT2Engine(const GPT2Engine&) = delete;
  GPT2Engine& operator=(const GPT2Engine&) = delete;

  // ids: row-major [B, L] token ids.  Fills the caches; returns the greedy next token of every row (and, when asked,
  // the last-position logits [B, vocab] as floats).
  std::vector<int32_t> prefill(const std::vector<int32_t>& ids, size_t B, size_t L, std::vector<float>* logits = nullptr);
  // tokens [B] at position cur_len(); returns the next greedy token per row.
  std::vector<int32_t> decode(const std::vector<int32_t>& tokens, std::vector<float>* logits = nullptr);
  // `steps` greedy steps as replays of ONE captured CUDA graph (token ids and the position stay on the device).
  // Returns row-major [B, steps].  On the CPU: a loop over decode().
  std::vector<int32_t> decode_graph(const std::vector<int32_t>& tokens, size_t steps);
  // Greedy generation (examples/gpt2_inference.cpp:71-121): returns row-major [B, n] with n <= max_new_tokens.
  std::vector<int32_t> generate(const std::vector<int32_t>& prompt, size_t B, size_t L, size_t max_new_tokens, long eos_token = 50256,
                                size_t* n_out = nullptr);
  // Oracle: the whole model over the whole sequence through the layers' own forward (what the reference does per token).
  std::vector<float> full_forward_logits(const std::vector<int32_t>& ids, size_t B, size_t L);

  void enable_fp8();                 // quantise the block weights once, switch prefill (>= 128 tokens) to MXFP8
  const std::string& precision() const { return precision_; }
  size_t cur_len() const { return cur_len_; }
  size_t cur_batch() const { return cur_batch_; }
  size_t max_len() const { return max_len_; }
  size_t max_batch() const { return max_batch_; }
  size_t vocab() const { return V_; }
  size_t embed_dim() const { return E_; }
  size_t num_layers() const { return blocks_.size(); }
  const Device& device() const { return dev_; }
  // true when decode_graph() replays the single persistent decode kernel (B <= 32, E <= 1024; TNN_DECODE_MEGA=0 disables)
  bool uses_persistent_decode() const { return use_mega_; }
  bool uses_fp8_kv_cache() const { return use_mega_ && graph_kv8_; }
  // logits [cur_batch, vocab] of the most recent prefill / decode / decode_graph step
  std::vector<float> last_logits();
  // TNN_DECODE_PROF=1: per-phase nanoseconds of CTA 0 in the last persistent decode step ({embed, [qkv, attn, out, fc1, fc2] x layers, head, argmax})
  std::vector<double> persistent_decode_profile();
  // device-timed helpers for the bench CLIs (CUDA events on the engine's flow, after `warmup` untimed runs)
  double time_prefill_ms(size_t B, size_t L, int iters, int warmup = 2);
  double time_decode_ms(size_t B, size_t context, size_t steps, bool graph = true);

 private:
  struct Block {
    LayerNormLayer* ln1 = nullptr;
    AttentionBlock* attn = nullptr;
    LayerNormLayer* ln2 = nullptr;
    DenseLayer* fc1 = nullptr;
    DenseLayer* fc2 = nullptr;
    Tensor kc, vc;                                   // [max_batch, max_len, E]
    // MXFP8 weights (q bytes + scale atoms)
    Tensor qkv_q, qkv_sf, out_q, out_sf, fc1_q, fc1_sf, fc2_q, fc2_sf;
    Tensor k8, v8, ks8, vs8;                         // fp8 shadow KV cache: e4m3 [max_batch, max_len, E] + ue8m0 [.., E / 32]
  };
  void embed(const int32_t* ids_dev_or_host, size_t B, size_t L, size_t pos);
  void layers_gpu(size_t B, size_t L, size_t pos, bool fp8);
  void layers_cpu(size_t B, size_t L, size_t pos);
  std::vector<int32_t> finish(size_t B, size_t L, std::vector<float>* logits);
  void decode_step_device(size_t B);                 // graph-replayable: position / tokens in device memory
  void ensure_buffers(size_t tokens);
  cudaStream_t stream() const;

  Layer& model_;
  Device dev_;
  bool gpu_;
  EmbeddingLayer* wte_ = nullptr;
  PositionalEmbeddingLayer* wpe_ = nullptr;
  LayerNormLayer* ln_f_ = nullptr;
  DenseLayer* head_ = nullptr;
  std::vector<Block> blocks_;
  size_t E_ = 0, H_ = 0, V_ = 0, F_ = 0, max_len_ = 0, max_batch_ = 0, cur_len_ = 0, cur_batch_ = 0;
  std::string precision_ = "bf16";
  // activation buffers (GPU: bf16; CPU: fp32), sized for `cap_tokens_`
  size_t cap_tokens_ = 0;
  Tensor ids_, x_, x2_, y_, qkv_, o_, h_, last_, logits_, next_;
  Tensor q8_, q8_sf_, h8_, h8_sf_;                    // MXFP8 activations
  // device decode state
  Tensor d_tok_, d_pos_, d_step_, d_hist_;
  // persistent whole-model decode kernel (csrc/kernels/decode_mega.cu): device array of per-layer pointers + scratch
  Tensor mega_layers_, mega_part_, mega_cnt_, mega_bar_, mega_prof_, mega_stats_, mega_amax_;
  std::vector<Tensor> mega_fold_;                     // LayerNorm-folded weights, column sums, biases
  const void* mega_head_w_ = nullptr;
  const float* mega_head_c_ = nullptr;
  const float* mega_head_b_ = nullptr;
  bool use_mega_ = false;
  // fp8 KV cache (precision "fp8" with TNN_KV_FP8=1; off by default: measured slower, see enable_fp8()): written by the fp8 QKV epilogue at prefill and by the persistent
  // decode kernel per token; kv8_len_ = positions it covers (the bf16 caches are always complete)
  bool kv8_enabled_ = false, graph_kv8_ = false;
  size_t kv8_len_ = 0;
  void build_mega();
  size_t hist_cap_ = 0, graph_batch_ = 0;
  cudaGraph_t graph_ = nullptr;
  cudaGraphExec_t graph_exec_ = nullptr;
};

}  // namespace tnn
