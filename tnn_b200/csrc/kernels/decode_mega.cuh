// Whole-model persistent decode step (decode_mega.cu): ONE kernel launch per generated token.
//
// A GPT-2 decode step at batch <= 32 is a chain of ~90 tiny kernels (LayerNorm, four weight-streaming GEMMs and one
// KV-cache attention per layer, LM head, arg-max); each is far below the launch + drain latency floor, so the step
// costs ~12 us per kernel no matter how fast the kernels are.  Here one persistent grid (2 CTAs per SM) walks all
// layers; the phases of a layer are separated by a device-wide arrive/wait barrier instead of a kernel boundary, the
// weights of the next phase are already in flight while a CTA waits at the barrier, the K/V append and the
// split-sequence attention merge happen inside the attention phase, and the greedy arg-max + token bookkeeping close
// the step.  LayerNorm costs nothing at run time: gamma / beta are folded into the following weight matrix once
// (decode_fold_ln), the row statistics are accumulated by the epilogue of the GEMM that PRODUCES the row, and the
// consuming GEMM runs on the raw activations and applies  rstd * (acc - mean * colsum(W')) + b'  in its epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace tnn {

struct DecodeLayerParams {
  const void* w_qkv; const float* c_qkv; const float* b_qkv;   // LN1-folded bf16 [3E, E], column sums [3E], folded bias [3E]
  const void* w_out; const float* b_out;                        // bf16 [E, E]
  const void* w_fc1; const float* c_fc1; const float* b_fc1;   // LN2-folded bf16 [F, E], [F], [F]
  const void* w_fc2; const float* b_fc2;                        // bf16 [E, F]
  void* kc; void* vc;                                           // bf16 KV cache [B, Lmax, E]
  void* k8; void* v8; void* ks8; void* vs8;                     // optional fp8 shadow: e4m3 [B, Lmax, E] + ue8m0 scales [B, Lmax, E/32]
  float eps1, eps2;
};

struct DecodeParams {
  const DecodeLayerParams* layers;             // DEVICE array [n_layers]
  int n_layers;
  int B, E, H, F, V, Lmax;
  long long vpad;                              // row stride of logits (multiple of 8)
  const float* wte; const float* wpe;          // fp32 [V, E], [Lmax, E]
  const void* w_head; const float* c_head; const float* b_head; float epsf;   // LN_f-folded bf16 [V, E], [V], [V]
  int* tok; int* pos; int* step; int* hist; int hist_cap;
  void* x; void* x2; void* qkv; void* o; void* h;   // bf16 [32, E | E | 3E | E | F] scratch (global, L2 resident)
  float* stats;                                // [2][32][2] fp32 (sum, sum of squares) of the rows of x / x2; zeroed once
  float* attn_part;                            // [B*H][16][72] fp32 split-sequence partials
  int* attn_cnt;                               // [B*H] arrival counters (zero-initialised once)
  void* logits;                                // bf16 [B, vpad]
  unsigned long long* amax;                    // [32] packed (ordered logit, ~column) running arg-max (zeroed once)
  unsigned int* bar;                           // grid barrier counter (zeroed by the launcher every step)
  int kv_fp8;                                  // attend over the fp8 shadow caches (they must cover [0, pos)); both caches are appended
  unsigned long long* prof;                    // optional [4 + 5 * n_layers] %globaltimer stamps of CTA 0 at the phase ends
};

// B <= 32, head_dim 64, E % 32 == 0, E <= 1024, F % 32 == 0
bool decode_mega_supported(int B, int E, int H, int F);
size_t decode_mega_part_bytes(int B, int H);
// Wf[n, k] = bf16(W[n, k] * gamma[k]);  csum[n] = sum_k Wf[n, k];  bias_f[n] = bias[n] + sum_k W[n, k] * beta[k]
void decode_fold_ln(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K, void* Wf, float* csum, float* bias_f,
                    cudaStream_t stream);
void decode_step_mega(const DecodeParams& p, cudaStream_t stream);

}  // namespace tnn
