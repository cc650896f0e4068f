// Host interface of the fused attention kernels (attention.cu).  All strides are in elements.
#pragma once
#include <cuda_runtime.h>

namespace tnn {

// q/k/v: bf16 views [B, L, H, 64] with token stride *_tok and batch stride *_batch (so the packed
// QKV projection [B, L, 3, H, 64] and the KV cache [B, Lmax, H, 64] are both addressed in place).
// out: bf16 [B, Lq, H*64]; lse: fp32 [B, H, Lq] or null.  q_offset = absolute position of query 0.
void attention_fwd_bf16(const void* q, long long q_tok, long long q_batch, const void* k, long long k_tok,
                        long long k_batch, const void* v, long long v_tok, long long v_batch, void* out,
                        float* lse, int B, int H, int Lq, int Lk, int causal, int q_offset, cudaStream_t stream, int f16 = 0);   // f16: fp16 q / k / v / out

// Flash-attention backward (attention_bwd.cu).  q/k/v: bf16 [B, L, H, 64] views sharing one token /
// batch stride (the packed projection); dout/out: contiguous bf16 [B, L, H*64]; lse: fp32 [B, H, L] from
// the forward.  Scratch: delta fp32 [B, H, L], dq_acc fp32 [B, L, H*64] (zeroed here).  dq/dk/dv: bf16
// [B, L, H, 64] views sharing dqkv_tok / dqkv_batch (normally the three slots of one packed tensor).
void attention_bwd_bf16(const void* q, const void* k, const void* v, long long qkv_tok, long long qkv_batch,
                        const void* dout, const void* out, const float* lse, float* delta, float* dq_acc, void* dq,
                        void* dk, void* dv, long long dqkv_tok, long long dqkv_batch, int B, int H, int L, int causal,
                        cudaStream_t stream, int f16 = 0);                          // f16: every 16-bit tensor is fp16

// Single-token decode: q bf16 [B, H*64] (batch stride q_batch), caches [B, Lmax, H, 64], attends to
// the first `len` cached positions.  out: bf16 [B, H*64].
void attention_decode_bf16(const void* q, long long q_batch, const void* kcache, const void* vcache,
                           long long tok_stride, long long batch_stride, void* out, int B, int H, int len,
                           cudaStream_t stream);

// Prefill: K / V slots of the packed projection qkv [B, L, 3, E] -> caches [B, Lmax, E] at [pos, pos+L)
void kv_cache_store_bf16(const void* qkv, void* kcache, void* vcache, long long tok_stride, long long batch_stride,
                         int B, int L, int E, int pos, cudaStream_t stream);

// Graph-replayable decode step pieces: the position lives in device memory.
void attention_decode_append_bf16(const void* qkv, void* kcache, void* vcache, long long tok_stride,
                                  long long batch_stride, void* out, int B, int H, const int* pos_ptr,
                                  cudaStream_t stream);
void embed_decode(const int* tok, const float* wte, const float* wpe, void* x, int B, int D, int V, const int* pos_ptr,
                  cudaStream_t stream);
void decode_epilogue(const int* next, int* tok, int* history, int B, int hist_cap, int* pos_ptr, int* step_ptr,
                     cudaStream_t stream);

}  // namespace tnn
