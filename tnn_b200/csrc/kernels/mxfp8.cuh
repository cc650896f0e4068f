// Block-scaled fp8 (MXFP8: e4m3 elements, one ue8m0 scale per 32 elements along K) -- host interface.
// Plain C++ (no torch).  Reference: the GPT-2 inference path the reference runs in fp32/fp16 through
// cuBLAS (examples/gpt2_inference.cpp:44-121, src/math/cuda/gemm.cu:74); BASELINE config 4 names fp8.
//
// Layouts
//   q   : uint8 (e4m3) row-major [rows, K], row stride ld (bytes), K % 128 == 0
//   sf  : uint8 (ue8m0) in the tcgen05 scale-factor atom layout: blocks of 128 rows x 4 scales (K = 128)
//         = 512 bytes each, block (rb, kb) at ((rb * (K/128) + kb) * 512); inside a block the scale of
//         row r (0..127), k-group g (0..3) lives at (r % 32) * 16 + (r / 32) * 4 + g -- i.e. exactly the
//         32 x 16-byte image that `tcgen05.cp.32x128b.warpx4` drops into 4 TMEM columns.
//         Allocate mxfp8_sf_bytes(rows, K) (one spare row block: tiles of 192 columns read two blocks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tnn {

inline long long mxfp8_sf_bytes(long long rows, long long K) {
  return ((rows + 127) / 128 + 1) * (K / 128) * 512;
}

// bf16 [rows, K] (row stride ldx elements) -> e4m3 q [rows, K] (row stride ldq bytes) + scales.
// Scale = 2^(floor(log2(amax)) - 8) per 32-element block (OCP MX rule for e4m3, emax = 8), elements saturate
// at +-448.  Rows >= `rows` of the last scale block are written as 1.0 (127).
void mxfp8_quantize_bf16(const void* x, long long ldx, long long rows, int K, void* q, long long ldq, void* sf,
                         cudaStream_t s);
// fp32 source (weights are fp32 masters)
void mxfp8_quantize_f32(const float* x, long long ldx, long long rows, int K, void* q, long long ldq, void* sf,
                        cudaStream_t s);
// q * 2^(sf - 127) -> fp32 [rows, K] (dense); test oracle and debugging aid
void mxfp8_dequantize(const void* q, long long ldq, const void* sf, long long rows, int K, float* out, cudaStream_t s);

// LayerNorm over rows of length C (C % 128 == 0, C <= 1024; bf16 in, fp32 affine) that writes the normalised row
// directly as MXFP8 (q + scales) for the following GEMM; `y_bf16` (nullable) also receives the bf16 result and
// `sum_out` (nullable) the residual stream x + residual.
void layernorm_fwd_mxfp8(const void* x, const void* residual, void* sum_out, long long rows, int C, const float* gamma,
                         const float* beta, float eps, void* q, void* sf, void* y_bf16, cudaStream_t s, bool pad_scales = true);
// (pad_scales = false: the caller keeps the scale rows beyond `rows` at 1.0 (byte 127) itself, e.g. a reused buffer that was
//  filled once -- saves one tiny launch per call)

struct Fp8Epilogue {
  void* out = nullptr;            // bf16 [M, N] row stride ldc (elements); null when only the fp8 output is wanted
  long long ldc = 0;
  const float* bias = nullptr;    // [N]
  const void* residual = nullptr; // bf16, same addressing as out
  int act = 0;                    // EpiAct (0 none, 1 relu, 2 gelu)
  float alpha = 1.f;
  // optional second output: the activation re-quantised to MXFP8 as the A operand of the NEXT GEMM
  void* out_q = nullptr;          // e4m3 [M, N], row stride ldq bytes
  long long ldq = 0;
  void* out_sf = nullptr;         // scales of out_q in the atom layout (K' = N)
  // optional segmented bf16 output (instead of `out`): the N columns are `seg_cols`-wide segments, segment s goes to
  //   seg_out[s] + (m / seg_rows) * seg_batch[s] + (m % seg_rows) * seg_tok[s] + (n - s * seg_cols)      (elements)
  // -- the fused QKV projection writes Q to its buffer and K / V straight into the KV caches at the prompt's positions
  // (no separate cache-store pass).  seg_cols % 32 == 0; no residual with segments.
  int seg_cols = 0, seg_rows = 0;
  void* seg_out[3] = {nullptr, nullptr, nullptr};
  long long seg_tok[3] = {0, 0, 0};
  long long seg_batch[3] = {0, 0, 0};
  // optional fp8 shadow of a segment (the fp8 KV cache of the decode kernel): e4m3 bytes at the SAME element offsets in
  // seg_q[s], one ue8m0 scale per 32 columns at (row offset / 32 + column / 32) in seg_qs[s]  (seg_tok / seg_batch % 32 == 0)
  void* seg_q[3] = {nullptr, nullptr, nullptr};
  void* seg_qs[3] = {nullptr, nullptr, nullptr};
};

// out[M, N] = act(alpha * (A . B^T) + bias + residual);  A [M, K] e4m3 (ld lda bytes) with scales sfa,
// B [N, K] e4m3 (ld ldb bytes) with scales sfb.  K % 128 == 0.  Persistent tcgen05 kernel
// (kind::mxf8f6f4.block_scale, scales staged smem -> TMEM with tcgen05.cp).
void gemm_mxfp8(const void* A, long long lda, const void* sfa, const void* B, long long ldb, const void* sfb, int M, int N,
                int K, const Fp8Epilogue& epi, cudaStream_t stream);

}  // namespace tnn
