// Host interface of the GEMMs / casts for the element types beside bf16 (fp16, fp32-as-tf32, fp64).  Plain C++.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tnn {

// dtype codes shared by the generic tensor ops (tensor_ops.cu) and the casts: 0 / 1 keep the meaning of the old `is_f32` flag
enum XtDType : int { XT_BF16 = 0, XT_F32 = 1, XT_F16 = 2, XT_F64 = 3 };

// v = alpha * acc (+ bias[n]) (+ residual[m, n]) -> act (0 none, 1 relu, 2 gelu-tanh) -> out[m, n] (= or +=)
struct XtEpilogue {
  void* out = nullptr;                // operand type, or fp32 when out_f32 (always fp32 for the tf32 GEMM)
  int out_f32 = 0;
  long long ldc = 0;                  // elements between output rows
  const float* bias = nullptr;        // [N] fp32
  const void* residual = nullptr;     // operand type
  long long ldr = 0;                  // 0: same as ldc
  int act = 0;
  float alpha = 1.f;
  int accumulate = 0;                 // out += v (weight gradients over micro-batches)
};

// C[M, N] = A x B^T-ish: A is [M, K] (K-major) or [K, M] (a_mn), B is [N, K] or [K, N] (b_mn); lda / ldb in elements.
void gemm_f16(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb, bool a_mn, bool b_mn, const XtEpilogue& epi,
              cudaStream_t stream);
void gemm_tf32(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb, bool a_mn, bool b_mn, const XtEpilogue& epi,
               cudaStream_t stream);
void gemm_f64(const double* A, const double* B, int M, int N, int K, long long lda, long long ldb, bool a_mn, bool b_mn, double alpha,
              const double* bias, double* C, long long ldc, bool accumulate, cudaStream_t stream);

// out[cols] (= or +=) alpha * column sums of x[rows, cols] (row stride ld); out is fp32 (fp64 for fp64 input)
void colsum_any(const void* x, int dt, long long rows, int cols, long long ld, double alpha, void* out, bool accumulate, cudaStream_t stream);
// LayerNorm for fp16 / fp32 / fp64 rows (bf16 has the tuned kernels of norm.cu): optional residual add (sum_out keeps x + residual
// for the backward pass), fp32 gamma / beta / statistics; backward adds `extra` (same dtype) to dx and accumulates dgamma / dbeta
void layernorm_fwd_any(const void* x, const void* residual, void* y, void* sum_out, int dt, long long rows, int C, const float* gamma,
                       const float* beta, float eps, float* mean, float* rstd, cudaStream_t stream);
void layernorm_bwd_any(const void* dy, const void* x, int dt, long long rows, int C, const float* gamma, const float* mean, const float* rstd,
                       void* dx, float* dgamma, float* dbeta, const void* extra, cudaStream_t stream);
void tensor_cast_any(const void* src, int src_dt, void* dst, int dst_dt, long long n, cudaStream_t stream);
// dst[c, r] = src[r, c]; src_ld / dst_ld = elements between rows (0: dense)
void tensor_transpose_any(const void* src, void* dst, int dt, int rows, int cols, cudaStream_t stream, long long src_ld = 0,
                          long long dst_ld = 0);

}  // namespace tnn
