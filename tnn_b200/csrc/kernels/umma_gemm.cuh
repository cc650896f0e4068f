// Public (host) interface of the tcgen05 GEMM / implicit-GEMM convolution family.
// Plain C++ (no torch) so that every .cu compiles in seconds.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tnn {

enum EpiAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };
enum EpiOut : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ATOMIC = 2 };

// Epilogue applied to the fp32 accumulator tile while it is read out of TMEM:
//   v = alpha * acc (+ bias[n]) (+ residual[m,n]) -> act -> store / atomic-add
// Row addressing is either plain (m * ldc) or an NHW grid with arbitrary strides
// (used for phase-strided stores of stride-2 dgrad).
struct Epilogue {
  void* out = nullptr;
  int out_mode = OUT_BF16;
  int act = ACT_NONE;
  float alpha = 1.f;
  const float* bias = nullptr;          // [N] fp32
  const void* residual = nullptr;       // bf16, same addressing as out
  long long ldc = 0;                    // plain mode: elements between rows
  int grid_mode = 0;                    // 1: rows are (n, oh, ow) of an output grid
  int OH = 1, OW = 1;
  long long sN = 0, sH = 0, sW = 0, base = 0;
  // BatchNorm statistics fusion: fp32 workspace [col_repl][2][N] (sum, sum of squares of the stored
  // values per column), replica row = blockIdx.x % col_repl; folded by bn_finalize_stats (norm.cu)
  float* col_sum = nullptr;
  int col_repl = 1;
  // BatchNorm BACKWARD fusion, for the dgrad of a convolution whose input is BN(+ReLU)(+dropout)(x):
  // the value this GEMM produces is d(BN output); with bnb_x set the epilogue turns it into
  //   g = mask(x) * dy * bnb_dscale        mask = [x*scale + shift > 0] (ReLU) and the dropout keep bit
  // stores g, and col_sum receives (sum g, sum g * xhat) per channel -- exactly what the BatchNorm
  // backward needs, so its two-tensor reduction pass disappears (norm.cu: bn_bwd_coef + bn_bwd_apply).
  const void* bnb_x = nullptr;          // bf16 BN input, same addressing as out
  const float* bnb_pack = nullptr;      // [N][4] fp32: invstd, -mean*invstd, scale, shift
  int bnb_relu = 0;
  float bnb_drop = 0.f;                 // dropout probability (0 = none)
  float bnb_dscale = 1.f;               // 1 / (1 - p)
  unsigned long long bnb_seed = 0;
  const unsigned long long* bnb_seed_ptr = nullptr;   // device step counter mixed into the seed
};

struct ConvTap {
  int dh, dw;      // spatial shift of the input window relative to the output pixel
  int nb;          // image-index offset (selects a stride-2 phase plane)
  int koff;        // column offset of this tap inside the [N, Ktot] weight matrix
};
static constexpr int kMaxTaps = 9;

struct ConvGeom {
  // input activation grid, NHWC, C contiguous
  int IN, IH, IW, C;
  // output grid (GEMM M = ON*OH*OW); OW must divide 128 (see pick_tile in .cu)
  int ON, OH, OW;
  int ntaps;
  ConvTap taps[kMaxTaps];
};

// C[M,N] = A * B^T-ish.  A is [M,K] (K-major, a_mn=false) or [K,M] (MN-major, a_mn=true),
// row strides lda/ldb in elements (must be multiples of 8).  B likewise ([N,K] or [K,N]).
// splits>1 => split-K, requires out_mode == OUT_F32_ATOMIC.
void gemm_bf16(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb,
               bool a_mn, bool b_mn, const Epilogue& epi, int splits, cudaStream_t stream);

// Implicit-GEMM convolution forward / data-gradient: out[m, n] = sum_t sum_c
//   X[n_img + nb_t, oh + dh_t, ow + dw_t, c] * Wt[n, koff_t + c]
// X bf16 NHWC, Wt bf16 [N, ldw] K-major.
void conv_fprop_bf16(const void* X, const void* Wt, long long ldw, int N, const ConvGeom& g,
                     const Epilogue& epi, cudaStream_t stream);

// Implicit-GEMM weight gradient for one group of (up to 3) taps:
//   dW[k, koff_t + c] += sum_pixels dY[pix, k] * X[pix shifted by tap t, c]
// dY bf16 [ON*OH*OW, Kout] (ld = ldy), X bf16 NHWC, dW fp32 [Kout, ldw] (atomic accumulate).
void conv_wgrad_bf16(const void* dY, long long ldy, int Kout, const void* X, const ConvGeom& g,
                     float* dW, long long ldw, float alpha, cudaStream_t stream);

bool conv_geom_supported(const ConvGeom& g);

}  // namespace tnn
