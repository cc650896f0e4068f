// Host launchers of the SIMT (non tensor-core) kernels.  All activation tensors are
// NHWC / row-major with the channel (feature) dimension contiguous; "bf16" buffers are
// __nv_bfloat16, parameters / statistics / gradients of parameters are fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace tnn {

extern std::atomic<unsigned long long> g_launch_count;

// ---- elementwise (elementwise.cu) ------------------------------------------------
enum ActKind : int {
  AK_NONE = 0, AK_RELU = 1, AK_LEAKY_RELU = 2, AK_SIGMOID = 3, AK_TANH = 4, AK_ELU = 5, AK_GELU = 6
};
void cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t s);
void cast_bf16_to_f32(const void* in, float* out, long long n, cudaStream_t s);
// out = a + b (bf16), optional relu
void add_bf16(const void* a, const void* b, void* out, long long n, int relu, cudaStream_t s);
// N-ary elementwise on bf16 arrays: op 0 add, 1 sub, 2 mul
void binary_bf16(const void* a, const void* b, void* out, long long n, int op, cudaStream_t s);
void act_fwd(const void* x, void* y, long long n, int kind, float slope, int dt, cudaStream_t s);   // dt: XtDType code
// dx = dy * act'(x)   (x = pre-activation input)
void act_bwd(const void* x, const void* dy, void* dx, long long n, int kind, float slope, int is_f32,
             cudaStream_t s);
// inverted dropout with a counter-based RNG: the mask is a pure function of (seed, index), so
// backward recomputes it instead of storing it.
// effective seed = seed ^ (*seed_ptr * golden) when seed_ptr != null (device-resident step counter)
void dropout_fwd(const void* x, void* y, long long n, float p, uint64_t seed, const uint64_t* seed_ptr,
                 const void* residual, cudaStream_t s);
void fill_f32(float* x, long long n, float v, cudaStream_t s);
void axpy_f32(float* y, const float* x, float a, long long n, cudaStream_t s);
void scale_f32(float* x, float a, long long n, cudaStream_t s);
// column sums of a [M, N] bf16 matrix accumulated into fp32 out[N] (bias gradient)
void colsum_bf16(const void* x, long long M, int N, long long ld, float alpha, float* out, cudaStream_t s);
// stride-2 phase split: X[N,H,W,C] -> Xp[4][N,H/2,W/2,C] (phase = (h&1)*2 + (w&1)) and inverse
void phase_split(const void* x, void* xp, int N, int H, int W, int C, cudaStream_t s);
void phase_merge(const void* xp, void* x, int N, int H, int W, int C, cudaStream_t s);
void nchw_to_nhwc(const void* x, void* y, int N, int C, int H, int W, int is_f32, cudaStream_t s);
void nhwc_to_nchw(const void* x, void* y, int N, int C, int H, int W, int is_f32, cudaStream_t s);
// conv weights: fp32 [K,R,S,C] -> bf16 [K,R,S,C] and bf16 [C,R,S,K] with (r,s) flipped (dgrad)
// batched dgrad-weight flip: one descriptor per convolution, block_begin = running sum of ceil(C/32)*ceil(K/32)*R*S
struct ConvFlipDesc {
  const float* w;      // fp32 [K, R, S, C]
  void* wt;            // bf16 [C, R, S, K], taps flipped
  int K, R, S, C;
  int block_begin;
  int pad_;
};
void conv_weight_flip_batch(const ConvFlipDesc* descs_dev, int n, int total_blocks, cudaStream_t s);
void conv_weight_prep(const float* w, void* w_bf16, void* wt_bf16, int K, int R, int S, int C,
                      cudaStream_t s);
void philox_uniform_f32(float* x, long long n, float lo, float hi, uint64_t seed, cudaStream_t s);
void philox_normal_f32(float* x, long long n, float mean, float std, uint64_t seed, cudaStream_t s);
// generic [rows, cols] transposes (bf16)
void transpose_bf16(const void* x, void* y, int rows, int cols, cudaStream_t s);

// ---- BatchNorm over NCHW tensors, native layout (norm_nchw.cu); bf16 or fp32 activations, fp32 statistics --------
size_t bn_nchw_ws_bytes(int C);                      // fp64 (sum, sum^2) pairs; zero it once, the kernels reset it
void bn_nchw_train_stats(const void* x, int is_f32, int N, int C, int HW, const float* gamma, const float* beta, float eps, float momentum,
                         float* running_mean, float* running_var, float* save_mean, float* save_invstd, float* scale_shift, void* ws,
                         cudaStream_t s);
void bn_nchw_eval_scale_shift(int C, const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                              float* scale_shift, cudaStream_t s);
void bn_nchw_apply(const void* x, void* y, int is_f32, int N, int C, int HW, const float* scale_shift, int relu, cudaStream_t s);
// dx, dgamma += , dbeta += ; coef = [C][3] scratch, ws as above
void bn_nchw_backward(const void* dy, const void* x, void* dx, int is_f32, int N, int C, int HW, const float* gamma, const float* save_mean,
                      const float* save_invstd, const float* scale_shift, int relu, float* dgamma, float* dbeta, float* coef, void* ws,
                      cudaStream_t s);

// ---- generic tensor ops (tensor_ops.cu) -------------------------------------------
enum TensorBinaryOp : int { TOP_ADD = 0, TOP_SUB = 1, TOP_MUL = 2, TOP_DIV = 3, TOP_MIN = 4, TOP_MAX = 5, TOP_EQ = 6, TOP_GT = 7 };
enum TensorUnaryOp : int { UOP_SQRT = 0, UOP_RSQRT = 1, UOP_RCP = 2, UOP_ABS = 3, UOP_CLAMP = 4, UOP_NEG = 5, UOP_EXP = 6, UOP_LOG = 7, UOP_SQUARE = 8 };
// out = a (op) b  (b == nullptr: a (op) scalar); fp32 or bf16
void tensor_binary(const void* a, const void* b, float scalar, void* out, long long n, int op, int dt, cudaStream_t s);
void tensor_unary(const void* a, void* out, long long n, int op, float lo, float hi, int dt, cudaStream_t s);
// op 0 fmadd a*b+c, 1 fmsub a*b-c, 2 fnmadd -(a*b)+c
void tensor_ternary(const void* a, const void* b, const void* c, void* out, long long n, int op, int dt, cudaStream_t s);
// mode 0 sum(a), 1 dot(a,b), 2 sum((a-mu)^2); *out (device fp32) is overwritten; finishes on the device
void tensor_reduce(const void* a, const void* b, float mu, long long n, int mode, int dt, float* out, cudaStream_t s);
// dst[n,h,w,c] = src[n,h+off_h,w+off_w,c] or fill outside (pad: negative offsets; crop/unpad: positive)
void tensor_window_copy(const void* src, void* dst, int N, int SH, int SW, int DH, int DW, int C, int off_h, int off_w,
                        float fill, int dt, cudaStream_t s);
void softmax_rows(const void* x, void* y, long long rows, int C, int dt, cudaStream_t s);
void softmax_rows_bwd(const void* y, const void* dy, void* dx, long long rows, int C, int dt, cudaStream_t s);
void tensor_bswap(void* p, long long n_elems, int elem_size, cudaStream_t s);

// ---- skinny GEMM for decode (skinny_gemm.cu): out[M<=32,N] = act(x W^T + bias (+ residual)), act 0/1 relu/2 gelu
void skinny_gemm_bf16(const void* x, long long ldx, const void* w, long long ldw, const float* bias,
                      const void* residual, long long ldr, void* out, long long ldo, int M, int N, int K, int act, int static_weights,
                      cudaStream_t s);

// ---- normalisation (norm.cu) -------------------------------------------------------
// BatchNorm over a [M, C] bf16 view.  ws is a persistent fp32 workspace of 2*C+4 floats
// (zero-initialised once); scale_shift receives [2*C] = {gamma*invstd, beta - mean*gamma*invstd}.
void bn_train_stats(const void* x, long long M, int C, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    float* save_mean, float* save_invstd, float* scale_shift, float* ws, float* pack,
                    cudaStream_t s);
// finalize statistics accumulated by a producer epilogue into ws = [replicas][2][C] (see umma_gemm.cuh)
void bn_finalize_stats(float* ws, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                       float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                       float* scale_shift, float* pack, cudaStream_t s);
// `pack` (nullable, [C][4] fp32 = invstd, -mean*invstd, gamma*invstd, beta - mean*gamma*invstd) is the
// per-channel record the BatchNorm-backward fusion in the dgrad epilogue reads (umma_gemm.cuh).
// bn_bwd_coef folds the (sum g, sum g*xhat) statistics such an epilogue left in ws.
void bn_bwd_coef(float* ws, long long M, int C, const float* gamma, const float* save_invstd, float* dgamma,
                 float* dbeta, float* coef, cudaStream_t s);
int bn_stat_replicas();
void bn_eval_scale_shift(int C, const float* gamma, const float* beta, const float* running_mean,
                         const float* running_var, float eps, float* scale_shift, cudaStream_t s);
// y = act(x*scale + shift) [* dropout]; residual-free.  drop_p == 0 disables dropout.
// `mask` (nullable, M*C/8 bytes, C % 8 == 0): one byte of "gradient passes" bits per 8-channel vector.
void bn_apply(const void* x, void* y, long long M, int C, const float* scale_shift, int relu,
              float drop_p, uint64_t seed, const uint64_t* seed_ptr, uint8_t* mask, cudaStream_t s);
// Backward.  g = dy * mask(y) (mask = y != 0 when relu/dropout were fused, scaled by 1/(1-p)).
// Accumulates dgamma/dbeta (fp32, +=) and writes coef[3*C] for bn_bwd_apply.
void bn_bwd_reduce(const void* dy, const void* x, const void* y, long long M, int C,
                   const float* gamma, const float* save_mean, const float* save_invstd,
                   int masked, float drop_p, float* dgamma, float* dbeta, float* coef, float* ws,
                   const float* beta, int relu, uint64_t seed, const uint64_t* seed_ptr, const uint8_t* bits,
                   cudaStream_t s);
// dx = coef0*(g - coef1 - xhat*coef2) (+ extra, an optional residual gradient to add)
void bn_bwd_apply(const void* dy, const void* x, const void* y, const void* extra, void* dx,
                  long long M, int C, const float* save_mean, const float* save_invstd,
                  const float* coef, int masked, float drop_p, float* colsum_out, float* colsum_ws,
                  const float* gamma, const float* beta, int relu, uint64_t seed, const uint64_t* seed_ptr,
                  const uint8_t* bits, cudaStream_t s);
// floats needed by the `ws` / `colsum_ws` workspaces of the BN kernels (zero-initialised once)
int bn_ws_floats(int C);
// LayerNorm over rows of length C (bf16 io, fp32 affine).  One warp per row.
void layernorm_fwd(const void* x, const void* residual, void* y, void* sum_out, long long rows, int C,
                   const float* gamma, const float* beta, float eps, float* mean, float* rstd,
                   cudaStream_t s);
void layernorm_bwd(const void* dy, const void* x, long long rows, int C, const float* gamma,
                   const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                   const void* extra, cudaStream_t s);
// GroupNorm on NHWC [N, HW, C] with G groups.
void groupnorm_fwd(const void* x, void* y, int N, int HW, int C, int G, const float* gamma,
                   const float* beta, float eps, float* mean, float* rstd, cudaStream_t s);
void groupnorm_bwd(const void* dy, const void* x, int N, int HW, int C, int G, const float* gamma,
                   const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                   cudaStream_t s);

// ---- pooling (pool.cu) -----------------------------------------------------------
void maxpool_fwd(const void* x, void* y, uint8_t* mask, int N, int H, int W, int C, int kh, int kw,
                 int sh, int sw, int ph, int pw, int OH, int OW, cudaStream_t s);
void maxpool_bwd(const void* dy, const uint8_t* mask, void* dx, int N, int H, int W, int C, int kh,
                 int kw, int sh, int sw, int ph, int pw, int OH, int OW, cudaStream_t s);
void avgpool_fwd(const void* x, void* y, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                 int ph, int pw, int OH, int OW, cudaStream_t s);
void avgpool_bwd(const void* dy, void* dx, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                 int ph, int pw, int OH, int OW, cudaStream_t s);

// ---- loss / accuracy (loss.cu) ------------------------------------------------------
// Fused log-softmax cross entropy.  logits bf16 or fp32 [rows, V] (row stride ld), labels int32.
// out[0] += sum of per-row losses * loss_scale; out[1] += #correct (argmax == label).
// grad (same dtype/stride as logits, may be null) = (softmax - onehot) * grad_scale.
void softmax_ce(const void* logits, int is_f32, long long ld, const int* labels, long long rows, int V,
                float loss_scale, float grad_scale, void* grad, float* out, cudaStream_t s);
// regression losses on fp32/bf16 flattened tensors: kind 0 mse, 1 mae, 2 huber(delta)
void regression_loss(const void* pred, const void* target, int is_f32, long long n, int kind,
                     float delta, float grad_scale, void* grad, float* out, cudaStream_t s);

// ---- optimizers (optim.cu) ----------------------------------------------------------
struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay;
  int decoupled;
  float bias_c1, bias_c2;  // 1 - beta^t
  float grad_scale;        // applied to the gradient before use (1/world, 1/accum)
};
// One launch over the whole parameter slab.  Also refreshes the bf16 compute copy and zeroes
// the gradient slab (so no separate memset / cast kernels exist on the step path).
void adam_step_flat(float* p, float* g, float* m, float* v, void* p_bf16, long long n,
                    const AdamHyper& h, int zero_grad, cudaStream_t s);
void sgd_step_flat(float* p, float* g, float* mom, void* p_bf16, long long n, float lr,
                   float momentum, float weight_decay, float grad_scale, int zero_grad,
                   cudaStream_t s);
// Fused data-parallel step (the product path): barrier -> reduce my 1/W gradient slice over
// NVLink peer pointers (or one multimem.ld_reduce through NVLS when *_mc are given) -> scale ->
// Adam/SGD on my slice (sharded optimizer state) -> broadcast fp32 + bf16 params to every
// replica -> barrier -> zero my gradient slab.  All pointer arrays are HOST arrays of device
// pointers (they are copied into the kernel argument block).
void dp_fused_step(int world, int rank, long long n, long long chunk, float* const* g_peers,
                   float* const* p_peers, void* const* pb_peers, uint32_t* const* flag_peers,
                   const float* g_mc, float* p_mc, void* pb_mc, float* m, float* v,
                   uint32_t* sync, const float* state, int use_adam, int ctas, long long lo0, long long hi0, long long st_off,
                   cudaStream_t s);
// budget of the cross-GPU flag spins inside dp_fused_step before the kernel reports who is missing and traps (default 20 s)
void dp_set_watchdog_ms(unsigned long long ms);
// Graph-replayable variants: hyper-parameters come from a 16-float device "step state" that the
// host refreshes with one small H2D copy per step (outside the captured region).
void adam_step_flat_dev(float* p, float* g, float* m, float* v, void* p_bf16, long long n,
                        const float* state, int zero_grad, cudaStream_t s);
void sgd_step_flat_dev(float* p, float* g, float* mom, void* p_bf16, long long n, const float* state,
                       int zero_grad, cudaStream_t s);
// Device-side barrier across ranks through peer-mapped flag words (st.release.sys /
// ld.acquire.sys).  `flags` is a DEVICE array of pointers; flags[r] points at rank r's flag
// array (world uint32 each); `epoch` must increase on every call.
void p2p_barrier(uint32_t* const* flags, int world, int rank, uint32_t epoch, cudaStream_t s);
// Zero a gradient slab (used after the post-step barrier in the fused DP path).
void zero_f32(float* x, long long n, cudaStream_t s);

// ---- direct / fallback convolution (conv_simt.cu) -----------------------------------
// Small-C_in direct convolution (first layer, C_in < 8): x bf16 NHWC, w fp32 [K,R,S,C].
void conv_direct_fwd(const void* x, const float* w, const float* bias, void* y, int N, int H, int W,
                     int C, int K, int R, int S, int stride, int pad, int OH, int OW, cudaStream_t s);
void conv_direct_wgrad(const void* x, const void* dy, float* dw, float* dbias, int N, int H, int W,
                       int C, int K, int R, int S, int stride, int pad, int OH, int OW,
                       cudaStream_t s);
// im2col / col2im for geometries the TMA window path does not cover (bf16, NHWC):
// col is [N*OH*OW, R*S*C] with row stride ldcol.
void im2col_nhwc(const void* x, void* col, long long ldcol, int N, int H, int W, int C, int R, int S,
                 int stride, int pad, int OH, int OW, cudaStream_t s);
void col2im_nhwc(const void* col, long long ldcol, void* dx, int N, int H, int W, int C, int R, int S,
                 int stride, int pad, int OH, int OW, cudaStream_t s);
// anisotropic stride / padding (height, width)
// dt: element type code (0 bf16, 1 fp32, 2 fp16)
void im2col_nhwc2(const void* x, void* col, long long ldcol, int N, int H, int W, int C, int R, int S, int sh, int sw,
                  int ph, int pw, int OH, int OW, cudaStream_t s, int dt = 0);
void col2im_nhwc2(const void* col, long long ldcol, void* dx, int N, int H, int W, int C, int R, int S, int sh, int sw,
                  int ph, int pw, int OH, int OW, cudaStream_t s, int dt = 0);

// ---- embedding & friends (embed.cu) ---------------------------------------------------
void embedding_fwd(const int* ids, const float* table, const float* pos_table, void* y, long long T,
                   int L, int D, int V, cudaStream_t s);
void embedding_bwd(const int* ids, const void* dy, float* dtable, float* dpos, long long T, int L,
                   int D, int V, cudaStream_t s);
void argmax_rows(const void* x, int is_f32, long long ld, long long rows, int V, int* out,
                 cudaStream_t s);

}  // namespace tnn
