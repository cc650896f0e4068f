// Fused log-softmax cross-entropy (loss + gradient + accuracy in one pass over the logits)
// and regression losses.  The loss scalar stays on the device (out[0] accumulates), so the
// host never synchronises on it.  Reference: src/nn/loss_impl/cuda/loss_ops.cu:76-447 and
// src/nn/accuracy_impl/cuda/accuracy.cu:18 (three kernels + host reads there).
#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

template <typename T>
__device__ __forceinline__ float ldv(const T* p, long long i);
template <>
__device__ __forceinline__ float ldv<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float ldv<bf16>(const bf16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void stv(float* p, long long i, float v) { p[i] = v; }
__device__ __forceinline__ void stv(bf16* p, long long i, float v) { p[i] = __float2bfloat16(v); }

template <typename T>
__global__ void k_softmax_ce(const T* __restrict__ logits, long long ld, const int* __restrict__ labels,
                             long long rows, int V, float loss_scale, float grad_scale,
                             T* __restrict__ grad, float* __restrict__ out) {
  __shared__ float red[32];
  __shared__ int redi[32];
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* x = logits + r * ld;
    const int label = labels[r];
    // max + argmax
    float mx = -INFINITY;
    int am = 0;
    for (int c = threadIdx.x; c < V; c += blockDim.x) {
      float v = ldv(x, c);
      if (v > mx) {
        mx = v;
        am = c;
      }
    }
    // warp argmax
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, mx, o);
      int oi = __shfl_xor_sync(0xffffffffu, am, o);
      if (ov > mx || (ov == mx && oi < am)) {
        mx = ov;
        am = oi;
      }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) {
      red[w] = mx;
      redi[w] = am;
    }
    __syncthreads();
    mx = lane < nw ? red[lane] : -INFINITY;
    am = lane < nw ? redi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, mx, o);
      int oi = __shfl_xor_sync(0xffffffffu, am, o);
      if (ov > mx || (ov == mx && oi < am)) {
        mx = ov;
        am = oi;
      }
    }
    float se = 0.f;
    for (int c = threadIdx.x; c < V; c += blockDim.x) se += __expf(ldv(x, c) - mx);
    se = block_sum(se, red);
    const float lse = mx + __logf(se);
    if (threadIdx.x == 0) {
      const bool valid = label >= 0 && label < V;
      if (valid) atomicAdd(out, (lse - ldv(x, label)) * loss_scale);
      if (valid && am == label) atomicAdd(out + 1, 1.f);
    }
    if (grad) {
      T* g = grad + r * ld;
      const float inv = 1.f / se;
      for (int c = threadIdx.x; c < V; c += blockDim.x) {
        float pr = __expf(ldv(x, c) - mx) * inv;
        stv(g, c, (pr - (c == label ? 1.f : 0.f)) * grad_scale);
      }
    }
    __syncthreads();
  }
}
void softmax_ce(const void* logits, int is_f32, long long ld, const int* labels, long long rows, int V,
                float loss_scale, float grad_scale, void* grad, float* out, cudaStream_t s) {
  if (rows <= 0) return;
  int threads = V <= 1024 ? 128 : 512;
  int blocks = (int)(rows < kSMs * 8 ? rows : kSMs * 8);
  if (is_f32)
    k_softmax_ce<float><<<blocks, threads, 0, s>>>((const float*)logits, ld, labels, rows, V, loss_scale,
                                                   grad_scale, (float*)grad, out);
  else
    k_softmax_ce<bf16><<<blocks, threads, 0, s>>>((const bf16*)logits, ld, labels, rows, V, loss_scale,
                                                  grad_scale, (bf16*)grad, out);
  TNN_CHECK_LAUNCH();
}

template <typename T>
__global__ void k_regression(const T* __restrict__ pred, const T* __restrict__ tgt, long long n, int kind,
                             float delta, float grad_scale, T* __restrict__ grad, float* out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float d = ldv(pred, i) - ldv(tgt, i);
    float l, g;
    if (kind == 0) {
      l = d * d;
      g = 2.f * d;
    } else if (kind == 1) {
      l = fabsf(d);
      g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    } else {
      float a = fabsf(d);
      l = a <= delta ? 0.5f * d * d : delta * (a - 0.5f * delta);
      g = a <= delta ? d : (d > 0.f ? delta : -delta);
    }
    acc += l;
    if (grad) stv(grad, i, g * grad_scale);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc / (float)n);
}
void regression_loss(const void* pred, const void* target, int is_f32, long long n, int kind,
                     float delta, float grad_scale, void* grad, float* out, cudaStream_t s) {
  if (n <= 0) return;
  int g = grid_for(n, 256, kSMs * 4);
  if (is_f32)
    k_regression<float><<<g, 256, 0, s>>>((const float*)pred, (const float*)target, n, kind, delta,
                                          grad_scale, (float*)grad, out);
  else
    k_regression<bf16><<<g, 256, 0, s>>>((const bf16*)pred, (const bf16*)target, n, kind, delta,
                                         grad_scale, (bf16*)grad, out);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
