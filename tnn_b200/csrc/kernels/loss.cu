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
// bf16 rows with 16-byte aligned starts (ld % 8 == 0): two passes of 16-byte loads instead of three
// passes of 2-byte loads.  Pass 1 is an online softmax (running max / rescaled sum / arg-max per
// thread, merged across the CTA), pass 2 re-reads the row (L2 resident: 100 KB at GPT-2's vocabulary)
// and writes the gradient as 16-byte vectors.  [8192, 50257] logits: 1.39 ms -> bandwidth bound.
__global__ void __launch_bounds__(512)
    k_softmax_ce_bf16v(const bf16* __restrict__ logits, long long ld, const int* __restrict__ labels,
                       long long rows, int V, float loss_scale, float grad_scale, bf16* __restrict__ grad,
                       float* __restrict__ out) {
  __shared__ float s_m[16], s_s[16], s_bv[16];
  __shared__ int s_bi[16];
  __shared__ float s_M, s_S;
  __shared__ int s_A;
  const int nv = V >> 3;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const bf16* x = logits + r * ld;
    const int label = labels[r];
    float m = -INFINITY, sum = 0.f, bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      float f[8];
      bf8_unpack(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
      float lm = f[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) lm = fmaxf(lm, f[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (f[j] > bv) {
          bv = f[j];
          bi = i * 8 + j;
        }
      const float mn = fmaxf(m, lm);
      float add = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) add += __expf(f[j] - mn);
      sum = sum * __expf(m - mn) + add;
      m = mn;
    }
    for (int c = (nv << 3) + threadIdx.x; c < V; c += blockDim.x) {      // ragged tail (V % 8 elements)
      const float v = __bfloat162float(x[c]);
      if (v > bv || (v == bv && c < bi)) {
        bv = v;
        bi = c;
      }
      const float mn = fmaxf(m, v);
      sum = sum * __expf(m - mn) + __expf(v - mn);
      m = mn;
    }
    // warp merge of (m, sum) and (bv, bi)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, sum, o);
      const float mn = fmaxf(m, om);
      sum = (mn == -INFINITY) ? 0.f : sum * __expf(m - mn) + os * __expf(om - mn);
      m = mn;
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    __syncthreads();
    if (lane == 0) {
      s_m[w] = m;
      s_s[w] = sum;
      s_bv[w] = bv;
      s_bi[w] = bi;
    }
    __syncthreads();
    if (w == 0) {
      m = lane < nw ? s_m[lane] : -INFINITY;
      sum = lane < nw ? s_s[lane] : 0.f;
      bv = lane < nw ? s_bv[lane] : -INFINITY;
      bi = lane < nw ? s_bi[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, sum, o);
        const float mn = fmaxf(m, om);
        sum = (mn == -INFINITY) ? 0.f : sum * __expf(m - mn) + os * __expf(om - mn);
        m = mn;
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_M = m;
        s_S = sum;
        s_A = bi;
        const bool valid = label >= 0 && label < V;
        if (valid) atomicAdd(out, (m + __logf(sum) - __bfloat162float(x[label])) * loss_scale);
        if (valid && bi == label) atomicAdd(out + 1, 1.f);
      }
    }
    __syncthreads();
    if (grad != nullptr) {
      const float M = s_M, inv = 1.f / s_S;
      bf16* g = grad + r * ld;
      for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        float f[8];
        bf8_unpack(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (__expf(f[j] - M) * inv - ((i * 8 + j) == label ? 1.f : 0.f)) * grad_scale;
        reinterpret_cast<uint4*>(g)[i] = bf8_pack(f);
      }
      for (int c = (nv << 3) + threadIdx.x; c < V; c += blockDim.x)
        g[c] = __float2bfloat16((__expf(__bfloat162float(x[c]) - M) * inv - (c == label ? 1.f : 0.f)) * grad_scale);
    }
    __syncthreads();
  }
}

void softmax_ce(const void* logits, int is_f32, long long ld, const int* labels, long long rows, int V,
                float loss_scale, float grad_scale, void* grad, float* out, cudaStream_t s) {
  if (rows <= 0) return;
  if (!is_f32 && V >= 2048 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
      (grad == nullptr || (reinterpret_cast<uintptr_t>(grad) & 15) == 0)) {
    const int blocks = (int)(rows < kSMs * 4 ? rows : kSMs * 4);
    k_softmax_ce_bf16v<<<blocks, 512, 0, s>>>((const bf16*)logits, ld, labels, rows, V, loss_scale, grad_scale,
                                             (bf16*)grad, out);
    TNN_CHECK_LAUNCH();
    return;
  }
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
