// Generic elementwise / layout / RNG kernels (128-bit accesses, grid-stride).
// Reference equivalents: src/ops/cuda/kernels.cu, src/nn/activations_impl/cuda/*.cu,
// src/nn/layers_impl/cuda/dropout.cu, src/tensor/cuda/tensor_ops.cu -- here they are one
// templated family, and most call sites vanish through epilogue fusion.
#include <cuda_fp16.h>

#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

std::atomic<unsigned long long> g_launch_count{0};

// ---------------------------------------------------------------- casts
__global__ void k_cast_f32_bf16(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  long long nv = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<const float4*>(in)[2 * i], b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    reinterpret_cast<uint4*>(out)[i] = bf8_pack(f);
  }
  for (long long i = nv * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}
__global__ void k_cast_bf16_f32(const bf16* __restrict__ in, float* __restrict__ out, long long n) {
  long long nv = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    bf8_unpack(reinterpret_cast<const uint4*>(in)[i], f);
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  for (long long i = nv * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}
void cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t s) {
  if (n <= 0) return;
  k_cast_f32_bf16<<<grid_for(n / 8 + 1, 256), 256, 0, s>>>(in, (bf16*)out, n);
  TNN_CHECK_LAUNCH();
}
void cast_bf16_to_f32(const void* in, float* out, long long n, cudaStream_t s) {
  if (n <= 0) return;
  k_cast_bf16_f32<<<grid_for(n / 8 + 1, 256), 256, 0, s>>>((const bf16*)in, out, n);
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- binary
template <int OP, bool RELU>
__global__ void k_binary_bf16(const bf16* __restrict__ a, const bf16* __restrict__ b,
                              bf16* __restrict__ o, long long n) {
  long long nv = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    bf8_unpack(reinterpret_cast<const uint4*>(a)[i], x);
    bf8_unpack(reinterpret_cast<const uint4*>(b)[i], y);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = OP == 0 ? x[j] + y[j] : (OP == 1 ? x[j] - y[j] : x[j] * y[j]);
      x[j] = RELU ? fmaxf(v, 0.f) : v;
    }
    reinterpret_cast<uint4*>(o)[i] = bf8_pack(x);
  }
  for (long long i = nv * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float x = __bfloat162float(a[i]), y = __bfloat162float(b[i]);
    float v = OP == 0 ? x + y : (OP == 1 ? x - y : x * y);
    o[i] = __float2bfloat16(RELU ? fmaxf(v, 0.f) : v);
  }
}
void add_bf16(const void* a, const void* b, void* out, long long n, int relu, cudaStream_t s) {
  if (n <= 0) return;
  int g = grid_for(n / 8 + 1, 256);
  if (relu) k_binary_bf16<0, true><<<g, 256, 0, s>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n);
  else k_binary_bf16<0, false><<<g, 256, 0, s>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n);
  TNN_CHECK_LAUNCH();
}
void binary_bf16(const void* a, const void* b, void* out, long long n, int op, cudaStream_t s) {
  if (n <= 0) return;
  int g = grid_for(n / 8 + 1, 256);
  if (op == 0) k_binary_bf16<0, false><<<g, 256, 0, s>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n);
  else if (op == 1) k_binary_bf16<1, false><<<g, 256, 0, s>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n);
  else k_binary_bf16<2, false><<<g, 256, 0, s>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n);
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- activations
__device__ __forceinline__ float act_f(float x, int kind, float slope) {
  switch (kind) {
    case AK_RELU: return fmaxf(x, 0.f);
    case AK_LEAKY_RELU: return x > 0.f ? x : slope * x;
    case AK_SIGMOID: return 1.f / (1.f + __expf(-x));
    case AK_TANH: return tanhf(x);
    case AK_ELU: return x > 0.f ? x : slope * (__expf(x) - 1.f);
    case AK_GELU: {
      float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.f + tanhf(u));
    }
    default: return x;
  }
}
__device__ __forceinline__ float act_df(float x, int kind, float slope) {
  switch (kind) {
    case AK_RELU: return x > 0.f ? 1.f : 0.f;
    case AK_LEAKY_RELU: return x > 0.f ? 1.f : slope;
    case AK_SIGMOID: {
      float s = 1.f / (1.f + __expf(-x));
      return s * (1.f - s);
    }
    case AK_TANH: {
      float t = tanhf(x);
      return 1.f - t * t;
    }
    case AK_ELU: return x > 0.f ? 1.f : slope * __expf(x);
    case AK_GELU: {
      float x2 = x * x;
      float u = 0.7978845608028654f * (x + 0.044715f * x * x2);
      float t = tanhf(u);
      float du = 0.7978845608028654f * (1.f + 3.f * 0.044715f * x2);
      return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * du;
    }
    default: return 1.f;
  }
}
// fp64 tensors evaluate their activations in double precision
__device__ __forceinline__ double act_f(double x, int kind, double slope) {
  switch (kind) {
    case AK_RELU: return fmax(x, 0.0);
    case AK_LEAKY_RELU: return x > 0.0 ? x : slope * x;
    case AK_SIGMOID: return 1.0 / (1.0 + exp(-x));
    case AK_TANH: return tanh(x);
    case AK_ELU: return x > 0.0 ? x : slope * (exp(x) - 1.0);
    case AK_GELU: {
      const double u = 0.7978845608028654 * (x + 0.044715 * x * x * x);
      return 0.5 * x * (1.0 + tanh(u));
    }
    default: return x;
  }
}
__device__ __forceinline__ double act_df(double x, int kind, double slope) {
  switch (kind) {
    case AK_RELU: return x > 0.0 ? 1.0 : 0.0;
    case AK_LEAKY_RELU: return x > 0.0 ? 1.0 : slope;
    case AK_SIGMOID: {
      const double s = 1.0 / (1.0 + exp(-x));
      return s * (1.0 - s);
    }
    case AK_TANH: {
      const double t = tanh(x);
      return 1.0 - t * t;
    }
    case AK_ELU: return x > 0.0 ? 1.0 : slope * exp(x);
    case AK_GELU: {
      const double x2 = x * x;
      const double u = 0.7978845608028654 * (x + 0.044715 * x * x2);
      const double t = tanh(u);
      const double du = 0.7978845608028654 * (1.0 + 3.0 * 0.044715 * x2);
      return 0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * du;
    }
    default: return 1.0;
  }
}
__global__ void k_act_fwd_f64(const double* __restrict__ x, double* __restrict__ y, long long n, int kind, double slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = act_f(x[i], kind, slope);
}
__global__ void k_act_bwd_f64(const double* __restrict__ x, const double* __restrict__ dy, double* __restrict__ dx, long long n, int kind,
                              double slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * act_df(x[i], kind, slope);
}
template <typename T>
__device__ __forceinline__ float ld1(const T* p, long long i);
template <>
__device__ __forceinline__ float ld1<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float ld1<bf16>(const bf16* p, long long i) { return __bfloat162float(p[i]); }
template <>
__device__ __forceinline__ float ld1<__half>(const __half* p, long long i) { return __half2float(p[i]); }
__device__ __forceinline__ void st1(float* p, long long i, float v) { p[i] = v; }
__device__ __forceinline__ void st1(bf16* p, long long i, float v) { p[i] = __float2bfloat16(v); }
__device__ __forceinline__ void st1(__half* p, long long i, float v) { p[i] = __float2half_rn(v); }

template <typename T>
__global__ void k_act_fwd(const T* __restrict__ x, T* __restrict__ y, long long n, int kind, float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    st1(y, i, act_f(ld1(x, i), kind, slope));
}
template <typename T>
__global__ void k_act_bwd(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                          long long n, int kind, float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    st1(dx, i, ld1(dy, i) * act_df(ld1(x, i), kind, slope));
}
// vectorised bf16 specialisations
__global__ void k_act_fwd_bf8(const bf16* __restrict__ x, bf16* __restrict__ y, long long nv, int kind,
                              float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    bf8_unpack(reinterpret_cast<const uint4*>(x)[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_f(f[j], kind, slope);
    reinterpret_cast<uint4*>(y)[i] = bf8_pack(f);
  }
}
__global__ void k_act_bwd_bf8(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                              bf16* __restrict__ dx, long long nv, int kind, float slope) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8], g[8];
    bf8_unpack(reinterpret_cast<const uint4*>(x)[i], f);
    bf8_unpack(reinterpret_cast<const uint4*>(dy)[i], g);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = g[j] * act_df(f[j], kind, slope);
    reinterpret_cast<uint4*>(dx)[i] = bf8_pack(f);
  }
}
// dt: 0 bf16, 1 fp32, 2 fp16, 3 fp64 (the XtDType codes of gemm_xt.cuh)
void act_fwd(const void* x, void* y, long long n, int kind, float slope, int dt, cudaStream_t s) {
  if (n <= 0) return;
  if (dt == 1) {
    k_act_fwd<float><<<grid_for(n, 256), 256, 0, s>>>((const float*)x, (float*)y, n, kind, slope);
  } else if (dt == 2) {
    k_act_fwd<__half><<<grid_for(n, 256), 256, 0, s>>>((const __half*)x, (__half*)y, n, kind, slope);
  } else if (dt == 3) {
    k_act_fwd_f64<<<grid_for(n, 256), 256, 0, s>>>((const double*)x, (double*)y, n, kind, (double)slope);
  } else {
    long long nv = n / 8;
    if (nv) k_act_fwd_bf8<<<grid_for(nv, 256), 256, 0, s>>>((const bf16*)x, (bf16*)y, nv, kind, slope);
    if (n - nv * 8)
      k_act_fwd<bf16><<<1, 32, 0, s>>>((const bf16*)x + nv * 8, (bf16*)y + nv * 8, n - nv * 8, kind, slope);
  }
  TNN_CHECK_LAUNCH();
}
void act_bwd(const void* x, const void* dy, void* dx, long long n, int kind, float slope, int dt,
             cudaStream_t s) {
  if (n <= 0) return;
  if (dt == 1) {
    k_act_bwd<float><<<grid_for(n, 256), 256, 0, s>>>((const float*)x, (const float*)dy, (float*)dx, n, kind, slope);
  } else if (dt == 2) {
    k_act_bwd<__half><<<grid_for(n, 256), 256, 0, s>>>((const __half*)x, (const __half*)dy, (__half*)dx, n, kind, slope);
  } else if (dt == 3) {
    k_act_bwd_f64<<<grid_for(n, 256), 256, 0, s>>>((const double*)x, (const double*)dy, (double*)dx, n, kind, (double)slope);
  } else {
    long long nv = n / 8;
    if (nv)
      k_act_bwd_bf8<<<grid_for(nv, 256), 256, 0, s>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, nv, kind, slope);
    if (n - nv * 8)
      k_act_bwd<bf16><<<1, 32, 0, s>>>((const bf16*)x + nv * 8, (const bf16*)dy + nv * 8,
                                       (bf16*)dx + nv * 8, n - nv * 8, kind, slope);
  }
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- dropout
// y = dropout(x) (+ res: the residual branch of the enclosing block, so "x + dropout(f(x))" is one kernel)
__global__ void k_dropout(const bf16* __restrict__ x, bf16* __restrict__ y, long long n, float p,
                          float scale, uint64_t seed, const uint64_t* __restrict__ seed_ptr,
                          const bf16* __restrict__ res) {
  if (seed_ptr) seed ^= (*seed_ptr) * 0x9E3779B97F4A7C15ull;
  long long nv = (n + 7) / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    bool keep[8];
    dropout_keep8(seed, i, p, keep);
    if (i * 8 + 8 <= n) {
      float f[8];
      bf8_unpack(reinterpret_cast<const uint4*>(x)[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = keep[j] ? f[j] * scale : 0.f;
      if (res != nullptr) {
        float r[8];
        bf8_unpack(__ldg(reinterpret_cast<const uint4*>(res) + i), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += r[j];
      }
      reinterpret_cast<uint4*>(y)[i] = bf8_pack(f);
    } else {
      for (int j = 0; j < 8 && i * 8 + j < n; ++j) {
        float v = keep[j] ? __bfloat162float(x[i * 8 + j]) * scale : 0.f;
        if (res != nullptr) v += __bfloat162float(res[i * 8 + j]);
        y[i * 8 + j] = __float2bfloat16(v);
      }
    }
  }
}
void dropout_fwd(const void* x, void* y, long long n, float p, uint64_t seed, const uint64_t* seed_ptr,
                 const void* residual, cudaStream_t s) {
  if (n <= 0) return;
  k_dropout<<<grid_for((n + 7) / 8, 256), 256, 0, s>>>((const bf16*)x, (bf16*)y, n, p, 1.f / (1.f - p), seed,
                                                      seed_ptr, (const bf16*)residual);
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- fp32 vector ops
__global__ void k_fill(float* x, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    x[i] = v;
}
__global__ void k_axpy(float* y, const float* x, float a, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}
__global__ void k_scale(float* x, float a, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    x[i] *= a;
}
void fill_f32(float* x, long long n, float v, cudaStream_t s) {
  if (n <= 0) return;
  k_fill<<<grid_for(n, 256), 256, 0, s>>>(x, n, v);
  TNN_CHECK_LAUNCH();
}
void zero_f32(float* x, long long n, cudaStream_t s) { fill_f32(x, n, 0.f, s); }
void axpy_f32(float* y, const float* x, float a, long long n, cudaStream_t s) {
  if (n <= 0) return;
  k_axpy<<<grid_for(n, 256), 256, 0, s>>>(y, x, a, n);
  TNN_CHECK_LAUNCH();
}
void scale_f32(float* x, float a, long long n, cudaStream_t s) {
  if (n <= 0) return;
  k_scale<<<grid_for(n, 256), 256, 0, s>>>(x, a, n);
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- column sums (bias grad)
// block = 256 threads = (256/LPR) row-lanes x LPR column-groups of 8 channels.
__global__ void k_colsum_bf16(const bf16* __restrict__ x, long long M, int N, long long ld, float alpha,
                              float* __restrict__ out) {
  extern __shared__ float sm[];
  const int groups = (N + 7) / 8;            // 8-wide column groups
  const int lpr = min(groups, 256);          // column groups handled per row pass
  const int rpb = 256 / lpr;                 // rows in flight per block
  const int cg = threadIdx.x % lpr, rl = threadIdx.x / lpr;
  const bool vec = (N % 8 == 0) && (ld % 8 == 0);
  for (int g0 = 0; g0 < groups; g0 += lpr) {
    const int g = g0 + cg;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (g < groups && rl < rpb) {
      const long long stride = (long long)gridDim.x * rpb;
      long long r = (long long)blockIdx.x * rpb + rl;
      if (vec) {
        for (; r + 3 * stride < M; r += 4 * stride) {
          uint4 q[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * stride) * ld + g * 8));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float f[8];
            bf8_unpack(q[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
          }
        }
      }
      for (; r < M; r += stride) {
        if (vec) {
          float f[8];
          bf8_unpack(*reinterpret_cast<const uint4*>(x + r * ld + g * 8), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        } else {
          for (int j = 0; j < 8 && g * 8 + j < N; ++j) acc[j] += __bfloat162float(x[r * ld + g * 8 + j]);
        }
      }
    }
    // reduce across row-lanes through shared memory
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[(rl * lpr + cg) * 8 + j] = (rl < rpb) ? acc[j] : 0.f;
    __syncthreads();
    if (rl == 0 && g < groups) {
      for (int j = 0; j < 8 && g * 8 + j < N; ++j) {
        float t = 0.f;
        for (int q = 0; q < rpb; ++q) t += sm[(q * lpr + cg) * 8 + j];
        atomicAdd(out + g * 8 + j, t * alpha);
      }
    }
  }
}
// Tall matrices (bias gradients of the transformer GEMMs: [8192, 768 .. 50257]): one CTA per
// (64-column tile, row split); 8 column groups x 32 row-lanes, eight 16-byte loads in flight per
// thread, one shared-memory reduction and 64 atomics per CTA.  (The generic kernel above keeps the
// CTA count <= 148 and 2..8 rows in flight: 0.23 TB/s on these shapes, 11 % of a GPT-2 training step.)
__global__ void __launch_bounds__(256)
    k_colsum_tall(const bf16* __restrict__ x, long long M, int N, long long ld, float alpha, float* __restrict__ out,
                  long long rows_per_split) {
  __shared__ float sm[32][65];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(M, r_begin + rows_per_split);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < N) {
    const bool fullv = c0 + 8 <= N;
    long long r = r_begin + rl;
    if (fullv) {
      for (; r + 7 * 32 < r_end; r += 8 * 32) {
        uint4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * 32) * ld + c0));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float f[8];
          bf8_unpack(q[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
      }
      for (; r < r_end; r += 32) {
        float f[8];
        bf8_unpack(__ldg(reinterpret_cast<const uint4*>(x + r * ld + c0)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    } else {
      for (; r < r_end; r += 32)
        for (int j = 0; j < 8 && c0 + j < N; ++j) acc[j] += __bfloat162float(x[r * ld + c0 + j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += sm[q][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < N) atomicAdd(out + c, t * alpha);
  }
}

void colsum_bf16(const void* x, long long M, int N, long long ld, float alpha, float* out, cudaStream_t s) {
  if (M <= 0 || N <= 0) return;
  if (M >= 2048 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int col_tiles = (N + 63) / 64;
    int splits = (4 * kSMs + col_tiles - 1) / col_tiles;
    const long long max_splits = M / 256 > 0 ? M / 256 : 1;        // >= 256 rows (one unrolled pass) per CTA
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    const long long rps = ((M + splits - 1) / splits + 31) / 32 * 32;
    splits = (int)((M + rps - 1) / rps);
    dim3 grid(col_tiles, splits);
    k_colsum_tall<<<grid, 256, 0, s>>>((const bf16*)x, M, N, ld, alpha, out, rps);
    TNN_CHECK_LAUNCH();
    return;
  }
  int groups = (N + 7) / 8;
  int lpr = groups < 256 ? groups : 256;
  int rpb = 256 / lpr;
  int blocks = (int)((M + rpb * 4 - 1) / (rpb * 4));
  if (blocks > kSMs) blocks = kSMs;   // same-address atomics: keep the CTA count low
  if (blocks < 1) blocks = 1;
  k_colsum_bf16<<<blocks, 256, 256 * 8 * sizeof(float), s>>>((const bf16*)x, M, N, ld, alpha, out);
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- stride-2 phase split
template <bool SPLIT>
__global__ void k_phase(const bf16* __restrict__ src, bf16* __restrict__ dst, int N, int H, int W, int C8) {
  // one thread per 8-channel vector of the full-resolution tensor
  const int H2 = H / 2, W2 = W / 2;
  long long total = (long long)N * H * W * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C8);
    long long t = i / C8;
    int w = (int)(t % W);
    t /= W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    int ph = (h & 1) * 2 + (w & 1);
    long long j = ((((long long)ph * N + n) * H2 + (h >> 1)) * W2 + (w >> 1)) * C8 + c;
    if (SPLIT) reinterpret_cast<uint4*>(dst)[j] = reinterpret_cast<const uint4*>(src)[i];
    else reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[j];
  }
}
void phase_split(const void* x, void* xp, int N, int H, int W, int C, cudaStream_t s) {
  long long total = (long long)N * H * W * (C / 8);
  k_phase<true><<<grid_for(total, 256), 256, 0, s>>>((const bf16*)x, (bf16*)xp, N, H, W, C / 8);
  TNN_CHECK_LAUNCH();
}
void phase_merge(const void* xp, void* x, int N, int H, int W, int C, cudaStream_t s) {
  long long total = (long long)N * H * W * (C / 8);
  k_phase<false><<<grid_for(total, 256), 256, 0, s>>>((const bf16*)xp, (bf16*)x, N, H, W, C / 8);
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- layout transposes
// [B, R, Cc] -> [B, Cc, R] tiled 32x32 through shared memory
template <typename T>
__global__ void k_transpose_batched(const T* __restrict__ x, T* __restrict__ y, int R, int Cc) {
  __shared__ T tile[32][33];
  const long long b = blockIdx.z;
  const T* xb = x + b * (long long)R * Cc;
  T* yb = y + b * (long long)R * Cc;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[j][threadIdx.x] = xb[(long long)r * Cc + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Cc) yb[(long long)c * R + r] = tile[threadIdx.x][j];
  }
}
template <typename T>
static void transpose_batched(const T* x, T* y, int B, int R, int Cc, cudaStream_t s) {
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, B), block(32, 8);
  k_transpose_batched<T><<<grid, block, 0, s>>>(x, y, R, Cc);
  TNN_CHECK_LAUNCH();
}
void nchw_to_nhwc(const void* x, void* y, int N, int C, int H, int W, int is_f32, cudaStream_t s) {
  if (is_f32) transpose_batched<float>((const float*)x, (float*)y, N, C, H * W, s);
  else transpose_batched<bf16>((const bf16*)x, (bf16*)y, N, C, H * W, s);
}
void nhwc_to_nchw(const void* x, void* y, int N, int C, int H, int W, int is_f32, cudaStream_t s) {
  if (is_f32) transpose_batched<float>((const float*)x, (float*)y, N, H * W, C, s);
  else transpose_batched<bf16>((const bf16*)x, (bf16*)y, N, H * W, C, s);
}
void transpose_bf16(const void* x, void* y, int rows, int cols, cudaStream_t s) {
  transpose_batched<bf16>((const bf16*)x, (bf16*)y, 1, rows, cols, s);
}

// ---------------------------------------------------------------- conv weight prep
__global__ void k_conv_weight_prep(const float* __restrict__ w, bf16* __restrict__ wb,
                                   bf16* __restrict__ wt, int K, int R, int S, int C) {
  long long total = (long long)K * R * S * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int ss = (int)(t % S);
    t /= S;
    int r = (int)(t % R);
    int k = (int)(t / R);
    bf16 v = __float2bfloat16(w[i]);
    if (wb) wb[i] = v;
    if (wt) wt[(((long long)c * R + (R - 1 - r)) * S + (S - 1 - ss)) * K + k] = v;
  }
}
// dgrad weights only: wt[c][R-1-r][S-1-s][k] = bf16(w[k][r][s][c]) as a tiled shared-memory transpose of
// the [K][C] slice of every filter tap -- reads and writes both coalesced (the element-wise kernel above
// writes 2-byte elements K apart: 27 us for the 512x3x3x512 WRN weights that this one moves in ~4 us).
__global__ void __launch_bounds__(256)
    k_conv_weight_flip(const float* __restrict__ w, bf16* __restrict__ wt, int K, int R, int S, int C) {
  __shared__ float tile[32][33];
  const int rs = blockIdx.z;
  const int r = rs / S, ss = rs - r * S;
  const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  const long long rsc = (long long)R * S * C, rsk = (long long)R * S * K;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int k = k0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (k < K && c < C) ? w[(long long)k * rsc + (long long)rs * C + c] : 0.f;
  }
  __syncthreads();
  const long long tap_out = ((long long)(R - 1 - r) * S + (S - 1 - ss)) * K;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, k = k0 + tx;
    if (c < C && k < K) wt[(long long)c * rsk + tap_out + k] = __float2bfloat16(tile[tx][ty + i]);
  }
}
// Every convolution's flipped dgrad weights in ONE launch (the per-layer launches were 15 x ~6 us of latency-bound
// kernels on the backward critical path of the WRN step): block -> (conv, c tile, k tile, tap) through the descriptors'
// running block offsets.
__global__ void __launch_bounds__(256) k_conv_weight_flip_batch(const ConvFlipDesc* __restrict__ descs, int n) {
  __shared__ float tile[32][33];
  int d = 0;
  while (d + 1 < n && (int)blockIdx.x >= descs[d + 1].block_begin) ++d;
  const ConvFlipDesc D = descs[d];
  int b = blockIdx.x - D.block_begin;
  const int ctiles = (D.C + 31) / 32, ktiles = (D.K + 31) / 32;
  const int bx = b % ctiles;
  b /= ctiles;
  const int by = b % ktiles;
  const int rs = b / ktiles;
  const int R = D.R, S = D.S, K = D.K, C = D.C;
  const int r = rs / S, ss = rs - r * S;
  const int c0 = bx * 32, k0 = by * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long rsc = (long long)R * S * C, rsk = (long long)R * S * K;
  bf16* wt = reinterpret_cast<bf16*>(D.wt);
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int k = k0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (k < K && c < C) ? D.w[(long long)k * rsc + (long long)rs * C + c] : 0.f;
  }
  __syncthreads();
  const long long tap_out = ((long long)(R - 1 - r) * S + (S - 1 - ss)) * K;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, k = k0 + tx;
    if (c < C && k < K) wt[(long long)c * rsk + tap_out + k] = __float2bfloat16(tile[tx][ty + i]);
  }
}
void conv_weight_flip_batch(const ConvFlipDesc* descs_dev, int n, int total_blocks, cudaStream_t s) {
  if (n <= 0 || total_blocks <= 0) return;
  k_conv_weight_flip_batch<<<total_blocks, 256, 0, s>>>(descs_dev, n);
  TNN_CHECK_LAUNCH();
}

void conv_weight_prep(const float* w, void* w_bf16, void* wt_bf16, int K, int R, int S, int C,
                      cudaStream_t s) {
  if (w_bf16 == nullptr && wt_bf16 != nullptr) {
    dim3 grid((C + 31) / 32, (K + 31) / 32, R * S);
    k_conv_weight_flip<<<grid, 256, 0, s>>>(w, (bf16*)wt_bf16, K, R, S, C);
    TNN_CHECK_LAUNCH();
    return;
  }
  long long total = (long long)K * R * S * C;
  k_conv_weight_prep<<<grid_for(total, 256), 256, 0, s>>>(w, (bf16*)w_bf16, (bf16*)wt_bf16, K, R, S, C);
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- RNG fills
__global__ void k_philox_uniform(float* x, long long n, float lo, float hi, uint64_t seed) {
  long long nv = (n + 3) / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    uint4 r = philox4(seed, (uint64_t)i);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    for (int j = 0; j < 4 && i * 4 + j < n; ++j) x[i * 4 + j] = lo + (hi - lo) * u32_to_unit(rr[j]);
  }
}
__global__ void k_philox_normal(float* x, long long n, float mean, float std, uint64_t seed) {
  long long nv = (n + 3) / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    uint4 r = philox4(seed, (uint64_t)i);
    float u0 = u32_to_unit(r.x), u1 = u32_to_unit(r.y), u2 = u32_to_unit(r.z), u3 = u32_to_unit(r.w);
    float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
    float z[4] = {r0 * cospif(2.f * u1), r0 * sinpif(2.f * u1), r1 * cospif(2.f * u3), r1 * sinpif(2.f * u3)};
    for (int j = 0; j < 4 && i * 4 + j < n; ++j) x[i * 4 + j] = mean + std * z[j];
  }
}
void philox_uniform_f32(float* x, long long n, float lo, float hi, uint64_t seed, cudaStream_t s) {
  if (n <= 0) return;
  k_philox_uniform<<<grid_for((n + 3) / 4, 256), 256, 0, s>>>(x, n, lo, hi, seed);
  TNN_CHECK_LAUNCH();
}
void philox_normal_f32(float* x, long long n, float mean, float std, uint64_t seed, cudaStream_t s) {
  if (n <= 0) return;
  k_philox_normal<<<grid_for((n + 3) / 4, 256), 256, 0, s>>>(x, n, mean, std, seed);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
