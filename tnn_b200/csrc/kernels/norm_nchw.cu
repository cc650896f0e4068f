// BatchNorm over NCHW tensors, native layout (no NHWC round trip): the legacy layer family of the reference keeps
// activations as [N, C, H, W] (src/nn/layers_impl/legacy_batchnorm_layer.cpp, cuda/batchnorm_nchw_ops.cu:119-550: one
// block per channel, three passes).  Here a channel's data are N contiguous planes of H*W elements, so
//   * statistics / backward reductions: grid (C, SPLIT) -- every block owns one channel and a slice of its N planes,
//     16-byte loads along the plane, fp32 per-thread partials, one fp64 atomic pair per block (SPLIT sized so that
//     C * SPLIT fills the 148 SMs a few times over even for C = 16),
//   * apply / backward apply: one block per (n, c) plane slice, the channel's scale / shift (or the three backward
//     coefficients) are block-uniform scalars, 16-byte loads and stores.
// bf16 or fp32 activations, fp32 parameters and statistics.  Training forward = 3 launches (stats, finalize, apply),
// backward = 3 (reduce, coefficients, apply); the finalize / coefficient kernels are C threads.
#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {
namespace {

template <typename T>
struct NV;
template <>
struct NV<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static void store(float* p, const float (&f)[4]) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
  __device__ static float get(const float* p) { return *p; }
  __device__ static void put(float* p, float v) { *p = v; }
};
template <>
struct NV<bf16> {
  static constexpr int N = 8;
  __device__ static void load(const bf16* p, float (&f)[8]) { bf8_unpack(*reinterpret_cast<const uint4*>(p), f); }
  __device__ static void store(bf16* p, const float (&f)[8]) { *reinterpret_cast<uint4*>(p) = bf8_pack(f); }
  __device__ static float get(const bf16* p) { return __bfloat162float(*p); }
  __device__ static void put(bf16* p, float v) { *p = __float2bfloat16(v); }
};

__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  // red: >= 64 floats of shared memory
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    red[w] = a;
    red[32 + w] = b;
  }
  __syncthreads();
  if (w == 0) {
    float x = lane < nw ? red[lane] : 0.f, y = lane < nw ? red[32 + lane] : 0.f;
    a = warp_sum(x);
    b = warp_sum(y);
  }
}

// ws[c] += (sum x, sum x^2) over this block's slice of channel c.  Planes n = blockIdx.y, +gridDim.y, ...
template <typename T>
__global__ void __launch_bounds__(256) k_bn_nchw_stats(const T* __restrict__ x, int N, int C, int HW, double* __restrict__ ws) {
  __shared__ float red[64];
  constexpr int V = NV<T>::N;
  const int c = blockIdx.x;
  const bool vec = (HW % V) == 0;
  float s = 0.f, q = 0.f;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const T* p = x + ((long long)n * C + c) * HW;
    if (vec) {
      for (int i = threadIdx.x * V; i < HW; i += blockDim.x * V) {
        float f[V];
        NV<T>::load(p + i, f);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          s += f[j];
          q = fmaf(f[j], f[j], q);
        }
      }
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float v = NV<T>::get(p + i);
        s += v;
        q = fmaf(v, v, q);
      }
    }
  }
  block_sum2(s, q, red);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 2 * c, (double)s);
    atomicAdd(ws + 2 * c + 1, (double)q);
  }
}

// mean / invstd / (scale, shift) per channel, running statistics (unbiased variance), workspace reset for the next call
__global__ void k_bn_nchw_finalize(double* __restrict__ ws, long long M, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ scale_shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = ws[2 * c] / (double)M;
  double var = ws[2 * c + 1] / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  ws[2 * c] = 0.0;
  ws[2 * c + 1] = 0.0;
  const float invstd = rsqrtf((float)var + eps);
  save_mean[c] = (float)mean;
  save_invstd[c] = invstd;
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_shift[2 * c] = g * invstd;
  scale_shift[2 * c + 1] = b - (float)mean * g * invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * ((double)M / (double)(M > 1 ? M - 1 : 1)));
}
__global__ void k_bn_nchw_eval_ss(int C, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ running_mean,
                                  const float* __restrict__ running_var, float eps, float* __restrict__ scale_shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(running_var[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_shift[2 * c] = g * invstd;
  scale_shift[2 * c + 1] = b - running_mean[c] * g * invstd;
}

// y = x * scale[c] + shift[c] (ReLU optional).  grid (N * C, chunks of the plane)
template <typename T>
__global__ void __launch_bounds__(256)
    k_bn_nchw_apply(const T* __restrict__ x, T* __restrict__ y, int C, int HW, const float* __restrict__ scale_shift, int relu) {
  constexpr int V = NV<T>::N;
  const long long plane = blockIdx.x;
  const int c = (int)(plane % C);
  const float sc = scale_shift[2 * c], sh = scale_shift[2 * c + 1];
  const T* p = x + plane * HW;
  T* o = y + plane * HW;
  if ((HW % V) == 0) {
    for (int i = (blockIdx.y * blockDim.x + threadIdx.x) * V; i < HW; i += gridDim.y * blockDim.x * V) {
      float f[V];
      NV<T>::load(p + i, f);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        f[j] = fmaf(f[j], sc, sh);
        if (relu) f[j] = fmaxf(f[j], 0.f);
      }
      NV<T>::store(o + i, f);
    }
  } else {
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < HW; i += gridDim.y * blockDim.x) {
      float v = fmaf(NV<T>::get(p + i), sc, sh);
      if (relu) v = fmaxf(v, 0.f);
      NV<T>::put(o + i, v);
    }
  }
}

// ws[c] += (sum g, sum g * xhat), g = dy (masked by the ReLU of the forward output when relu)
template <typename T>
__global__ void __launch_bounds__(256)
    k_bn_nchw_bwd_reduce(const T* __restrict__ dy, const T* __restrict__ x, int N, int C, int HW, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ scale_shift, int relu, double* __restrict__ ws) {
  __shared__ float red[64];
  constexpr int V = NV<T>::N;
  const int c = blockIdx.x;
  const float mu = mean[c], is = invstd[c], sc = scale_shift[2 * c], sh = scale_shift[2 * c + 1];
  const bool vec = (HW % V) == 0;
  float s = 0.f, q = 0.f;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const long long base = ((long long)n * C + c) * HW;
    if (vec) {
      for (int i = threadIdx.x * V; i < HW; i += blockDim.x * V) {
        float fx[V], fg[V];
        NV<T>::load(x + base + i, fx);
        NV<T>::load(dy + base + i, fg);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float g = (relu && fmaf(fx[j], sc, sh) <= 0.f) ? 0.f : fg[j];
          s += g;
          q = fmaf(g, (fx[j] - mu) * is, q);
        }
      }
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float xv = NV<T>::get(x + base + i);
        const float g = (relu && fmaf(xv, sc, sh) <= 0.f) ? 0.f : NV<T>::get(dy + base + i);
        s += g;
        q = fmaf(g, (xv - mu) * is, q);
      }
    }
  }
  block_sum2(s, q, red);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 2 * c, (double)s);
    atomicAdd(ws + 2 * c + 1, (double)q);
  }
}
// dgamma += sum g*xhat, dbeta += sum g; coef[c] = (a, b, d): dx = a * g + b * x + d
__global__ void k_bn_nchw_bwd_coef(double* __restrict__ ws, long long M, int C, const float* __restrict__ gamma, const float* __restrict__ mean,
                                   const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sg = ws[2 * c], sgx = ws[2 * c + 1];
  ws[2 * c] = 0.0;
  ws[2 * c + 1] = 0.0;
  if (dgamma) dgamma[c] += (float)sgx;
  if (dbeta) dbeta[c] += (float)sg;
  const float g = gamma ? gamma[c] : 1.f, is = invstd[c], mu = mean[c];
  const float mg = (float)(sg / (double)M), mgx = (float)(sgx / (double)M);
  // dx = g*is*(gr - mg - xhat*mgx), xhat = (x - mu)*is  ->  a*gr + b*x + d
  const float a = g * is;
  const float b = -g * is * is * mgx;
  coef[3 * c] = a;
  coef[3 * c + 1] = b;
  coef[3 * c + 2] = -a * mg - b * mu;
}
template <typename T>
__global__ void __launch_bounds__(256)
    k_bn_nchw_bwd_apply(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, int C, int HW, const float* __restrict__ coef,
                        const float* __restrict__ scale_shift, int relu) {
  constexpr int V = NV<T>::N;
  const long long plane = blockIdx.x;
  const int c = (int)(plane % C);
  const float a = coef[3 * c], b = coef[3 * c + 1], d = coef[3 * c + 2], sc = scale_shift[2 * c], sh = scale_shift[2 * c + 1];
  const long long base = plane * HW;
  if ((HW % V) == 0) {
    for (int i = (blockIdx.y * blockDim.x + threadIdx.x) * V; i < HW; i += gridDim.y * blockDim.x * V) {
      float fx[V], fg[V];
      NV<T>::load(x + base + i, fx);
      NV<T>::load(dy + base + i, fg);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float g = (relu && fmaf(fx[j], sc, sh) <= 0.f) ? 0.f : fg[j];
        fg[j] = fmaf(a, g, fmaf(b, fx[j], d));
      }
      NV<T>::store(dx + base + i, fg);
    }
  } else {
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < HW; i += gridDim.y * blockDim.x) {
      const float xv = NV<T>::get(x + base + i);
      const float g = (relu && fmaf(xv, sc, sh) <= 0.f) ? 0.f : NV<T>::get(dy + base + i);
      NV<T>::put(dx + base + i, fmaf(a, g, fmaf(b, xv, d)));
    }
  }
}

int split_for(int N, int C) {
  // enough blocks for ~4 waves of 148 SMs x 2, never more than one block per plane
  int s = (4 * 2 * kSMs + C - 1) / C;
  if (s > N) s = N;
  return s < 1 ? 1 : s;
}
int chunks_for(long long planes, int HW, int vec) {
  // planes * chunks >= ~8 blocks per SM; a chunk is at least one pass of 256 threads
  int per = (HW + 256 * vec - 1) / (256 * vec);
  long long want = (8LL * kSMs + planes - 1) / planes;
  if (want > per) want = per;
  if (want > 65535) want = 65535;
  return want < 1 ? 1 : (int)want;
}

}  // namespace

size_t bn_nchw_ws_bytes(int C) { return sizeof(double) * 2 * (size_t)C; }

#define NCHW_T(is_f32, ...)                         \
  do {                                              \
    if (is_f32) { using T = float; __VA_ARGS__; }   \
    else { using T = bf16; __VA_ARGS__; }           \
  } while (0)

void bn_nchw_train_stats(const void* x, int is_f32, int N, int C, int HW, const float* gamma, const float* beta, float eps, float momentum,
                         float* running_mean, float* running_var, float* save_mean, float* save_invstd, float* scale_shift, void* ws,
                         cudaStream_t s) {
  if (N <= 0 || C <= 0 || HW <= 0) return;
  dim3 grid(C, split_for(N, C));
  NCHW_T(is_f32, k_bn_nchw_stats<T><<<grid, 256, 0, s>>>((const T*)x, N, C, HW, (double*)ws));
  TNN_CHECK_LAUNCH();
  k_bn_nchw_finalize<<<(C + 127) / 128, 128, 0, s>>>((double*)ws, (long long)N * HW, C, gamma, beta, eps, momentum, running_mean, running_var,
                                                     save_mean, save_invstd, scale_shift);
  TNN_CHECK_LAUNCH();
}
void bn_nchw_eval_scale_shift(int C, const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                              float* scale_shift, cudaStream_t s) {
  k_bn_nchw_eval_ss<<<(C + 127) / 128, 128, 0, s>>>(C, gamma, beta, running_mean, running_var, eps, scale_shift);
  TNN_CHECK_LAUNCH();
}
void bn_nchw_apply(const void* x, void* y, int is_f32, int N, int C, int HW, const float* scale_shift, int relu, cudaStream_t s) {
  if (N <= 0 || C <= 0 || HW <= 0) return;
  const long long planes = (long long)N * C;
  dim3 grid((unsigned)planes, chunks_for(planes, HW, is_f32 ? 4 : 8));
  NCHW_T(is_f32, k_bn_nchw_apply<T><<<grid, 256, 0, s>>>((const T*)x, (T*)y, C, HW, scale_shift, relu));
  TNN_CHECK_LAUNCH();
}
void bn_nchw_backward(const void* dy, const void* x, void* dx, int is_f32, int N, int C, int HW, const float* gamma, const float* save_mean,
                      const float* save_invstd, const float* scale_shift, int relu, float* dgamma, float* dbeta, float* coef, void* ws,
                      cudaStream_t s) {
  if (N <= 0 || C <= 0 || HW <= 0) return;
  dim3 grid(C, split_for(N, C));
  NCHW_T(is_f32, k_bn_nchw_bwd_reduce<T><<<grid, 256, 0, s>>>((const T*)dy, (const T*)x, N, C, HW, save_mean, save_invstd, scale_shift, relu,
                                                             (double*)ws));
  TNN_CHECK_LAUNCH();
  k_bn_nchw_bwd_coef<<<(C + 127) / 128, 128, 0, s>>>((double*)ws, (long long)N * HW, C, gamma, save_mean, save_invstd, dgamma, dbeta, coef);
  TNN_CHECK_LAUNCH();
  const long long planes = (long long)N * C;
  dim3 grid2((unsigned)planes, chunks_for(planes, HW, is_f32 ? 4 : 8));
  NCHW_T(is_f32, k_bn_nchw_bwd_apply<T><<<grid2, 256, 0, s>>>((const T*)dy, (const T*)x, (T*)dx, C, HW, coef, scale_shift, relu));
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
