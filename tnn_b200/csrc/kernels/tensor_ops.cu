// Generic tensor op family: the `ops::` surface of the reference (include/ops/ops.hpp:17-835,
// src/ops/cuda/kernels.cu:30-565, src/tensor/cuda/tensor_ops.cu:136-300) re-done for sm_100a.
//   * one templated, 128-bit vectorised, grid-stride elementwise kernel family for bf16, fp32, fp16 and fp64 (the `dt`
//     argument of every entry point is an XtDType code; 0 = bf16 and 1 = fp32 keep the old `is_f32` meaning; fp64 does
//     its arithmetic in double, the others in float)
//     (binary, binary-with-scalar, unary, ternary fused multiply-add forms),
//   * reductions (sum, dot, sum of squared differences) that FINISH ON THE DEVICE: warp shuffles,
//     one fp32 atomic per block into a double-free accumulator -- no cudaMalloc, no host
//     summation (the reference copies per-block partials to the host, kernels.cu:278-315),
//   * pad / unpad / crop over NHWC, channel softmax, dtype casts, byte swap.
#include <cuda_fp16.h>

#include "common.cuh"
#include "gemm_xt.cuh"
#include "simt_ops.cuh"

namespace tnn {

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
  using A = float;
  using V = float4;
  __device__ static void load(const float* p, float (&f)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
  __device__ static float get(const float* p) { return *p; }
  __device__ static void put(float* p, float v) { *p = v; }
};
template <>
struct Vec<bf16> {
  static constexpr int N = 8;
  using A = float;
  __device__ static void load(const bf16* p, float (&f)[8]) { bf8_unpack(*reinterpret_cast<const uint4*>(p), f); }
  __device__ static void store(bf16* p, const float (&f)[8]) { *reinterpret_cast<uint4*>(p) = bf8_pack(f); }
  __device__ static float get(const bf16* p) { return __bfloat162float(*p); }
  __device__ static void put(bf16* p, float v) { *p = __float2bfloat16(v); }
};

template <>
struct Vec<__half> {
  static constexpr int N = 8;
  using A = float;
  __device__ static void load(const __half* p, float (&f)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ static void store(__half* p, const float (&f)[8]) {
    uint4 q;
    __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = q;
  }
  __device__ static float get(const __half* p) { return __half2float(*p); }
  __device__ static void put(__half* p, float v) { *p = __float2half_rn(v); }
};
template <>
struct Vec<double> {
  static constexpr int N = 2;
  using A = double;
  __device__ static void load(const double* p, double (&f)[2]) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    f[0] = v.x; f[1] = v.y;
  }
  __device__ static void store(double* p, const double (&f)[2]) { *reinterpret_cast<double2*>(p) = make_double2(f[0], f[1]); }
  __device__ static double get(const double* p) { return *p; }
  __device__ static void put(double* p, double v) { *p = v; }
};

// run BODY with T bound to the element type of dtype code `dt`
#define TNN_DT_SWITCH(dt, ...)                         \
  switch (dt) {                                        \
    case XT_F32: { using T = float; __VA_ARGS__; } break;   \
    case XT_F16: { using T = __half; __VA_ARGS__; } break;  \
    case XT_F64: { using T = double; __VA_ARGS__; } break;  \
    default: { using T = bf16; __VA_ARGS__; } break;        \
  }

__device__ __forceinline__ double apply_binary(int op, double a, double b) {
  switch (op) {
    case TOP_ADD: return a + b;
    case TOP_SUB: return a - b;
    case TOP_MUL: return a * b;
    case TOP_DIV: return a / b;
    case TOP_MIN: return fmin(a, b);
    case TOP_MAX: return fmax(a, b);
    case TOP_EQ: return a == b ? 1.0 : 0.0;
    case TOP_GT: return a > b ? 1.0 : 0.0;
    default: return a;
  }
}
__device__ __forceinline__ double apply_unary(int op, double a, double lo, double hi) {
  switch (op) {
    case UOP_SQRT: return sqrt(a);
    case UOP_RSQRT: return 1.0 / sqrt(a);
    case UOP_RCP: return 1.0 / a;
    case UOP_ABS: return fabs(a);
    case UOP_CLAMP: return fmin(fmax(a, lo), hi);
    case UOP_NEG: return -a;
    case UOP_EXP: return exp(a);
    case UOP_LOG: return log(a);
    case UOP_SQUARE: return a * a;
    default: return a;
  }
}
__device__ __forceinline__ float apply_binary(int op, float a, float b) {
  switch (op) {
    case TOP_ADD: return a + b;
    case TOP_SUB: return a - b;
    case TOP_MUL: return a * b;
    case TOP_DIV: return a / b;
    case TOP_MIN: return fminf(a, b);
    case TOP_MAX: return fmaxf(a, b);
    case TOP_EQ: return a == b ? 1.f : 0.f;
    case TOP_GT: return a > b ? 1.f : 0.f;
    default: return a;
  }
}
__device__ __forceinline__ float apply_unary(int op, float a, float lo, float hi) {
  switch (op) {
    case UOP_SQRT: return sqrtf(a);
    case UOP_RSQRT: return rsqrtf(a);
    case UOP_RCP: return 1.f / a;
    case UOP_ABS: return fabsf(a);
    case UOP_CLAMP: return fminf(fmaxf(a, lo), hi);
    case UOP_NEG: return -a;
    case UOP_EXP: return __expf(a);
    case UOP_LOG: return __logf(a);
    case UOP_SQUARE: return a * a;
    default: return a;
  }
}

// out = a (op) b, or a (op) scalar when b == nullptr
template <typename T>
__global__ void __launch_bounds__(256)
    k_binary(const T* __restrict__ a, const T* __restrict__ b, float scalar, T* __restrict__ out, long long n, int op,
             int vec_ok) {
  constexpr int N = Vec<T>::N;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const long long nv = vec_ok ? n / N : 0;
  for (long long i = tid; i < nv; i += stride) {
    using A = typename Vec<T>::A;
    A fa[N], fb[N], fo[N];
    Vec<T>::load(a + i * N, fa);
    if (b) Vec<T>::load(b + i * N, fb);
#pragma unroll
    for (int j = 0; j < N; ++j) fo[j] = apply_binary(op, fa[j], b ? fb[j] : (A)scalar);
    Vec<T>::store(out + i * N, fo);
  }
  for (long long i = nv * N + tid; i < n; i += stride)
    Vec<T>::put(out + i, apply_binary(op, Vec<T>::get(a + i), b ? Vec<T>::get(b + i) : (typename Vec<T>::A)scalar));
}
template <typename T>
__global__ void __launch_bounds__(256)
    k_unary(const T* __restrict__ a, T* __restrict__ out, long long n, int op, float lo, float hi, int vec_ok) {
  constexpr int N = Vec<T>::N;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const long long nv = vec_ok ? n / N : 0;
  for (long long i = tid; i < nv; i += stride) {
    using A = typename Vec<T>::A;
    A fa[N], fo[N];
    Vec<T>::load(a + i * N, fa);
#pragma unroll
    for (int j = 0; j < N; ++j) fo[j] = apply_unary(op, fa[j], (A)lo, (A)hi);
    Vec<T>::store(out + i * N, fo);
  }
  for (long long i = nv * N + tid; i < n; i += stride)
    Vec<T>::put(out + i, apply_unary(op, Vec<T>::get(a + i), (typename Vec<T>::A)lo, (typename Vec<T>::A)hi));
}
// fmadd: a*b + c ; fmsub: a*b - c ; fnmadd: -(a*b) + c
template <typename T>
__global__ void __launch_bounds__(256)
    k_ternary(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, T* __restrict__ out, long long n,
              int op) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n; i += stride) {
    const typename Vec<T>::A x = Vec<T>::get(a + i), y = Vec<T>::get(b + i), z = Vec<T>::get(c + i);
    Vec<T>::put(out + i, op == 0 ? fma(x, y, z) : (op == 1 ? fma(x, y, -z) : fma(-x, y, z)));
  }
}

static int vec_ok_ptrs(const void* a, const void* b, const void* c) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

void tensor_binary(const void* a, const void* b, float scalar, void* out, long long n, int op, int dt, cudaStream_t s) {
  if (n <= 0) return;
  const int v = vec_ok_ptrs(a, b, out);
  TNN_DT_SWITCH(dt, k_binary<T><<<grid_for(n / Vec<T>::N + 1, 256), 256, 0, s>>>((const T*)a, (const T*)b, scalar, (T*)out, n, op, v));
  TNN_CHECK_LAUNCH();
}
void tensor_unary(const void* a, void* out, long long n, int op, float lo, float hi, int dt, cudaStream_t s) {
  if (n <= 0) return;
  const int v = vec_ok_ptrs(a, out, nullptr);
  TNN_DT_SWITCH(dt, k_unary<T><<<grid_for(n / Vec<T>::N + 1, 256), 256, 0, s>>>((const T*)a, (T*)out, n, op, lo, hi, v));
  TNN_CHECK_LAUNCH();
}
void tensor_ternary(const void* a, const void* b, const void* c, void* out, long long n, int op, int dt, cudaStream_t s) {
  if (n <= 0) return;
  TNN_DT_SWITCH(dt, k_ternary<T><<<grid_for(n, 256), 256, 0, s>>>((const T*)a, (const T*)b, (const T*)c, (T*)out, n, op));
  TNN_CHECK_LAUNCH();
}

// ---- reductions: mode 0 sum(a), 1 dot(a,b), 2 sum((a - mu)^2); result accumulated into *out (fp32)
template <typename T>
__global__ void __launch_bounds__(256)
    k_reduce(const T* __restrict__ a, const T* __restrict__ b, float mu, long long n, int mode, float* __restrict__ out) {
  using A = typename Vec<T>::A;
  A acc_t = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const A x = Vec<T>::get(a + i);
    if (mode == 0) acc_t += x;
    else if (mode == 1) acc_t += x * Vec<T>::get(b + i);
    else acc_t += (x - (A)mu) * (x - (A)mu);
  }
  float acc = warp_sum((float)acc_t);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = part[threadIdx.x];
    v += __shfl_xor_sync(0xffu, v, 4);
    v += __shfl_xor_sync(0xffu, v, 2);
    v += __shfl_xor_sync(0xffu, v, 1);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}
void tensor_reduce(const void* a, const void* b, float mu, long long n, int mode, int dt, float* out, cudaStream_t s) {
  cudaMemsetAsync(out, 0, sizeof(float), s);
  if (n <= 0) return;
  TNN_DT_SWITCH(dt, k_reduce<T><<<grid_for(n, 256, kSMs * 4), 256, 0, s>>>((const T*)a, (const T*)b, mu, n, mode, out));
  TNN_CHECK_LAUNCH();
}

// ---- pad / unpad / crop over [N, H, W, C]
template <typename T>
__global__ void __launch_bounds__(256)
    k_window_copy(const T* __restrict__ src, T* __restrict__ dst, int N, int SH, int SW, int DH, int DW, int C, int oh,
                  int ow, float fill) {
  // dst[n, h, w, c] = src[n, h + oh, w + ow, c] when inside src, else fill   (oh/ow may be negative = padding)
  const long long total = (long long)N * DH * DW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int w = (int)(r % DW);
    r /= DW;
    const int h = (int)(r % DH);
    const int n = (int)(r / DH);
    const int sh = h + oh, sw = w + ow;
    typename Vec<T>::A v = fill;
    if (sh >= 0 && sh < SH && sw >= 0 && sw < SW) v = Vec<T>::get(src + (((long long)n * SH + sh) * SW + sw) * C + c);
    Vec<T>::put(dst + i, v);
  }
}
void tensor_window_copy(const void* src, void* dst, int N, int SH, int SW, int DH, int DW, int C, int off_h, int off_w,
                        float fill, int dt, cudaStream_t s) {
  const long long total = (long long)N * DH * DW * C;
  if (total <= 0) return;
  TNN_DT_SWITCH(dt, k_window_copy<T><<<grid_for(total, 256), 256, 0, s>>>((const T*)src, (T*)dst, N, SH, SW, DH, DW, C, off_h, off_w, fill));
  TNN_CHECK_LAUNCH();
}

// ---- row softmax (last dim): one warp per row, fwd and bwd
__device__ __forceinline__ float exp_t(float x) { return __expf(x); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__global__ void __launch_bounds__(256) k_softmax_rows(const T* __restrict__ x, T* __restrict__ y, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    using A = typename Vec<T>::A;
    const T* xr = x + r * C;
    A mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmax(mx, (A)Vec<T>::get(xr + c));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    A sum = 0;
    for (int c = lane; c < C; c += 32) sum += exp_t(Vec<T>::get(xr + c) - mx);
    sum = warp_sum(sum);
    const A inv = (A)1 / sum;
    for (int c = lane; c < C; c += 32) Vec<T>::put(y + r * C + c, exp_t(Vec<T>::get(xr + c) - mx) * inv);
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
    k_softmax_rows_bwd(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    using A = typename Vec<T>::A;
    A dot = 0;
    for (int c = lane; c < C; c += 32) dot += Vec<T>::get(y + r * C + c) * Vec<T>::get(dy + r * C + c);
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) {
      const A yy = Vec<T>::get(y + r * C + c);
      Vec<T>::put(dx + r * C + c, yy * (Vec<T>::get(dy + r * C + c) - dot));
    }
  }
}
void softmax_rows(const void* x, void* y, long long rows, int C, int dt, cudaStream_t s) {
  if (rows <= 0) return;
  const int blocks = grid_for((rows + 7) / 8, 1, kSMs * 8);
  TNN_DT_SWITCH(dt, k_softmax_rows<T><<<blocks, 256, 0, s>>>((const T*)x, (T*)y, rows, C));
  TNN_CHECK_LAUNCH();
}
void softmax_rows_bwd(const void* y, const void* dy, void* dx, long long rows, int C, int dt, cudaStream_t s) {
  if (rows <= 0) return;
  const int blocks = grid_for((rows + 7) / 8, 1, kSMs * 8);
  TNN_DT_SWITCH(dt, k_softmax_rows_bwd<T><<<blocks, 256, 0, s>>>((const T*)y, (const T*)dy, (T*)dx, rows, C));
  TNN_CHECK_LAUNCH();
}

// ---- byte swap of 2/4/8-byte elements (endianness conversion of received tensors)
__global__ void __launch_bounds__(256) k_bswap(uint8_t* p, long long n_elems, int elem) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_elems; i += (long long)gridDim.x * blockDim.x) {
    uint8_t* e = p + i * elem;
    for (int a = 0, b = elem - 1; a < b; ++a, --b) {
      const uint8_t t = e[a];
      e[a] = e[b];
      e[b] = t;
    }
  }
}
void tensor_bswap(void* p, long long n_elems, int elem_size, cudaStream_t s) {
  if (n_elems <= 0 || elem_size < 2) return;
  k_bswap<<<grid_for(n_elems, 256), 256, 0, s>>>((uint8_t*)p, n_elems, elem_size);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
