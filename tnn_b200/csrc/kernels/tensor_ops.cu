// Generic tensor op family: the `ops::` surface of the reference (include/ops/ops.hpp:17-835,
// src/ops/cuda/kernels.cu:30-565, src/tensor/cuda/tensor_ops.cu:136-300) re-done for sm_100a.
//   * one templated, 128-bit vectorised, grid-stride elementwise kernel family for fp32 and bf16
//     (binary, binary-with-scalar, unary, ternary fused multiply-add forms),
//   * reductions (sum, dot, sum of squared differences) that FINISH ON THE DEVICE: warp shuffles,
//     one fp32 atomic per block into a double-free accumulator -- no cudaMalloc, no host
//     summation (the reference copies per-block partials to the host, kernels.cu:278-315),
//   * pad / unpad / crop over NHWC, channel softmax, dtype casts, byte swap.
#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
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
  __device__ static void load(const bf16* p, float (&f)[8]) { bf8_unpack(*reinterpret_cast<const uint4*>(p), f); }
  __device__ static void store(bf16* p, const float (&f)[8]) { *reinterpret_cast<uint4*>(p) = bf8_pack(f); }
  __device__ static float get(const bf16* p) { return __bfloat162float(*p); }
  __device__ static void put(bf16* p, float v) { *p = __float2bfloat16(v); }
};

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
    float fa[N], fb[N], fo[N];
    Vec<T>::load(a + i * N, fa);
    if (b) Vec<T>::load(b + i * N, fb);
#pragma unroll
    for (int j = 0; j < N; ++j) fo[j] = apply_binary(op, fa[j], b ? fb[j] : scalar);
    Vec<T>::store(out + i * N, fo);
  }
  for (long long i = nv * N + tid; i < n; i += stride)
    Vec<T>::put(out + i, apply_binary(op, Vec<T>::get(a + i), b ? Vec<T>::get(b + i) : scalar));
}
template <typename T>
__global__ void __launch_bounds__(256)
    k_unary(const T* __restrict__ a, T* __restrict__ out, long long n, int op, float lo, float hi, int vec_ok) {
  constexpr int N = Vec<T>::N;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const long long nv = vec_ok ? n / N : 0;
  for (long long i = tid; i < nv; i += stride) {
    float fa[N], fo[N];
    Vec<T>::load(a + i * N, fa);
#pragma unroll
    for (int j = 0; j < N; ++j) fo[j] = apply_unary(op, fa[j], lo, hi);
    Vec<T>::store(out + i * N, fo);
  }
  for (long long i = nv * N + tid; i < n; i += stride) Vec<T>::put(out + i, apply_unary(op, Vec<T>::get(a + i), lo, hi));
}
// fmadd: a*b + c ; fmsub: a*b - c ; fnmadd: -(a*b) + c
template <typename T>
__global__ void __launch_bounds__(256)
    k_ternary(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, T* __restrict__ out, long long n,
              int op) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n; i += stride) {
    const float x = Vec<T>::get(a + i), y = Vec<T>::get(b + i), z = Vec<T>::get(c + i);
    Vec<T>::put(out + i, op == 0 ? fmaf(x, y, z) : (op == 1 ? fmaf(x, y, -z) : fmaf(-x, y, z)));
  }
}

static int vec_ok_ptrs(const void* a, const void* b, const void* c) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

void tensor_binary(const void* a, const void* b, float scalar, void* out, long long n, int op, int is_f32, cudaStream_t s) {
  if (n <= 0) return;
  const int v = vec_ok_ptrs(a, b, out);
  if (is_f32)
    k_binary<float><<<grid_for(n / 4 + 1, 256), 256, 0, s>>>((const float*)a, (const float*)b, scalar, (float*)out, n, op, v);
  else
    k_binary<bf16><<<grid_for(n / 8 + 1, 256), 256, 0, s>>>((const bf16*)a, (const bf16*)b, scalar, (bf16*)out, n, op, v);
  TNN_CHECK_LAUNCH();
}
void tensor_unary(const void* a, void* out, long long n, int op, float lo, float hi, int is_f32, cudaStream_t s) {
  if (n <= 0) return;
  const int v = vec_ok_ptrs(a, out, nullptr);
  if (is_f32)
    k_unary<float><<<grid_for(n / 4 + 1, 256), 256, 0, s>>>((const float*)a, (float*)out, n, op, lo, hi, v);
  else
    k_unary<bf16><<<grid_for(n / 8 + 1, 256), 256, 0, s>>>((const bf16*)a, (bf16*)out, n, op, lo, hi, v);
  TNN_CHECK_LAUNCH();
}
void tensor_ternary(const void* a, const void* b, const void* c, void* out, long long n, int op, int is_f32, cudaStream_t s) {
  if (n <= 0) return;
  if (is_f32)
    k_ternary<float><<<grid_for(n, 256), 256, 0, s>>>((const float*)a, (const float*)b, (const float*)c, (float*)out, n, op);
  else
    k_ternary<bf16><<<grid_for(n, 256), 256, 0, s>>>((const bf16*)a, (const bf16*)b, (const bf16*)c, (bf16*)out, n, op);
  TNN_CHECK_LAUNCH();
}

// ---- reductions: mode 0 sum(a), 1 dot(a,b), 2 sum((a - mu)^2); result accumulated into *out (fp32)
template <typename T>
__global__ void __launch_bounds__(256)
    k_reduce(const T* __restrict__ a, const T* __restrict__ b, float mu, long long n, int mode, float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = Vec<T>::get(a + i);
    if (mode == 0) acc += x;
    else if (mode == 1) acc += x * Vec<T>::get(b + i);
    else acc += (x - mu) * (x - mu);
  }
  acc = warp_sum(acc);
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
void tensor_reduce(const void* a, const void* b, float mu, long long n, int mode, int is_f32, float* out, cudaStream_t s) {
  cudaMemsetAsync(out, 0, sizeof(float), s);
  if (n <= 0) return;
  if (is_f32)
    k_reduce<float><<<grid_for(n, 256, kSMs * 4), 256, 0, s>>>((const float*)a, (const float*)b, mu, n, mode, out);
  else
    k_reduce<bf16><<<grid_for(n, 256, kSMs * 4), 256, 0, s>>>((const bf16*)a, (const bf16*)b, mu, n, mode, out);
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
    float v = fill;
    if (sh >= 0 && sh < SH && sw >= 0 && sw < SW) v = Vec<T>::get(src + (((long long)n * SH + sh) * SW + sw) * C + c);
    Vec<T>::put(dst + i, v);
  }
}
void tensor_window_copy(const void* src, void* dst, int N, int SH, int SW, int DH, int DW, int C, int off_h, int off_w,
                        float fill, int is_f32, cudaStream_t s) {
  const long long total = (long long)N * DH * DW * C;
  if (total <= 0) return;
  if (is_f32)
    k_window_copy<float><<<grid_for(total, 256), 256, 0, s>>>((const float*)src, (float*)dst, N, SH, SW, DH, DW, C, off_h, off_w, fill);
  else
    k_window_copy<bf16><<<grid_for(total, 256), 256, 0, s>>>((const bf16*)src, (bf16*)dst, N, SH, SW, DH, DW, C, off_h, off_w, fill);
  TNN_CHECK_LAUNCH();
}

// ---- row softmax (last dim): one warp per row, fwd and bwd
template <typename T>
__global__ void __launch_bounds__(256) k_softmax_rows(const T* __restrict__ x, T* __restrict__ y, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    const T* xr = x + r * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, Vec<T>::get(xr + c));
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += __expf(Vec<T>::get(xr + c) - mx);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int c = lane; c < C; c += 32) Vec<T>::put(y + r * C + c, __expf(Vec<T>::get(xr + c) - mx) * inv);
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
    k_softmax_rows_bwd(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) dot += Vec<T>::get(y + r * C + c) * Vec<T>::get(dy + r * C + c);
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) {
      const float yy = Vec<T>::get(y + r * C + c);
      Vec<T>::put(dx + r * C + c, yy * (Vec<T>::get(dy + r * C + c) - dot));
    }
  }
}
void softmax_rows(const void* x, void* y, long long rows, int C, int is_f32, cudaStream_t s) {
  if (rows <= 0) return;
  const int blocks = grid_for((rows + 7) / 8, 1, kSMs * 8);
  if (is_f32) k_softmax_rows<float><<<blocks, 256, 0, s>>>((const float*)x, (float*)y, rows, C);
  else k_softmax_rows<bf16><<<blocks, 256, 0, s>>>((const bf16*)x, (bf16*)y, rows, C);
  TNN_CHECK_LAUNCH();
}
void softmax_rows_bwd(const void* y, const void* dy, void* dx, long long rows, int C, int is_f32, cudaStream_t s) {
  if (rows <= 0) return;
  const int blocks = grid_for((rows + 7) / 8, 1, kSMs * 8);
  if (is_f32) k_softmax_rows_bwd<float><<<blocks, 256, 0, s>>>((const float*)y, (const float*)dy, (float*)dx, rows, C);
  else k_softmax_rows_bwd<bf16><<<blocks, 256, 0, s>>>((const bf16*)y, (const bf16*)dy, (bf16*)dx, rows, C);
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
