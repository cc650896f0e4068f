// Shared device helpers for the SIMT kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

namespace tnn {

// number of kernels launched by this library in the current process (bench.py reports it)
extern std::atomic<unsigned long long> g_launch_count;   // atomic: the C++ data-parallel host launches from one thread per GPU

#define TNN_CHECK_LAUNCH()                                                                    \
  do {                                                                                        \
    ++::tnn::g_launch_count;                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess) {                                                                  \
      fprintf(stderr, "[tnn_b200] launch error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, \
              __LINE__);                                                                      \
      abort();                                                                                \
    }                                                                                         \
  } while (0)

// cudaFuncSetAttribute (dynamic shared-memory limit, carve-out) applies to ONE device.  A launcher that caches
// "already configured" therefore caches it per device: with a per-process flag the second GPU of a single-process
// multi-GPU run (cpp DataParallel, the in-process pipeline) launches with the default 48 KB limit and fails.
//   static PerDeviceOnce once;  if (once.begin()) { cudaFuncSetAttribute(...); once.done(); }
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  std::mutex mu;
  static int device() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 63;
  }
  // true exactly once per device; the caller holds the lock until done()
  bool begin() {
    const unsigned long long bit = 1ull << device();
    if (mask.load(std::memory_order_acquire) & bit) return false;
    mu.lock();
    if (mask.load(std::memory_order_relaxed) & bit) {
      mu.unlock();
      return false;
    }
    return true;
  }
  void done() {
    mask.fetch_or(1ull << device(), std::memory_order_release);
    mu.unlock();
  }
};

static constexpr int kSMs = 148;
typedef __nv_bfloat16 bf16;

__host__ inline int grid_for(long long n, int block, int max_blocks = kSMs * 16) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

// 8 bf16 <-> 8 floats through one 16-byte register quad
struct alignas(16) BF8 {
  uint4 q;
};
__device__ __forceinline__ void bf8_unpack(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 bf8_pack(const float (&f)[8]) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return q;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

// Philox4x32-R counter RNG: (seed, counter) -> 4 x uint32.  R = 10 is the standard strength;
// R = 7 is the shortest variant that still passes BigCrush and is what the dropout masks use.
template <int ROUNDS = 10>
__device__ __forceinline__ uint4 philox4(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x2B7E1516u, c3 = 0x3C6EF372u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) {  // (0,1)
  return ((x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
// keep-mask for 8 consecutive elements starting at index 8*vec_idx: ONE Philox block per vector,
// 16 random bits per element (drop probability resolved to 1/65536).
__device__ __forceinline__ void dropout_keep8(uint64_t seed, long long vec_idx, float p, bool (&keep)[8]) {
  const uint4 a = philox4<7>(seed, (uint64_t)vec_idx);
  const uint32_t thr = (uint32_t)(p * 65536.0f + 0.5f);
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    keep[2 * i] = (w[i] & 0xFFFFu) >= thr;
    keep[2 * i + 1] = (w[i] >> 16) >= thr;
  }
}

}  // namespace tnn
