// BatchNorm / LayerNorm / GroupNorm kernels (bf16 activations, fp32 statistics + affine).
//
// BatchNorm works on the [M, C] NHWC view.  Unlike the reference's block-per-channel
// kernels (src/nn/layers_impl/cuda/batchnorm_nchw_ops.cu:119-449, which launch only C
// blocks and under-fill a 148-SM part), rows are spread over ~4 CTAs per SM, each thread
// owns 8 contiguous channels (16-byte loads), CTAs combine through fp32 atomics and the
// LAST CTA to finish folds the statistics into (scale, shift) and the running averages --
// one launch per pass, no host round trip.
#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

// ---------------------------------------------------------------------------------------
// Mapping shared by all four BatchNorm kernels: a thread owns ONE group of 8 consecutive
// channels (so every per-channel coefficient sits in registers) and strides over rows, UNROLL
// rows in flight per iteration (UNROLL x 2-3 independent 16-byte loads per thread keep enough
// bytes in flight to saturate HBM3e).  block = 256 threads = lpr channel-groups x rpb row-lanes.
// ---------------------------------------------------------------------------------------
static constexpr int BN_UNROLL = 4;
static constexpr int BN_STATS_UNROLL = 8;     // the statistics pass holds nothing but its accumulators
// Partial sums go to one of BN_REPL replica rows and the last CTA folds the replicas.  Each atomic
// instruction touches 32 different channels, so per-address contention is one lane per CTA; 8
// replicas keep it at ~74 CTAs per address while the single-CTA fold at the end of every BN
// kernel (a serial tail that was ~half of the small-tensor kernels' time with 64 replicas) stays short.
static constexpr int BN_REPL = 8;

struct BnMap {
  int groups, lpr, rpb, cg, rl;
  __device__ __forceinline__ BnMap(int C) {
    groups = (C + 7) / 8;
    lpr = min(groups, 256);
    rpb = 256 / lpr;
    cg = threadIdx.x % lpr;
    rl = threadIdx.x / lpr;
  }
};

__device__ __forceinline__ void load8(const bf16* x, long long r, int C, int g, float (&f)[8]) {
  if ((C & 7) == 0) {
    bf8_unpack(__ldg(reinterpret_cast<const uint4*>(x + r * C + g * 8)), f);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (g * 8 + j < C) ? __bfloat162float(x[r * C + g * 8 + j]) : 0.f;
  }
}
__device__ __forceinline__ uint4 load8raw(const bf16* x, long long r, int C, int g) {
  return __ldg(reinterpret_cast<const uint4*>(x + r * C + g * 8));
}
__device__ __forceinline__ void store8(bf16* y, long long r, int C, int g, const float (&f)[8]) {
  if ((C & 7) == 0) {
    *reinterpret_cast<uint4*>(y + r * C + g * 8) = bf8_pack(f);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (g * 8 + j < C) y[r * C + g * 8 + j] = __float2bfloat16(f[j]);
  }
}
__device__ __forceinline__ void load_coef8(const float* p, int C, int g, float (&f)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = (p != nullptr && g * 8 + j < C) ? __ldg(p + g * 8 + j) : 0.f;
}

// cross-row-lane reduction of K x 8 per-thread partials -> atomicAdd into ws[k*C + c]
template <int K>
__device__ __forceinline__ void block_channel_reduce(const BnMap& m, int g, int C, float (&acc)[K][8], float* ws,
                                                     float* sm) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[(m.rl * m.lpr + m.cg) * 8 + j] = (m.rl < m.rpb) ? acc[k][j] : 0.f;
    __syncthreads();
    if (m.rl == 0 && g < m.groups) {
      for (int j = 0; j < 8 && g * 8 + j < C; ++j) {
        float t = 0.f;
        for (int q = 0; q < m.rpb; ++q) t += sm[(q * m.lpr + m.cg) * 8 + j];
        atomicAdd(ws + ((size_t)(blockIdx.x % BN_REPL) * K + k) * C + g * 8 + j, t);
      }
    }
  }
}
// sum of replica rows for quantity k, channel c (called by the last CTA), and reset.  All
// BN_REPL loads are issued before any store so they overlap (one L2 round trip, not 64).
template <int K>
__device__ __forceinline__ float fold_replicas(float* ws, int k, int C, int c) {
  float v[BN_REPL];
#pragma unroll
  for (int r = 0; r < BN_REPL; ++r) v[r] = __ldcg(ws + ((size_t)r * K + k) * C + c);
  float t = 0.f;
#pragma unroll
  for (int r = 0; r < BN_REPL; ++r) t += v[r];
#pragma unroll
  for (int r = 0; r < BN_REPL; ++r) ws[((size_t)r * K + k) * C + c] = 0.f;
  return t;
}

// returns true in exactly one CTA: the last one to arrive (all partial sums visible)
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// Grid size: at most 4 CTAs per SM, and chosen so that every row-lane gets a WHOLE number of unrolled
// passes (`unroll` rows each): the rows left over for the one-load-at-a-time tail loop cost a DRAM
// round trip apiece (ncu: 20-27 us for a 17 MB tensor, long-scoreboard + barrier stalls).
static int bn_blocks(long long M, int C, int unroll = BN_UNROLL) {
  int groups = (C + 7) / 8;
  int lpr = groups < 256 ? groups : 256;
  int rpb = 256 / lpr;
  const long long cap = kSMs * 4;
  long long rpl = (M + (long long)rpb * cap - 1) / ((long long)rpb * cap);     // rows per lane at the cap
  rpl = (rpl + unroll - 1) / unroll * unroll;
  long long b = (M + (long long)rpb * rpl - 1) / ((long long)rpb * rpl);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int bn_ws_floats(int C) { return BN_REPL * 2 * C + 4; }

// ------------------------------------------------------------------ BN forward statistics
__global__ void __launch_bounds__(256)
    k_bn_stats(const bf16* __restrict__ x, long long M, int C, const float* __restrict__ gamma,
               const float* __restrict__ beta, float eps, float momentum, float* running_mean,
               float* running_var, float* save_mean, float* save_invstd, float* scale_shift, float* ws,
               float4* pack) {
  extern __shared__ float sm[];
  const BnMap m(C);
  const bool vec = (C & 7) == 0;
  for (int g0 = 0; g0 < m.groups; g0 += m.lpr) {
    const int g = g0 + m.cg;
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
    if (g < m.groups && m.rl < m.rpb) {
      const long long stride = (long long)gridDim.x * m.rpb;
      long long r = (long long)blockIdx.x * m.rpb + m.rl;
      if (vec) {
        for (; r + (BN_STATS_UNROLL - 1) * stride < M; r += BN_STATS_UNROLL * stride) {
          uint4 q[BN_STATS_UNROLL];
#pragma unroll
          for (int u = 0; u < BN_STATS_UNROLL; ++u) q[u] = load8raw(x, r + u * stride, C, g);
#pragma unroll
          for (int u = 0; u < BN_STATS_UNROLL; ++u) {
            float f[8];
            bf8_unpack(q[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[0][j] += f[j];
              acc[1][j] += f[j] * f[j];
            }
          }
        }
      }
      for (; r < M; r += stride) {
        float f[8];
        load8(x, r, C, g, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += f[j];
          acc[1][j] += f[j] * f[j];
        }
      }
    }
    block_channel_reduce<2>(m, g, C, acc, ws, sm);
  }
  if (last_block_done(reinterpret_cast<unsigned int*>(ws + (size_t)BN_REPL * 2 * C))) {
    const float inv_m = 1.f / (float)M;
    const float unbias = M > 1 ? (float)M / (float)(M - 1) : 1.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = fold_replicas<2>(ws, 0, C, c), ss = fold_replicas<2>(ws, 1, C, c);
      float mean = s * inv_m;
      float var = fmaxf(ss * inv_m - mean * mean, 0.f);
      float invstd = rsqrtf(var + eps);
      save_mean[c] = mean;
      save_invstd[c] = invstd;
      float gm = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
      scale_shift[c] = gm * invstd;
      scale_shift[C + c] = bt - mean * gm * invstd;
      if (pack) pack[c] = make_float4(invstd, -mean * invstd, gm * invstd, bt - mean * gm * invstd);
      if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
      }
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned int*>(ws + (size_t)BN_REPL * 2 * C) = 0u;
  }
}
// ---- per-thread asynchronous rings (cp.async): every thread streams ITS OWN rows through a private ring of
// 16-byte shared-memory slots, RING stages deep, so 2..3 x RING x 16 bytes per thread are in flight without
// holding them in registers (ncu on the register-staged tall kernels: 30-43 % of DRAM bandwidth,
// long-scoreboard stalls, 3 CTAs/SM by registers).  No barrier is involved: a thread only ever reads the
// slots it filled itself, after cp.async.wait_group.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// "Tall" layout of the statistics pass: a CTA owns 64 channels (8 vector groups) and a slab of rows,
// 32 row-lanes x 8 rows in flight per thread, one shared-memory reduction and 128 atomics per CTA straight
// into ws[0:2C] (no replicas: a column is touched once per row slab).  The generic kernel above gives
// every CTA ALL channels, i.e. 2*C atomics per CTA -- 0.5 M atomics and a 16-way replica fold for the
// 512-channel WRN stage, which is why it needs 20-27 us whatever the tensor size (ncu: long-scoreboard +
// barrier stalls, 7 % DRAM utilisation).
__global__ void __launch_bounds__(256)
    k_bn_stats_tall(const bf16* __restrict__ x, long long M, int C, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                    float* running_var, float* save_mean, float* save_invstd, float* scale_shift, float* ws,
                    float4* pack, long long rows_per_split, unsigned int total_ctas) {
  __shared__ float s1[32][65], s2[32][65];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(M, r_begin + rows_per_split);
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C) {
    long long r = r_begin + rl;
    for (; r + 7 * 32 < r_end; r += 8 * 32) {
      uint4 q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * 32) * C + c0));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float f[8];
        bf8_unpack(q[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] += f[j];
          b[j] += f[j] * f[j];
        }
      }
    }
    for (; r < r_end; r += 32) {
      float f[8];
      bf8_unpack(__ldg(reinterpret_cast<const uint4*>(x + r * C + c0)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] += f[j];
        b[j] += f[j] * f[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[rl][cg * 8 + j] = a[j];
    s2[rl][cg * 8 + j] = b[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = threadIdx.x & 63;
    const bool sq = threadIdx.x >= 64;
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += sq ? s2[q][col] : s1[q][col];
    const int c = blockIdx.x * 64 + col;
    if (c < C) atomicAdd(ws + (sq ? C : 0) + c, t);
  }
  // last CTA: derive the statistics, reset the accumulators
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws + (size_t)BN_REPL * 2 * C);
    is_last = (atomicAdd(counter, 1u) == total_ctas - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float inv_m = 1.f / (float)M;
  const float unbias = M > 1 ? (float)M / (float)(M - 1) : 1.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s = __ldcg(ws + c), ss = __ldcg(ws + C + c);
    ws[c] = 0.f;
    ws[C + c] = 0.f;
    const float mean = s * inv_m;
    const float var = fmaxf(ss * inv_m - mean * mean, 0.f);
    const float invstd = rsqrtf(var + eps);
    save_mean[c] = mean;
    save_invstd[c] = invstd;
    const float gm = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    scale_shift[c] = gm * invstd;
    scale_shift[C + c] = bt - mean * gm * invstd;
    if (pack) pack[c] = make_float4(invstd, -mean * invstd, gm * invstd, bt - mean * gm * invstd);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
    }
  }
}

// cp.async version of k_bn_stats_tall (same grid, same outputs)
template <int RING>
__global__ void __launch_bounds__(256)
    k_bn_stats_async(const bf16* __restrict__ x, long long M, int C, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                     float* running_var, float* save_mean, float* save_invstd, float* scale_shift, float* ws,
                     float4* pack, long long rows_per_split, unsigned int total_ctas) {
  extern __shared__ __align__(16) uint8_t bn_dyn[];
  uint4* ring = reinterpret_cast<uint4*>(bn_dyn);                        // [RING][256]
  const int tid = threadIdx.x, cg = tid & 7, rl = tid >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(M, r_begin + rows_per_split);
  const long long r0 = r_begin + rl;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C && r0 < r_end) {
    const int n = (int)((r_end - r0 + 31) / 32);
#pragma unroll
    for (int st = 0; st < RING; ++st) {
      if (st < n) cp_async16(&ring[st * 256 + tid], x + (r0 + 32LL * st) * C + c0);
      cp_async_commit();
    }
    for (int base = 0; base < n; base += RING) {
#pragma unroll
      for (int st = 0; st < RING; ++st) {
        const int i = base + st;
        cp_async_wait<RING - 1>();
        if (i < n) {
          float f[8];
          bf8_unpack(ring[st * 256 + tid], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a[j] += f[j];
            b[j] += f[j] * f[j];
          }
        }
        if (i + RING < n) cp_async16(&ring[st * 256 + tid], x + (r0 + 32LL * (i + RING)) * C + c0);
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  float* s1 = reinterpret_cast<float*>(bn_dyn);             // [32][65] (the ring is at least 2 * 32 * 65 floats)
  float* s2 = s1 + 32 * 65;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[rl * 65 + cg * 8 + j] = a[j];
    s2[rl * 65 + cg * 8 + j] = b[j];
  }
  __syncthreads();
  if (tid < 128) {
    const int col = tid & 63;
    const bool sq = tid >= 64;
    const float* src = sq ? s2 : s1;
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += src[q * 65 + col];
    const int c = blockIdx.x * 64 + col;
    if (c < C) atomicAdd(ws + (sq ? C : 0) + c, t);
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws + (size_t)BN_REPL * 2 * C);
    is_last = (atomicAdd(counter, 1u) == total_ctas - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float inv_m = 1.f / (float)M;
  const float unbias = M > 1 ? (float)M / (float)(M - 1) : 1.f;
  for (int c = tid; c < C; c += blockDim.x) {
    const float sm_ = __ldcg(ws + c), ss = __ldcg(ws + C + c);
    ws[c] = 0.f;
    ws[C + c] = 0.f;
    const float mean = sm_ * inv_m;
    const float var = fmaxf(ss * inv_m - mean * mean, 0.f);
    const float invstd = rsqrtf(var + eps);
    save_mean[c] = mean;
    save_invstd[c] = invstd;
    const float gm = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    scale_shift[c] = gm * invstd;
    scale_shift[C + c] = bt - mean * gm * invstd;
    if (pack) pack[c] = make_float4(invstd, -mean * invstd, gm * invstd, bt - mean * gm * invstd);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
    }
  }
}

void bn_train_stats(const void* x, long long M, int C, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    float* save_mean, float* save_invstd, float* scale_shift, float* ws, float* pack,
                    cudaStream_t s) {
  {
    static int tall = -1;
    if (tall < 0) {
      const char* e = getenv("TNN_BN_STATS_TALL");
      tall = (e != nullptr) ? atoi(e) : 1;
    }
    if (tall && (C % 8) == 0 && M >= 256 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
      const int col_tiles = (C + 63) / 64;
      int splits = (4 * kSMs + col_tiles - 1) / col_tiles;
      const long long max_splits = M / 256 > 0 ? M / 256 : 1;              // >= one 8-deep pass per thread
      if (splits > max_splits) splits = (int)max_splits;
      if (splits < 1) splits = 1;
      const long long rps = ((M + splits - 1) / splits + 31) / 32 * 32;
      splits = (int)((M + rps - 1) / rps);
      dim3 grid(col_tiles, splits);
      static int use_async = -1;
      if (use_async < 0) {
        const char* e = getenv("TNN_BN_ASYNC");
        use_async = (e != nullptr) ? atoi(e) : 1;
      }
      if (use_async) {
        constexpr int RING = 8;
        constexpr int dyn = RING * 256 * 16;                    // 32 KB
        static PerDeviceOnce cfg;
        if (cfg.begin()) {
          cudaFuncSetAttribute(k_bn_stats_async<RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
          cfg.done();
        }
        k_bn_stats_async<RING><<<grid, 256, dyn, s>>>((const bf16*)x, M, C, gamma, beta, eps, momentum, running_mean,
                                                      running_var, save_mean, save_invstd, scale_shift, ws,
                                                      reinterpret_cast<float4*>(pack), rps, (unsigned int)(col_tiles * splits));
        TNN_CHECK_LAUNCH();
        return;
      }
      k_bn_stats_tall<<<grid, 256, 0, s>>>((const bf16*)x, M, C, gamma, beta, eps, momentum, running_mean, running_var,
                                           save_mean, save_invstd, scale_shift, ws, reinterpret_cast<float4*>(pack), rps,
                                           (unsigned int)(col_tiles * splits));
      TNN_CHECK_LAUNCH();
      return;
    }
  }
  static PerDeviceOnce configured;
  if (configured.begin()) {
    cudaFuncSetAttribute(k_bn_stats, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    configured.done();
  }
  k_bn_stats<<<bn_blocks(M, C, BN_STATS_UNROLL), 256, 256 * 8 * sizeof(float), s>>>(
      (const bf16*)x, M, C, gamma, beta, eps, momentum, running_mean, running_var, save_mean,
      save_invstd, scale_shift, ws, reinterpret_cast<float4*>(pack));
  TNN_CHECK_LAUNCH();
}

// Statistics that were accumulated by a producer kernel's epilogue (conv / GEMM `col_sum`
// workspace, replica layout [BN_REPL][2][C]): fold the replicas, derive mean / invstd /
// (scale, shift), update the running averages and reset the workspace.  One small CTA.
__global__ void __launch_bounds__(256)
    k_bn_finalize(float* ws, long long M, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float eps, float momentum, float* running_mean, float* running_var, float* save_mean,
                  float* save_invstd, float* scale_shift, float4* pack) {
  const float inv_m = 1.f / (float)M;
  const float unbias = M > 1 ? (float)M / (float)(M - 1) : 1.f;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float s = fold_replicas<2>(ws, 0, C, c), ss = fold_replicas<2>(ws, 1, C, c);
    float mean = s * inv_m;
    float var = fmaxf(ss * inv_m - mean * mean, 0.f);
    float invstd = rsqrtf(var + eps);
    save_mean[c] = mean;
    save_invstd[c] = invstd;
    float gm = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    scale_shift[c] = gm * invstd;
    scale_shift[C + c] = bt - mean * gm * invstd;
    if (pack) pack[c] = make_float4(invstd, -mean * invstd, gm * invstd, bt - mean * gm * invstd);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
    }
  }
}
void bn_finalize_stats(float* ws, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                       float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                       float* scale_shift, float* pack, cudaStream_t s) {
  k_bn_finalize<<<(C + 255) / 256, 256, 0, s>>>(ws, M, C, gamma, beta, eps, momentum, running_mean, running_var, save_mean,
                                             save_invstd, scale_shift, reinterpret_cast<float4*>(pack));
  TNN_CHECK_LAUNCH();
}
int bn_stat_replicas() { return BN_REPL; }

__global__ void k_bn_eval_ss(int C, const float* gamma, const float* beta, const float* rm,
                             const float* rv, float eps, float* ss) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    float invstd = rsqrtf(rv[c] + eps);
    float gm = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    ss[c] = gm * invstd;
    ss[C + c] = bt - rm[c] * gm * invstd;
  }
}
void bn_eval_scale_shift(int C, const float* gamma, const float* beta, const float* running_mean,
                         const float* running_var, float eps, float* scale_shift, cudaStream_t s) {
  k_bn_eval_ss<<<(C + 127) / 128, 128, 0, s>>>(C, gamma, beta, running_mean, running_var, eps, scale_shift);
  TNN_CHECK_LAUNCH();
}

// ------------------------------------------------------------------ BN apply
// The dropout mask index is the element's 8-wide vector index r*groups + g (stable across
// thread mappings), so forward and any recomputation agree.
template <bool RELU, bool DROP>
__global__ void __launch_bounds__(256)
    k_bn_apply(const bf16* __restrict__ x, bf16* __restrict__ y, long long M, int C,
               const float* __restrict__ ss, float p, float dscale, uint64_t seed,
               const uint64_t* __restrict__ seed_ptr, uint8_t* __restrict__ mask) {
  if (DROP && seed_ptr) seed ^= (*seed_ptr) * 0x9E3779B97F4A7C15ull;
  const BnMap m(C);
  for (int g0 = 0; g0 < m.groups; g0 += m.lpr) {
    const int g = g0 + m.cg;
    if (g >= m.groups || m.rl >= m.rpb) continue;
    float sc[8], sh[8];
    load_coef8(ss, C, g, sc);
    load_coef8(ss + C, C, g, sh);
    const long long stride = (long long)gridDim.x * m.rpb;
    long long r = (long long)blockIdx.x * m.rpb + m.rl;
    auto one = [&](long long row, float (&f)[8]) {
      bool keep[8];
      if (DROP) dropout_keep8(seed, row * m.groups + g, p, keep);
#pragma unroll
      unsigned int bits = 0u;
      for (int j = 0; j < 8; ++j) {
        float v = f[j] * sc[j] + sh[j];
        if (RELU) v = fmaxf(v, 0.f);
        if (DROP) v = keep[j] ? v * dscale : 0.f;
        f[j] = v;
        bits |= (v != 0.f ? 1u : 0u) << j;
      }
      store8(y, row, C, g, f);
      // one byte per 8-channel vector = the "gradient passes" bits: backward reads this instead of y
      if (mask != nullptr) mask[row * m.groups + g] = (uint8_t)bits;
    };
    if ((C & 7) == 0) {
      for (; r + (BN_UNROLL - 1) * stride < M; r += BN_UNROLL * stride) {
        uint4 q[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) q[u] = load8raw(x, r + u * stride, C, g);
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
          float f[8];
          bf8_unpack(q[u], f);
          one(r + u * stride, f);
        }
      }
    }
    for (; r < M; r += stride) {
      float f[8];
      load8(x, r, C, g, f);
      one(r, f);
    }
  }
}
// Tall / cp.async version of k_bn_apply (no reduction): a CTA owns 64 channels x a slab of rows, every thread
// streams its x rows through a private ring and writes y (+ the keep-bit byte) straight to global memory.
template <bool RELU, bool DROP, int RING>
__global__ void __launch_bounds__(256)
    k_bn_apply_async(const bf16* __restrict__ x, bf16* __restrict__ y, long long M, int C,
                     const float* __restrict__ ss, float p, float dscale, uint64_t seed,
                     const uint64_t* __restrict__ seed_ptr, uint8_t* __restrict__ mask, long long rows_per_split) {
  extern __shared__ __align__(16) uint8_t bn_dyn[];
  uint4* ring = reinterpret_cast<uint4*>(bn_dyn);                        // [RING][256]
  if (DROP && seed_ptr) seed ^= (*seed_ptr) * 0x9E3779B97F4A7C15ull;
  const int tid = threadIdx.x, cg = tid & 7, rl = tid >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const int groups = C >> 3, g = c0 >> 3;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(M, r_begin + rows_per_split);
  const long long r0 = r_begin + rl;
  if (c0 >= C || r0 >= r_end) return;
  const int n = (int)((r_end - r0 + 31) / 32);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(ss + c0 + j);
    sh[j] = __ldg(ss + C + c0 + j);
  }
#pragma unroll
  for (int st = 0; st < RING; ++st) {
    if (st < n) cp_async16(&ring[st * 256 + tid], x + (r0 + 32LL * st) * C + c0);
    cp_async_commit();
  }
  for (int base = 0; base < n; base += RING) {
#pragma unroll
    for (int st = 0; st < RING; ++st) {
      const int i = base + st;
      cp_async_wait<RING - 1>();
      if (i < n) {
        const long long row = r0 + 32LL * i;
        float f[8];
        bf8_unpack(ring[st * 256 + tid], f);
        bool keep[8];
        if (DROP) dropout_keep8(seed, row * groups + g, p, keep);
        unsigned int bits = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = f[j] * sc[j] + sh[j];
          if (RELU) v = fmaxf(v, 0.f);
          if (DROP) v = keep[j] ? v * dscale : 0.f;
          f[j] = v;
          bits |= (v != 0.f ? 1u : 0u) << j;
        }
        *reinterpret_cast<uint4*>(y + row * C + c0) = bf8_pack(f);
        if (mask != nullptr) mask[row * groups + g] = (uint8_t)bits;
      }
      if (i + RING < n) cp_async16(&ring[st * 256 + tid], x + (r0 + 32LL * (i + RING)) * C + c0);
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
}

void bn_apply(const void* x, void* y, long long M, int C, const float* scale_shift, int relu,
              float drop_p, uint64_t seed, const uint64_t* seed_ptr, uint8_t* mask, cudaStream_t s) {
  {
    static int use_async = -1;
    if (use_async < 0) {
      const char* e = getenv("TNN_BN_ASYNC");
      const char* t = getenv("TNN_BN_STATS_TALL");
      use_async = ((e != nullptr) ? atoi(e) : 1) && ((t != nullptr) ? atoi(t) : 1);
    }
    if (use_async && (C % 8) == 0 && M >= 256 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
      constexpr int RING = 8;
      constexpr int dyn = RING * 256 * 16;
      const int col_tiles = (C + 63) / 64;
      int splits = (4 * kSMs + col_tiles - 1) / col_tiles;
      const long long max_splits = M / 256 > 0 ? M / 256 : 1;
      if (splits > max_splits) splits = (int)max_splits;
      if (splits < 1) splits = 1;
      const long long rps = ((M + splits - 1) / splits + 31) / 32 * 32;
      splits = (int)((M + rps - 1) / rps);
      dim3 grid(col_tiles, splits);
      const float dsc = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#define TNN_BN_FWD_ASYNC(R_, D_)                                                                                  \
  k_bn_apply_async<R_, D_, RING><<<grid, 256, dyn, s>>>((const bf16*)x, (bf16*)y, M, C, scale_shift, drop_p, dsc, seed, \
                                                        seed_ptr, mask, rps)
      if (relu && drop_p > 0.f) TNN_BN_FWD_ASYNC(true, true);
      else if (relu) TNN_BN_FWD_ASYNC(true, false);
      else if (drop_p > 0.f) TNN_BN_FWD_ASYNC(false, true);
      else TNN_BN_FWD_ASYNC(false, false);
#undef TNN_BN_FWD_ASYNC
      TNN_CHECK_LAUNCH();
      return;
    }
  }
  int g = bn_blocks(M, C);
  float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const bf16* xi = (const bf16*)x;
  bf16* yo = (bf16*)y;
  if (relu && drop_p > 0.f) k_bn_apply<true, true><<<g, 256, 0, s>>>(xi, yo, M, C, scale_shift, drop_p, ds, seed, seed_ptr, mask);
  else if (relu) k_bn_apply<true, false><<<g, 256, 0, s>>>(xi, yo, M, C, scale_shift, drop_p, ds, seed, seed_ptr, mask);
  else if (drop_p > 0.f) k_bn_apply<false, true><<<g, 256, 0, s>>>(xi, yo, M, C, scale_shift, drop_p, ds, seed, seed_ptr, mask);
  else k_bn_apply<false, false><<<g, 256, 0, s>>>(xi, yo, M, C, scale_shift, drop_p, ds, seed, seed_ptr, mask);
  TNN_CHECK_LAUNCH();
}

// ------------------------------------------------------------------ BN backward
// Mask recomputation (y == nullptr with masked != 0): instead of re-reading the forward output to
// find which elements survived ReLU / dropout, the mask is rebuilt from x and the saved statistics
// (v = x*gamma*invstd + beta - mean*gamma*invstd, the exact forward expression) and from the
// counter-based dropout generator -- one tensor less to read in each of the two backward passes.
struct BnMaskArgs {
  const float* beta;
  int relu;
  float p;
  uint64_t seed;
  const uint64_t* seed_ptr;
  const uint8_t* bits;       // MODE 3: one byte of keep bits per 8-channel vector, written by the forward apply
};

// MODE 0: dy is used as is (no mask, or dy is already the masked g of the fused dgrad epilogue)
// MODE 1: mask = (y != 0), scaled by 1/(1-p)        MODE 2: mask rebuilt from x, statistics and RNG
// MODE 3: mask from the forward's keep-bit bytes (1/16 of the bytes of y: the default when C % 8 == 0)
// The modes are separate instantiations so the common paths do not carry the recompute state in
// registers: ncu showed the single do-everything kernel at 128 registers = 2 CTAs / SM and 16-22 % of
// HBM bandwidth (latency bound); these run 3 CTAs / SM with the per-channel constants pre-combined.
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 2 ? 2 : 3)
    k_bn_bwd_reduce(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ y,
                    long long M, int C, const float* __restrict__ gamma,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    float dscale, float* dgamma, float* dbeta, float* coef, float* ws, BnMaskArgs ma) {
  extern __shared__ float sm[];
  const BnMap m(C);
  const bool vec = (C & 7) == 0;
  uint64_t seed = ma.seed;
  if (MODE == 2 && ma.p > 0.f && ma.seed_ptr) seed ^= (*ma.seed_ptr) * 0x9E3779B97F4A7C15ull;
  // rows in flight per thread: the bit-mask / no-mask modes read two tensors per row and can afford four
  constexpr int U = (MODE == 0 || MODE == 3) ? 4 : 2;
  for (int g0 = 0; g0 < m.groups; g0 += m.lpr) {
    const int g = g0 + m.cg;
    float acc[2][8];       // sum g, sum g * (x - mean); the invstd factor is applied once at the end
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
    if (g < m.groups && m.rl < m.rpb) {
      float mu[8], sc[8], sh[8];
      load_coef8(mean, C, g, mu);
      if (MODE == 2) {
        float gm[8], bt[8], is[8];
        load_coef8(invstd, C, g, is);
        load_coef8(gamma, C, g, gm);
        load_coef8(ma.beta, C, g, bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (gamma == nullptr) gm[j] = 1.f;
          sc[j] = gm[j] * is[j];
          sh[j] = bt[j] - mu[j] * gm[j] * is[j];
        }
      }
      const long long stride = (long long)gridDim.x * m.rpb;
      long long r = (long long)blockIdx.x * m.rpb + m.rl;
      auto one = [&](long long row, const float (&d)[8], const float (&xv)[8], const float (&yv)[8], unsigned int bits) {
        bool keep[8];
        if (MODE == 2 && ma.p > 0.f) dropout_keep8(seed, row * m.groups + g, ma.p, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float gg = d[j];
          if (MODE == 2) {
            bool on = ma.relu ? (xv[j] * sc[j] + sh[j] > 0.f) : true;
            if (ma.p > 0.f) on = on && keep[j];
            gg = on ? gg * dscale : 0.f;
          } else if (MODE == 1) {
            gg = (yv[j] != 0.f) ? gg * dscale : 0.f;
          } else if (MODE == 3) {
            gg = ((bits >> j) & 1u) ? gg * dscale : 0.f;
          }
          acc[0][j] += gg;
          acc[1][j] += gg * (xv[j] - mu[j]);
        }
      };
      if (vec) {
        for (; r + (U - 1) * stride < M; r += U * stride) {
          uint4 qd[U], qx[U], qy[U];
          unsigned int qb[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            qd[u] = load8raw(dy, r + u * stride, C, g);
            qx[u] = load8raw(x, r + u * stride, C, g);
            if (MODE == 1) qy[u] = load8raw(y, r + u * stride, C, g);
            qb[u] = (MODE == 3) ? __ldg(ma.bits + (r + u * stride) * m.groups + g) : 0u;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float d[8], xv[8], yv[8];
            bf8_unpack(qd[u], d);
            bf8_unpack(qx[u], xv);
            if (MODE == 1) bf8_unpack(qy[u], yv);
            one(r + u * stride, d, xv, yv, qb[u]);
          }
        }
      }
      for (; r < M; r += stride) {
        float d[8], xv[8], yv[8];
        load8(dy, r, C, g, d);
        load8(x, r, C, g, xv);
        if (MODE == 1) load8(y, r, C, g, yv);
        one(r, d, xv, yv, (MODE == 3) ? __ldg(ma.bits + r * m.groups + g) : 0u);
      }
    }
    block_channel_reduce<2>(m, g, C, acc, ws, sm);
  }
  if (last_block_done(reinterpret_cast<unsigned int*>(ws + (size_t)BN_REPL * 2 * C))) {
    const float inv_m = 1.f / (float)M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float is = invstd[c];
      float db = fold_replicas<2>(ws, 0, C, c), dg = fold_replicas<2>(ws, 1, C, c) * is;
      if (dbeta) dbeta[c] += db;
      if (dgamma) dgamma[c] += dg;
      float gm = gamma ? gamma[c] : 1.f;
      coef[c] = gm * is;
      coef[C + c] = db * inv_m;
      coef[2 * C + c] = dg * inv_m;
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned int*>(ws + (size_t)BN_REPL * 2 * C) = 0u;
  }
}
// Tall layout of the backward reduction (same scheme as k_bn_stats_tall): 64 channels x a slab of rows per
// CTA, 4 rows (dy, x, keep bits) in flight per thread, 128 atomics per CTA into ws[0:2C].
template <int MODE>     // 0: dy as is, 3: keep bits
__global__ void __launch_bounds__(256)
    k_bn_bwd_reduce_tall(const bf16* __restrict__ dy, const bf16* __restrict__ x, long long M, int C,
                         const float* __restrict__ gamma, const float* __restrict__ mean,
                         const float* __restrict__ invstd, float dscale, float* dgamma, float* dbeta, float* coef,
                         float* ws, const uint8_t* __restrict__ bits, long long rows_per_split, unsigned int total_ctas) {
  __shared__ float s1[32][65], s2[32][65];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const int groups = C >> 3, g = c0 >> 3;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(M, r_begin + rows_per_split);
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C) {
    float mu[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mu[j] = __ldg(mean + c0 + j);
    auto one = [&](const uint4& qd, const uint4& qx, unsigned int kb) {
      float d[8], xv[8];
      bf8_unpack(qd, d);
      bf8_unpack(qx, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gg = d[j];
        if (MODE == 3) gg = ((kb >> j) & 1u) ? gg * dscale : 0.f;
        a[j] += gg;
        b[j] += gg * (xv[j] - mu[j]);
      }
    };
    long long r = r_begin + rl;
    for (; r + 3 * 32 < r_end; r += 4 * 32) {
      uint4 qd[4], qx[4];
      unsigned int kb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        qd[u] = __ldg(reinterpret_cast<const uint4*>(dy + (r + u * 32) * C + c0));
        qx[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * 32) * C + c0));
        kb[u] = (MODE == 3) ? __ldg(bits + (r + u * 32) * groups + g) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) one(qd[u], qx[u], kb[u]);
    }
    for (; r < r_end; r += 32)
      one(__ldg(reinterpret_cast<const uint4*>(dy + r * C + c0)), __ldg(reinterpret_cast<const uint4*>(x + r * C + c0)),
          (MODE == 3) ? __ldg(bits + r * groups + g) : 0u);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[rl][cg * 8 + j] = a[j];
    s2[rl][cg * 8 + j] = b[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = threadIdx.x & 63;
    const bool second = threadIdx.x >= 64;
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += second ? s2[q][col] : s1[q][col];
    const int c = blockIdx.x * 64 + col;
    if (c < C) atomicAdd(ws + (second ? C : 0) + c, t);
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws + (size_t)BN_REPL * 2 * C);
    is_last = (atomicAdd(counter, 1u) == total_ctas - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float inv_m = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float is = invstd[c];
    const float db = __ldcg(ws + c), dg = __ldcg(ws + C + c) * is;
    ws[c] = 0.f;
    ws[C + c] = 0.f;
    if (dbeta) dbeta[c] += db;
    if (dgamma) dgamma[c] += dg;
    coef[c] = (gamma ? gamma[c] : 1.f) * is;
    coef[C + c] = db * inv_m;
    coef[2 * C + c] = dg * inv_m;
  }
}

// cp.async version of k_bn_bwd_reduce_tall (same grid, same outputs)
template <int MODE, int RING>
__global__ void __launch_bounds__(256)
    k_bn_bwd_reduce_async(const bf16* __restrict__ dy, const bf16* __restrict__ x, long long M, int C,
                          const float* __restrict__ gamma, const float* __restrict__ mean,
                          const float* __restrict__ invstd, float dscale, float* dgamma, float* dbeta, float* coef,
                          float* ws, const uint8_t* __restrict__ bits, long long rows_per_split, unsigned int total_ctas) {
  extern __shared__ __align__(16) uint8_t bn_dyn[];
  uint4* ring = reinterpret_cast<uint4*>(bn_dyn);                        // [RING][2][256]
  const int tid = threadIdx.x, cg = tid & 7, rl = tid >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const int groups = C >> 3, g = c0 >> 3;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(M, r_begin + rows_per_split);
  const long long r0 = r_begin + rl;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C && r0 < r_end) {
    const int n = (int)((r_end - r0 + 31) / 32);
    float mu[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mu[j] = __ldg(mean + c0 + j);
    unsigned int kbr[RING];
#pragma unroll
    for (int st = 0; st < RING; ++st) {
      kbr[st] = 0u;
      if (st < n) {
        const long long row = r0 + 32LL * st;
        cp_async16(&ring[(st * 2 + 0) * 256 + tid], dy + row * C + c0);
        cp_async16(&ring[(st * 2 + 1) * 256 + tid], x + row * C + c0);
        if (MODE == 3) kbr[st] = __ldg(bits + row * groups + g);
      }
      cp_async_commit();
    }
    for (int base = 0; base < n; base += RING) {
#pragma unroll
      for (int st = 0; st < RING; ++st) {
        const int i = base + st;
        cp_async_wait<RING - 1>();
        if (i < n) {
          float d[8], xv[8];
          bf8_unpack(ring[(st * 2 + 0) * 256 + tid], d);
          bf8_unpack(ring[(st * 2 + 1) * 256 + tid], xv);
          const unsigned int kb = kbr[st];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float gg = d[j];
            if (MODE == 3) gg = ((kb >> j) & 1u) ? gg * dscale : 0.f;
            a[j] += gg;
            b[j] += gg * (xv[j] - mu[j]);
          }
        }
        const int nx = i + RING;
        if (nx < n) {
          const long long row = r0 + 32LL * nx;
          cp_async16(&ring[(st * 2 + 0) * 256 + tid], dy + row * C + c0);
          cp_async16(&ring[(st * 2 + 1) * 256 + tid], x + row * C + c0);
          if (MODE == 3) kbr[st] = __ldg(bits + row * groups + g);
        }
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                          // every thread is done with its ring: reuse it
  float* s1 = reinterpret_cast<float*>(bn_dyn);             // [32][65]
  float* s2 = s1 + 32 * 65;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[rl * 65 + cg * 8 + j] = a[j];
    s2[rl * 65 + cg * 8 + j] = b[j];
  }
  __syncthreads();
  if (tid < 128) {
    const int col = tid & 63;
    const bool second = tid >= 64;
    const float* src = second ? s2 : s1;
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += src[q * 65 + col];
    const int c = blockIdx.x * 64 + col;
    if (c < C) atomicAdd(ws + (second ? C : 0) + c, t);
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws + (size_t)BN_REPL * 2 * C);
    is_last = (atomicAdd(counter, 1u) == total_ctas - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float inv_m = 1.f / (float)M;
  for (int c = tid; c < C; c += blockDim.x) {
    const float is = invstd[c];
    const float db = __ldcg(ws + c), dg = __ldcg(ws + C + c) * is;
    ws[c] = 0.f;
    ws[C + c] = 0.f;
    if (dbeta) dbeta[c] += db;
    if (dgamma) dgamma[c] += dg;
    coef[c] = (gamma ? gamma[c] : 1.f) * is;
    coef[C + c] = db * inv_m;
    coef[2 * C + c] = dg * inv_m;
  }
}

void bn_bwd_reduce(const void* dy, const void* x, const void* y, long long M, int C,
                   const float* gamma, const float* save_mean, const float* save_invstd,
                   int masked, float drop_p, float* dgamma, float* dbeta, float* coef, float* ws,
                   const float* beta, int relu, uint64_t seed, const uint64_t* seed_ptr, const uint8_t* bits,
                   cudaStream_t s) {
  float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  BnMaskArgs ma{beta, relu, drop_p, seed, seed_ptr, bits};
  static PerDeviceOnce configured;
  if (configured.begin()) {   // ncu: the default carveout admitted ONE 8 KB-smem CTA per SM (occupancy limit "shared memory")
    cudaFuncSetAttribute(k_bn_bwd_reduce<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaFuncSetAttribute(k_bn_bwd_reduce<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaFuncSetAttribute(k_bn_bwd_reduce<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaFuncSetAttribute(k_bn_bwd_reduce<3>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    configured.done();
  }
  const int mode = !masked ? 0 : ((bits != nullptr && (C & 7) == 0) ? 3 : (y != nullptr ? 1 : 2));
  {
    static int tall = -1;
    if (tall < 0) {
      const char* e = getenv("TNN_BN_STATS_TALL");
      tall = (e != nullptr) ? atoi(e) : 1;
    }
    if (tall && (mode == 0 || mode == 3) && (C % 8) == 0 && M >= 256 &&
        ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
      const int col_tiles = (C + 63) / 64;
      int splits = (4 * kSMs + col_tiles - 1) / col_tiles;
      const long long max_splits = M / 128 > 0 ? M / 128 : 1;              // >= one 4-deep pass per thread
      if (splits > max_splits) splits = (int)max_splits;
      if (splits < 1) splits = 1;
      const long long rps = ((M + splits - 1) / splits + 31) / 32 * 32;
      splits = (int)((M + rps - 1) / rps);
      dim3 grid2(col_tiles, splits);
      const unsigned int total = (unsigned int)(col_tiles * splits);
      static int use_async = -1;
      if (use_async < 0) {
        const char* e = getenv("TNN_BN_ASYNC");
        use_async = (e != nullptr) ? atoi(e) : 1;
      }
      if (use_async) {
        constexpr int RING = 8;
        constexpr int dyn = RING * 2 * 256 * 16;                  // 64 KB: three CTAs per SM
        static PerDeviceOnce cfg;
        if (cfg.begin()) {
          cudaFuncSetAttribute(k_bn_bwd_reduce_async<0, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
          cudaFuncSetAttribute(k_bn_bwd_reduce_async<3, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
          cfg.done();
        }
        if (mode == 0)
          k_bn_bwd_reduce_async<0, RING><<<grid2, 256, dyn, s>>>((const bf16*)dy, (const bf16*)x, M, C, gamma, save_mean,
                                                                 save_invstd, ds, dgamma, dbeta, coef, ws, bits, rps, total);
        else
          k_bn_bwd_reduce_async<3, RING><<<grid2, 256, dyn, s>>>((const bf16*)dy, (const bf16*)x, M, C, gamma, save_mean,
                                                                 save_invstd, ds, dgamma, dbeta, coef, ws, bits, rps, total);
        TNN_CHECK_LAUNCH();
        return;
      }
      if (mode == 0)
        k_bn_bwd_reduce_tall<0><<<grid2, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, M, C, gamma, save_mean, save_invstd, ds,
                                                     dgamma, dbeta, coef, ws, bits, rps, total);
      else
        k_bn_bwd_reduce_tall<3><<<grid2, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, M, C, gamma, save_mean, save_invstd, ds,
                                                     dgamma, dbeta, coef, ws, bits, rps, total);
      TNN_CHECK_LAUNCH();
      return;
    }
  }
  const int grid = bn_blocks(M, C, (mode == 0 || mode == 3) ? 4 : 2);
  const size_t smem = 256 * 8 * sizeof(float);
#define TNN_BN_RED(MODE_)                                                                                         \
  k_bn_bwd_reduce<MODE_><<<grid, 256, smem, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)y, M, C, gamma,    \
                                                 save_mean, save_invstd, ds, dgamma, dbeta, coef, ws, ma)
  if (mode == 0) TNN_BN_RED(0);
  else if (mode == 1) TNN_BN_RED(1);
  else if (mode == 3) TNN_BN_RED(3);
  else TNN_BN_RED(2);
#undef TNN_BN_RED
  TNN_CHECK_LAUNCH();
}

// dx = c0*(g - c1 - xhat*c2) (+ extra) with xhat = (x - mean)*invstd, evaluated as A*g + B*x + D with
// A = c0, B = -c0*c2*invstd, D = c0*(c2*invstd*mean - c1): three constants per channel instead of five.
template <int MODE, bool EXTRA>
__global__ void __launch_bounds__(256, MODE == 2 ? 2 : 3)
    k_bn_bwd_apply(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ y,
                   const bf16* __restrict__ extra, bf16* __restrict__ dx, long long M, int C,
                   const float* __restrict__ mean, const float* __restrict__ invstd,
                   const float* __restrict__ coef, float dscale, float* colsum_out,
                   float* colsum_ws, const float* __restrict__ gamma, BnMaskArgs ma) {
  extern __shared__ float sm[];
  const BnMap m(C);
  const bool vec = (C & 7) == 0;
  uint64_t seed = ma.seed;
  if (MODE == 2 && ma.p > 0.f && ma.seed_ptr) seed ^= (*ma.seed_ptr) * 0x9E3779B97F4A7C15ull;
  for (int g0 = 0; g0 < m.groups; g0 += m.lpr) {
    const int g = g0 + m.cg;
    float csum[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) csum[0][j] = 0.f;
    if (g < m.groups && m.rl < m.rpb) {
      float ca[8], cb[8], cd[8], sc[8], sh[8];
      {
        float mu[8], is[8], c1[8], c2[8];
        load_coef8(mean, C, g, mu);
        load_coef8(invstd, C, g, is);
        load_coef8(coef, C, g, ca);
        load_coef8(coef + C, C, g, c1);
        load_coef8(coef + 2 * C, C, g, c2);
        if (MODE == 2) {
          float gm[8], bt[8];
          load_coef8(gamma, C, g, gm);
          load_coef8(ma.beta, C, g, bt);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (gamma == nullptr) gm[j] = 1.f;
            sc[j] = gm[j] * is[j];
            sh[j] = bt[j] - mu[j] * gm[j] * is[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          cb[j] = -ca[j] * c2[j] * is[j];
          cd[j] = ca[j] * (c2[j] * is[j] * mu[j] - c1[j]);
        }
      }
      const long long stride = (long long)gridDim.x * m.rpb;
      long long r = (long long)blockIdx.x * m.rpb + m.rl;
      auto one = [&](long long row, float (&d)[8], const float (&xv)[8], const float (&yv)[8], const float (&ev)[8],
                     unsigned int bits) {
        bool keep[8];
        if (MODE == 2 && ma.p > 0.f) dropout_keep8(seed, row * m.groups + g, ma.p, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float gg = d[j];
          if (MODE == 2) {
            bool on = ma.relu ? (xv[j] * sc[j] + sh[j] > 0.f) : true;
            if (ma.p > 0.f) on = on && keep[j];
            gg = on ? gg * dscale : 0.f;
          } else if (MODE == 1) {
            gg = (yv[j] != 0.f) ? gg * dscale : 0.f;
          } else if (MODE == 3) {
            gg = ((bits >> j) & 1u) ? gg * dscale : 0.f;
          }
          float v = ca[j] * gg + cb[j] * xv[j] + cd[j];
          if (EXTRA) v += ev[j];
          d[j] = v;
          csum[0][j] += v;
        }
        store8(dx, row, C, g, d);
      };
      if (vec) {
        constexpr int U = 2;      // (3 rows in flight measured slower: 278 vs 254 us per step over the 7 launches)
        for (; r + (U - 1) * stride < M; r += U * stride) {
          uint4 qd[U], qx[U], qy[U], qe[U];
          unsigned int qb[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            qd[u] = load8raw(dy, r + u * stride, C, g);
            qx[u] = load8raw(x, r + u * stride, C, g);
            if (MODE == 1) qy[u] = load8raw(y, r + u * stride, C, g);
            if (EXTRA) qe[u] = load8raw(extra, r + u * stride, C, g);
            qb[u] = (MODE == 3) ? __ldg(ma.bits + (r + u * stride) * m.groups + g) : 0u;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float d[8], xv[8], yv[8], ev[8];
            bf8_unpack(qd[u], d);
            bf8_unpack(qx[u], xv);
            if (MODE == 1) bf8_unpack(qy[u], yv);
            if (EXTRA) bf8_unpack(qe[u], ev);
            one(r + u * stride, d, xv, yv, ev, qb[u]);
          }
        }
      }
      for (; r < M; r += stride) {
        float d[8], xv[8], yv[8], ev[8];
        load8(dy, r, C, g, d);
        load8(x, r, C, g, xv);
        if (MODE == 1) load8(y, r, C, g, yv);
        if (EXTRA) load8(extra, r, C, g, ev);
        one(r, d, xv, yv, ev, (MODE == 3) ? __ldg(ma.bits + r * m.groups + g) : 0u);
      }
    }
    if (colsum_out) block_channel_reduce<1>(m, g, C, csum, colsum_ws, sm);
  }
  // fused bias gradient of the producer convolution: column sums of dx
  if (colsum_out && last_block_done(reinterpret_cast<unsigned int*>(colsum_ws + (size_t)BN_REPL * C))) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) colsum_out[c] = fold_replicas<1>(colsum_ws, 0, C, c);
    if (threadIdx.x == 0) *reinterpret_cast<unsigned int*>(colsum_ws + (size_t)BN_REPL * C) = 0u;
  }
}
// Backward statistics that a dgrad epilogue accumulated (ws = [BN_REPL][2][C]: sum g, sum g*xhat):
// fold, add into dgamma / dbeta, write the apply coefficients, reset the workspace.  One small CTA.
__global__ void __launch_bounds__(256)
    k_bn_bwd_coef(float* ws, long long M, int C, const float* __restrict__ gamma, const float* __restrict__ invstd,
                  float* dgamma, float* dbeta, float* coef) {
  const float inv_m = 1.f / (float)M;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float db = fold_replicas<2>(ws, 0, C, c), dg = fold_replicas<2>(ws, 1, C, c);
    if (dbeta) dbeta[c] += db;
    if (dgamma) dgamma[c] += dg;
    coef[c] = (gamma ? gamma[c] : 1.f) * invstd[c];
    coef[C + c] = db * inv_m;
    coef[2 * C + c] = dg * inv_m;
  }
}
void bn_bwd_coef(float* ws, long long M, int C, const float* gamma, const float* save_invstd, float* dgamma,
                 float* dbeta, float* coef, cudaStream_t s) {
  k_bn_bwd_coef<<<(C + 255) / 256, 256, 0, s>>>(ws, M, C, gamma, save_invstd, dgamma, dbeta, coef);
  TNN_CHECK_LAUNCH();
}

// Tall layout of the backward apply (see k_bn_stats_tall): 64 channels x a slab of rows per CTA; the
// per-channel constants are loaded once per thread, the fused bias-gradient column sums cost 64 atomics
// per CTA instead of C.
template <int MODE, bool EXTRA>     // MODE 0: dy as is, 3: keep bits
__global__ void __launch_bounds__(256)
    k_bn_bwd_apply_tall(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ extra,
                        bf16* __restrict__ dx, long long M, int C, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ coef, float dscale,
                        float* colsum_out, float* colsum_ws, const uint8_t* __restrict__ bits,
                        long long rows_per_split, unsigned int total_ctas) {
  __shared__ float s1[32][65];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const int groups = C >> 3, g = c0 >> 3;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(M, r_begin + rows_per_split);
  float csum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C) {
    float ca[8], cb[8], cd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mu = __ldg(mean + c0 + j), is = __ldg(invstd + c0 + j);
      const float c1 = __ldg(coef + C + c0 + j), c2 = __ldg(coef + 2 * C + c0 + j);
      ca[j] = __ldg(coef + c0 + j);
      cb[j] = -ca[j] * c2 * is;
      cd[j] = ca[j] * (c2 * is * mu - c1);
    }
    auto one = [&](long long row, const uint4& qd, const uint4& qx, const uint4& qe, unsigned int kb) {
      float d[8], xv[8], ev[8];
      bf8_unpack(qd, d);
      bf8_unpack(qx, xv);
      if (EXTRA) bf8_unpack(qe, ev);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gg = d[j];
        if (MODE == 3) gg = ((kb >> j) & 1u) ? gg * dscale : 0.f;
        float v = ca[j] * gg + cb[j] * xv[j] + cd[j];
        if (EXTRA) v += ev[j];
        d[j] = v;
        csum[j] += v;
      }
      *reinterpret_cast<uint4*>(dx + row * C + c0) = bf8_pack(d);
    };
    constexpr int U = (EXTRA || MODE == 3) ? 2 : 4;
    long long r = r_begin + rl;
    for (; r + (U - 1) * 32 < r_end; r += U * 32) {
      uint4 qd[U], qx[U], qe[U];
      unsigned int kb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        qd[u] = __ldg(reinterpret_cast<const uint4*>(dy + (r + u * 32) * C + c0));
        qx[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + u * 32) * C + c0));
        if (EXTRA) qe[u] = __ldg(reinterpret_cast<const uint4*>(extra + (r + u * 32) * C + c0));
        kb[u] = (MODE == 3) ? __ldg(bits + (r + u * 32) * groups + g) : 0u;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) one(r + u * 32, qd[u], qx[u], qe[u], kb[u]);
    }
    for (; r < r_end; r += 32) {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      one(r, __ldg(reinterpret_cast<const uint4*>(dy + r * C + c0)), __ldg(reinterpret_cast<const uint4*>(x + r * C + c0)),
          EXTRA ? __ldg(reinterpret_cast<const uint4*>(extra + r * C + c0)) : z, (MODE == 3) ? __ldg(bits + r * groups + g) : 0u);
    }
  }
  if (colsum_out == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[rl][cg * 8 + j] = csum[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += s1[q][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < C) atomicAdd(colsum_ws + c, t);
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(colsum_ws + (size_t)BN_REPL * C);
    is_last = (atomicAdd(counter, 1u) == total_ctas - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    colsum_out[c] = __ldcg(colsum_ws + c);
    colsum_ws[c] = 0.f;
  }
}

// cp.async version of k_bn_bwd_apply_tall (same grid, same outputs)
template <int MODE, bool EXTRA, int RING>
__global__ void __launch_bounds__(256)
    k_bn_bwd_apply_async(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ extra,
                         bf16* __restrict__ dx, long long M, int C, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ coef, float dscale,
                         float* colsum_out, float* colsum_ws, const uint8_t* __restrict__ bits,
                         long long rows_per_split, unsigned int total_ctas) {
  constexpr int NT = EXTRA ? 3 : 2;
  extern __shared__ __align__(16) uint8_t bn_dyn[];
  uint4* ring = reinterpret_cast<uint4*>(bn_dyn);                        // [RING][NT][256]
  const int tid = threadIdx.x, cg = tid & 7, rl = tid >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const int groups = C >> 3, g = c0 >> 3;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(M, r_begin + rows_per_split);
  const long long r0 = r_begin + rl;
  float csum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C && r0 < r_end) {
    const int n = (int)((r_end - r0 + 31) / 32);
    float ca[8], cb[8], cd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mu = __ldg(mean + c0 + j), is = __ldg(invstd + c0 + j);
      const float c1 = __ldg(coef + C + c0 + j), c2 = __ldg(coef + 2 * C + c0 + j);
      ca[j] = __ldg(coef + c0 + j);
      cb[j] = -ca[j] * c2 * is;
      cd[j] = ca[j] * (c2 * is * mu - c1);
    }
    unsigned int kbr[RING];
    auto issue = [&](int st, int i) {
      const long long row = r0 + 32LL * i;
      cp_async16(&ring[(st * NT + 0) * 256 + tid], dy + row * C + c0);
      cp_async16(&ring[(st * NT + 1) * 256 + tid], x + row * C + c0);
      if (EXTRA) cp_async16(&ring[(st * NT + 2) * 256 + tid], extra + row * C + c0);
    };
#pragma unroll
    for (int st = 0; st < RING; ++st) {
      kbr[st] = 0u;
      if (st < n) {
        issue(st, st);
        if (MODE == 3) kbr[st] = __ldg(bits + (r0 + 32LL * st) * groups + g);
      }
      cp_async_commit();
    }
    for (int base = 0; base < n; base += RING) {
#pragma unroll
      for (int st = 0; st < RING; ++st) {
        const int i = base + st;
        cp_async_wait<RING - 1>();
        if (i < n) {
          float d[8], xv[8], ev[8];
          bf8_unpack(ring[(st * NT + 0) * 256 + tid], d);
          bf8_unpack(ring[(st * NT + 1) * 256 + tid], xv);
          if (EXTRA) bf8_unpack(ring[(st * NT + 2) * 256 + tid], ev);
          const unsigned int kb = kbr[st];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float gg = d[j];
            if (MODE == 3) gg = ((kb >> j) & 1u) ? gg * dscale : 0.f;
            float v = ca[j] * gg + cb[j] * xv[j] + cd[j];
            if (EXTRA) v += ev[j];
            d[j] = v;
            csum[j] += v;
          }
          *reinterpret_cast<uint4*>(dx + (r0 + 32LL * i) * C + c0) = bf8_pack(d);
        }
        const int nx = i + RING;
        if (nx < n) {
          issue(st, nx);
          if (MODE == 3) kbr[st] = __ldg(bits + (r0 + 32LL * nx) * groups + g);
        }
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();
  if (colsum_out == nullptr) return;
  __syncthreads();
  float* s1 = reinterpret_cast<float*>(bn_dyn);             // [32][65]
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[rl * 65 + cg * 8 + j] = csum[j];
  __syncthreads();
  if (tid < 64) {
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += s1[q * 65 + tid];
    const int c = blockIdx.x * 64 + tid;
    if (c < C) atomicAdd(colsum_ws + c, t);
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(colsum_ws + (size_t)BN_REPL * C);
    is_last = (atomicAdd(counter, 1u) == total_ctas - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = tid; c < C; c += blockDim.x) {
    colsum_out[c] = __ldcg(colsum_ws + c);
    colsum_ws[c] = 0.f;
  }
}

void bn_bwd_apply(const void* dy, const void* x, const void* y, const void* extra, void* dx,
                  long long M, int C, const float* save_mean, const float* save_invstd,
                  const float* coef, int masked, float drop_p, float* colsum_out, float* colsum_ws,
                  const float* gamma, const float* beta, int relu, uint64_t seed, const uint64_t* seed_ptr,
                  const uint8_t* bits, cudaStream_t s) {
  float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  BnMaskArgs ma{beta, relu, drop_p, seed, seed_ptr, bits};
  const int mode = !masked ? 0 : ((bits != nullptr && (C & 7) == 0) ? 3 : (y != nullptr ? 1 : 2));
  {
    static int tall = -1;
    if (tall < 0) {
      const char* e = getenv("TNN_BN_STATS_TALL");
      tall = (e != nullptr) ? atoi(e) : 1;
    }
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (tall && (mode == 0 || mode == 3) && (C % 8) == 0 && M >= 256 && al(dy) && al(x) && al(extra) && al(dx)) {
      const int col_tiles = (C + 63) / 64;
      int splits = (4 * kSMs + col_tiles - 1) / col_tiles;
      const long long max_splits = M / 128 > 0 ? M / 128 : 1;
      if (splits > max_splits) splits = (int)max_splits;
      if (splits < 1) splits = 1;
      const long long rps = ((M + splits - 1) / splits + 31) / 32 * 32;
      splits = (int)((M + rps - 1) / rps);
      dim3 grid2(col_tiles, splits);
      const unsigned int total = (unsigned int)(col_tiles * splits);
      static int use_async = -1;
      if (use_async < 0) {
        const char* e = getenv("TNN_BN_ASYNC");
        use_async = (e != nullptr) ? atoi(e) : 1;
      }
      if (use_async) {
#define TNN_BN_APPLY_ASYNC(MODE_, EXTRA_, RING_)                                                                     \
  do {                                                                                                               \
    constexpr int dyn = RING_ * ((EXTRA_) ? 3 : 2) * 256 * 16;                                                       \
    static PerDeviceOnce cfg;                                                                                         \
    if (cfg.begin()) {                                                                                                      \
      cudaFuncSetAttribute(k_bn_bwd_apply_async<MODE_, EXTRA_, RING_>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn); \
      cfg.done();                                                                                                    \
    }                                                                                                                \
    k_bn_bwd_apply_async<MODE_, EXTRA_, RING_><<<grid2, 256, dyn, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)extra, \
                                                                       (bf16*)dx, M, C, save_mean, save_invstd, coef, ds, \
                                                                       colsum_out, colsum_ws, bits, rps, total);      \
  } while (0)
        if (extra != nullptr) {
          if (mode == 0) TNN_BN_APPLY_ASYNC(0, true, 5);
          else TNN_BN_APPLY_ASYNC(3, true, 5);
        } else {
          if (mode == 0) TNN_BN_APPLY_ASYNC(0, false, 8);
          else TNN_BN_APPLY_ASYNC(3, false, 8);
        }
#undef TNN_BN_APPLY_ASYNC
        TNN_CHECK_LAUNCH();
        return;
      }
#define TNN_BN_APPLY_TALL(MODE_, EXTRA_)                                                                          \
  k_bn_bwd_apply_tall<MODE_, EXTRA_><<<grid2, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)extra,   \
                                                           (bf16*)dx, M, C, save_mean, save_invstd, coef, ds,      \
                                                           colsum_out, colsum_ws, bits, rps, total)
      if (extra != nullptr) {
        if (mode == 0) TNN_BN_APPLY_TALL(0, true);
        else TNN_BN_APPLY_TALL(3, true);
      } else {
        if (mode == 0) TNN_BN_APPLY_TALL(0, false);
        else TNN_BN_APPLY_TALL(3, false);
      }
#undef TNN_BN_APPLY_TALL
      TNN_CHECK_LAUNCH();
      return;
    }
  }
  const int grid = bn_blocks(M, C, 2);
  const size_t smem = 256 * 8 * sizeof(float);
#define TNN_BN_APPLY(MODE_, EXTRA_)                                                                              \
  do {                                                                                                           \
    static PerDeviceOnce configured;                                                                              \
    if (configured.begin()) {                                                                                           \
      cudaFuncSetAttribute(k_bn_bwd_apply<MODE_, EXTRA_>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);   \
      configured.done();                                                                                         \
    }                                                                                                            \
    k_bn_bwd_apply<MODE_, EXTRA_><<<grid, 256, smem, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)y,       \
                                                          (const bf16*)extra, (bf16*)dx, M, C, save_mean,        \
                                                          save_invstd, coef, ds, colsum_out, colsum_ws, gamma, ma); \
  } while (0)
  if (extra != nullptr) {
    if (mode == 0) TNN_BN_APPLY(0, true);
    else if (mode == 1) TNN_BN_APPLY(1, true);
    else if (mode == 3) TNN_BN_APPLY(3, true);
    else TNN_BN_APPLY(2, true);
  } else {
    if (mode == 0) TNN_BN_APPLY(0, false);
    else if (mode == 1) TNN_BN_APPLY(1, false);
    else if (mode == 3) TNN_BN_APPLY(3, false);
    else TNN_BN_APPLY(2, false);
  }
#undef TNN_BN_APPLY
  TNN_CHECK_LAUNCH();
}

// ------------------------------------------------------------------ LayerNorm (warp per row)
__global__ void __launch_bounds__(256)
    k_ln_fwd(const bf16* __restrict__ x, const bf16* __restrict__ res, bf16* __restrict__ y,
             bf16* __restrict__ sum_out, long long rows, int C, const float* __restrict__ gamma,
             const float* __restrict__ beta, float eps, float* mean, float* rstd) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: dependents may start their prologue
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const bf16* xr = x + r * C;
    const bf16* rr = res ? res + r * C : nullptr;
    float s = 0.f, ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      float v = __bfloat162float(xr[c]);
      if (rr) v += __bfloat162float(rr[c]);
      if (sum_out) {
        bf16 h = __float2bfloat16(v);
        sum_out[r * C + c] = h;
        v = __bfloat162float(h);
      }
      s += v;
      ss += v * v;
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    float m = s / C;
    float var = fmaxf(ss / C - m * m, 0.f);
    float rs = rsqrtf(var + eps);
    if (lane == 0) {
      if (mean) mean[r] = m;
      if (rstd) rstd[r] = rs;
    }
    const bf16* src = sum_out ? sum_out + r * C : xr;
    for (int c = lane; c < C; c += 32) {
      float v = __bfloat162float(src[c]);
      if (!sum_out && rr) v += __bfloat162float(rr[c]);
      float o = (v - m) * rs;
      if (gamma) o = o * gamma[c] + (beta ? beta[c] : 0.f);
      y[r * C + c] = __float2bfloat16(o);
    }
  }
}
// Rows of up to 32 lanes x NV 16-byte vectors (C % 8 == 0, C <= 256 * NV) stay in registers between the
// statistics and the normalisation: one global read per operand, 16-byte loads and stores
// (the scalar kernel above re-reads the row with 2-byte loads: 15 us for [8192, 768], 2 TB/s).
template <int NV>
__global__ void __launch_bounds__(256)
    k_ln_fwd_vec(const bf16* __restrict__ x, const bf16* __restrict__ res, bf16* __restrict__ y,
                 bf16* __restrict__ sum_out, long long rows, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, float* mean, float* rstd) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = C >> 3;
  for (long long r = warp; r < rows; r += nwarps) {
    float v[NV][8];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        bf8_unpack(__ldg(reinterpret_cast<const uint4*>(x + r * C) + vi), v[i]);
        if (res) {
          float t[8];
          bf8_unpack(__ldg(reinterpret_cast<const uint4*>(res + r * C) + vi), t);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += t[j];
        }
        if (sum_out) {
          const uint4 q = bf8_pack(v[i]);
          reinterpret_cast<uint4*>(sum_out + r * C)[vi] = q;
          bf8_unpack(q, v[i]);                       // statistics of the ROUNDED sum, as the scalar path
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s += v[i][j];
          ss += v[i][j] * v[i][j];
        }
      }
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    const float m = s / C;
    const float rs = rsqrtf(fmaxf(ss / C - m * m, 0.f) + eps);
    if (lane == 0) {
      if (mean) mean[r] = m;
      if (rstd) rstd[r] = rs;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = (v[i][j] - m) * rs;
          if (gamma) o[j] = o[j] * __ldg(gamma + vi * 8 + j) + (beta ? __ldg(beta + vi * 8 + j) : 0.f);
        }
        reinterpret_cast<uint4*>(y + r * C)[vi] = bf8_pack(o);
      }
    }
  }
}
static bool ln_vec_ok(const void* a, const void* b, const void* c, const void* d, int C) {
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return (C % 8) == 0 && C <= 256 * 8 && al(a) && al(b) && al(c) && al(d);
}

void layernorm_fwd(const void* x, const void* residual, void* y, void* sum_out, long long rows, int C,
                   const float* gamma, const float* beta, float eps, float* mean, float* rstd,
                   cudaStream_t s) {
  if (rows <= 0) return;
  if (ln_vec_ok(x, residual, y, sum_out, C)) {
    const int g = grid_for(rows * 32, 256);
    const int nv = (C / 8 + 31) / 32;
#define TNN_LN_FWD(NV_)                                                                                      \
  k_ln_fwd_vec<NV_><<<g, 256, 0, s>>>((const bf16*)x, (const bf16*)residual, (bf16*)y, (bf16*)sum_out, rows, C, \
                                      gamma, beta, eps, mean, rstd)
    if (nv <= 1) TNN_LN_FWD(1);
    else if (nv <= 2) TNN_LN_FWD(2);
    else if (nv <= 4) TNN_LN_FWD(4);
    else TNN_LN_FWD(8);
#undef TNN_LN_FWD
    TNN_CHECK_LAUNCH();
    return;
  }
  int g = grid_for(rows * 32, 256);
  k_ln_fwd<<<g, 256, 0, s>>>((const bf16*)x, (const bf16*)residual, (bf16*)y, (bf16*)sum_out, rows, C,
                             gamma, beta, eps, mean, rstd);
  TNN_CHECK_LAUNCH();
}
__global__ void __launch_bounds__(256)
    k_ln_bwd_dx(const bf16* __restrict__ dy, const bf16* __restrict__ x, long long rows, int C,
                const float* __restrict__ gamma, const float* __restrict__ mean,
                const float* __restrict__ rstd, bf16* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float m = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      float g = __bfloat162float(dy[r * C + c]) * (gamma ? gamma[c] : 1.f);
      float xh = (__bfloat162float(x[r * C + c]) - m) * rs;
      s1 += g;
      s2 += g * xh;
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    for (int c = lane; c < C; c += 32) {
      float g = __bfloat162float(dy[r * C + c]) * (gamma ? gamma[c] : 1.f);
      float xh = (__bfloat162float(x[r * C + c]) - m) * rs;
      dx[r * C + c] = __float2bfloat16(rs * (g - s1 - xh * s2));
    }
  }
}
__global__ void __launch_bounds__(256)
    k_ln_bwd_params(const bf16* __restrict__ dy, const bf16* __restrict__ x, long long rows, int C,
                    const float* __restrict__ mean, const float* __restrict__ rstd, float* dgamma,
                    float* dbeta) {
  // thread per column, block strides over a slab of rows
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float dg = 0.f, db = 0.f;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    float g = __bfloat162float(dy[r * C + c]);
    float xh = (__bfloat162float(x[r * C + c]) - mean[r]) * rstd[r];
    dg += g * xh;
    db += g;
  }
  if (dgamma) atomicAdd(dgamma + c, dg);
  if (dbeta) atomicAdd(dbeta + c, db);
}
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ extra: the gradient that bypassed this layer through
// a residual connection, so the block's "main + shortcut" add needs no kernel of its own)
template <int NV>
__global__ void __launch_bounds__(256)
    k_ln_bwd_dx_vec(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ extra,
                    long long rows, int C, const float* __restrict__ gamma, const float* __restrict__ mean,
                    const float* __restrict__ rstd, bf16* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = C >> 3;
  for (long long r = warp; r < rows; r += nwarps) {
    const float m = mean[r], rs = rstd[r];
    float g[NV][8], xh[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        bf8_unpack(__ldg(reinterpret_cast<const uint4*>(dy + r * C) + vi), g[i]);
        bf8_unpack(__ldg(reinterpret_cast<const uint4*>(x + r * C) + vi), xh[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (gamma) g[i][j] *= __ldg(gamma + vi * 8 + j);
          xh[i][j] = (xh[i][j] - m) * rs;
          s1 += g[i][j];
          s2 += g[i][j] * xh[i][j];
        }
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float o[8], e[8];
        if (extra) bf8_unpack(__ldg(reinterpret_cast<const uint4*>(extra + r * C) + vi), e);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = rs * (g[i][j] - s1 - xh[i][j] * s2);
          if (extra) o[j] += e[j];
        }
        reinterpret_cast<uint4*>(dx + r * C)[vi] = bf8_pack(o);
      }
    }
  }
}
// dgamma / dbeta: a CTA owns 64 columns (8 vector groups) and a slab of rows; 32 row-lanes x 16-byte loads,
// shared-memory reduction, one atomic per column and quantity
__global__ void __launch_bounds__(256)
    k_ln_bwd_params_vec(const bf16* __restrict__ dy, const bf16* __restrict__ x, long long rows, int C,
                        const float* __restrict__ mean, const float* __restrict__ rstd, float* dgamma,
                        float* dbeta, long long rows_per_split) {
  __shared__ float sg[32][65], sb[32][65];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cg * 8;
  const long long r_begin = (long long)blockIdx.y * rows_per_split;
  const long long r_end = min(rows, r_begin + rows_per_split);
  float ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C) {
    for (long long r = r_begin + rl; r < r_end; r += 32) {
      float g[8], xv[8];
      bf8_unpack(__ldg(reinterpret_cast<const uint4*>(dy + r * C + c0)), g);
      bf8_unpack(__ldg(reinterpret_cast<const uint4*>(x + r * C + c0)), xv);
      const float m = __ldg(mean + r), rs = __ldg(rstd + r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ag[j] += g[j] * (xv[j] - m) * rs;
        ab[j] += g[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sg[rl][cg * 8 + j] = ag[j];
    sb[rl][cg * 8 + j] = ab[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = threadIdx.x & 63;
    const bool is_b = threadIdx.x >= 64;
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += is_b ? sb[q][col] : sg[q][col];
    const int c = blockIdx.x * 64 + col;
    if (c < C) {
      if (is_b) {
        if (dbeta) atomicAdd(dbeta + c, t);
      } else if (dgamma) {
        atomicAdd(dgamma + c, t);
      }
    }
  }
}

void layernorm_bwd(const void* dy, const void* x, long long rows, int C, const float* gamma,
                   const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                   const void* extra, cudaStream_t s) {
  if (rows <= 0) return;
  if (ln_vec_ok(dy, x, dx, extra, C)) {
    const int g = grid_for(rows * 32, 256);
    const int nv = (C / 8 + 31) / 32;
#define TNN_LN_BWD(NV_)                                                                                          \
  k_ln_bwd_dx_vec<NV_><<<g, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)extra, rows, C, gamma, mean, \
                                         rstd, (bf16*)dx)
    if (nv <= 1) TNN_LN_BWD(1);
    else if (nv <= 2) TNN_LN_BWD(2);
    else if (nv <= 4) TNN_LN_BWD(4);
    else TNN_LN_BWD(8);
#undef TNN_LN_BWD
    if (dgamma || dbeta) {
      const int col_tiles = (C + 63) / 64;
      int splits = (4 * kSMs + col_tiles - 1) / col_tiles;
      const long long max_splits = rows / 64 > 0 ? rows / 64 : 1;
      if (splits > max_splits) splits = (int)max_splits;
      const long long rps = ((rows + splits - 1) / splits + 31) / 32 * 32;
      splits = (int)((rows + rps - 1) / rps);
      k_ln_bwd_params_vec<<<dim3(col_tiles, splits), 256, 0, s>>>((const bf16*)dy, (const bf16*)x, rows, C, mean, rstd,
                                                                  dgamma, dbeta, rps);
    }
    TNN_CHECK_LAUNCH();
    return;
  }
  if (extra != nullptr) {
    fprintf(stderr, "[tnn_b200] layernorm_bwd: the fused residual-gradient add needs C %% 8 == 0 and 16-byte aligned rows\n");
    abort();
  }
  k_ln_bwd_dx<<<grid_for(rows * 32, 256), 256, 0, s>>>((const bf16*)dy, (const bf16*)x, rows, C, gamma,
                                                      mean, rstd, (bf16*)dx);
  if (dgamma || dbeta) {
    int ysplit = (int)(rows < 512 ? rows : 512);
    dim3 grid((C + 127) / 128, ysplit);
    k_ln_bwd_params<<<grid, 128, 0, s>>>((const bf16*)dy, (const bf16*)x, rows, C, mean, rstd, dgamma, dbeta);
  }
  TNN_CHECK_LAUNCH();
}

// ------------------------------------------------------------------ GroupNorm (block per (n, g))
__global__ void __launch_bounds__(256)
    k_gn_fwd(const bf16* __restrict__ x, bf16* __restrict__ y, int HW, int C, int G,
             const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* mean,
             float* rstd) {
  __shared__ float red[32];
  const int n = blockIdx.x / G, g = blockIdx.x % G, Cg = C / G;
  const long long base = (long long)n * HW * C + g * Cg;
  const int cnt = HW * Cg;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int p = i / Cg, c = i - p * Cg;
    float v = __bfloat162float(x[base + (long long)p * C + c]);
    s += v;
    ss += v * v;
  }
  s = block_sum(s, red);
  ss = block_sum(ss, red);
  float m = s / cnt, var = fmaxf(ss / cnt - m * m, 0.f), rs = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    mean[blockIdx.x] = m;
    rstd[blockIdx.x] = rs;
  }
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int p = i / Cg, c = i - p * Cg;
    float v = (__bfloat162float(x[base + (long long)p * C + c]) - m) * rs;
    int ch = g * Cg + c;
    if (gamma) v = v * gamma[ch] + (beta ? beta[ch] : 0.f);
    y[base + (long long)p * C + c] = __float2bfloat16(v);
  }
}
void groupnorm_fwd(const void* x, void* y, int N, int HW, int C, int G, const float* gamma,
                   const float* beta, float eps, float* mean, float* rstd, cudaStream_t s) {
  k_gn_fwd<<<N * G, 256, 0, s>>>((const bf16*)x, (bf16*)y, HW, C, G, gamma, beta, eps, mean, rstd);
  TNN_CHECK_LAUNCH();
}
__global__ void __launch_bounds__(256)
    k_gn_bwd(const bf16* __restrict__ dy, const bf16* __restrict__ x, int HW, int C, int G,
             const float* __restrict__ gamma, const float* __restrict__ mean,
             const float* __restrict__ rstd, bf16* __restrict__ dx, float* dgamma, float* dbeta) {
  __shared__ float red[32];
  const int n = blockIdx.x / G, g = blockIdx.x % G, Cg = C / G;
  const long long base = (long long)n * HW * C + g * Cg;
  const int cnt = HW * Cg;
  const float m = mean[blockIdx.x], rs = rstd[blockIdx.x];
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int p = i / Cg, c = i - p * Cg;
    float gg = __bfloat162float(dy[base + (long long)p * C + c]) * (gamma ? gamma[g * Cg + c] : 1.f);
    float xh = (__bfloat162float(x[base + (long long)p * C + c]) - m) * rs;
    s1 += gg;
    s2 += gg * xh;
  }
  s1 = block_sum(s1, red) / cnt;
  s2 = block_sum(s2, red) / cnt;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int p = i / Cg, c = i - p * Cg;
    float d = __bfloat162float(dy[base + (long long)p * C + c]);
    float gg = d * (gamma ? gamma[g * Cg + c] : 1.f);
    float xh = (__bfloat162float(x[base + (long long)p * C + c]) - m) * rs;
    dx[base + (long long)p * C + c] = __float2bfloat16(rs * (gg - s1 - xh * s2));
  }
  // per-channel affine grads: thread c sums over the pixels of this (n, g)
  for (int c = threadIdx.x; c < Cg; c += blockDim.x) {
    float dg = 0.f, db = 0.f;
    for (int p = 0; p < HW; ++p) {
      float d = __bfloat162float(dy[base + (long long)p * C + c]);
      float xh = (__bfloat162float(x[base + (long long)p * C + c]) - m) * rs;
      dg += d * xh;
      db += d;
    }
    if (dgamma) atomicAdd(dgamma + g * Cg + c, dg);
    if (dbeta) atomicAdd(dbeta + g * Cg + c, db);
  }
}
void groupnorm_bwd(const void* dy, const void* x, int N, int HW, int C, int G, const float* gamma,
                   const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                   cudaStream_t s) {
  k_gn_bwd<<<N * G, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, HW, C, G, gamma, mean, rstd,
                                 (bf16*)dx, dgamma, dbeta);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
