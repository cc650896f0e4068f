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

// Generic "per-channel sums of K quantities over M rows" skeleton.
// Each functor call yields K values per channel for one 8-wide vector.
template <int K, typename F>
__device__ __forceinline__ void channel_reduce(long long M, int C, float* ws, float* sm, F f) {
  const int groups = (C + 7) / 8;
  const int lpr = min(groups, 256);
  const int rpb = 256 / lpr;
  const int cg = threadIdx.x % lpr, rl = threadIdx.x / lpr;
  for (int g0 = 0; g0 < groups; g0 += lpr) {
    const int g = g0 + cg;
    float acc[K][8];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
    if (g < groups && rl < rpb) {
      for (long long r = (long long)blockIdx.x * rpb + rl; r < M; r += (long long)gridDim.x * rpb)
        f(r, g, acc);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) sm[(rl * lpr + cg) * 8 + j] = (rl < rpb) ? acc[k][j] : 0.f;
      __syncthreads();
      if (rl == 0 && g < groups) {
        for (int j = 0; j < 8 && g * 8 + j < C; ++j) {
          float t = 0.f;
          for (int q = 0; q < rpb; ++q) t += sm[(q * lpr + cg) * 8 + j];
          atomicAdd(ws + k * C + g * 8 + j, t);
        }
      }
    }
  }
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

__device__ __forceinline__ void load8(const bf16* x, long long r, int C, int g, float (&f)[8]) {
  if ((C & 7) == 0) {
    bf8_unpack(*reinterpret_cast<const uint4*>(x + r * C + g * 8), f);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (g * 8 + j < C) ? __bfloat162float(x[r * C + g * 8 + j]) : 0.f;
  }
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

// ------------------------------------------------------------------ BN forward statistics
__global__ void __launch_bounds__(256)
    k_bn_stats(const bf16* __restrict__ x, long long M, int C, const float* __restrict__ gamma,
               const float* __restrict__ beta, float eps, float momentum, float* running_mean,
               float* running_var, float* save_mean, float* save_invstd, float* scale_shift, float* ws) {
  extern __shared__ float sm[];
  channel_reduce<2>(M, C, ws, sm, [&](long long r, int g, float (&acc)[2][8]) {
    float f[8];
    load8(x, r, C, g, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[0][j] += f[j];
      acc[1][j] += f[j] * f[j];
    }
  });
  if (last_block_done(reinterpret_cast<unsigned int*>(ws + 2 * C))) {
    const float inv_m = 1.f / (float)M;
    const float unbias = M > 1 ? (float)M / (float)(M - 1) : 1.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = __ldcg(ws + c), ss = __ldcg(ws + C + c);
      float mean = s * inv_m;
      float var = fmaxf(ss * inv_m - mean * mean, 0.f);
      float invstd = rsqrtf(var + eps);
      save_mean[c] = mean;
      save_invstd[c] = invstd;
      float gm = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
      scale_shift[c] = gm * invstd;
      scale_shift[C + c] = bt - mean * gm * invstd;
      if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
      }
      ws[c] = 0.f;
      ws[C + c] = 0.f;
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned int*>(ws + 2 * C) = 0u;
  }
}
static int bn_blocks(long long M, int C) {
  int groups = (C + 7) / 8;
  int lpr = groups < 256 ? groups : 256;
  int rpb = 256 / lpr;
  long long b = (M + rpb - 1) / rpb;
  if (b > kSMs * 4) b = kSMs * 4;
  if (b < 1) b = 1;
  return (int)b;
}
void bn_train_stats(const void* x, long long M, int C, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    float* save_mean, float* save_invstd, float* scale_shift, float* ws,
                    cudaStream_t s) {
  k_bn_stats<<<bn_blocks(M, C), 256, 256 * 8 * sizeof(float), s>>>(
      (const bf16*)x, M, C, gamma, beta, eps, momentum, running_mean, running_var, save_mean,
      save_invstd, scale_shift, ws);
  TNN_CHECK_LAUNCH();
}

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
template <bool RELU, bool DROP>
__global__ void __launch_bounds__(256)
    k_bn_apply(const bf16* __restrict__ x, bf16* __restrict__ y, long long M, int C,
               const float* __restrict__ ss, float p, float dscale, uint64_t seed) {
  const int groups = (C + 7) / 8;
  const long long total = M * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int g = (int)(i - r * groups);
    float f[8];
    load8(x, r, C, g, f);
    bool keep[8];
    if (DROP) dropout_keep8(seed, i, p, keep);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = min(g * 8 + j, C - 1);
      float v = f[j] * __ldg(ss + c) + __ldg(ss + C + c);
      if (RELU) v = fmaxf(v, 0.f);
      if (DROP) v = keep[j] ? v * dscale : 0.f;
      f[j] = v;
    }
    store8(y, r, C, g, f);
  }
}
void bn_apply(const void* x, void* y, long long M, int C, const float* scale_shift, int relu,
              float drop_p, uint64_t seed, cudaStream_t s) {
  long long total = M * ((C + 7) / 8);
  int g = grid_for(total, 256);
  float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const bf16* xi = (const bf16*)x;
  bf16* yo = (bf16*)y;
  if (relu && drop_p > 0.f) k_bn_apply<true, true><<<g, 256, 0, s>>>(xi, yo, M, C, scale_shift, drop_p, ds, seed);
  else if (relu) k_bn_apply<true, false><<<g, 256, 0, s>>>(xi, yo, M, C, scale_shift, drop_p, ds, seed);
  else if (drop_p > 0.f) k_bn_apply<false, true><<<g, 256, 0, s>>>(xi, yo, M, C, scale_shift, drop_p, ds, seed);
  else k_bn_apply<false, false><<<g, 256, 0, s>>>(xi, yo, M, C, scale_shift, drop_p, ds, seed);
  TNN_CHECK_LAUNCH();
}

// ------------------------------------------------------------------ BN backward
__global__ void __launch_bounds__(256)
    k_bn_bwd_reduce(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ y,
                    long long M, int C, const float* __restrict__ gamma,
                    const float* __restrict__ mean, const float* __restrict__ invstd, int masked,
                    float dscale, float* dgamma, float* dbeta, float* coef, float* ws) {
  extern __shared__ float sm[];
  channel_reduce<2>(M, C, ws, sm, [&](long long r, int g, float (&acc)[2][8]) {
    float d[8], xv[8], yv[8];
    load8(dy, r, C, g, d);
    load8(x, r, C, g, xv);
    if (masked) load8(y, r, C, g, yv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = min(g * 8 + j, C - 1);
      float gg = d[j];
      if (masked) gg = (yv[j] != 0.f) ? gg * dscale : 0.f;
      float xh = (xv[j] - __ldg(mean + c)) * __ldg(invstd + c);
      acc[0][j] += gg;
      acc[1][j] += gg * xh;
    }
  });
  if (last_block_done(reinterpret_cast<unsigned int*>(ws + 2 * C))) {
    const float inv_m = 1.f / (float)M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float db = __ldcg(ws + c), dg = __ldcg(ws + C + c);
      if (dbeta) dbeta[c] += db;
      if (dgamma) dgamma[c] += dg;
      float gm = gamma ? gamma[c] : 1.f;
      coef[c] = gm * invstd[c];
      coef[C + c] = db * inv_m;
      coef[2 * C + c] = dg * inv_m;
      ws[c] = 0.f;
      ws[C + c] = 0.f;
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned int*>(ws + 2 * C) = 0u;
  }
}
void bn_bwd_reduce(const void* dy, const void* x, const void* y, long long M, int C,
                   const float* gamma, const float* save_mean, const float* save_invstd,
                   int masked, float drop_p, float* dgamma, float* dbeta, float* coef, float* ws,
                   cudaStream_t s) {
  float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  k_bn_bwd_reduce<<<bn_blocks(M, C), 256, 256 * 8 * sizeof(float), s>>>(
      (const bf16*)dy, (const bf16*)x, (const bf16*)y, M, C, gamma, save_mean, save_invstd, masked,
      ds, dgamma, dbeta, coef, ws);
  TNN_CHECK_LAUNCH();
}

__global__ void __launch_bounds__(256)
    k_bn_bwd_apply(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ y,
                   const bf16* __restrict__ extra, bf16* __restrict__ dx, long long M, int C,
                   const float* __restrict__ mean, const float* __restrict__ invstd,
                   const float* __restrict__ coef, int masked, float dscale) {
  const int groups = (C + 7) / 8;
  const long long total = M * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int g = (int)(i - r * groups);
    float d[8], xv[8], yv[8], ev[8];
    load8(dy, r, C, g, d);
    load8(x, r, C, g, xv);
    if (masked) load8(y, r, C, g, yv);
    if (extra) load8(extra, r, C, g, ev);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = min(g * 8 + j, C - 1);
      float gg = d[j];
      if (masked) gg = (yv[j] != 0.f) ? gg * dscale : 0.f;
      float xh = (xv[j] - __ldg(mean + c)) * __ldg(invstd + c);
      float v = __ldg(coef + c) * (gg - __ldg(coef + C + c) - xh * __ldg(coef + 2 * C + c));
      if (extra) v += ev[j];
      d[j] = v;
    }
    store8(dx, r, C, g, d);
  }
}
void bn_bwd_apply(const void* dy, const void* x, const void* y, const void* extra, void* dx,
                  long long M, int C, const float* save_mean, const float* save_invstd,
                  const float* coef, int masked, float drop_p, cudaStream_t s) {
  long long total = M * ((C + 7) / 8);
  float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  k_bn_bwd_apply<<<grid_for(total, 256), 256, 0, s>>>((const bf16*)dy, (const bf16*)x, (const bf16*)y,
                                                      (const bf16*)extra, (bf16*)dx, M, C, save_mean,
                                                      save_invstd, coef, masked, ds);
  TNN_CHECK_LAUNCH();
}

// ------------------------------------------------------------------ LayerNorm (warp per row)
__global__ void __launch_bounds__(256)
    k_ln_fwd(const bf16* __restrict__ x, const bf16* __restrict__ res, bf16* __restrict__ y,
             bf16* __restrict__ sum_out, long long rows, int C, const float* __restrict__ gamma,
             const float* __restrict__ beta, float eps, float* mean, float* rstd) {
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
void layernorm_fwd(const void* x, const void* residual, void* y, void* sum_out, long long rows, int C,
                   const float* gamma, const float* beta, float eps, float* mean, float* rstd,
                   cudaStream_t s) {
  if (rows <= 0) return;
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
void layernorm_bwd(const void* dy, const void* x, long long rows, int C, const float* gamma,
                   const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                   cudaStream_t s) {
  if (rows <= 0) return;
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
