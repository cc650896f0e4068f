// NHWC pooling.  Max-pool stores an 8-bit window offset (kh*KW+kw) instead of the
// reference's int32 argmax index (src/nn/layers_impl/cuda/maxpool_ops.cu:19-94) and the
// backward is a gather over the (few) windows covering each input pixel, so no zero-fill
// pass and no atomics are needed.
#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

struct PoolP {
  int N, H, W, C8, kh, kw, sh, sw, ph, pw, OH, OW;
};

__global__ void __launch_bounds__(256)
    k_maxpool_fwd(const bf16* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ mask, PoolP p) {
  const long long total = (long long)p.N * p.OH * p.OW * p.C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % p.C8);
    long long t = i / p.C8;
    int ow = (int)(t % p.OW);
    t /= p.OW;
    int oh = (int)(t % p.OH);
    int n = (int)(t / p.OH);
    float best[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      arg[j] = 0;
    }
    for (int a = 0; a < p.kh; ++a) {
      int h = oh * p.sh - p.ph + a;
      if (h < 0 || h >= p.H) continue;
      for (int b = 0; b < p.kw; ++b) {
        int w = ow * p.sw - p.pw + b;
        if (w < 0 || w >= p.W) continue;
        float f[8];
        bf8_unpack(reinterpret_cast<const uint4*>(x)[(((long long)n * p.H + h) * p.W + w) * p.C8 + c], f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > best[j]) {
            best[j] = f[j];
            arg[j] = a * p.kw + b;
          }
      }
    }
    reinterpret_cast<uint4*>(y)[i] = bf8_pack(best);
    if (mask) {
      uint2 m;
      m.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      m.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      reinterpret_cast<uint2*>(mask)[i] = m;
    }
  }
}
__global__ void __launch_bounds__(256)
    k_maxpool_bwd(const bf16* __restrict__ dy, const uint8_t* __restrict__ mask, bf16* __restrict__ dx,
                  PoolP p) {
  const long long total = (long long)p.N * p.H * p.W * p.C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % p.C8);
    long long t = i / p.C8;
    int w = (int)(t % p.W);
    t /= p.W;
    int h = (int)(t % p.H);
    int n = (int)(t / p.H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // output windows that contain (h, w): oh*sh - ph + a == h
    for (int a = 0; a < p.kh; ++a) {
      int hh = h + p.ph - a;
      if (hh < 0 || hh % p.sh) continue;
      int oh = hh / p.sh;
      if (oh >= p.OH) continue;
      for (int b = 0; b < p.kw; ++b) {
        int ww = w + p.pw - b;
        if (ww < 0 || ww % p.sw) continue;
        int ow = ww / p.sw;
        if (ow >= p.OW) continue;
        long long o = (((long long)n * p.OH + oh) * p.OW + ow) * p.C8 + c;
        uint2 m = reinterpret_cast<const uint2*>(mask)[o];
        float g[8];
        bf8_unpack(reinterpret_cast<const uint4*>(dy)[o], g);
        const int code = a * p.kw + b;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int mj = ((j < 4 ? m.x : m.y) >> ((j & 3) * 8)) & 0xff;
          if (mj == code) acc[j] += g[j];
        }
      }
    }
    reinterpret_cast<uint4*>(dx)[i] = bf8_pack(acc);
  }
}
__global__ void __launch_bounds__(256)
    k_avgpool_fwd(const bf16* __restrict__ x, bf16* __restrict__ y, PoolP p) {
  const long long total = (long long)p.N * p.OH * p.OW * p.C8;
  const float inv = 1.f / (float)(p.kh * p.kw);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % p.C8);
    long long t = i / p.C8;
    int ow = (int)(t % p.OW);
    t /= p.OW;
    int oh = (int)(t % p.OH);
    int n = (int)(t / p.OH);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (p.ph == 0 && p.pw == 0 && oh * p.sh + p.kh <= p.H && ow * p.sw + p.kw <= p.W && p.kw == p.W) {
      // full-width windows without padding (global average pooling): one contiguous run of kh * W pixels, 8 loads in
      // flight instead of a serial bounds-checked walk (WRN 8x8x512 head: 25.6 us -> a few us)
      const uint4* src = reinterpret_cast<const uint4*>(x) + (((long long)n * p.H + (long long)oh * p.sh) * p.W) * p.C8 + c;
      const int cnt = p.kh * p.W;
      int q = 0;
      for (; q + 8 <= cnt; q += 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[(long long)(q + u) * p.C8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float f[8];
          bf8_unpack(v[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
      }
      for (; q < cnt; ++q) {
        float f[8];
        bf8_unpack(src[(long long)q * p.C8], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    } else
    for (int a = 0; a < p.kh; ++a) {
      int h = oh * p.sh - p.ph + a;
      if (h < 0 || h >= p.H) continue;
      for (int b = 0; b < p.kw; ++b) {
        int w = ow * p.sw - p.pw + b;
        if (w < 0 || w >= p.W) continue;
        float f[8];
        bf8_unpack(reinterpret_cast<const uint4*>(x)[(((long long)n * p.H + h) * p.W + w) * p.C8 + c], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    reinterpret_cast<uint4*>(y)[i] = bf8_pack(acc);
  }
}
__global__ void __launch_bounds__(256)
    k_avgpool_bwd(const bf16* __restrict__ dy, bf16* __restrict__ dx, PoolP p) {
  const long long total = (long long)p.N * p.H * p.W * p.C8;
  const float inv = 1.f / (float)(p.kh * p.kw);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % p.C8);
    long long t = i / p.C8;
    int w = (int)(t % p.W);
    t /= p.W;
    int h = (int)(t % p.H);
    int n = (int)(t / p.H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (p.ph == 0 && p.pw == 0 && p.sh >= p.kh && p.sw >= p.kw) {
      // non-overlapping windows (incl. global pooling): a pixel belongs to at most ONE window -- no 2-D search
      const int oh = h / p.sh, ow = w / p.sw;
      if (oh < p.OH && ow < p.OW && h - oh * p.sh < p.kh && w - ow * p.sw < p.kw)
        bf8_unpack(reinterpret_cast<const uint4*>(dy)[(((long long)n * p.OH + oh) * p.OW + ow) * p.C8 + c], acc);
    } else
    for (int a = 0; a < p.kh; ++a) {
      int hh = h + p.ph - a;
      if (hh < 0 || hh % p.sh) continue;
      int oh = hh / p.sh;
      if (oh >= p.OH) continue;
      for (int b = 0; b < p.kw; ++b) {
        int ww = w + p.pw - b;
        if (ww < 0 || ww % p.sw) continue;
        int ow = ww / p.sw;
        if (ow >= p.OW) continue;
        float g[8];
        bf8_unpack(reinterpret_cast<const uint4*>(dy)[(((long long)n * p.OH + oh) * p.OW + ow) * p.C8 + c], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    reinterpret_cast<uint4*>(dx)[i] = bf8_pack(acc);
  }
}

static PoolP mk(int N, int H, int W, int C, int kh, int kw, int sh, int sw, int ph, int pw, int OH, int OW) {
  if (C % 8 != 0) {
    fprintf(stderr, "[tnn_b200] pooling kernels need C %% 8 == 0 (got %d)\n", C);
    abort();
  }
  return PoolP{N, H, W, C / 8, kh, kw, sh, sw, ph, pw, OH, OW};
}
void maxpool_fwd(const void* x, void* y, uint8_t* mask, int N, int H, int W, int C, int kh, int kw,
                 int sh, int sw, int ph, int pw, int OH, int OW, cudaStream_t s) {
  PoolP p = mk(N, H, W, C, kh, kw, sh, sw, ph, pw, OH, OW);
  k_maxpool_fwd<<<grid_for((long long)N * OH * OW * p.C8, 256), 256, 0, s>>>((const bf16*)x, (bf16*)y, mask, p);
  TNN_CHECK_LAUNCH();
}
void maxpool_bwd(const void* dy, const uint8_t* mask, void* dx, int N, int H, int W, int C, int kh,
                 int kw, int sh, int sw, int ph, int pw, int OH, int OW, cudaStream_t s) {
  PoolP p = mk(N, H, W, C, kh, kw, sh, sw, ph, pw, OH, OW);
  k_maxpool_bwd<<<grid_for((long long)N * H * W * p.C8, 256), 256, 0, s>>>((const bf16*)dy, mask, (bf16*)dx, p);
  TNN_CHECK_LAUNCH();
}
void avgpool_fwd(const void* x, void* y, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                 int ph, int pw, int OH, int OW, cudaStream_t s) {
  PoolP p = mk(N, H, W, C, kh, kw, sh, sw, ph, pw, OH, OW);
  k_avgpool_fwd<<<grid_for((long long)N * OH * OW * p.C8, 256), 256, 0, s>>>((const bf16*)x, (bf16*)y, p);
  TNN_CHECK_LAUNCH();
}
void avgpool_bwd(const void* dy, void* dx, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                 int ph, int pw, int OH, int OW, cudaStream_t s) {
  PoolP p = mk(N, H, W, C, kh, kw, sh, sw, ph, pw, OH, OW);
  k_avgpool_bwd<<<grid_for((long long)N * H * W * p.C8, 256), 256, 0, s>>>((const bf16*)dy, (bf16*)dx, p);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
