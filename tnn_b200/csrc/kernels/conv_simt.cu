// SIMT convolution helpers.
//  * direct first-layer convolution (C_in < 8 cannot be described by a 16-byte-aligned TMA
//    tensor map and has K = R*S*C_in <= 75, far below one UMMA tile)
//  * im2col / col2im for geometries outside the TMA window path (odd widths, 5x5/7x7,
//    stride > 2); the GEMM itself still runs on tcgen05.
#include <cuda_fp16.h>

#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

struct DConvP {
  int N, H, W, C, K, R, S, stride, pad, OH, OW;
  int stride_w, pad_w1;      // im2col / col2im only: width stride (0 = same as `stride`) and width pad + 1 (0 = same as `pad`)
};

// thread = (output pixel, group of 8 output channels)
__global__ void __launch_bounds__(256)
    k_conv_direct_fwd(const bf16* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, bf16* __restrict__ y, DConvP p) {
  extern __shared__ float sw[];  // [R*S*C][K]  (k fastest so a thread reads 8 consecutive)
  const int rsc = p.R * p.S * p.C;
  for (int i = threadIdx.x; i < rsc * p.K; i += blockDim.x) {
    int k = i / rsc, j = i - k * rsc;
    sw[j * p.K + k] = w[i];
  }
  __syncthreads();
  const int kg = p.K / 8;
  const long long total = (long long)p.N * p.OH * p.OW * kg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int g = (int)(i % kg);
    long long t = i / kg;
    int ow = (int)(t % p.OW);
    t /= p.OW;
    int oh = (int)(t % p.OH);
    int n = (int)(t / p.OH);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[g * 8 + j] : 0.f;
    for (int r = 0; r < p.R; ++r) {
      int h = oh * p.stride - p.pad + r;
      if (h < 0 || h >= p.H) continue;
      for (int s = 0; s < p.S; ++s) {
        int ww = ow * p.stride - p.pad + s;
        if (ww < 0 || ww >= p.W) continue;
        const bf16* xp = x + (((long long)n * p.H + h) * p.W + ww) * p.C;
        for (int c = 0; c < p.C; ++c) {
          float xv = __bfloat162float(xp[c]);
          const float* wp = sw + ((r * p.S + s) * p.C + c) * p.K + g * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += xv * wp[j];
        }
      }
    }
    reinterpret_cast<uint4*>(y)[i] = bf8_pack(acc);
  }
}
void conv_direct_fwd(const void* x, const float* w, const float* bias, void* y, int N, int H, int W,
                     int C, int K, int R, int S, int stride, int pad, int OH, int OW, cudaStream_t s) {
  if (K % 8 != 0) {
    fprintf(stderr, "[tnn_b200] conv_direct_fwd needs K %% 8 == 0\n");
    abort();
  }
  DConvP p{N, H, W, C, K, R, S, stride, pad, OH, OW};
  size_t smem = (size_t)R * S * C * K * sizeof(float);
  static std::atomic<size_t> configured[64];   // per device (the attribute is per device)
  std::atomic<size_t>& conf = configured[PerDeviceOnce::device()];
  if (smem > 48 * 1024 && smem > conf.load()) {
    cudaFuncSetAttribute(k_conv_direct_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conf.store(smem);
  }
  long long total = (long long)N * OH * OW * (K / 8);
  k_conv_direct_fwd<<<grid_for(total, 256), 256, smem, s>>>((const bf16*)x, w, bias, (bf16*)y, p);
  TNN_CHECK_LAUNCH();
}

// wgrad: block stages PIX pixels of dy and the matching input patches in smem, every thread
// owns outputs o = tid, tid+256, ... of dW[K][R*S*C]; block partials go out by atomicAdd.
static constexpr int DW_PIX = 32;
__global__ void __launch_bounds__(256)
    k_conv_direct_wgrad(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw,
                        float* __restrict__ dbias, DConvP p, int pix_per_block) {
  extern __shared__ float sm[];
  const int rsc = p.R * p.S * p.C;
  float* s_dy = sm;                   // [DW_PIX][K]
  float* s_x = sm + DW_PIX * p.K;     // [DW_PIX][rsc]
  const int nout = p.K * rsc;
  const int per = (nout + 255) / 256;
  float acc[8];                        // supports nout <= 2048
  float bacc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const long long npix = (long long)p.N * p.OH * p.OW;
  const long long p_begin = (long long)blockIdx.x * pix_per_block;
  const long long p_end = min(npix, p_begin + pix_per_block);
  for (long long p0 = p_begin; p0 < p_end; p0 += DW_PIX) {
    const int np = (int)min((long long)DW_PIX, p_end - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < DW_PIX * p.K; i += blockDim.x) {
      int pp = i / p.K, k = i - pp * p.K;
      s_dy[i] = pp < np ? __bfloat162float(dy[(p0 + pp) * p.K + k]) : 0.f;
    }
    for (int i = threadIdx.x; i < DW_PIX * rsc; i += blockDim.x) {
      int pp = i / rsc, j = i - pp * rsc;
      float v = 0.f;
      if (pp < np) {
        long long pix = p0 + pp;
        int ow = (int)(pix % p.OW);
        long long t = pix / p.OW;
        int oh = (int)(t % p.OH);
        int n = (int)(t / p.OH);
        int c = j % p.C, rs = j / p.C, s = rs % p.S, r = rs / p.S;
        int h = oh * p.stride - p.pad + r, ww = ow * p.stride - p.pad + s;
        if (h >= 0 && h < p.H && ww >= 0 && ww < p.W)
          v = __bfloat162float(x[(((long long)n * p.H + h) * p.W + ww) * p.C + c]);
      }
      s_x[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q >= per) break;
      int o = threadIdx.x + q * 256;
      if (o < nout) {
        int k = o / rsc, j = o - k * rsc;
        float a = 0.f;
        for (int pp = 0; pp < DW_PIX; ++pp) a += s_dy[pp * p.K + k] * s_x[pp * rsc + j];
        acc[q] += a;
      }
    }
    if (dbias && threadIdx.x < p.K) {
      for (int pp = 0; pp < DW_PIX; ++pp) bacc += s_dy[pp * p.K + threadIdx.x];
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    int o = threadIdx.x + q * 256;
    if (q < per && o < nout) atomicAdd(dw + o, acc[q]);
  }
  if (dbias && threadIdx.x < p.K) atomicAdd(dbias + threadIdx.x, bacc);
}
// Tensor-core variant for the first layer (R*S*C + 1 <= 32, K <= 64): dW^T[j][k] = sum_pix P[pix][j] * dy[pix][k]
// with P the im2col patch matrix (27 columns for 3x3x3) plus a column of ones whose product is the
// bias gradient.  A CTA stages 1024 pixels of P and dy TRANSPOSED in shared memory (bf16, the
// reduction dimension contiguous, rows padded by 16 bytes so fragment loads are conflict free); each of
// the 8 warps reduces its 128 pixels with mma.sync m16n8k16 (M = 32 patch columns, N = K output
// channels), warps combine through shared memory and the CTA adds its [K][rsc] block to dW with one
// atomic per element.  (The scalar kernel above spent ~110 us on the 3->16 WRN stem: two shared-memory
// loads per FMA and a divide per staged element.)
static constexpr int DT_PIX = 1024;
static constexpr int DT_LD = DT_PIX + 8;
__device__ __forceinline__ void dt_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int NT>   // NT = K / 8 output-channel tiles
__global__ void __launch_bounds__(256)
    k_conv_stem_wgrad(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw,
                      float* __restrict__ dbias, DConvP p) {
  extern __shared__ __align__(16) uint8_t dt_smem[];
  bf16* sP = reinterpret_cast<bf16*>(dt_smem);                 // [32][DT_LD]   patch^T (row 31.. rsc = ones / zero)
  bf16* sD = sP + 32 * DT_LD;                                  // [K][DT_LD]    dy^T
  float* sAcc = reinterpret_cast<float*>(sD + p.K * DT_LD);    // [32][K]
  const int rsc = p.R * p.S * p.C;
  const int K = p.K;
  const long long npix = (long long)p.N * p.OH * p.OW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) sAcc[i] = 0.f;
  float acc[2][NT][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

  for (long long p0 = (long long)blockIdx.x * DT_PIX; p0 < npix; p0 += (long long)gridDim.x * DT_PIX) {
    __syncthreads();
    // ---- stage patch^T: lanes walk pixels (conflict-free 2-byte stores), each warp owns patch columns
    //      j = warp, warp + 8, ...; the pixel is decoded once and reused for the warp's 4 columns
    {
      int jr[4], js[4], jc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = warp + 8 * q;
        jc[q] = j % p.C;
        const int rs = j / p.C;
        js[q] = rs % p.S;
        jr[q] = rs / p.S;
      }
      for (int pp = lane; pp < DT_PIX; pp += 32) {
        const long long pix = p0 + pp;
        const bool ok = pix < npix;
        int n = 0, oh = 0, ow = 0;
        if (ok) {
          ow = (int)(pix % p.OW);
          const long long t = pix / p.OW;
          oh = (int)(t % p.OH);
          n = (int)(t / p.OH);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = warp + 8 * q;
          float v = 0.f;
          if (ok) {
            if (j < rsc) {
              const int h = oh * p.stride - p.pad + jr[q], ww = ow * p.stride - p.pad + js[q];
              if (h >= 0 && h < p.H && ww >= 0 && ww < p.W)
                v = __bfloat162float(x[(((long long)n * p.H + h) * p.W + ww) * p.C + jc[q]]);
            } else if (j == rsc) {
              v = 1.f;                                // ones column: its product is the bias gradient
            }
          }
          sP[j * DT_LD + pp] = __float2bfloat16(v);
        }
      }
    }
    // ---- stage dy^T: one pixel per thread per pass, 16 bytes (8 channels) at a time
    for (int pp = threadIdx.x; pp < DT_PIX; pp += blockDim.x) {
      const long long pix = p0 + pp;
      for (int k8 = 0; k8 < K; k8 += 8) {
        uint4 q = make_uint4(0, 0, 0, 0);
        if (pix < npix) q = __ldg(reinterpret_cast<const uint4*>(dy + pix * K + k8));
        const bf16* h = reinterpret_cast<const bf16*>(&q);
#pragma unroll
        for (int t = 0; t < 8; ++t) sD[(k8 + t) * DT_LD + pp] = h[t];
      }
    }
    __syncthreads();
    // ---- this warp's 128 pixels: 8 k-steps of 16
    const int pw = warp * 128;
#pragma unroll 2
    for (int ks = 0; ks < 8; ++ks) {
      const int k0 = pw + ks * 16 + (lane & 3) * 2;
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const bf16* base = sP + (mt * 16 + (lane >> 2)) * DT_LD + k0;
        a[mt][0] = *reinterpret_cast<const uint32_t*>(base);
        a[mt][1] = *reinterpret_cast<const uint32_t*>(base + 8 * DT_LD);
        a[mt][2] = *reinterpret_cast<const uint32_t*>(base + 8);
        a[mt][3] = *reinterpret_cast<const uint32_t*>(base + 8 * DT_LD + 8);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const bf16* bb = sD + (nt * 8 + (lane >> 2)) * DT_LD + k0;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(bb);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(bb + 8);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) dt_mma(acc[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
      }
    }
  }
  // ---- combine the warps, then one atomic per dW element
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int row = mt * 16 + (lane >> 2), col = nt * 8 + (lane & 3) * 2;
      atomicAdd(&sAcc[row * K + col], acc[mt][nt][0]);
      atomicAdd(&sAcc[row * K + col + 1], acc[mt][nt][1]);
      atomicAdd(&sAcc[(row + 8) * K + col], acc[mt][nt][2]);
      atomicAdd(&sAcc[(row + 8) * K + col + 1], acc[mt][nt][3]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
    const int j = i / K, k = i - j * K;
    if (j < rsc) atomicAdd(dw + k * rsc + j, sAcc[i]);
    else if (j == rsc && dbias) atomicAdd(dbias + k, sAcc[i]);
  }
}

void conv_direct_wgrad(const void* x, const void* dy, float* dw, float* dbias, int N, int H, int W,
                       int C, int K, int R, int S, int stride, int pad, int OH, int OW,
                       cudaStream_t s) {
  const int rsc = R * S * C;
  {
    static int use_tc = -1;
    if (use_tc < 0) {
      const char* e = getenv("TNN_STEM_WGRAD_TC");
      use_tc = (e != nullptr) ? atoi(e) : 1;
    }
    const int NT = K / 8;
    if (use_tc && rsc + 1 <= 32 && K % 8 == 0 && (NT == 1 || NT == 2 || NT == 4 || NT == 8)) {
      DConvP p{N, H, W, C, K, R, S, stride, pad, OH, OW};
      const long long npix = (long long)N * OH * OW;
      const int blocks = (int)min((long long)kSMs * 2, (npix + DT_PIX - 1) / DT_PIX);
      const int smem = (32 + K) * DT_LD * 2 + 32 * K * 4;
#define TNN_STEM(NT_)                                                                                          \
  do {                                                                                                         \
    static PerDeviceOnce configured;                                                                            \
    if (configured.begin()) {                                                                                         \
      cudaFuncSetAttribute(k_conv_stem_wgrad<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   \
      configured.done();                                                                                       \
    }                                                                                                          \
    k_conv_stem_wgrad<NT_><<<blocks, 256, smem, s>>>((const bf16*)x, (const bf16*)dy, dw, dbias, p);           \
  } while (0)
      if (NT == 1) TNN_STEM(1);
      else if (NT == 2) TNN_STEM(2);
      else if (NT == 4) TNN_STEM(4);
      else TNN_STEM(8);
#undef TNN_STEM
      TNN_CHECK_LAUNCH();
      return;
    }
  }
  if (K * rsc > 2048 || K > 256) {
    fprintf(stderr, "[tnn_b200] conv_direct_wgrad: K*R*S*C too large (%d)\n", K * rsc);
    abort();
  }
  DConvP p{N, H, W, C, K, R, S, stride, pad, OH, OW};
  long long npix = (long long)N * OH * OW;
  int blocks = (int)min((long long)kSMs * 4, (npix + DW_PIX - 1) / DW_PIX);
  int ppb = (int)((npix + blocks - 1) / blocks);
  ppb = (ppb + DW_PIX - 1) / DW_PIX * DW_PIX;
  blocks = (int)((npix + ppb - 1) / ppb);
  size_t smem = (size_t)DW_PIX * (K + rsc) * sizeof(float);
  k_conv_direct_wgrad<<<blocks, 256, smem, s>>>((const bf16*)x, (const bf16*)dy, dw, dbias, p, ppb);
  TNN_CHECK_LAUNCH();
}

// --------------------------------------------------------------------------- im2col / col2im
// templated on the element type: bf16 (the conv paths outside the TMA-window kernels), fp16 and fp32 (convolutions with an
// explicit io dtype: im2col + the fp16 / tf32 tcgen05 GEMMs of gemm_xt.cu)
__device__ __forceinline__ float cs_get(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float cs_get(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ float cs_get(const float* p) { return *p; }
__device__ __forceinline__ void cs_put(bf16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void cs_put(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void cs_put(float* p, float v) { *p = v; }
template <typename T>
__global__ void __launch_bounds__(256)
    k_im2col(const T* __restrict__ x, T* __restrict__ col, long long ldcol, DConvP p) {
  const int rsc = p.R * p.S * p.C;
  const long long total = (long long)p.N * p.OH * p.OW * rsc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int j = (int)(i % rsc);
    long long pix = i / rsc;
    int ow = (int)(pix % p.OW);
    long long t = pix / p.OW;
    int oh = (int)(t % p.OH);
    int n = (int)(t / p.OH);
    int c = j % p.C, rs = j / p.C, s = rs % p.S, r = rs / p.S;
    const int sw_ = p.stride_w ? p.stride_w : p.stride, pw_ = p.pad_w1 ? p.pad_w1 - 1 : p.pad;
    int h = oh * p.stride - p.pad + r, w = ow * sw_ - pw_ + s;
    T v;
    cs_put(&v, 0.f);
    if (h >= 0 && h < p.H && w >= 0 && w < p.W) v = x[(((long long)n * p.H + h) * p.W + w) * p.C + c];
    col[pix * ldcol + j] = v;
  }
}
void im2col_nhwc(const void* x, void* col, long long ldcol, int N, int H, int W, int C, int R, int S,
                 int stride, int pad, int OH, int OW, cudaStream_t s) {
  im2col_nhwc2(x, col, ldcol, N, H, W, C, R, S, stride, stride, pad, pad, OH, OW, s);
}
void im2col_nhwc2(const void* x, void* col, long long ldcol, int N, int H, int W, int C, int R, int S, int sh, int sw,
                  int ph, int pw, int OH, int OW, cudaStream_t s, int dt) {
  DConvP p{N, H, W, C, 0, R, S, sh, ph, OH, OW, sw, pw + 1};
  long long total = (long long)N * OH * OW * R * S * C;
  if (dt == 1) k_im2col<float><<<grid_for(total, 256), 256, 0, s>>>((const float*)x, (float*)col, ldcol, p);
  else if (dt == 2) k_im2col<__half><<<grid_for(total, 256), 256, 0, s>>>((const __half*)x, (__half*)col, ldcol, p);
  else k_im2col<bf16><<<grid_for(total, 256), 256, 0, s>>>((const bf16*)x, (bf16*)col, ldcol, p);
  TNN_CHECK_LAUNCH();
}
// gather formulation: one thread per input element sums the col entries that map onto it
template <typename T>
__global__ void __launch_bounds__(256)
    k_col2im(const T* __restrict__ col, long long ldcol, T* __restrict__ dx, DConvP p) {
  const long long total = (long long)p.N * p.H * p.W * p.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % p.C);
    long long t = i / p.C;
    int w = (int)(t % p.W);
    t /= p.W;
    int h = (int)(t % p.H);
    int n = (int)(t / p.H);
    float acc = 0.f;
    const int sw_ = p.stride_w ? p.stride_w : p.stride, pw_ = p.pad_w1 ? p.pad_w1 - 1 : p.pad;
    for (int r = 0; r < p.R; ++r) {
      int hh = h + p.pad - r;
      if (hh < 0 || hh % p.stride) continue;
      int oh = hh / p.stride;
      if (oh >= p.OH) continue;
      for (int s = 0; s < p.S; ++s) {
        int ww = w + pw_ - s;
        if (ww < 0 || ww % sw_) continue;
        int ow = ww / sw_;
        if (ow >= p.OW) continue;
        long long pix = ((long long)n * p.OH + oh) * p.OW + ow;
        acc += cs_get(col + pix * ldcol + (r * p.S + s) * p.C + c);
      }
    }
    cs_put(dx + i, acc);
  }
}
void col2im_nhwc(const void* col, long long ldcol, void* dx, int N, int H, int W, int C, int R, int S,
                 int stride, int pad, int OH, int OW, cudaStream_t s) {
  col2im_nhwc2(col, ldcol, dx, N, H, W, C, R, S, stride, stride, pad, pad, OH, OW, s);
}
void col2im_nhwc2(const void* col, long long ldcol, void* dx, int N, int H, int W, int C, int R, int S, int sh, int sw,
                  int ph, int pw, int OH, int OW, cudaStream_t s, int dt) {
  DConvP p{N, H, W, C, 0, R, S, sh, ph, OH, OW, sw, pw + 1};
  long long total = (long long)N * H * W * C;
  if (dt == 1) k_col2im<float><<<grid_for(total, 256), 256, 0, s>>>((const float*)col, ldcol, (float*)dx, p);
  else if (dt == 2) k_col2im<__half><<<grid_for(total, 256), 256, 0, s>>>((const __half*)col, ldcol, (__half*)dx, p);
  else k_col2im<bf16><<<grid_for(total, 256), 256, 0, s>>>((const bf16*)col, ldcol, (bf16*)dx, p);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
