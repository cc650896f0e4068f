// GEMMs for the element types beside bf16: the reference's layers are templated on fp16 / fp32 / fp64 IO
// (include/type/type.hpp, src/math/cuda/cublas_gemm.cu:20-164 picks cublasHgemm / cublasSgemm / cublasDgemm /
// cublasGemmEx per dtype); here each type gets the tensor-core mode Blackwell has for it:
//   * fp16  -> tcgen05.mma kind::f16 with fp16 operands (a_format = b_format = 0), fp32 accumulation in TMEM,
//   * fp32  -> tcgen05.mma kind::tf32: fp32 tensors go through TMA untouched, the tensor core reads the upper 19 bits
//              (what cuBLAS calls CUBLAS_COMPUTE_32F_FAST_TF32), fp32 accumulation and fp32 output; K-major operands
//              (an MN-major one is transposed into a stream-ordered scratch buffer first),
//   * fp64  -> a register-tiled DFMA kernel (sm_100a has no fp64 tensor-core path worth the name).
// One warp-specialised kernel template serves fp16 and tf32: warp 0 = TMA producer (128-byte swizzled boxes, K-major or
// MN-major operands), warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue (tcgen05.ld 32x32b, alpha / bias / residual /
// activation / accumulate, 16-byte stores).  It is the plain library GEMM of those dtypes, not a tuned hot path: the
// bf16 / MXFP8 kernels (umma_gemm.cu, mxfp8.cu) are the ones the flagship models run on.
#include "gemm_xt.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cudaTypedefs.h>
#include <math.h>

#include <mutex>

#include "common.cuh"
#include "sm100.cuh"

#define TNN_XT_CUDA(call)                                                                                  \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) {                                                                               \
      fprintf(stderr, "[tnn_b200] gemm_xt: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
      abort();                                                                                             \
    }                                                                                                      \
  } while (0)

namespace tnn {
using namespace sm100;

namespace {

constexpr int XT_BM = 128;
constexpr int XT_THREADS = 192;

template <typename T>
struct XtTraits;
template <>
struct XtTraits<__half> {
  static constexpr int BK = 64;          // elements per 128-byte swizzle row
  static constexpr int UK = 16;          // K per tcgen05.mma (32 bytes)
  static constexpr int MN_ATOM = 64;     // MN-major: elements per 128-byte row
  static constexpr uint32_t FMT = 0u;    // instruction-descriptor operand format: f16
  static constexpr bool TF32 = false;
  static constexpr CUtensorMapDataType TM = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
};
template <>
struct XtTraits<float> {
  static constexpr int BK = 32;
  static constexpr int UK = 8;
  static constexpr int MN_ATOM = 32;
  static constexpr uint32_t FMT = 2u;    // tf32
  static constexpr bool TF32 = true;
  static constexpr CUtensorMapDataType TM = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};

// [4,6) c_format = f32, [7,10) a_format, [10,13) b_format, [15] a_major, [16] b_major, [17,23) N >> 3, [24,29) M >> 4
__host__ __device__ constexpr uint32_t xt_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct XtParams {
  int M, N, K, kb_total;
  XtEpilogue e;
};

template <typename T>
__device__ __forceinline__ float xt_load(const T* p);
template <>
__device__ __forceinline__ float xt_load<__half>(const __half* p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float xt_load<float>(const float* p) { return *p; }

template <typename T, int BLOCK_N, bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(XT_THREADS)
    k_gemm_xt(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const XtParams p) {
  using TR = XtTraits<T>;
  constexpr int BK = TR::BK, UK = TR::UK, ATOM = TR::MN_ATOM;
  constexpr int A_BYTES = XT_BM * 128;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ATOM_BYTES = BK * 128;             // one MN-major atom: BK k-rows of 128 bytes
  constexpr int KSTEP_MN = UK * 128;               // MN-major: UK k-rows per MMA
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;

  extern __shared__ uint8_t xt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xt_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * XT_BM, n0 = blockIdx.y * BLOCK_N;
  const int nkb = p.kb_total;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        const int kB = i * BK;
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < XT_BM / ATOM; ++j) tma_load_2d(sa + j * ATOM_BYTES, &tmA, &full[s], m0 + j * ATOM, kB);
        } else {
          tma_load_2d(sa, &tmA, &full[s], kB, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / ATOM; ++j) tma_load_2d(sb + j * ATOM_BYTES, &tmB, &full[s], n0 + j * ATOM, kB);
        } else {
          tma_load_2d(sb, &tmB, &full[s], kB, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = xt_idesc(TR::FMT, XT_BM, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          const uint64_t ad = A_MN ? make_smem_desc_sw128(a_addr + k * KSTEP_MN, ATOM_BYTES, 1024) : make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
          const uint64_t bd = B_MN ? make_smem_desc_sw128(b_addr + k * KSTEP_MN, ATOM_BYTES, 1024) : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
          if (TR::TF32) umma_tf32(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
          else umma_f16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const XtEpilogue& e = p.e;
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    const long long row_off = (long long)m * e.ldc;
    const bool out_f32 = TR::TF32 || e.out_f32;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      const int n = n0 + c;
      if (n >= p.N) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * e.alpha;
      if (e.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n + j < p.N) v[j] += __ldg(e.bias + n + j);
      }
      if (e.residual != nullptr) {
        const T* res = reinterpret_cast<const T*>(e.residual) + (long long)m * e.ldr + n;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n + j < p.N) v[j] += xt_load<T>(res + j);
      }
      if (e.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (e.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
      }
      const bool full_w = (n + 32 <= p.N);
      if (out_f32) {
        float* o = reinterpret_cast<float*>(e.out) + row_off + n;
        const bool vec = full_w && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
        if (vec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 t = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (e.accumulate) {
              const float4 old = reinterpret_cast<const float4*>(o)[j];
              t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
            }
            reinterpret_cast<float4*>(o)[j] = t;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < p.N) o[j] = e.accumulate ? o[j] + v[j] : v[j];
        }
      } else {
        __half* o = reinterpret_cast<__half*>(e.out) + row_off + n;
        const bool vec = full_w && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
        if (vec) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 qv;
            __half2* h = reinterpret_cast<__half2*>(&qv);
            if (e.accumulate) {
              const uint4 old = reinterpret_cast<const uint4*>(o)[j];
              const __half2* oh = reinterpret_cast<const __half2*>(&old);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __half22float2(oh[t]);
                h[t] = __floats2half2_rn(v[8 * j + 2 * t] + f.x, v[8 * j + 2 * t + 1] + f.y);
              }
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
            }
            reinterpret_cast<uint4*>(o)[j] = qv;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < p.N) o[j] = __float2half_rn(e.accumulate ? __half2float(o[j]) + v[j] : v[j]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

PFN_cuTensorMapEncodeTiled_v12000 xt_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || ptr == nullptr || qres != cudaDriverEntryPointSuccess) {
      fprintf(stderr, "[tnn_b200] gemm_xt: cuTensorMapEncodeTiled unavailable\n");
      abort();
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// operand [outer, inner] with `ld` elements between outer rows; box {box_inner, box_outer}, 128-byte swizzle
template <typename T>
CUtensorMap xt_map(const void* ptr, long long inner, long long outer, long long ld, uint32_t box_inner, uint32_t box_outer) {
  CUtensorMap m;
  cuuint64_t gdims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld * (long long)sizeof(T))};
  cuuint32_t bx[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = xt_encode_fn()(&m, XtTraits<T>::TM, 2, const_cast<void*>(ptr), gdims, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[tnn_b200] gemm_xt tensor map failed (%d): inner=%lld outer=%lld ld=%lld box=%ux%u ptr=%p\n", (int)r, inner, outer, ld,
            box_inner, box_outer, ptr);
    abort();
  }
  return m;
}

template <typename T, int BLOCK_N, bool A_MN, bool B_MN>
void xt_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const XtParams& p, cudaStream_t stream) {
  constexpr int STAGES = 4;
  constexpr int smem = STAGES * (XT_BM * 128 + BLOCK_N * 128) + 1024 + 256;
  auto kern = k_gemm_xt<T, BLOCK_N, A_MN, B_MN, STAGES>;
  static PerDeviceOnce configured;
  if (configured.begin()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      fprintf(stderr, "[tnn_b200] gemm_xt smem attribute: %s\n", cudaGetErrorString(e));
      abort();
    }
    configured.done();
  }
  dim3 grid((p.M + XT_BM - 1) / XT_BM, (p.N + BLOCK_N - 1) / BLOCK_N, 1);
  kern<<<grid, XT_THREADS, smem, stream>>>(tmA, tmB, p);
  TNN_CHECK_LAUNCH();
}

template <typename T>
void xt_gemm(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb, bool a_mn, bool b_mn, const XtEpilogue& epi,
             cudaStream_t stream) {
  using TR = XtTraits<T>;
  constexpr int ALIGN = 16 / (int)sizeof(T);
  if (M <= 0 || N <= 0) return;
  if (K <= 0 || epi.out == nullptr || lda % ALIGN != 0 || ldb % ALIGN != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(B) & 15) != 0) {
    fprintf(stderr, "[tnn_b200] gemm_xt: needs K > 0, 16-byte aligned operands and row strides that are multiples of %d elements\n", ALIGN);
    abort();
  }
  XtParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.kb_total = (K + TR::BK - 1) / TR::BK;
  p.e = epi;
  if (p.e.ldr == 0) p.e.ldr = p.e.ldc;
  const bool wide = N > 64;
  const uint32_t bn = wide ? 128u : 64u;
  // K-major operand [rows, K]: dims {K, rows}, box {BK, rows of the tile}; MN-major operand [K, rows]: dims {rows, K}, box {atom, BK}
  const CUtensorMap tmA = a_mn ? xt_map<T>(A, M, K, lda, TR::MN_ATOM, TR::BK) : xt_map<T>(A, K, M, lda, TR::BK, XT_BM);
  const CUtensorMap tmB = b_mn ? xt_map<T>(B, N, K, ldb, TR::MN_ATOM, TR::BK) : xt_map<T>(B, K, N, ldb, TR::BK, bn);
#define XT_GO(BN_, AM_, BM_) xt_launch<T, BN_, AM_, BM_>(tmA, tmB, p, stream)
  if constexpr (TR::TF32) {
    if (a_mn || b_mn) {
      fprintf(stderr, "[tnn_b200] gemm_xt: the tf32 kernel takes K-major operands (gemm_tf32 transposes MN-major ones)\n");
      abort();
    }
    if (wide) XT_GO(128, false, false); else XT_GO(64, false, false);
  } else if (wide) {
    if (a_mn) { if (b_mn) XT_GO(128, true, true); else XT_GO(128, true, false); }
    else { if (b_mn) XT_GO(128, false, true); else XT_GO(128, false, false); }
  } else {
    if (a_mn) { if (b_mn) XT_GO(64, true, true); else XT_GO(64, true, false); }
    else { if (b_mn) XT_GO(64, false, true); else XT_GO(64, false, false); }
  }
#undef XT_GO
}

// ------------------------------------------------------------------------------------------------ fp64
// C[m, n] = alpha * sum_k A(m, k) * B(n, k) (+ bias[n]) (+ C when accumulate); A(m, k) = A[m * sam + k * sak] (either major).
// 64 x 64 tile, 256 threads x (4 x 4) outputs, K in slabs of 16 staged k-major in shared memory.
constexpr int D_TILE = 64, D_BK = 16;
__global__ void __launch_bounds__(256)
    k_gemm_f64(const double* __restrict__ A, const double* __restrict__ B, int M, int N, int K, long long sam, long long sak, long long sbn,
               long long sbk, double alpha, const double* __restrict__ bias, double* __restrict__ C, long long ldc, int accumulate) {
  __shared__ double sA[D_BK][D_TILE + 1];
  __shared__ double sB[D_BK][D_TILE + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * D_TILE, n0 = blockIdx.y * D_TILE;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += D_BK) {
    for (int i = threadIdx.x; i < D_TILE * D_BK; i += 256) {
      // the contiguous dimension of the operand is the fast index of the staging loop
      int r, k;
      if (sak == 1) { k = i % D_BK; r = i / D_BK; } else { r = i % D_TILE; k = i / D_TILE; }
      const int m = m0 + r, kk = k0 + k;
      sA[k][r] = (m < M && kk < K) ? A[m * sam + kk * sak] : 0.0;
    }
    for (int i = threadIdx.x; i < D_TILE * D_BK; i += 256) {
      int r, k;
      if (sbk == 1) { k = i % D_BK; r = i / D_BK; } else { r = i % D_TILE; k = i / D_TILE; }
      const int n = n0 + r, kk = k0 + k;
      sB[k][r] = (n < N && kk < K) ? B[n * sbn + kk * sbk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < D_BK; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      double v = alpha * acc[i][j];
      if (bias != nullptr) v += bias[n];
      double* o = C + m * ldc + n;
      *o = accumulate ? *o + v : v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ casts
template <typename T>
__device__ __forceinline__ double cv_load(const void* p, long long i);
template <>
__device__ __forceinline__ double cv_load<bf16>(const void* p, long long i) { return (double)__bfloat162float(reinterpret_cast<const bf16*>(p)[i]); }
template <>
__device__ __forceinline__ double cv_load<float>(const void* p, long long i) { return (double)reinterpret_cast<const float*>(p)[i]; }
template <>
__device__ __forceinline__ double cv_load<__half>(const void* p, long long i) { return (double)__half2float(reinterpret_cast<const __half*>(p)[i]); }
template <>
__device__ __forceinline__ double cv_load<double>(const void* p, long long i) { return reinterpret_cast<const double*>(p)[i]; }

__device__ __forceinline__ double cv_get(const void* p, int dt, long long i) {
  switch (dt) {
    case XT_BF16: return cv_load<bf16>(p, i);
    case XT_F32: return cv_load<float>(p, i);
    case XT_F16: return cv_load<__half>(p, i);
    default: return cv_load<double>(p, i);
  }
}
__device__ __forceinline__ void cv_put(void* p, int dt, long long i, double v) {
  switch (dt) {
    case XT_BF16: reinterpret_cast<bf16*>(p)[i] = __float2bfloat16((float)v); break;
    case XT_F32: reinterpret_cast<float*>(p)[i] = (float)v; break;
    case XT_F16: reinterpret_cast<__half*>(p)[i] = __float2half_rn((float)v); break;
    default: reinterpret_cast<double*>(p)[i] = v; break;
  }
}
__global__ void __launch_bounds__(256) k_cast_any(const void* __restrict__ src, int sdt, void* __restrict__ dst, int ddt, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    cv_put(dst, ddt, i, cv_get(src, sdt, i));
}
// [rows, cols] -> [cols, rows], 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
    k_transpose_any(const void* __restrict__ src, void* __restrict__ dst, int dt, int rows, int cols, long long src_ld, long long dst_ld) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    if (r < rows && c < cols) tile[j][tx] = cv_get(src, dt, (long long)r * src_ld + c);
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < rows && c < cols) cv_put(dst, dt, (long long)c * dst_ld + r, tile[tx][j]);
  }
}

// out[n] (+)= alpha * sum_m x[m, n]: 32 columns per block, 8 row lanes, shared-memory fold.  out is fp32, or fp64 for fp64 input.
__global__ void __launch_bounds__(256)
    k_colsum_any(const void* __restrict__ x, int dt, long long rows, int cols, long long ld, double alpha, void* __restrict__ out, int accumulate) {
  __shared__ double part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double acc = 0.0;
  if (c < cols)
    for (long long r = ty; r < rows; r += 8) acc += cv_get(x, dt, r * ld + c);
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += part[j][tx];
    t *= alpha;
    if (dt == XT_F64) {
      double* o = reinterpret_cast<double*>(out) + c;
      *o = accumulate ? *o + t : t;
    } else {
      float* o = reinterpret_cast<float*>(out) + c;
      *o = accumulate ? *o + (float)t : (float)t;
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm, any dtype
// One warp per row (fp32 statistics, two-pass variance over registers-free strided loads): the fp16 / fp32 / fp64
// counterpart of the bf16 warp-per-row kernels in norm.cu.  y = (x (+ residual) - mean) * rstd * gamma + beta.
__global__ void __launch_bounds__(256)
    k_layernorm_fwd_any(const void* __restrict__ x, const void* __restrict__ residual, void* __restrict__ y, void* __restrict__ sum_out, int dt,
                        long long rows, int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    const long long base = r * C;
    double s = 0.0;
    for (int c = lane; c < C; c += 32) {
      double v = cv_get(x, dt, base + c);
      if (residual) {
        v += cv_get(residual, dt, base + c);
        if (sum_out) cv_put(sum_out, dt, base + c, v);
      }
      s += v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double mean = s / C;
    double q = 0.0;
    const void* src = (residual && sum_out) ? sum_out : x;        // the stored sum is what backward sees
    for (int c = lane; c < C; c += 32) {
      double v = cv_get(src, dt, base + c);
      if (residual && !sum_out) v += cv_get(residual, dt, base + c);
      q += (v - mean) * (v - mean);
    }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const double rstd = 1.0 / sqrt(q / C + (double)eps);
    if (lane == 0) {
      mean_out[r] = (float)mean;
      rstd_out[r] = (float)rstd;
    }
    for (int c = lane; c < C; c += 32) {
      double v = cv_get(src, dt, base + c);
      if (residual && !sum_out) v += cv_get(residual, dt, base + c);
      double o = (v - mean) * rstd;
      if (gamma) o = o * gamma[c] + (beta ? beta[c] : 0.f);
      cv_put(y, dt, base + c, o);
    }
  }
}
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; dgamma / dbeta through fp32 atomics (one per warp and column)
__global__ void __launch_bounds__(256)
    k_layernorm_bwd_any(const void* __restrict__ dy, const void* __restrict__ x, int dt, long long rows, int C, const float* __restrict__ gamma,
                        const float* __restrict__ mean, const float* __restrict__ rstd, void* __restrict__ dx, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, const void* __restrict__ extra) {
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8LL + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8LL) {
    const long long base = r * C;
    const double mu = mean[r], is = rstd[r];
    double sg = 0.0, sgx = 0.0;
    for (int c = lane; c < C; c += 32) {
      const double d = cv_get(dy, dt, base + c), xh = (cv_get(x, dt, base + c) - mu) * is;
      const double g = gamma ? d * gamma[c] : d;
      sg += g;
      sgx += g * xh;
      if (dgamma) atomicAdd(dgamma + c, (float)(d * xh));
      if (dbeta) atomicAdd(dbeta + c, (float)d);
    }
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const double mg = sg / C, mgx = sgx / C;
    for (int c = lane; c < C; c += 32) {
      const double d = cv_get(dy, dt, base + c), xh = (cv_get(x, dt, base + c) - mu) * is;
      const double g = gamma ? d * gamma[c] : d;
      double v = is * (g - mg - xh * mgx);
      if (extra) v += cv_get(extra, dt, base + c);
      cv_put(dx, dt, base + c, v);
    }
  }
}

}  // namespace

void layernorm_fwd_any(const void* x, const void* residual, void* y, void* sum_out, int dt, long long rows, int C, const float* gamma,
                       const float* beta, float eps, float* mean, float* rstd, cudaStream_t stream) {
  if (rows <= 0 || C <= 0) return;
  k_layernorm_fwd_any<<<grid_for((rows + 7) / 8, 1, kSMs * 8), 256, 0, stream>>>(x, residual, y, sum_out, dt, rows, C, gamma, beta, eps, mean, rstd);
  TNN_CHECK_LAUNCH();
}
void layernorm_bwd_any(const void* dy, const void* x, int dt, long long rows, int C, const float* gamma, const float* mean, const float* rstd,
                       void* dx, float* dgamma, float* dbeta, const void* extra, cudaStream_t stream) {
  if (rows <= 0 || C <= 0) return;
  k_layernorm_bwd_any<<<grid_for((rows + 7) / 8, 1, kSMs * 8), 256, 0, stream>>>(dy, x, dt, rows, C, gamma, mean, rstd, dx, dgamma, dbeta, extra);
  TNN_CHECK_LAUNCH();
}

void colsum_any(const void* x, int dt, long long rows, int cols, long long ld, double alpha, void* out, bool accumulate, cudaStream_t stream) {
  if (cols <= 0) return;
  k_colsum_any<<<(cols + 31) / 32, 256, 0, stream>>>(x, dt, rows, cols, ld, alpha, out, accumulate ? 1 : 0);
  TNN_CHECK_LAUNCH();
}

void gemm_f16(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb, bool a_mn, bool b_mn, const XtEpilogue& epi,
              cudaStream_t stream) {
  xt_gemm<__half>(A, B, M, N, K, lda, ldb, a_mn, b_mn, epi, stream);
}
void gemm_tf32(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb, bool a_mn, bool b_mn, const XtEpilogue& epi,
               cudaStream_t stream) {
  // The tf32 kernel takes K-major operands (the MN-major shared-memory descriptors that work for 16-bit elements did
  // not produce correct products for 32-bit ones on the B200: tests/test_dtypes_gpu.py history), so an MN-major
  // operand [K, rows] is first transposed into a stream-ordered scratch buffer [rows, K4] (K4 = K rounded to 4).
  if (M <= 0 || N <= 0) return;
  const long long K4 = (K + 3) / 4 * 4;
  float* ta = nullptr;
  float* tb = nullptr;
  if (a_mn) {
    TNN_XT_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ta), sizeof(float) * (size_t)M * K4, stream));
    tensor_transpose_any(A, ta, XT_F32, K, M, stream, lda, K4);
  }
  if (b_mn) {
    TNN_XT_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tb), sizeof(float) * (size_t)N * K4, stream));
    tensor_transpose_any(B, tb, XT_F32, K, N, stream, ldb, K4);
  }
  xt_gemm<float>(a_mn ? ta : A, b_mn ? tb : B, M, N, K, a_mn ? K4 : lda, b_mn ? K4 : ldb, false, false, epi, stream);
  if (ta) TNN_XT_CUDA(cudaFreeAsync(ta, stream));
  if (tb) TNN_XT_CUDA(cudaFreeAsync(tb, stream));
}
void gemm_f64(const double* A, const double* B, int M, int N, int K, long long lda, long long ldb, bool a_mn, bool b_mn, double alpha,
              const double* bias, double* C, long long ldc, bool accumulate, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((M + D_TILE - 1) / D_TILE, (N + D_TILE - 1) / D_TILE, 1);
  // K-major [rows, K]: (row stride ld, k stride 1); MN-major [K, rows]: (row stride 1, k stride ld)
  k_gemm_f64<<<grid, 256, 0, stream>>>(A, B, M, N, K, a_mn ? 1 : lda, a_mn ? lda : 1, b_mn ? 1 : ldb, b_mn ? ldb : 1, alpha, bias, C, ldc,
                                       accumulate ? 1 : 0);
  TNN_CHECK_LAUNCH();
}
void tensor_cast_any(const void* src, int src_dt, void* dst, int dst_dt, long long n, cudaStream_t stream) {
  if (n <= 0) return;
  k_cast_any<<<grid_for(n, 256), 256, 0, stream>>>(src, src_dt, dst, dst_dt, n);
  TNN_CHECK_LAUNCH();
}
void tensor_transpose_any(const void* src, void* dst, int dt, int rows, int cols, cudaStream_t stream, long long src_ld, long long dst_ld) {
  if (rows <= 0 || cols <= 0) return;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, 1);
  k_transpose_any<<<grid, 256, 0, stream>>>(src, dst, dt, rows, cols, src_ld > 0 ? src_ld : cols, dst_ld > 0 ? dst_ld : rows);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
