// tcgen05 GEMM + implicit-GEMM convolution kernels for sm_100a.
//
// One warp-specialised CTA computes a 128 x BLOCK_N fp32 tile in TMEM:
//   warp 0      : TMA producer (one elected lane) -> 128B-swizzled smem ring
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  : epilogue, tcgen05.ld TMEM -> registers -> fused epilogue -> global
// Two CTAs are co-resident per SM (<= 256 TMEM columns, <= ~100 KB smem each) so one
// CTA's epilogue overlaps the other's main loop without a persistent scheduler.
//
// The same main loop serves
//   * Dense fwd / dgrad / wgrad  (any K-major / MN-major operand combination),
//   * Conv2d fprop and dgrad as implicit GEMM: the activation operand is fetched by
//     4-D TMA boxes shifted per filter tap, the zero halo comes from TMA OOB fill, no
//     im2col buffer exists (reference: im2col + cuBLAS, src/nn/layers_impl/
//     legacy_conv2d_layer.cpp:137-227, or cuDNN graphs, cuda/cudnn_conv2d_ops.cu),
//   * Conv2d wgrad (second kernel): both operands MN-major, K = pixels, one TMEM
//     accumulator per filter tap, split-K over pixel blocks with fp32 red.add
//     (beta = 1 accumulate semantics of the reference, conv2d_nchw_ops.cu:85-87).
#include "umma_gemm.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "sm100.cuh"

namespace tnn {
using namespace sm100;
extern std::atomic<unsigned long long> g_launch_count;

#define TNN_CUDA_CHECK(x)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "[tnn_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_),   \
              __FILE__, __LINE__);                                                     \
      abort();                                                                         \
    }                                                                                  \
  } while (0)

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
static constexpr int kThreads = 192;
static constexpr int ATOM_BYTES = 64 * 128;  // 64 rows x 128 B

struct KParams {
  int M, N, K;
  int kc_per_tap;
  int kb_total;
  int kb_per_split;
  int OH, OW;       // output grid (conv) for tile -> (n, h) mapping
  int BH, BN;       // TMA window box: {64, OW, BH, BN}
  int ntaps;
  int vec_ok;       // 16-byte vector stores legal
  ConvTap taps[kMaxTaps];
  Epilogue epi;
};

// ---------------------------------------------------------------------------
// epilogue helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long epi_row_offset(const Epilogue& e, int m) {
  if (e.grid_mode) {
    int hw = e.OH * e.OW;
    int n = m / hw;
    int r = m - n * hw;
    int oh = r / e.OW;
    int ow = r - oh * e.OW;
    return e.base + (long long)n * e.sN + (long long)oh * e.sH + (long long)ow * e.sW;
  }
  return e.base + (long long)m * e.ldc;
}

// Epilogue math for 32 consecutive columns [n, n+32) of one output row held in r[]: alpha, bias,
// residual, activation and (optionally) the fused BatchNorm column statistics.  Result in v[].
__device__ __forceinline__ void epi_compute_32(const KParams& p, const uint32_t (&r)[32], bool row_ok,
                                               long long row_off, int n, float (&v)[32]) {
  const Epilogue& e = p.epi;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * e.alpha;
  if (e.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < p.N) v[j] += __ldg(e.bias + n + j);
  }
  const bool full = (n + 32 <= p.N) && p.vec_ok;
  if (row_ok && e.residual != nullptr) {
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(e.residual) + row_off + n;
    if (full) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(res) + j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float2 f = __bfloat1622float2(h[t]);
          v[j * 8 + t * 2] += f.x;
          v[j * 8 + t * 2 + 1] += f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < p.N) v[j] += __bfloat162float(res[j]);
    }
  }
  if (e.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (e.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
  }
  float xh[32];
  const bool bnb = e.bnb_x != nullptr;
  if (bnb) {
    // BatchNorm backward fusion (see Epilogue): v is d(BN output) here
    const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(e.bnb_x) + row_off + n;
    float xv[32];
    if (row_ok && full) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(xs) + j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float2 f = __bfloat1622float2(h[t]);
          xv[j * 8 + t * 2] = f.x;
          xv[j * 8 + t * 2 + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) xv[j] = (row_ok && n + j < p.N) ? __bfloat162float(xs[j]) : 0.f;
    }
    unsigned long long seed = e.bnb_seed;
    if (e.bnb_drop > 0.f && e.bnb_seed_ptr != nullptr) seed ^= (*e.bnb_seed_ptr) * 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      bool keep[8];
      if (e.bnb_drop > 0.f) {
        // same element -> mask mapping as the forward apply kernel: vector index = element offset / 8
        dropout_keep8(seed, (row_off + n + j8 * 8) >> 3, e.bnb_drop, keep);
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int j = j8 * 8 + t;
        const float4 P = (n + j < p.N) ? __ldg(reinterpret_cast<const float4*>(e.bnb_pack) + n + j)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        bool on = e.bnb_relu ? (xv[j] * P.z + P.w > 0.f) : true;
        if (e.bnb_drop > 0.f) on = on && keep[t];
        v[j] = on ? v[j] * e.bnb_dscale : 0.f;
        xh[j] = xv[j] * P.x + P.y;
      }
    }
  }
  if (e.col_sum != nullptr) {
    // BatchNorm statistics of the STORED tile (sum and sum of squares per output channel), so the
    // consumer BatchNorm never re-reads this activation: a butterfly transpose-reduce over the
    // warp's 32 rows leaves column (n + lane) in lane `lane` after 31 shuffles per quantity, then
    // one fp32 atomic per lane into a replica row (same replica layout as norm.cu).
    float a[32], b[32];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float t = v[j];
      if (e.out_mode == OUT_BF16) t = __bfloat162float(__float2bfloat16(t));
      t = row_ok ? t : 0.f;
      a[j] = t;
      b[j] = bnb ? t * xh[j] : t * t;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < off; ++j) {
        const float sa = upper ? a[j] : a[j + off], ka = upper ? a[j + off] : a[j];
        const float sb = upper ? b[j] : b[j + off], kb = upper ? b[j + off] : b[j];
        a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
        b[j] = kb + __shfl_xor_sync(0xffffffffu, sb, off);
      }
    }
    if (n + lane < p.N) {
      float* ws = e.col_sum + (size_t)(blockIdx.x % e.col_repl) * 2 * p.N;
      atomicAdd(ws + n + lane, a[0]);
      atomicAdd(ws + p.N + n + lane, b[0]);
    }
  }
}

// Direct (per-thread) global store of the 32 values computed above.
__device__ __forceinline__ void epi_write_32(const KParams& p, const float (&v)[32], bool row_ok, long long row_off,
                                             int n) {
  const Epilogue& e = p.epi;
  const bool full = (n + 32 <= p.N) && p.vec_ok;
  if (!row_ok) return;
  if (e.out_mode == OUT_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + row_off + n;
    if (full) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q;
        q.x = pack_bf16(v[j * 8 + 0], v[j * 8 + 1]);
        q.y = pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
        q.z = pack_bf16(v[j * 8 + 4], v[j * 8 + 5]);
        q.w = pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
        reinterpret_cast<uint4*>(o)[j] = q;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < p.N) o[j] = __float2bfloat16(v[j]);
    }
  } else if (e.out_mode == OUT_F32) {
    float* o = reinterpret_cast<float*>(e.out) + row_off + n;
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(o)[j] = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < p.N) o[j] = v[j];
    }
  } else {  // OUT_F32_ATOMIC
    float* o = reinterpret_cast<float*>(e.out) + row_off + n;
    if ((n + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) red_add_v4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < p.N) atomicAdd(o + j, v[j]);
    }
  }
}

__device__ __forceinline__ void epi_store_32(const KParams& p, const uint32_t (&r)[32], bool row_ok,
                                             long long row_off, int n) {
  float v[32];
  epi_compute_32(p, r, row_ok, row_off, n, v);
  epi_write_32(p, v, row_ok, row_off, n);
}

// ---------------------------------------------------------------------------
// GEMM / conv-fprop kernel
// ---------------------------------------------------------------------------
template <int BLOCK_N, bool A_MN, bool B_MN, bool CONV_A, int STAGES>
__global__ void __launch_bounds__(kThreads)
    umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmB, const KParams p) {
  constexpr int A_BYTES = BLOCK_M * 128;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  static_assert(!(CONV_A && A_MN), "conv window operand is K-major");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n0 = blockIdx.y * BLOCK_N;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
  const int nkb = kb_end - kb_begin;

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
      // ----------------------------- TMA producer -----------------------------
      int img0 = 0, h0 = 0;
      if (CONV_A) {
        const int hw = p.OH * p.OW;
        if (p.BN > 1) {
          img0 = blockIdx.x * p.BN;
        } else {
          img0 = m0 / hw;
          h0 = (m0 - img0 * hw) / p.OW;
        }
      }
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb_begin + i;
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        int kB;
        if (CONV_A) {
          const int t = kb / p.kc_per_tap;
          const int kc = kb - t * p.kc_per_tap;
          const ConvTap tap = p.taps[t];
          tma_load_4d(sa, &tmA, &full[s], kc * BLOCK_K, tap.dw, h0 + tap.dh, img0 + tap.nb);
          kB = tap.koff + kc * BLOCK_K;
        } else {
          kB = kb * BLOCK_K;
          if (A_MN) {
            tma_load_2d(sa, &tmA, &full[s], m0, kB);
            tma_load_2d(sa + ATOM_BYTES, &tmA, &full[s], m0 + 64, kB);
          } else {
            tma_load_2d(sa, &tmA, &full[s], kB, m0);
          }
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(sb + j * ATOM_BYTES, &tmB, &full[s], n0 + j * 64, kB);
        } else {
          tma_load_2d(sb, &tmB, &full[s], kB, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ----------------------------- MMA issuer -----------------------------
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          const uint64_t ad = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, ATOM_BYTES, 1024)
                                   : make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
          const uint64_t bd = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, ATOM_BYTES, 1024)
                                   : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
          umma_f16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ----------------------------- epilogue -----------------------------
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const long long row_off = row_ok ? epi_row_offset(p.epi, m) : 0;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      if (n0 + c >= p.N) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      epi_store_32(p, r, row_ok, row_off, n0 + c);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------
// Persistent variant of the kernel above: grid = min(#tiles, 2 x #SMs); every CTA walks tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ... with
//   * the TMA producer running ahead across tile boundaries (one smem ring for all tiles),
//   * TWO accumulator buffers in TMEM (2 x BLOCK_N columns): the MMA warp fills buffer it&1 while
//     the epilogue warps drain the other one, so the epilogue (and the launch / TMEM-alloc /
//     descriptor-fetch prologue, paid once per CTA instead of once per tile) is off the tensor
//     core's critical path.  Short-K shapes (GPT-2: K = 768, early conv layers) were dominated
//     by exactly those per-tile costs.
// Tile order is n-fastest so concurrently running CTAs share the A row block through L2.
// ---------------------------------------------------------------------------
template <int BLOCK_N, bool A_MN, bool B_MN, bool CONV_A, int STAGES, bool TMA_STORE>
__global__ void __launch_bounds__(kThreads)
    umma_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const __grid_constant__ CUtensorMap tmC, const KParams p, int m_tiles, int n_tiles,
                             int num_tiles) {
  constexpr int A_BYTES = BLOCK_M * 128;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // TMA_STORE: the epilogue stages a [128 rows x 64 cols] bf16 sub-tile in 128B-swizzled shared
  // memory and one thread issues a coalesced TMA store (cp.async.bulk.tensor ... global.shared::cta);
  // the per-thread 16-byte row stores of the direct path cost one LSU wavefront each and, for
  // short-K tiles, the store stream was as long as the MMA main loop.
  constexpr int STORE_BYTES = TMA_STORE ? BLOCK_M * 128 : 0;
  constexpr uint32_t ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS <= 512, "two accumulator buffers must fit in TMEM");
  static_assert(!(CONV_A && A_MN), "conv window operand is K-major");

  // two CTAs per SM only fit without alignment slack: the window itself is declared 1024-byte aligned
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sC = smem + STAGES * STAGE_BYTES;      // store staging (TMA_STORE only)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STORE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TMA_STORE) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ----------------------------- TMA producer -----------------------------
      int kit = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tile_n = t % n_tiles;
        const int tile_m = (t / n_tiles) % m_tiles;
        const int z = t / (n_tiles * m_tiles);
        const int m0 = tile_m * BLOCK_M, n0 = tile_n * BLOCK_N;
        const int kb_begin = z * p.kb_per_split;
        const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
        int img0 = 0, h0 = 0;
        if (CONV_A) {
          const int hw = p.OH * p.OW;
          if (p.BN > 1) {
            img0 = tile_m * p.BN;
          } else {
            img0 = m0 / hw;
            h0 = (m0 - img0 * hw) / p.OW;
          }
        }
        for (int kb = kb_begin; kb < kb_end; ++kb, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full[s], STAGE_BYTES);
          int kB;
          if (CONV_A) {
            const int tp = kb / p.kc_per_tap;
            const int kc = kb - tp * p.kc_per_tap;
            const ConvTap tap = p.taps[tp];
            tma_load_4d(sa, &tmA, &full[s], kc * BLOCK_K, tap.dw, h0 + tap.dh, img0 + tap.nb);
            kB = tap.koff + kc * BLOCK_K;
          } else {
            kB = kb * BLOCK_K;
            if (A_MN) {
              tma_load_2d(sa, &tmA, &full[s], m0, kB);
              tma_load_2d(sa + ATOM_BYTES, &tmA, &full[s], m0 + 64, kB);
            } else {
              tma_load_2d(sa, &tmA, &full[s], kB, m0);
            }
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * ATOM_BYTES, &tmB, &full[s], n0 + j * 64, kB);
          } else {
            tma_load_2d(sb, &tmB, &full[s], kB, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ----------------------------- MMA issuer -----------------------------
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      int kit = 0, it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int z = t / (n_tiles * m_tiles);
        const int kb_begin = z * p.kb_per_split;
        const int nkb = min(kb_begin + p.kb_per_split, p.kb_total) - kb_begin;
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);      // epilogue drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int i = 0; i < nkb; ++i, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t ad = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, ATOM_BYTES, 1024)
                                     : make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bd = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, ATOM_BYTES, 1024)
                                     : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_f16(d_tmem, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ----------------------------- epilogue -----------------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int tile_n = t % n_tiles;
      const int tile_m = (t / n_tiles) % m_tiles;
      const int m0 = tile_m * BLOCK_M, n0 = tile_n * BLOCK_N;
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const int m = m0 + row;
      const bool row_ok = m < p.M;
      const long long row_off = row_ok ? epi_row_offset(p.epi, m) : 0;
      if (TMA_STORE) {
        const bool store_thread = (threadIdx.x == 64);          // first lane of the first epilogue warp
        uint8_t* srow = sC + row * 128;
        const int sw = row & 7;
#pragma unroll 1
        for (int h = 0; h < BLOCK_N; h += 64) {
          if (n0 + h >= p.N) break;
          if (store_thread) tma_store_wait_read();              // previous sub-tile left the staging buffer
          named_bar_sync(1, 128);
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ACC_COLS + h + c, r);
            tmem_ld_wait();
            float v[32];
            epi_compute_32(p, r, row_ok, row_off, n0 + h + c, v);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint4 o;
              o.x = pack_bf16(v[t * 8 + 0], v[t * 8 + 1]);
              o.y = pack_bf16(v[t * 8 + 2], v[t * 8 + 3]);
              o.z = pack_bf16(v[t * 8 + 4], v[t * 8 + 5]);
              o.w = pack_bf16(v[t * 8 + 6], v[t * 8 + 7]);
              *reinterpret_cast<uint4*>(srow + ((((c >> 3) + t) ^ sw) << 4)) = o;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (store_thread) {
            tma_store_2d(&tmC, sC, n0 + h, m0);                 // rows >= M / cols >= N are clipped by TMA
            tma_store_commit();
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          if (n0 + c >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ACC_COLS + c, r);
          tmem_ld_wait();
          epi_store_32(p, r, row_ok, row_off, n0 + c);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
    if (TMA_STORE && threadIdx.x == 64) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------
// CTA-pair GEMM / conv-fprop kernel (cta_group::2): a (2,1,1) cluster computes a 256 x BLOCK_N
// tile.  Each CTA TMA-loads its own 128 rows of A (or its own conv window) and its own
// BLOCK_N/2 rows of B, so per-SM shared-memory traffic per MMA halves compared with the
// single-CTA kernel (the SS-mode 128x128 tile is smem-bandwidth bound at ~50% of tensor peak).
// The leader issues tcgen05.mma.cta_group::2 with M = 256; accumulator rows 0..127 sit in the
// leader's TMEM, 128..255 in the peer's; commits are multicast to both CTAs' barriers.
// K-major A (plain or conv window) x K-major B only (dense fwd, conv fprop, conv dgrad).
// ---------------------------------------------------------------------------
template <int BLOCK_N, bool CONV_A, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads)
    umma_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const KParams p) {
  constexpr int A_BYTES = BLOCK_M * 128;
  constexpr int B_BYTES = (BLOCK_N / 2) * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BLOCK_N;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m0 = blockIdx.x * BLOCK_M;                 // this CTA's 128 rows (pair = 256 rows)
  const int n0 = blockIdx.y * BLOCK_N;
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
  if (warp == 1) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ----------------------------- TMA producer (both CTAs; bytes land on the leader's barrier)
      int img0 = 0, h0 = 0;
      if (CONV_A) {
        const int hw = p.OH * p.OW;
        if (p.BN > 1) {
          img0 = blockIdx.x * p.BN;
        } else {
          img0 = m0 / hw;
          h0 = (m0 - img0 * hw) / p.OW;
        }
      }
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        if (leader) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
        int kB;
        if (CONV_A) {
          const int t = i / p.kc_per_tap;
          const int kc = i - t * p.kc_per_tap;
          const ConvTap tap = p.taps[t];
          tma_load_4d_2sm(sa, &tmA, &full[s], kc * BLOCK_K, tap.dw, h0 + tap.dh, img0 + tap.nb);
          kB = tap.koff + kc * BLOCK_K;
        } else {
          kB = i * BLOCK_K;
          tma_load_2d_2sm(sa, &tmA, &full[s], kB, m0);
        }
        tma_load_2d_2sm(sb, &tmB, &full[s], kB, n0 + (int)rank * (BLOCK_N / 2));
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ----------------------------- MMA issuer (leader CTA only)
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BLOCK_N, 0u, 0u);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k)
          umma_f16_2sm(tmem_base, make_smem_desc_sw128(a_addr + k * 32, 0, 1024),
                       make_smem_desc_sw128(b_addr + k * 32, 0, 1024), idesc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit_2sm(&empty[s]);
      }
      umma_commit_2sm(tmem_full);
    }
  } else {
    // ----------------------------- epilogue (each CTA drains its own 128 accumulator rows)
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const long long row_off = row_ok ? epi_row_offset(p.epi, m) : 0;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      if (n0 + c >= p.N) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      epi_store_32(p, r, row_ok, row_off, n0 + c);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------
// Persistent CTA-pair kernel: the (2,1,1) cluster walks 256 x BLOCK_N tiles, double-buffered
// accumulators in both CTAs' TMEM.  The leader may only overwrite accumulator buffer `acc` once
// BOTH CTAs' epilogues drained their 128 rows of it, so all 256 epilogue threads of the pair
// arrive on the LEADER's tmem_empty barrier (remote mbarrier arrive through the cluster window).
// ---------------------------------------------------------------------------
template <int BLOCK_N, bool CONV_A, int STAGES, bool TMA_STORE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads)
    umma_gemm2_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const __grid_constant__ CUtensorMap tmC, const KParams p,
                              int m_tiles /* 256-row tiles */, int n_tiles, int num_tiles) {
  constexpr int A_BYTES = BLOCK_M * 128;
  constexpr int B_BYTES = (BLOCK_N / 2) * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STORE_BYTES = TMA_STORE ? BLOCK_M * 128 : 0;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  static_assert(TMEM_COLS <= 512, "two accumulator buffers must fit in TMEM");

  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sC = smem + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STORE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2] (only the leader's copy is used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int nkb = p.kb_total;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TMA_STORE) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 256);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int kit = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int tile_n = t % n_tiles;
        const int tile_m = t / n_tiles;
        const int m0 = (tile_m * 2 + (int)rank) * BLOCK_M;
        const int n0 = tile_n * BLOCK_N;
        int img0 = 0, h0 = 0;
        if (CONV_A) {
          const int hw = p.OH * p.OW;
          if (p.BN > 1) {
            img0 = (tile_m * 2 + (int)rank) * p.BN;
          } else {
            img0 = m0 / hw;
            h0 = (m0 - img0 * hw) / p.OW;
          }
        }
        for (int i = 0; i < nkb; ++i, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (leader) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
          int kB;
          if (CONV_A) {
            const int tp = i / p.kc_per_tap;
            const int kc = i - tp * p.kc_per_tap;
            const ConvTap tap = p.taps[tp];
            tma_load_4d_2sm(sa, &tmA, &full[s], kc * BLOCK_K, tap.dw, h0 + tap.dh, img0 + tap.nb);
            kB = tap.koff + kc * BLOCK_K;
          } else {
            kB = i * BLOCK_K;
            tma_load_2d_2sm(sa, &tmA, &full[s], kB, m0);
          }
          tma_load_2d_2sm(sb, &tmB, &full[s], kB, n0 + (int)rank * (BLOCK_N / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BLOCK_N, 0u, 0u);
      int kit = 0, it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int i = 0; i < nkb; ++i, ++kit) {
          const int s = kit % STAGES;
          const uint32_t ph = (kit / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_f16_2sm(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 0, 1024),
                         make_smem_desc_sw128(b_addr + k * 32, 0, 1024), idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty[s]);
        }
        umma_commit_2sm(&tmem_full[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int tile_n = t % n_tiles;
      const int tile_m = t / n_tiles;
      const int m0 = (tile_m * 2 + (int)rank) * BLOCK_M;
      const int n0 = tile_n * BLOCK_N;
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const int m = m0 + row;
      const bool row_ok = m < p.M;
      const long long row_off = row_ok ? epi_row_offset(p.epi, m) : 0;
      if (TMA_STORE) {
        const bool store_thread = (threadIdx.x == 64);
        uint8_t* srow = sC + row * 128;
        const int sw = row & 7;
#pragma unroll 1
        for (int h = 0; h < BLOCK_N; h += 64) {
          if (n0 + h >= p.N) break;
          if (store_thread) tma_store_wait_read();
          named_bar_sync(1, 128);
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + h + c, r);
            tmem_ld_wait();
            float v[32];
            epi_compute_32(p, r, row_ok, row_off, n0 + h + c, v);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint4 o;
              o.x = pack_bf16(v[t * 8 + 0], v[t * 8 + 1]);
              o.y = pack_bf16(v[t * 8 + 2], v[t * 8 + 3]);
              o.z = pack_bf16(v[t * 8 + 4], v[t * 8 + 5]);
              o.w = pack_bf16(v[t * 8 + 6], v[t * 8 + 7]);
              *reinterpret_cast<uint4*>(srow + ((((c >> 3) + t) ^ sw) << 4)) = o;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (store_thread) {
            tma_store_2d(&tmC, sC, n0 + h, m0);
            tma_store_commit();
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          if (n0 + c >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + c, r);
          tmem_ld_wait();
          epi_store_32(p, r, row_ok, row_off, n0 + c);
        }
      }
      tc_fence_before();
      mbar_arrive_leader(&tmem_empty[acc]);
    }
    if (TMA_STORE && threadIdx.x == 64) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------
// conv wgrad kernel: M = Kout tile (128), N = 64 input channels per tap, T taps
// ---------------------------------------------------------------------------
template <int T, int STAGES>
__global__ void __launch_bounds__(kThreads)
    umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmA,
                      const __grid_constant__ CUtensorMap tmB, const KParams p, int c_tiles,
                      float* __restrict__ dW, long long ldw) {
  constexpr int A_BYTES = 2 * ATOM_BYTES;
  constexpr int B_BYTES = T * ATOM_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (T * 64 <= 64) ? 64 : ((T * 64 <= 128) ? 128 : ((T * 64 <= 256) ? 256 : 512));

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kout0 = (blockIdx.y / c_tiles) * BLOCK_M;
  const int c0 = (blockIdx.y % c_tiles) * 64;
  const int tap0 = blockIdx.z * T;
  const int kb_begin = blockIdx.x * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
  const int nkb = kb_end - kb_begin;

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

  if (nkb > 0) {
    if (warp == 0) {
      if (elect_one()) {
        const int hw = p.OH * p.OW;
        for (int i = 0; i < nkb; ++i) {
          const int kb = kb_begin + i;
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full[s], STAGE_BYTES);
          const int pix0 = kb * BLOCK_K;
          int img0, h0;
          if (p.BN > 1) {
            img0 = kb * p.BN;
            h0 = 0;
          } else {
            img0 = pix0 / hw;
            h0 = (pix0 - img0 * hw) / p.OW;
          }
          tma_load_2d(sa, &tmA, &full[s], kout0, pix0);
          tma_load_2d(sa + ATOM_BYTES, &tmA, &full[s], kout0 + 64, pix0);
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const ConvTap tap = p.taps[tap0 + t];
            tma_load_4d(sb + t * ATOM_BYTES, &tmB, &full[s], c0, tap.dw, h0 + tap.dh,
                        img0 + tap.nb);
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        // ONE MMA of N = T*64 per k-step: the T tap tiles are consecutive 64-wide MN-major atoms
        // (LBO = ATOM_BYTES), so the dY operand is fetched from shared memory once per k-step
        // instead of once per tap -- the kernel was shared-memory-bandwidth bound.
        constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, T * 64, 1u, 1u);
        constexpr uint32_t idesc1 = make_idesc_bf16(BLOCK_M, 64, 1u, 1u);
        const bool wide = p.vec_ok != 0;      // host switch (TNN_WGRAD_WIDE): one N = T*64 MMA vs T MMAs of N = 64
        for (int i = 0; i < nkb; ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
          if (wide) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * 2048, ATOM_BYTES, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 2048, ATOM_BYTES, 1024);
              umma_f16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int t = 0; t < T; ++t) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k) {
                const uint64_t ad = make_smem_desc_sw128(a_addr + k * 2048, ATOM_BYTES, 1024);
                const uint64_t bd = make_smem_desc_sw128(b_addr + t * ATOM_BYTES + k * 2048, ATOM_BYTES, 1024);
                umma_f16(tmem_base + t * 64, ad, bd, idesc1, (i > 0 || k > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(tmem_full);
      }
    } else {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int ko = kout0 + row;
      const bool row_ok = ko < p.M;
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        const int koff = p.taps[tap0 + t].koff;
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
          if (c0 + c >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 64 + c, r);
          tmem_ld_wait();
          if (row_ok) {
            float* o = dW + (long long)ko * ldw + koff + c0 + c;
            const float al = p.epi.alpha;
            if ((c0 + c + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
              // split-K partial sums: 8 vector reductions instead of 32 scalar atomics per thread
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                red_add_v4(o + j, __uint_as_float(r[j]) * al, __uint_as_float(r[j + 1]) * al,
                           __uint_as_float(r[j + 2]) * al, __uint_as_float(r[j + 3]) * al);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + c + j < p.N) atomicAdd(o + j, __uint_as_float(r[j]) * al);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TNN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (p == nullptr || qres != cudaDriverEntryPointSuccess) {
      fprintf(stderr, "[tnn_b200] cuTensorMapEncodeTiled unavailable\n");
      abort();
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// dims / box innermost-first; strides (bytes) for dims 1..rank-1.
static CUtensorMap make_tmap_bf16(const void* ptr, int rank, const uint64_t* dims,
                                  const uint64_t* strides_bytes, const uint32_t* box) {
  CUtensorMap m;
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr),
                               gdims, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr,
            "[tnn_b200] cuTensorMapEncodeTiled failed (%d): rank=%d ptr=%p dims=[%llu,%llu,%llu,%llu] "
            "stride1=%llu box=[%u,%u,%u,%u]\n",
            (int)r, rank, ptr, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
            (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
            (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0,
            rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    abort();
  }
  return m;
}

// 2-D operand map.  K-major: global [rows, K] -> dims {K, rows}, box {64, box_rows}.
// MN-major: global [K, mn] -> dims {mn, K}, box {64, 64}.
static CUtensorMap make_operand_map(const void* ptr, long long rows, long long cols, long long ld,
                                    uint32_t box_rows) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)ld * 2};
  uint32_t box[2] = {64, box_rows};
  return make_tmap_bf16(ptr, 2, dims, str, box);
}

static CUtensorMap make_window_map(const void* X, const ConvGeom& g, int BH, int BN) {
  uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.IN};
  uint64_t str[3] = {(uint64_t)g.C * 2, (uint64_t)g.IW * g.C * 2, (uint64_t)g.IH * g.IW * g.C * 2};
  uint32_t box[4] = {64, (uint32_t)g.OW, (uint32_t)BH, (uint32_t)BN};
  return make_tmap_bf16(X, 4, dims, str, box);
}

static bool pick_window(const ConvGeom& g, int rows, int* BH, int* BN) {
  // window of `rows` output pixels = BN images x BH rows x OW columns
  if (g.OW <= 0 || rows % g.OW != 0) return false;
  int r = rows / g.OW;  // rows of width OW
  if (r <= g.OH) {
    if (g.OH % r != 0) return false;
    *BH = r;
    *BN = 1;
  } else {
    if (r % g.OH != 0) return false;
    *BH = g.OH;
    *BN = r / g.OH;
  }
  return *BH <= 256 && *BN <= 256 && g.OW <= 256;
}

bool conv_geom_supported(const ConvGeom& g) {
  int bh, bn;
  if (g.C % 8 != 0) return false;
  if (g.ntaps < 1 || g.ntaps > kMaxTaps) return false;
  return pick_window(g, 128, &bh, &bn) && pick_window(g, 64, &bh, &bn);
}

template <typename Kern>
static void set_smem(Kern k, int bytes) {
  TNN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool CONV_A>
static void launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, dim3 grid,
                        cudaStream_t stream) {
  constexpr int STAGES = (BLOCK_N <= 64) ? 4 : ((BLOCK_N <= 128) ? 3 : 4);
  constexpr int smem = STAGES * (BLOCK_M * 128 + BLOCK_N * 128) + 1024 + 256;
  auto kern = umma_gemm_kernel<BLOCK_N, A_MN, B_MN, CONV_A, STAGES>;
  static PerDeviceOnce configured;
  if (configured.begin()) {
    set_smem(kern, smem);
    configured.done();
  }
  kern<<<grid, kThreads, smem, stream>>>(tmA, tmB, p);
  ++g_launch_count;
  TNN_CUDA_CHECK(cudaGetLastError());
}

template <int BLOCK_N, bool CONV_A>
static void launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, dim3 grid,
                         cudaStream_t stream) {
  constexpr int STAGES = (BLOCK_N <= 128) ? 4 : 3;
  constexpr int smem = STAGES * (BLOCK_M * 128 + (BLOCK_N / 2) * 128) + 1024 + 256;
  auto kern = umma_gemm2_kernel<BLOCK_N, CONV_A, STAGES>;
  static PerDeviceOnce configured;
  if (configured.begin()) {
    set_smem(kern, smem);
    configured.done();
  }
  kern<<<grid, kThreads, smem, stream>>>(tmA, tmB, p);   // cluster dims are a kernel attribute
  ++g_launch_count;
  TNN_CUDA_CHECK(cudaGetLastError());
}

template <int BLOCK_N, bool CONV_A, bool TMA_STORE>
static void launch_gemm2_persist_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                   const KParams& p, int m_tiles256, int n_tiles, cudaStream_t stream) {
  constexpr int STAGES = (BLOCK_N <= 128) ? 4 : 6;     // BLOCK_N = 256 owns the SM (TMEM full): deeper ring
  // barriers: (2*STAGES + 4) mbarriers + the TMEM slot = 132 bytes at 6 stages -- 128 was 4 bytes short
  // (tmem_slot landed past the dynamic window: "illegal memory access" on the first pair launch)
  constexpr int kBarBytes = ((2 * STAGES + 4) * 8 + 4 + 127) / 128 * 128;
  constexpr int smem = STAGES * (BLOCK_M * 128 + (BLOCK_N / 2) * 128) + (TMA_STORE ? BLOCK_M * 128 : 0) + kBarBytes;
  auto kern = umma_gemm2_persist_kernel<BLOCK_N, CONV_A, STAGES, TMA_STORE>;
  static PerDeviceOnce configured;
  if (configured.begin()) {
    set_smem(kern, smem);
    configured.done();
  }
  const int num_tiles = m_tiles256 * n_tiles;
  const int max_pairs = (BLOCK_N <= 128) ? 148 : 74;   // 2 resident CTAs per SM, or 1
  const int pairs = num_tiles < max_pairs ? num_tiles : max_pairs;
  kern<<<2 * pairs, kThreads, smem, stream>>>(tmA, tmB, tmC, p, m_tiles256, n_tiles, num_tiles);
  ++g_launch_count;
  TNN_CUDA_CHECK(cudaGetLastError());
}

// CTA-pair path.  TNN_GEMM_2CTA: unset/-1 = auto (measured on B200, profiles/gemm2_check_r1.log and
// gemm_persist_check_r1.log: the persistent pair kernel wins for N >= 256 with >= 32 k-blocks, i.e. the
// 16x16x256 and 8x8x512 conv layers (+9..17%), and loses on short-K GEMMs where two resident 1-CTA
// persistent CTAs per SM overlap better), 0 = never, 1 = non-persistent pair kernel, 2 = persistent pair.
static int mode_2cta() {
  static int mode = -2;
  if (mode == -2) {
    const char* e = getenv("TNN_GEMM_2CTA");
    mode = (e != nullptr) ? atoi(e) : -1;
  }
  return mode;
}
static bool use_2cta(int M, int N, bool a_mn, bool b_mn, int splits, int kb_total, bool conv = false) {
  const int mode = mode_2cta();
  if (mode == 0 || a_mn || b_mn || splits > 1 || M < 256 || N < 128) return false;
  if (mode < 0) {
    // auto rule; the two thresholds are tunable for A/B runs (TNN_GEMM_2CTA_MIN_N / _MIN_KB)
    // Dense GEMM: N >= 256 and >= 32 k-blocks (short-K GEMMs overlap better as two resident 1-CTA
    // persistent CTAs per SM).  Implicit-GEMM convolutions: N >= 128 and >= 8 k-blocks -- measured in the
    // same run on WRN-16-8 (profiles/ab_r1_b20.md): 4.300 ms/step vs 4.330 with the dense rule.
    static int min_n = -1, min_kb = -1;
    if (min_n < 0) {
      const char* a = getenv("TNN_GEMM_2CTA_MIN_N");
      const char* b = getenv("TNN_GEMM_2CTA_MIN_KB");
      min_n = (a != nullptr) ? atoi(a) : 0;
      min_kb = (b != nullptr) ? atoi(b) : 0;
    }
    const int need_n = min_n > 0 ? min_n : (conv ? 128 : 256);
    const int need_kb = min_kb > 0 ? min_kb : (conv ? 8 : 32);
    return N >= need_n && kb_total >= need_kb;
  }
  return true;
}
static bool use_2cta_persist() { return mode_2cta() != 1; }

static bool use_persistent() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("TNN_GEMM_PERSIST");
    mode = (e != nullptr) ? atoi(e) : 1;
  }
  return mode > 0;
}

// TMA-store epilogue is legal for a plain row-major bf16 output with 16-byte aligned rows
static bool tma_store_ok(const Epilogue& e, int N) {
  static int mode = -1;
  if (mode < 0) {
    const char* env = getenv("TNN_GEMM_TMA_STORE");
    mode = (env != nullptr) ? atoi(env) : 1;
  }
  // N itself need not be a multiple of 8 (GPT-2's 50257-wide LM head writes into rows padded to 50264):
  // the store map clips the last box and the epilogue guards bias / residual loads per column
  (void)N;
  return mode > 0 && e.out_mode == OUT_BF16 && e.grid_mode == 0 && e.base == 0 && (e.ldc % 8) == 0 &&
         (e.residual == nullptr || N % 8 == 0) && (reinterpret_cast<uintptr_t>(e.out) & 15) == 0;
}
static CUtensorMap make_store_map(const Epilogue& e, int M, int N) {
  uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
  uint64_t str[1] = {(uint64_t)e.ldc * 2};
  uint32_t box[2] = {64, (uint32_t)BLOCK_M};
  return make_tmap_bf16(e.out, 2, dims, str, box);
}

template <int BLOCK_N, bool CONV_A>
static void launch_gemm2_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, int m_tiles256,
                                 int n_tiles, cudaStream_t stream) {
  if (tma_store_ok(p.epi, p.N)) {
    CUtensorMap tmC = make_store_map(p.epi, p.M, p.N);
    launch_gemm2_persist_t<BLOCK_N, CONV_A, true>(tmA, tmB, tmC, p, m_tiles256, n_tiles, stream);
  } else {
    launch_gemm2_persist_t<BLOCK_N, CONV_A, false>(tmA, tmB, tmA, p, m_tiles256, n_tiles, stream);
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool CONV_A, bool TMA_STORE>
static void launch_gemm_persist_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const KParams& p,
                                  int m_tiles, int n_tiles, int splits, cudaStream_t stream) {
  constexpr int STAGES = (BLOCK_N <= 64) ? 4 : ((BLOCK_N <= 128) ? 3 : 4);
  constexpr int kBarBytes = ((2 * STAGES + 4) * 8 + 4 + 127) / 128 * 128;
  constexpr int smem = STAGES * (BLOCK_M * 128 + BLOCK_N * 128) + (TMA_STORE ? BLOCK_M * 128 : 0) + kBarBytes;
  auto kern = umma_gemm_persist_kernel<BLOCK_N, A_MN, B_MN, CONV_A, STAGES, TMA_STORE>;
  static PerDeviceOnce configured;
  if (configured.begin()) {
    set_smem(kern, smem);
    configured.done();
  }
  const int num_tiles = m_tiles * n_tiles * splits;
  const int max_ctas = (BLOCK_N <= 128) ? 2 * 148 : 148;
  const int grid = num_tiles < max_ctas ? num_tiles : max_ctas;
  kern<<<grid, kThreads, smem, stream>>>(tmA, tmB, tmC, p, m_tiles, n_tiles, num_tiles);
  ++g_launch_count;
  TNN_CUDA_CHECK(cudaGetLastError());
}
template <int BLOCK_N, bool A_MN, bool B_MN, bool CONV_A>
static void launch_gemm_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, int m_tiles,
                                int n_tiles, int splits, cudaStream_t stream) {
  if (splits == 1 && tma_store_ok(p.epi, p.N)) {
    CUtensorMap tmC = make_store_map(p.epi, p.M, p.N);
    launch_gemm_persist_t<BLOCK_N, A_MN, B_MN, CONV_A, true>(tmA, tmB, tmC, p, m_tiles, n_tiles, splits, stream);
  } else {
    launch_gemm_persist_t<BLOCK_N, A_MN, B_MN, CONV_A, false>(tmA, tmB, tmA, p, m_tiles, n_tiles, splits, stream);
  }
}

static int pick_block_n(int N) {
  const char* e = getenv("TNN_GEMM_BLOCK_N");
  if (e != nullptr) {
    int v = atoi(e);
    if (v == 64 || v == 128 || v == 256) return v;
  }
  if (N <= 64) return 64;
  return 128;
}

static int epi_vec_ok(const Epilogue& e, int N) {
  // row alignment is what the 16-byte accesses need; a ragged last chunk (N % 8 != 0) takes the
  // per-element path through the `full` test in the epilogue
  (void)N;
  bool ok = (e.base % 8 == 0);
  if (e.grid_mode)
    ok = ok && (e.sN % 8 == 0) && (e.sH % 8 == 0) && (e.sW % 8 == 0);
  else
    ok = ok && (e.ldc % 8 == 0);
  ok = ok && ((reinterpret_cast<uintptr_t>(e.out) & 15) == 0);
  if (e.residual) ok = ok && ((reinterpret_cast<uintptr_t>(e.residual) & 15) == 0);
  return ok ? 1 : 0;
}

#define TNN_DISPATCH_BN(BN_, ...)            \
  if (BN_ == 64) {                           \
    constexpr int BNc = 64;                  \
    __VA_ARGS__                              \
  } else if (BN_ == 128) {                   \
    constexpr int BNc = 128;                 \
    __VA_ARGS__                              \
  } else {                                   \
    constexpr int BNc = 256;                 \
    __VA_ARGS__                              \
  }

void gemm_bf16(const void* A, const void* B, int M, int N, int K, long long lda, long long ldb,
               bool a_mn, bool b_mn, const Epilogue& epi, int splits, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  if (lda % 8 != 0 || ldb % 8 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(B) & 15)) {
    fprintf(stderr, "[tnn_b200] gemm_bf16: operands must be 16-byte aligned with ld %% 8 == 0\n");
    abort();
  }
  if (use_2cta(M, N, a_mn, b_mn, splits, (K + BLOCK_K - 1) / BLOCK_K)) {
    const int bn2 = N >= 256 ? 256 : 128;
    KParams p{};
    p.M = M;
    p.N = N;
    p.K = K;
    p.kc_per_tap = (K + BLOCK_K - 1) / BLOCK_K;
    p.kb_total = p.kc_per_tap;
    p.kb_per_split = p.kb_total;
    p.ntaps = 1;
    p.epi = epi;
    p.vec_ok = epi_vec_ok(epi, N);
    CUtensorMap tmA = make_operand_map(A, M, K, lda, BLOCK_M);
    CUtensorMap tmB = make_operand_map(B, N, K, ldb, bn2 / 2);
    dim3 grid(2 * ((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)), (N + bn2 - 1) / bn2, 1);
    if (use_2cta_persist()) {
      if (bn2 == 256) launch_gemm2_persist<256, false>(tmA, tmB, p, (int)grid.x / 2, (int)grid.y, stream);
      else launch_gemm2_persist<128, false>(tmA, tmB, p, (int)grid.x / 2, (int)grid.y, stream);
      return;
    }
    if (bn2 == 256) launch_gemm2<256, false>(tmA, tmB, p, grid, stream);
    else launch_gemm2<128, false>(tmA, tmB, p, grid, stream);
    return;
  }
  int bn = pick_block_n(N);
  {
    // wide K-major GEMMs (QKV / FC1 / LM head forward) as 128 x 256 tiles, one CTA per SM: faster in the
    // isolated gemm_bench (qkv 1110 -> 1180, fc1 1144 -> 1237 TFLOP/s) but SLOWER inside the models in a
    // same-run A/B (GPT-2 prefill 11.6 -> 12.9 ms, training step 14.5 -> 15.3 ms: bias/GELU epilogues of a
    // 256-wide tile on 4 warps, one CTA per SM) -> opt-in with TNN_GEMM_AUTO_BN256=1.
    static int auto256 = -1;
    if (auto256 < 0) {
      const char* e = getenv("TNN_GEMM_AUTO_BN256");
      auto256 = (e != nullptr) ? atoi(e) : 0;
    }
    if (auto256 && getenv("TNN_GEMM_BLOCK_N") == nullptr && !a_mn && !b_mn && splits <= 1 && N >= 1024 &&
        (long long)((M + BLOCK_M - 1) / BLOCK_M) * ((N + 255) / 256) >= 2 * 148)
      bn = 256;
  }
  KParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.kc_per_tap = (K + BLOCK_K - 1) / BLOCK_K;
  p.kb_total = p.kc_per_tap;
  if (splits < 1) splits = 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  if (splits > 1 && epi.out_mode != OUT_F32_ATOMIC) {
    fprintf(stderr, "[tnn_b200] gemm_bf16: split-K requires OUT_F32_ATOMIC\n");
    abort();
  }
  p.ntaps = 1;
  p.epi = epi;
  p.vec_ok = epi_vec_ok(epi, N);
  CUtensorMap tmA = a_mn ? make_operand_map(A, K, M, lda, 64) : make_operand_map(A, M, K, lda, BLOCK_M);
  CUtensorMap tmB = b_mn ? make_operand_map(B, K, N, ldb, 64) : make_operand_map(B, N, K, ldb, bn);
  dim3 grid((M + BLOCK_M - 1) / BLOCK_M, (N + bn - 1) / bn, splits);
  if (use_persistent() && bn == 256 && !a_mn && !b_mn && splits == 1) {
    // 128 x 256 tiles, one CTA per SM (both TMEM accumulator buffers = all 512 columns)
    launch_gemm_persist<256, false, false, false>(tmA, tmB, p, (int)grid.x, (int)grid.y, 1, stream);
    return;
  }
  if (use_persistent() && bn <= 128) {
    const int mt = (int)grid.x, nt = (int)grid.y;
    if (bn == 64) {
      if (!a_mn && !b_mn) launch_gemm_persist<64, false, false, false>(tmA, tmB, p, mt, nt, splits, stream);
      else if (!a_mn && b_mn) launch_gemm_persist<64, false, true, false>(tmA, tmB, p, mt, nt, splits, stream);
      else if (a_mn && !b_mn) launch_gemm_persist<64, true, false, false>(tmA, tmB, p, mt, nt, splits, stream);
      else launch_gemm_persist<64, true, true, false>(tmA, tmB, p, mt, nt, splits, stream);
    } else {
      if (!a_mn && !b_mn) launch_gemm_persist<128, false, false, false>(tmA, tmB, p, mt, nt, splits, stream);
      else if (!a_mn && b_mn) launch_gemm_persist<128, false, true, false>(tmA, tmB, p, mt, nt, splits, stream);
      else if (a_mn && !b_mn) launch_gemm_persist<128, true, false, false>(tmA, tmB, p, mt, nt, splits, stream);
      else launch_gemm_persist<128, true, true, false>(tmA, tmB, p, mt, nt, splits, stream);
    }
    return;
  }
  TNN_DISPATCH_BN(bn, {
    if (!a_mn && !b_mn) launch_gemm<BNc, false, false, false>(tmA, tmB, p, grid, stream);
    else if (!a_mn && b_mn) launch_gemm<BNc, false, true, false>(tmA, tmB, p, grid, stream);
    else if (a_mn && !b_mn) launch_gemm<BNc, true, false, false>(tmA, tmB, p, grid, stream);
    else launch_gemm<BNc, true, true, false>(tmA, tmB, p, grid, stream);
  })
}

void conv_fprop_bf16(const void* X, const void* Wt, long long ldw, int N, const ConvGeom& g,
                     const Epilogue& epi, cudaStream_t stream) {
  int BH, BN;
  if (!pick_window(g, BLOCK_M, &BH, &BN) || g.C % 8 != 0 || ldw % 8 != 0) {
    fprintf(stderr, "[tnn_b200] conv_fprop_bf16: unsupported geometry OH=%d OW=%d C=%d\n", g.OH,
            g.OW, g.C);
    abort();
  }
  const int M = g.ON * g.OH * g.OW;
  const int bn = pick_block_n(N);
  KParams p{};
  p.M = M;
  p.N = N;
  p.K = g.C;
  p.kc_per_tap = (g.C + BLOCK_K - 1) / BLOCK_K;
  p.kb_total = p.kc_per_tap * g.ntaps;
  p.kb_per_split = p.kb_total;
  p.OH = g.OH;
  p.OW = g.OW;
  p.BH = BH;
  p.BN = BN;
  p.ntaps = g.ntaps;
  for (int t = 0; t < g.ntaps; ++t) p.taps[t] = g.taps[t];
  p.epi = epi;
  p.vec_ok = epi_vec_ok(epi, N);
  CUtensorMap tmA = make_window_map(X, g, BH, BN);
  if (use_2cta(M, N, false, false, 1, p.kb_total, true)) {
    const int bn2 = N >= 256 ? 256 : 128;
    CUtensorMap tmB2 = make_operand_map(Wt, N, ldw, ldw, bn2 / 2);
    dim3 grid2(2 * ((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)), (N + bn2 - 1) / bn2, 1);
    if (use_2cta_persist()) {
      if (bn2 == 256) launch_gemm2_persist<256, true>(tmA, tmB2, p, (int)grid2.x / 2, (int)grid2.y, stream);
      else launch_gemm2_persist<128, true>(tmA, tmB2, p, (int)grid2.x / 2, (int)grid2.y, stream);
      return;
    }
    if (bn2 == 256) launch_gemm2<256, true>(tmA, tmB2, p, grid2, stream);
    else launch_gemm2<128, true>(tmA, tmB2, p, grid2, stream);
    return;
  }
  // weight matrix [N, Ktot]: the last tap may read past koff+C; the extent is ldw
  CUtensorMap tmB = make_operand_map(Wt, N, ldw, ldw, bn);
  dim3 grid((M + BLOCK_M - 1) / BLOCK_M, (N + bn - 1) / bn, 1);
  if (use_persistent() && bn <= 128) {
    if (bn == 64) launch_gemm_persist<64, false, false, true>(tmA, tmB, p, (int)grid.x, (int)grid.y, 1, stream);
    else launch_gemm_persist<128, false, false, true>(tmA, tmB, p, (int)grid.x, (int)grid.y, 1, stream);
    return;
  }
  TNN_DISPATCH_BN(bn, { launch_gemm<BNc, false, false, true>(tmA, tmB, p, grid, stream); })
}

void conv_wgrad_bf16(const void* dY, long long ldy, int Kout, const void* X, const ConvGeom& g,
                     float* dW, long long ldw, float alpha, cudaStream_t stream) {
  int BH, BN;
  if (!pick_window(g, BLOCK_K, &BH, &BN) || g.C % 8 != 0 || ldy % 8 != 0 ||
      (g.ntaps != 1 && g.ntaps % 3 != 0)) {
    fprintf(stderr, "[tnn_b200] conv_wgrad_bf16: unsupported geometry OH=%d OW=%d C=%d taps=%d\n",
            g.OH, g.OW, g.C, g.ntaps);
    abort();
  }
  const int Mpix = g.ON * g.OH * g.OW;
  KParams p{};
  p.M = Kout;
  p.N = g.C;
  p.K = Mpix;
  p.kb_total = (Mpix + BLOCK_K - 1) / BLOCK_K;
  p.OH = g.OH;
  p.OW = g.OW;
  p.BH = BH;
  p.BN = BN;
  p.ntaps = g.ntaps;
  for (int t = 0; t < g.ntaps; ++t) p.taps[t] = g.taps[t];
  p.epi.alpha = alpha;
  {
    static int wide = -1;
    if (wide < 0) {
      const char* e = getenv("TNN_WGRAD_WIDE");
      wide = (e != nullptr) ? atoi(e) : 1;
    }
    p.vec_ok = wide;          // reused as the wide-MMA switch by the wgrad kernel
  }
  const int T = (g.ntaps == 1) ? 1 : 3;
  const int groups = g.ntaps / T;
  const int c_tiles = (g.C + 63) / 64;
  const int k_tiles = (Kout + BLOCK_M - 1) / BLOCK_M;
  const int tiles = c_tiles * k_tiles * groups;
  // Split-K plan.  The kernel is not persistent, so the grid must FIT the resident slots: with the
  // 4-stage ring (160 KB) one CTA lives on an SM (148 slots), with the 2-stage ring two do (296).
  // tiles*splits just above a multiple of the slot count costs a whole extra wave (the old
  // ceil(296/tiles) rule launched 300 / 312 / 384 CTAs for the three WRN stages = 3 waves of 148),
  // so splits = floor(slots / tiles) and the ring depth whose slot count is filled best wins.
  static int forced_stages = -1;
  if (forced_stages < 0) {
    const char* e = getenv("TNN_WGRAD_STAGES");
    forced_stages = (e != nullptr) ? atoi(e) : 0;
  }
  int stages = 4, splits = 1;
  {
    double best = -1.0;
    for (int st : {4, 2}) {
      if (T == 1 && st == 2) continue;                   // the 1-tap kernel has one (6-stage) variant
      if (forced_stages != 0 && T != 1 && st != forced_stages) continue;
      const int slots = (st == 2) ? 296 : 148;
      int sp = slots / tiles;
      if (sp < 1) sp = 1;
      if (sp > p.kb_total) sp = p.kb_total;
      const int ctas = tiles * sp;
      const int waves = (ctas + slots - 1) / slots;
      const double eff = (double)ctas / ((double)waves * slots) * (st == 2 ? 0.96 : 1.0);
      if (eff > best) {
        best = eff;
        stages = st;
        splits = sp;
      }
    }
  }
  {
    static int legacy = -1;                 // A/B switch: the previous ceil(296 / tiles) rule
    if (legacy < 0) {
      const char* e = getenv("TNN_WGRAD_LEGACY_SPLITS");
      legacy = (e != nullptr) ? atoi(e) : 0;
    }
    if (legacy) {
      stages = 4;
      splits = (2 * 148 + tiles - 1) / tiles;
      if (splits > p.kb_total) splits = p.kb_total;
    }
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;

  CUtensorMap tmA = make_operand_map(dY, Mpix, Kout, ldy, 64);
  uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.IN};
  uint64_t str[3] = {(uint64_t)g.C * 2, (uint64_t)g.IW * g.C * 2, (uint64_t)g.IH * g.IW * g.C * 2};
  uint32_t box[4] = {64, (uint32_t)g.OW, (uint32_t)BH, (uint32_t)BN};
  CUtensorMap tmB = make_tmap_bf16(X, 4, dims, str, box);
  dim3 grid(splits, k_tiles * c_tiles, groups);
  if (T == 1) {
    constexpr int STAGES = 6;
    constexpr int smem = STAGES * (2 * ATOM_BYTES + ATOM_BYTES) + 1024 + 256;
    auto kern = umma_wgrad_kernel<1, STAGES>;
    static PerDeviceOnce configured;
    if (configured.begin()) {
      set_smem(kern, smem);
      configured.done();
    }
    kern<<<grid, kThreads, smem, stream>>>(tmA, tmB, p, c_tiles, dW, ldw);
    ++g_launch_count;
  } else {
    if (stages == 2) {
      constexpr int STAGES = 2;
      constexpr int smem = STAGES * (2 * ATOM_BYTES + 3 * ATOM_BYTES) + 1024 + 256;
      auto kern = umma_wgrad_kernel<3, STAGES>;
      static PerDeviceOnce configured;
      if (configured.begin()) {
        set_smem(kern, smem);
        configured.done();
      }
      kern<<<grid, kThreads, smem, stream>>>(tmA, tmB, p, c_tiles, dW, ldw);
    } else {
      constexpr int STAGES = 4;
      constexpr int smem = STAGES * (2 * ATOM_BYTES + 3 * ATOM_BYTES) + 1024 + 256;
      auto kern = umma_wgrad_kernel<3, STAGES>;
      static PerDeviceOnce configured;
      if (configured.begin()) {
        set_smem(kern, smem);
        configured.done();
      }
      kern<<<grid, kThreads, smem, stream>>>(tmA, tmB, p, c_tiles, dW, ldw);
    }
    ++g_launch_count;
  }
  TNN_CUDA_CHECK(cudaGetLastError());
}

}  // namespace tnn
