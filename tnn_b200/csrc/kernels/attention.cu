// Fused multi-head attention for sm_100a (head_dim 64, bf16).
//
// Forward (prefill / training forward): flash-attention on the 5th-gen tensor cores.
//   One CTA = one 128-query tile of one (batch, head).  192 threads:
//     warp 0      TMA producer: Q once, then K_j / V_j tiles into a 2-stage ring
//     warp 1      TMEM allocator + tcgen05.mma issuer:  S = Q K_j^T (128x128x64 into TMEM
//                 columns [0,128)),  PV = P_j V_j (128x64x128 into columns [128,192))
//     warps 2..5  softmax: thread r owns query row r (= TMEM lane r): tcgen05.ld S, running
//                 max / sum (online softmax, exp2 with the 1/sqrt(d)*log2e factor folded in),
//                 causal / length masking in registers, P -> bf16 -> 128B-swizzled smem (the
//                 A operand of the PV MMA), then O = O*alpha + PV with O in registers.
//   Q/K/V are addressed straight out of the packed QKV projection ([B, L, 3, H, D]) or the
//   KV cache ([B, Lmax, H, D]) by 4-D TMA descriptors; there is no permute_heads / mask /
//   softmax kernel and no [B*H, L, L] score tensor (reference: src/nn/blocks_impl/
//   attention_block.cpp:138-181, cuda/permute_heads.cu, causal_mask.cu, softmax.cu; flash
//   variant: cuDNN SDPA, cuda/cudnn_flash_attention_ops.cu:123).  Two CTAs are co-resident per
//   SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
//
// Decode (L_q = 1 against the KV cache): memory bound; a SIMT kernel, one CTA per (batch, head),
//   8 lanes per key (16-byte loads), online softmax per lane group, merged through smem.
#include "attention.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <math.h>

#include <mutex>

#include "common.cuh"
#include "sm100.cuh"

namespace tnn {
using namespace sm100;

static constexpr int AT_BM = 128;        // queries per CTA
static constexpr int AT_BN = 128;        // keys per KV tile
static constexpr int AT_D = 64;          // head dim
static constexpr int AT_TILE = 128 * 128;  // bytes of one [128 rows][64 bf16] tile
static constexpr int AT_THREADS = 192;
static constexpr int AT_SMEM = 16384 /*Q*/ + 2 * 2 * 16384 /*K,V x2*/ + 2 * 16384 /*P*/ + 128 /*barriers*/;

struct AttnParams {
  int B, H, Lq, Lk;
  int causal;
  int q_offset;          // absolute position of query row 0 (prefill continuation)
  float scale_log2;      // log2(e) / sqrt(D)
  bf16* O;               // [B, Lq, H*D]
  long long o_row;       // elements between consecutive query rows of O (= H*D)
  float* lse;            // [B, H, Lq] natural-log LSE, may be null
};

// SPLIT = true (TNN_ATTN_SPLIT=1, opt-in): eight softmax warps instead of four (320 threads): warps 2..5 own
// key columns [0,64) of their TMEM-lane quadrant, warps 6..9 columns [64,128), and each thread keeps 32 of
// the 64 output columns.  The two threads of a query row agree on the running maximum through a flag +
// (rarely) an exchange that aliases the dead P tile; their partial row sums meet once at the end.
// Measured in the same run (B200, B=32 L=1024 H=12 causal): 0.220 ms vs 0.202 ms for the 4-warp kernel,
// L=4096: 0.374 vs 0.327 ms -- halving the per-thread softmax work does NOT help, the per-tile chain
// S-MMA -> softmax -> PV-MMA -> O update inside a CTA is the bound (two CTAs per SM interleave only two such
// chains).  The fix is a two-tile ping-pong per CTA, not more softmax threads; kept for that experiment.
template <bool CAUSAL, bool SPLIT, bool F16>
__global__ void __launch_bounds__(SPLIT ? 320 : AT_THREADS, 2)
    attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  // 2 CTAs/SM need 2 x (112 KB + barriers) <= 227 KB: no alignment slack, so the dynamic window
  // itself is declared 1024-byte aligned (128B-swizzle atoms) and checked once.
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_TILE;           // 2 stages
  uint8_t* sV = sK + 2 * AT_TILE;       // 2 stages
  uint8_t* sP = sV + 2 * AT_TILE;       // [128 q][128 keys] bf16 = two 64-key atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_TILE);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_free = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qb = gridDim.x - 1 - blockIdx.x;   // heavy (late, causal) tiles first
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int q0 = qb * AT_BM;
  int kv_len = p.Lk;
  if (CAUSAL) kv_len = min(p.Lk, q0 + AT_BM + p.q_offset);
  const int nblk = (kv_len + AT_BN - 1) / AT_BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, SPLIT ? 256 : 128);
    mbar_init(p_full, SPLIT ? 256 : 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, SPLIT ? 256 : 128);
    tmem_slot[1] = 0u;                       // SPLIT: "some row needs the second sweep" flag (tile index + 1)
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_PV = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, AT_TILE);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * AT_TILE);
        tma_load_4d(sK + s * AT_TILE, &tmK, &kv_full[s], 0, h, j * AT_BN, b);
        tma_load_4d(sV + s * AT_TILE, &tmV, &kv_full[s], 0, h, j * AT_BN, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_16<F16>(128, 128, 0u, 0u);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = make_idesc_16<F16>(128, 64, 0u, 1u);   // P (K-major) x V (MN-major)
      mbar_wait(q_full, 0);
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        if (j > 0) mbar_wait(s_free, (j - 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * AT_TILE);
        const uint32_t v_addr = smem_u32(sV + s * AT_TILE);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_f16(tmem_S, make_smem_desc_sw128(q_addr + k * 32, 0, 1024),
                   make_smem_desc_sw128(k_addr + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        if (j > 0) mbar_wait(o_free, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k) {
          // P: keys [16k, 16k+16) live in atom k/4 at byte offset (k%4)*32 inside the 128-byte rows
          const uint64_t ad = make_smem_desc_sw128(p_addr + (k >> 2) * AT_TILE + (k & 3) * 32, 0, 1024);
          // V: 16 key-rows = 2048 bytes
          const uint64_t bd = make_smem_desc_sw128(v_addr + k * 2048, AT_TILE, 1024);
          umma_f16(tmem_PV, ad, bd, idesc_pv, k > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
      }
    }
  } else if (SPLIT) {
    // ------------------------------------------------------------------ softmax / output warps, two threads per row
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int qi = q0 + row + p.q_offset;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = sP + half * AT_TILE + row * 128;        // this thread's 64 key columns = atom `half`
    const int sw = row & 7;
    float* ex = reinterpret_cast<float*>(sP);               // [2][128] exchange area; aliases P while P is dead
    volatile uint32_t* flag = tmem_slot + 1;
    const int cbase = half * 64;

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * AT_BN;
      const bool need_mask = (kbase + AT_BN > p.Lk) || (CAUSAL && (kbase + AT_BN - 1 > q0 + p.q_offset));
      auto sweep = [&](float m_ref, float& lsum_out, float& mx_out, bool write) {
        float lsum = 0.f, mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_S + lane_off + cbase + c, r);
          tmem_ld_wait();
          float pr[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            float raw = __uint_as_float(r[t]);
            if (need_mask) {
              const int kj = kbase + cbase + c + t;
              const bool ok = (kj < p.Lk) && (!CAUSAL || kj <= qi);
              raw = ok ? raw : -INFINITY;
            }
            mx = fmaxf(mx, raw);
            pr[t] = exp2f(fmaf(raw, p.scale_log2, -m_ref));
          }
          if (write) {
            const int cc0 = (c >> 5) * 4;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint4 v;
              v.x = pack16<F16>(pr[t * 8 + 0], pr[t * 8 + 1]);
              v.y = pack16<F16>(pr[t * 8 + 2], pr[t * 8 + 3]);
              v.z = pack16<F16>(pr[t * 8 + 4], pr[t * 8 + 5]);
              v.w = pack16<F16>(pr[t * 8 + 6], pr[t * 8 + 7]);
              *reinterpret_cast<uint4*>(prow + (((cc0 + t) ^ sw) << 4)) = v;
            }
#pragma unroll
            for (int t = 0; t < 32; ++t) lsum += pr[t];
          }
        }
        lsum_out = lsum;
        mx_out = mx * p.scale_log2;
      };
      float lsum = 0.f, mx_tile = -INFINITY, alpha = 1.f, m_use = m;
      bool resweep = true;
      if (j > 0) {
        sweep((m == -INFINITY) ? 0.f : m, lsum, mx_tile, true);
        if (mx_tile > m + 16.f) *flag = (uint32_t)(j + 1);
        named_bar_sync(2, 256);                       // all first sweeps done, flag final for this tile
        resweep = (*flag == (uint32_t)(j + 1));
      } else {
        float dummy;
        sweep(0.f, dummy, mx_tile, false);            // first tile: maximum only (nothing written)
      }
      if (resweep) {
        // agree on the row maximum with the thread that owns the other 64 columns of this row
        ex[half * 128 + row] = mx_tile;
        named_bar_sync(2, 256);
        const float other = ex[(half ^ 1) * 128 + row];
        named_bar_sync(2, 256);                       // everybody has read before P overwrites the area
        const float mx = fmaxf(m, fmaxf(mx_tile, other));
        m_use = (mx == -INFINITY) ? 0.f : mx;
        alpha = (m == -INFINITY) ? 0.f : exp2f(m - m_use);
        float unused;
        sweep(m_use, lsum, unused, true);
      }
      tc_fence_before();
      mbar_arrive(s_free);
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      l = l * alpha + lsum;
      m = m_use;
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld_32x32(tmem_PV + lane_off + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) o[t] = o[t] * alpha + __uint_as_float(r[t]);
      }
      tc_fence_before();
      mbar_arrive(o_free);
    }
    // the last P V product has completed (o_full): P is dead, use it to add the two partial row sums
    ex[half * 128 + row] = l;
    named_bar_sync(2, 256);
    l += ex[(half ^ 1) * 128 + row];
    const int qrow = q0 + row;
    if (qrow < p.Lq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      bf16* dst = p.O + ((long long)b * p.Lq + qrow) * p.o_row + (long long)h * AT_D + half * 32;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint4 v;
        v.x = pack16<F16>(o[t * 8 + 0] * inv, o[t * 8 + 1] * inv);
        v.y = pack16<F16>(o[t * 8 + 2] * inv, o[t * 8 + 3] * inv);
        v.z = pack16<F16>(o[t * 8 + 4] * inv, o[t * 8 + 5] * inv);
        v.w = pack16<F16>(o[t * 8 + 6] * inv, o[t * 8 + 7] * inv);
        reinterpret_cast<uint4*>(dst)[t] = v;
      }
      if (half == 0 && p.lse != nullptr)
        p.lse[((long long)b * p.H + h) * p.Lq + qrow] = (m + log2f(fmaxf(l, 1e-30f))) * 0.6931471805599453f;
    }
  } else {
    // ------------------------------------------------------------------ softmax / output warps
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // TMEM lane == query row inside the tile
    const int qi = q0 + row + p.q_offset;          // absolute position (for the causal mask)
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float o[AT_D];
#pragma unroll
    for (int i = 0; i < AT_D; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = sP + row * 128;
    const int sw = row & 7;

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * AT_BN;
      const bool need_mask = (kbase + AT_BN > p.Lk) || (CAUSAL && (kbase + AT_BN - 1 > q0 + p.q_offset));
      // Probabilities for this tile.  One sweep over S: p = exp2(s*c - m_ref), written as bf16 into the
      // swizzled P tile, summed in fp32, while the tile's own row maximum is tracked on the side.  m_ref is
      // the running maximum; only when a row's tile maximum exceeds it by more than 2^16 (never after the
      // first tiles in practice) does the warp take the classic second sweep with the new maximum and
      // rescale O / l.  The first tile (m_ref = -inf) always takes both sweeps.  (The separate max sweep
      // over TMEM cost a quarter of the softmax warps' instructions, and they are the kernel's bottleneck.)
      auto sweep = [&](float m_ref, float& lsum_out, float& mx_out, bool write) {
        // four independent partial maxima / sums: one accumulator would be a 128-deep dependent chain per tile and row
        // (~4 cycles per link), longer than the tile's 512 tensor cycles with only two softmax warps per scheduler
        float ls4[4] = {0.f, 0.f, 0.f, 0.f}, mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        // (issuing the TMEM load of chunk c + 1 under the work on chunk c was measured SLOWER: 0.240 vs 0.203 ms at B32 L1024
        //  H12 -- the second 32-register buffer pushes the kernel to 168 registers with spills)
#pragma unroll 1
        for (int c = 0; c < AT_BN; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_S + lane_off + c, r);
          tmem_ld_wait();
          float pr[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            float raw = __uint_as_float(r[t]);
            if (need_mask) {
              const int kj = kbase + c + t;
              const bool ok = (kj < p.Lk) && (!CAUSAL || kj <= qi);
              raw = ok ? raw : -INFINITY;
            }
            mx4[t & 3] = fmaxf(mx4[t & 3], raw);                   // maximum of the RAW scores (scale > 0)
            pr[t] = exp2f(fmaf(raw, p.scale_log2, -m_ref));        // exp2(-inf) = 0 for masked columns
          }
          if (write) {
            uint8_t* atom = prow + (c >> 6) * AT_TILE;
            const int cc0 = ((c >> 5) & 1) * 4;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint4 v;
              v.x = pack16<F16>(pr[t * 8 + 0], pr[t * 8 + 1]);
              v.y = pack16<F16>(pr[t * 8 + 2], pr[t * 8 + 3]);
              v.z = pack16<F16>(pr[t * 8 + 4], pr[t * 8 + 5]);
              v.w = pack16<F16>(pr[t * 8 + 6], pr[t * 8 + 7]);
              *reinterpret_cast<uint4*>(atom + (((cc0 + t) ^ sw) << 4)) = v;
            }
#pragma unroll
            for (int t = 0; t < 32; ++t) ls4[t & 3] += pr[t];
          }
        }
        lsum_out = (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
        mx_out = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;          // (-inf stays -inf)
      };
      float lsum = 0.f, mx_tile = -INFINITY, alpha = 1.f, m_use = m;
      bool resweep = true;
      if (j > 0) {
        // every lane of the warp has a finite running maximum after the first tile unless its row is fully
        // masked so far (m == -inf is represented by m_use = 0 below, with l = 0)
        // (a row that has been fully masked so far has m = -inf: sweep against 0 and let the test below
        //  send the warp through the second sweep as soon as the row sees a finite score)
        sweep((m == -INFINITY) ? 0.f : m, lsum, mx_tile, true);
        resweep = __any_sync(0xffffffffu, mx_tile > m + 16.f) != 0;
      } else {
        float dummy;
        sweep(0.f, dummy, mx_tile, false);         // first tile: maximum only
      }
      if (resweep) {
        const float mx = fmaxf(m, mx_tile);
        m_use = (mx == -INFINITY) ? 0.f : mx;
        alpha = (m == -INFINITY) ? 0.f : exp2f(m - m_use);      // first tile: nothing accumulated yet
        float unused;
        sweep(m_use, lsum, unused, true);
      }
      tc_fence_before();
      mbar_arrive(s_free);               // S columns may be overwritten by the next Q K^T
      fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the MMA (async proxy)
      mbar_arrive(p_full);
      l = l * alpha + lsum;
      m = m_use;
      // O = O * alpha + P V
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < AT_D; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_PV + lane_off + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) o[c + t] = o[c + t] * alpha + __uint_as_float(r[t]);
      }
      tc_fence_before();
      mbar_arrive(o_free);
    }

    const int qrow = q0 + row;
    if (qrow < p.Lq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      bf16* dst = p.O + ((long long)b * p.Lq + qrow) * p.o_row + (long long)h * AT_D;
#pragma unroll
      for (int t = 0; t < AT_D / 8; ++t) {
        uint4 v;
        v.x = pack16<F16>(o[t * 8 + 0] * inv, o[t * 8 + 1] * inv);
        v.y = pack16<F16>(o[t * 8 + 2] * inv, o[t * 8 + 3] * inv);
        v.z = pack16<F16>(o[t * 8 + 4] * inv, o[t * 8 + 5] * inv);
        v.w = pack16<F16>(o[t * 8 + 6] * inv, o[t * 8 + 7] * inv);
        reinterpret_cast<uint4*>(dst)[t] = v;
      }
      if (p.lse != nullptr)
        p.lse[((long long)b * p.H + h) * p.Lq + qrow] = (m + log2f(fmaxf(l, 1e-30f))) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 attn_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || ptr == nullptr || qres != cudaDriverEntryPointSuccess) {
      fprintf(stderr, "[tnn_b200] attention: cuTensorMapEncodeTiled unavailable\n");
      abort();
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// [B, L, H, 64] view with arbitrary token / batch strides (elements): dims {64, H, L, B}.
CUtensorMap make_head_map(const void* ptr, int B, int L, int H, long long tok_stride, long long batch_stride, int f16) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)AT_D, (cuuint64_t)H, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t str[3] = {(cuuint64_t)AT_D * 2, (cuuint64_t)tok_stride * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[4] = {64, 1, 128, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = attn_encode_fn()(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, str, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[tnn_b200] attention: cuTensorMapEncodeTiled failed (%d) B=%d L=%d H=%d tok=%lld batch=%lld ptr=%p\n",
            (int)r, B, L, H, tok_stride, batch_stride, ptr);
    abort();
  }
  return m;
}

void attention_fwd_bf16(const void* q, long long q_tok, long long q_batch, const void* k, long long k_tok,
                        long long k_batch, const void* v, long long v_tok, long long v_batch, void* out,
                        float* lse, int B, int H, int Lq, int Lk, int causal, int q_offset, cudaStream_t stream, int f16) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0) return;
  if ((q_tok % 8) || (k_tok % 8) || (v_tok % 8) || (q_batch % 8) || (k_batch % 8) || (v_batch % 8) ||
      (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    fprintf(stderr, "[tnn_b200] attention_fwd_bf16: operands must be 16-byte aligned (strides %% 8 == 0)\n");
    abort();
  }
  CUtensorMap tmQ = make_head_map(q, B, Lq, H, q_tok, q_batch, f16);
  CUtensorMap tmK = make_head_map(k, B, Lk, H, k_tok, k_batch, f16);
  CUtensorMap tmV = make_head_map(v, B, Lk, H, v_tok, v_batch, f16);
  AttnParams p{};
  p.B = B;
  p.H = H;
  p.Lq = Lq;
  p.Lk = Lk;
  p.causal = causal;
  p.q_offset = q_offset;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)AT_D);
  p.O = reinterpret_cast<bf16*>(out);
  p.o_row = (long long)H * AT_D;
  p.lse = lse;
  static PerDeviceOnce configured;
  static int split = 0;
  if (configured.begin()) {
    cudaFuncSetAttribute(attn_fwd_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    cudaFuncSetAttribute(attn_fwd_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    cudaFuncSetAttribute(attn_fwd_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    cudaFuncSetAttribute(attn_fwd_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    cudaFuncSetAttribute(attn_fwd_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    cudaFuncSetAttribute(attn_fwd_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    const char* e = getenv("TNN_ATTN_SPLIT");
    split = (e != nullptr) ? atoi(e) : 0;
    configured.done();
  }
  dim3 grid((Lq + AT_BM - 1) / AT_BM, H, B);
  if (f16) {     // fp16 operands and output (the reference's flash attention is fp16-only): same kernel, other operand format
    if (causal)
      attn_fwd_kernel<true, false, true><<<grid, AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
    else
      attn_fwd_kernel<false, false, true><<<grid, AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  } else if (split) {
    if (causal)
      attn_fwd_kernel<true, true, false><<<grid, 320, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
    else
      attn_fwd_kernel<false, true, false><<<grid, 320, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  } else if (causal)
    attn_fwd_kernel<true, false, false><<<grid, AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  else
    attn_fwd_kernel<false, false, false><<<grid, AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  TNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------
// decode: one query token per (batch, head) against the KV cache
// ---------------------------------------------------------------------------------------------
static constexpr int DEC_THREADS = 256;
static constexpr int DEC_GROUPS = DEC_THREADS / 8;   // 8 lanes per key

__global__ void __launch_bounds__(DEC_THREADS)
    attn_decode_kernel(const bf16* __restrict__ q, long long q_batch, const bf16* __restrict__ kc,
                       const bf16* __restrict__ vc, long long tok_stride, long long batch_stride,
                       bf16* __restrict__ out, int H, int len, float scale_log2) {
  const int h = blockIdx.x, b = blockIdx.y;
  const int g = threadIdx.x >> 3;      // key group 0..31
  const int sub = threadIdx.x & 7;     // which 8 dims of the 64
  const bf16* qp = q + (long long)b * q_batch + h * AT_D + sub * 8;
  float qf[8];
  bf8_unpack(*reinterpret_cast<const uint4*>(qp), qf);
#pragma unroll
  for (int i = 0; i < 8; ++i) qf[i] *= scale_log2;
  const bf16* kb = kc + (long long)b * batch_stride + h * AT_D + sub * 8;
  const bf16* vb = vc + (long long)b * batch_stride + h * AT_D + sub * 8;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  for (int j = g; j < len; j += DEC_GROUPS) {
    float kf[8], vf[8];
    bf8_unpack(__ldg(reinterpret_cast<const uint4*>(kb + (long long)j * tok_stride)), kf);
    bf8_unpack(__ldg(reinterpret_cast<const uint4*>(vb + (long long)j * tok_stride)), vf);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += qf[i] * kf[i];
    // the 4 key groups of a warp run different trip counts: shuffle inside the 8-lane group only
    s += __shfl_xor_sync(gmask, s, 1);
    s += __shfl_xor_sync(gmask, s, 2);
    s += __shfl_xor_sync(gmask, s, 4);
    const float mn = fmaxf(m, s);
    const float a = exp2f(m - mn), e = exp2f(s - mn);
    l = l * a + e;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = acc[i] * a + e * vf[i];
    m = mn;
  }
  __shared__ float sm_m[DEC_GROUPS], sm_l[DEC_GROUPS], sm_o[DEC_GROUPS][AT_D];
  if (sub == 0) {
    sm_m[g] = m;
    sm_l[g] = l;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm_o[g][sub * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < AT_D) {
    float M = -INFINITY;
    for (int i = 0; i < DEC_GROUPS; ++i) M = fmaxf(M, sm_m[i]);
    float L = 0.f, O = 0.f;
    for (int i = 0; i < DEC_GROUPS; ++i) {
      const float w = (sm_m[i] == -INFINITY) ? 0.f : exp2f(sm_m[i] - M);
      L += sm_l[i] * w;
      O += sm_o[i][threadIdx.x] * w;
    }
    out[((long long)b * H + h) * AT_D + threadIdx.x] = __float2bfloat16(L > 0.f ? O / L : 0.f);
  }
}

// Decode step with the position held in DEVICE memory (so the whole step is CUDA-graph
// replayable): appends this token's K / V rows to the cache at *pos_ptr, then attends over
// positions [0, *pos_ptr].  qkv is the packed projection of the current token [B, 3, H, 64].
__global__ void __launch_bounds__(DEC_THREADS)
    attn_decode_append_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kc, bf16* __restrict__ vc,
                              long long tok_stride, long long batch_stride, bf16* __restrict__ out, int H,
                              const int* __restrict__ pos_ptr, float scale_log2) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: dependents may start their prologue
  const int h = blockIdx.x, b = blockIdx.y;
  const int pos = *pos_ptr;
  const int len = pos + 1;
  const bf16* row = qkv + (long long)b * 3 * H * AT_D;
  // threads 0..7 append K, 8..15 append V (16 bytes each)
  if (threadIdx.x < 16) {
    const int which = threadIdx.x >> 3, sub8 = threadIdx.x & 7;
    const uint4 val = *reinterpret_cast<const uint4*>(row + (long long)(1 + which) * H * AT_D + h * AT_D + sub8 * 8);
    bf16* dst = (which ? vc : kc) + (long long)b * batch_stride + (long long)pos * tok_stride + h * AT_D + sub8 * 8;
    *reinterpret_cast<uint4*>(dst) = val;
  }
  __syncthreads();
  const int g = threadIdx.x >> 3;
  const int sub = threadIdx.x & 7;
  float qf[8];
  bf8_unpack(*reinterpret_cast<const uint4*>(row + h * AT_D + sub * 8), qf);
#pragma unroll
  for (int i = 0; i < 8; ++i) qf[i] *= scale_log2;
  const bf16* kb = kc + (long long)b * batch_stride + h * AT_D + sub * 8;
  const bf16* vb = vc + (long long)b * batch_stride + h * AT_D + sub * 8;
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  for (int j = g; j < len; j += DEC_GROUPS) {
    float kf[8], vf[8];
    // plain loads: row `pos` was written by this CTA just above (not via the read-only path)
    bf8_unpack(*reinterpret_cast<const uint4*>(kb + (long long)j * tok_stride), kf);
    bf8_unpack(*reinterpret_cast<const uint4*>(vb + (long long)j * tok_stride), vf);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += qf[i] * kf[i];
    s += __shfl_xor_sync(gmask, s, 1);
    s += __shfl_xor_sync(gmask, s, 2);
    s += __shfl_xor_sync(gmask, s, 4);
    const float mn = fmaxf(m, s);
    const float a = exp2f(m - mn), e = exp2f(s - mn);
    l = l * a + e;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = acc[i] * a + e * vf[i];
    m = mn;
  }
  __shared__ float sm_m[DEC_GROUPS], sm_l[DEC_GROUPS], sm_o[DEC_GROUPS][AT_D];
  if (sub == 0) {
    sm_m[g] = m;
    sm_l[g] = l;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm_o[g][sub * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < AT_D) {
    float M = -INFINITY;
    for (int i = 0; i < DEC_GROUPS; ++i) M = fmaxf(M, sm_m[i]);
    float L = 0.f, O = 0.f;
    for (int i = 0; i < DEC_GROUPS; ++i) {
      const float w = (sm_m[i] == -INFINITY) ? 0.f : exp2f(sm_m[i] - M);
      L += sm_l[i] * w;
      O += sm_o[i][threadIdx.x] * w;
    }
    out[((long long)b * H + h) * AT_D + threadIdx.x] = __float2bfloat16(L > 0.f ? O / L : 0.f);
  }
}

void attention_decode_append_bf16(const void* qkv, void* kcache, void* vcache, long long tok_stride,
                                  long long batch_stride, void* out, int B, int H, const int* pos_ptr,
                                  cudaStream_t stream) {
  if (B <= 0 || H <= 0) return;
  dim3 grid(H, B);
  attn_decode_append_kernel<<<grid, DEC_THREADS, 0, stream>>>(
      reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(kcache), reinterpret_cast<bf16*>(vcache), tok_stride,
      batch_stride, reinterpret_cast<bf16*>(out), H, pos_ptr, 1.4426950408889634f / sqrtf((float)AT_D));
  TNN_CHECK_LAUNCH();
}

// Prefill: copy the K and V slots of the packed projection [B, L, 3, E] into the caches at [pos, pos+L)
// (16-byte vectors; replaces two strided framework copy kernels per layer, ~55 us each at B=32 L=1024)
__global__ void __launch_bounds__(256)
    k_kv_store(const bf16* __restrict__ qkv, bf16* __restrict__ kc, bf16* __restrict__ vc, long long n_vec, int E8,
               int L, long long tok_stride, long long batch_stride, int pos) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_vec) return;
  const int c = (int)(gid % E8);
  long long t = gid / E8;
  const int which = (int)(t & 1);            // 0 = K, 1 = V
  t >>= 1;
  const int l = (int)(t % L);
  const long long b = t / L;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(qkv + ((b * L + l) * 3 + 1 + which) * (long long)E8 * 8) + c);
  bf16* dst = (which ? vc : kc) + b * batch_stride + (long long)(pos + l) * tok_stride;
  reinterpret_cast<uint4*>(dst)[c] = v;
}
void kv_cache_store_bf16(const void* qkv, void* kcache, void* vcache, long long tok_stride, long long batch_stride,
                         int B, int L, int E, int pos, cudaStream_t stream) {
  if (B <= 0 || L <= 0) return;
  const long long n_vec = (long long)B * L * 2 * (E / 8);
  k_kv_store<<<(unsigned)((n_vec + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(kcache), reinterpret_cast<bf16*>(vcache), n_vec, E / 8, L,
      tok_stride, batch_stride, pos);
  TNN_CHECK_LAUNCH();
}

// x[b, :] = wte[tok[b], :] + wpe[*pos, :]  (bf16 out), one CTA per sequence
__global__ void __launch_bounds__(256)
    k_embed_decode(const int* __restrict__ tok, const float* __restrict__ wte, const float* __restrict__ wpe,
                   bf16* __restrict__ x, int D, int V, const int* __restrict__ pos_ptr) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: dependents may start their prologue
  const int b = blockIdx.x;
  const int pos = *pos_ptr;
  int id = tok[b];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    x[(long long)b * D + d] = __float2bfloat16(wte[(long long)id * D + d] + wpe[(long long)pos * D + d]);
}
void embed_decode(const int* tok, const float* wte, const float* wpe, void* x, int B, int D, int V, const int* pos_ptr,
                  cudaStream_t stream) {
  k_embed_decode<<<B, 256, 0, stream>>>(tok, wte, wpe, reinterpret_cast<bf16*>(x), D, V, pos_ptr);
  TNN_CHECK_LAUNCH();
}

// end of a decode step: publish the sampled tokens, log them, advance the device position
__global__ void k_decode_epilogue(const int* __restrict__ next, int* __restrict__ tok, int* __restrict__ history,
                                  int B, int hist_cap, int* pos_ptr, int* step_ptr) {
  const int step = *step_ptr;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    tok[b] = next[b];
    if (step < hist_cap) history[(long long)step * B + b] = next[b];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *pos_ptr += 1;
    *step_ptr = step + 1;
  }
}
void decode_epilogue(const int* next, int* tok, int* history, int B, int hist_cap, int* pos_ptr, int* step_ptr,
                     cudaStream_t stream) {
  k_decode_epilogue<<<1, 256, 0, stream>>>(next, tok, history, B, hist_cap, pos_ptr, step_ptr);
  TNN_CHECK_LAUNCH();
}

void attention_decode_bf16(const void* q, long long q_batch, const void* kcache, const void* vcache,
                           long long tok_stride, long long batch_stride, void* out, int B, int H, int len,
                           cudaStream_t stream) {
  if (B <= 0 || H <= 0 || len <= 0) return;
  dim3 grid(H, B);
  attn_decode_kernel<<<grid, DEC_THREADS, 0, stream>>>(
      reinterpret_cast<const bf16*>(q), q_batch, reinterpret_cast<const bf16*>(kcache),
      reinterpret_cast<const bf16*>(vcache), tok_stride, batch_stride, reinterpret_cast<bf16*>(out), H, len,
      1.4426950408889634f / sqrtf((float)AT_D));
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
