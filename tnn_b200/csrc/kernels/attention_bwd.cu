// Flash-attention backward for sm_100a (head_dim 64, bf16): dQ, dK, dV from dO without ever
// materialising the [L, L] probability matrix (reference: the non-flash block recomputes the full
// score tensor and runs four batched cuBLAS GEMMs + softmax/mask kernels,
// src/nn/blocks_impl/attention_block.cpp:183-297; the flash block calls cuDNN SDPA backward,
// cuda/cudnn_flash_attention_ops.cu:209).
//
// One CTA owns one 128-key tile (K_j, V_j stay in shared memory for the CTA's lifetime) of one
// (batch, head) and walks the query tiles i that can see it (i >= j when causal).  320 threads:
//   warp 0       TMA producer: K_j, V_j once; Q_i, dO_i into a 2-stage ring
//   warp 1       TMEM allocator + tcgen05.mma issuer, five MMAs per query tile
//                  S   = Q_i K_j^T            128x128x64   TMEM cols [0,128)
//                  dP  = dO_i V_j^T           128x128x64   TMEM cols [128,256)
//                  dV += P^T dO_i             128x64x128   TMEM cols [384,448)  (A = P^T, MN-major)
//                  dK += dS^T Q_i             128x64x128   TMEM cols [320,384)  (A = dS^T, MN-major)
//                  dQ_i = dS K_j              128x64x128   TMEM cols [256,320) / [448,512) (ping-pong)
//   warps 2..9   thread (row, half): P = exp2(S*c - lse), dS = P*(dP - delta)*scale for 64 of the
//                128 key columns of query row `row`, both written as bf16 into 128B-swizzled smem
//                tiles that serve as K-major A operand (dS K) AND, read through an MN-major
//                descriptor, as the transposed operand (P^T dO, dS^T Q) -- no transpose pass.
//                dQ_i leaves TMEM as fp32 vector reductions (red.global.add.v4.f32) into dq_acc,
//                overlapped with the next tile's MMAs; dK / dV are written once at the end.
// All 512 TMEM columns and 160 KB of shared memory are in use: one CTA per SM.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <math.h>

#include "attention.cuh"
#include "common.cuh"
#include "sm100.cuh"

namespace tnn {
using namespace sm100;

CUtensorMap make_head_map(const void* ptr, int B, int L, int H, long long tok_stride, long long batch_stride, int f16);

static constexpr int AB_TILE = 128 * 128;     // bytes of one [128 rows][64 bf16] tile
static constexpr int AB_THREADS = 320;
static constexpr int AB_SMEM = 2 * AB_TILE /*K,V*/ + 4 * AB_TILE /*Q,dO x2*/ + 2 * AB_TILE /*P*/ +
                               2 * AB_TILE /*dS*/ + 256 /*barriers*/;

struct AttnBwdParams {
  int B, H, L;
  float scale_log2;      // log2(e) / sqrt(D)
  float scale;           // 1 / sqrt(D)
  const float* lse;      // [B, H, L] natural log
  const float* delta;    // [B, H, L] rowsum(dO * O)
  float* dq_acc;         // fp32 [B, L, H, 64]
  bf16* dk;              // [B, L, H, 64] views (token / batch strides in elements)
  bf16* dv;
  long long dkv_tok, dkv_batch;
};

template <bool CAUSAL, bool F16>
__global__ void __launch_bounds__(AB_THREADS, 1)
    attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                    const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQ = sV + AB_TILE;            // 2 stages
  uint8_t* sdO = sQ + 2 * AB_TILE;       // 2 stages
  uint8_t* sP = sdO + 2 * AB_TILE;       // [128 q][128 keys] bf16 = two 64-key atoms
  uint8_t* sdS = sP + 2 * AB_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * AB_TILE);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* pds_full = bars + 7;
  uint64_t* dq_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int jb = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nq = (p.L + 127) / 128;
  const int i_start = CAUSAL ? jb : 0;
  const int n_it = nq - i_start;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(s_free, 256);
    mbar_init(pds_full, 256);
    mbar_init(dq_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_dP = tmem_base + 128;
  const uint32_t tmem_dK = tmem_base + 320;
  const uint32_t tmem_dV = tmem_base + 384;
  // dQ ping-pong: buffer 0 at +256, buffer 1 at +448

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_4d(sK, &tmK, kv_full, 0, h, jb * 128, b);
      tma_load_4d(sV, &tmV, kv_full, 0, h, jb * 128, b);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const int i = i_start + it;
        mbar_wait(&qdo_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[s], 2 * AB_TILE);
        tma_load_4d(sQ + s * AB_TILE, &tmQ, &qdo_full[s], 0, h, i * 128, b);
        tma_load_4d(sdO + s * AB_TILE, &tmdO, &qdo_full[s], 0, h, i * 128, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_16<F16>(128, 128, 0u, 0u);    // K-major x K-major
      constexpr uint32_t idesc_t = make_idesc_16<F16>(128, 64, 1u, 1u);     // MN-major A (transposed tile) x MN-major B
      constexpr uint32_t idesc_q = make_idesc_16<F16>(128, 64, 0u, 1u);     // K-major A x MN-major B
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      mbar_wait(kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t q_addr = smem_u32(sQ + s * AB_TILE);
        const uint32_t do_addr = smem_u32(sdO + s * AB_TILE);
        mbar_wait(&qdo_full[s], (it >> 1) & 1);
        if (it > 0) mbar_wait(s_free, (it - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_S, make_smem_desc_sw128(q_addr + k * 32, 0, 1024),
                   make_smem_desc_sw128(k_addr + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_dP, make_smem_desc_sw128(do_addr + k * 32, 0, 1024),
                   make_smem_desc_sw128(v_addr + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(sdp_full);
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
        const uint32_t tmem_dQ = tmem_base + ((it & 1) ? 448u : 256u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 query rows per step = 2048 bytes in every [q][64] atom
          umma_f16(tmem_dV, make_smem_desc_sw128(p_addr + k * 2048, AB_TILE, 1024),
                   make_smem_desc_sw128(do_addr + k * 2048, AB_TILE, 1024), idesc_t, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem_dK, make_smem_desc_sw128(ds_addr + k * 2048, AB_TILE, 1024),
                   make_smem_desc_sw128(q_addr + k * 2048, AB_TILE, 1024), idesc_t, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // keys [16k, 16k+16): atom k/4, byte offset (k%4)*32; K_j rows 16k..
          umma_f16(tmem_dQ, make_smem_desc_sw128(ds_addr + (k >> 2) * AB_TILE + (k & 3) * 32, 0, 1024),
                   make_smem_desc_sw128(k_addr + k * 2048, AB_TILE, 1024), idesc_q, k > 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(&qdo_empty[s]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                    // which 64 key columns / 32 output columns
    const int row = quad * 32 + lane;                    // TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int sw = row & 7;
    const long long bh = (long long)b * p.H + h;
    const float log2e = 1.4426950408889634f;
    uint8_t* prow = sP + half * AB_TILE + row * 128;
    uint8_t* dsrow = sdS + half * AB_TILE + row * 128;
    const int kbase = jb * 128 + half * 64;

    // dq_full(it_done) must already have been waited on by the caller (see the loop: the wait sits
    // BEFORE this thread's pds_full arrival, otherwise the MMA warp could complete the NEXT phase of
    // dq_full first and a parity wait two phases behind would never return)
    auto flush_dq = [&](int it_done) {
      tc_fence_after();
      const int qi = (i_start + it_done) * 128 + row;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((it_done & 1) ? 448u : 256u) + lane_off + half * 32, r);
      tmem_ld_wait();
      if (qi < p.L) {
        float* dst = p.dq_acc + (((long long)b * p.L + qi) * p.H + h) * 64 + half * 32;
#pragma unroll
        for (int t = 0; t < 8; ++t)
          red_add_v4(dst + 4 * t, __uint_as_float(r[4 * t]), __uint_as_float(r[4 * t + 1]),
                     __uint_as_float(r[4 * t + 2]), __uint_as_float(r[4 * t + 3]));
      }
    };

    for (int it = 0; it < n_it; ++it) {
      const int i = i_start + it;
      const int qi = i * 128 + row;
      float lse2 = 0.f, dl = 0.f;
      if (qi < p.L) {
        lse2 = __ldg(p.lse + bh * p.L + qi) * log2e;
        dl = __ldg(p.delta + bh * p.L + qi);
      }
      const bool need_mask = (i * 128 + 128 > p.L) || (jb * 128 + 128 > p.L) || (CAUSAL && i == jb);
      mbar_wait(sdp_full, it & 1);
      // S/dP of tile `it` complete => every earlier MMA (incl. dQ of tile it-1) has completed: this
      // wait returns at once, and dQ(it) cannot be issued before this thread arrives on pds_full
      if (it > 0) mbar_wait(dq_full, (it - 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t rs[32], rp[32];
        tmem_ld_32x32(tmem_S + lane_off + half * 64 + c, rs);
        tmem_ld_32x32(tmem_dP + lane_off + half * 64 + c, rp);
        tmem_ld_wait();
        float pr[32], ds[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          float e = exp2f(__uint_as_float(rs[t]) * p.scale_log2 - lse2);
          if (need_mask) {
            const int kj = kbase + c + t;
            const bool ok = (qi < p.L) && (kj < p.L) && (!CAUSAL || kj <= qi);
            e = ok ? e : 0.f;
          }
          pr[t] = e;
          ds[t] = e * (__uint_as_float(rp[t]) - dl) * p.scale;
        }
        const int cc0 = (c >> 5) * 4;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 v, w;
          v.x = pack16<F16>(pr[t * 8 + 0], pr[t * 8 + 1]);
          v.y = pack16<F16>(pr[t * 8 + 2], pr[t * 8 + 3]);
          v.z = pack16<F16>(pr[t * 8 + 4], pr[t * 8 + 5]);
          v.w = pack16<F16>(pr[t * 8 + 6], pr[t * 8 + 7]);
          w.x = pack16<F16>(ds[t * 8 + 0], ds[t * 8 + 1]);
          w.y = pack16<F16>(ds[t * 8 + 2], ds[t * 8 + 3]);
          w.z = pack16<F16>(ds[t * 8 + 4], ds[t * 8 + 5]);
          w.w = pack16<F16>(ds[t * 8 + 6], ds[t * 8 + 7]);
          const int off = ((cc0 + t) ^ sw) << 4;
          *reinterpret_cast<uint4*>(prow + off) = v;
          *reinterpret_cast<uint4*>(dsrow + off) = w;
        }
      }
      tc_fence_before();
      mbar_arrive(s_free);               // S / dP columns may be overwritten by the next tile
      fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the MMA (async proxy)
      mbar_arrive(pds_full);
      if (it > 0) {
        flush_dq(it - 1);                // overlaps this tile's dV / dK / dQ MMAs
        tc_fence_before();
      }
    }
    if (n_it > 0) {
      mbar_wait(dq_full, (n_it - 1) & 1);   // last commit: dK / dV / dQ accumulators are final
      flush_dq(n_it - 1);
      const int kj = jb * 128 + row;
      uint32_t rk[32], rv[32];
      tmem_ld_32x32(tmem_dK + lane_off + half * 32, rk);
      tmem_ld_32x32(tmem_dV + lane_off + half * 32, rv);
      tmem_ld_wait();
      if (kj < p.L) {
        const long long off = (long long)b * p.dkv_batch + (long long)kj * p.dkv_tok + (long long)h * 64 + half * 32;
        uint4* dk = reinterpret_cast<uint4*>(p.dk + off);
        uint4* dv = reinterpret_cast<uint4*>(p.dv + off);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 a, c;
          a.x = pack16<F16>(__uint_as_float(rk[8 * t + 0]), __uint_as_float(rk[8 * t + 1]));
          a.y = pack16<F16>(__uint_as_float(rk[8 * t + 2]), __uint_as_float(rk[8 * t + 3]));
          a.z = pack16<F16>(__uint_as_float(rk[8 * t + 4]), __uint_as_float(rk[8 * t + 5]));
          a.w = pack16<F16>(__uint_as_float(rk[8 * t + 6]), __uint_as_float(rk[8 * t + 7]));
          c.x = pack16<F16>(__uint_as_float(rv[8 * t + 0]), __uint_as_float(rv[8 * t + 1]));
          c.y = pack16<F16>(__uint_as_float(rv[8 * t + 2]), __uint_as_float(rv[8 * t + 3]));
          c.z = pack16<F16>(__uint_as_float(rv[8 * t + 4]), __uint_as_float(rv[8 * t + 5]));
          c.w = pack16<F16>(__uint_as_float(rv[8 * t + 6]), __uint_as_float(rv[8 * t + 7]));
          dk[t] = a;
          dv[t] = c;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// delta[b, h, l] = sum_d dO[b, l, h, d] * O[b, l, h, d]; 8 lanes per (b, l, h), 16-byte loads
template <bool F16>
__device__ __forceinline__ void unpack16x8(const uint4& q, float (&f)[8]) {
  if constexpr (F16) {
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  } else {
    bf8_unpack(q, f);
  }
}
template <bool F16>
__global__ void __launch_bounds__(256)
    k_attn_delta(const bf16* __restrict__ dO, const bf16* __restrict__ O, float* __restrict__ delta, long long rows,
                 int H, int L) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one per 8 elements
  const long long head_row = gid >> 3;                                      // (b*L + l)*H + h
  float s = 0.f;
  if (head_row < rows) {
    float a[8], c[8];
    unpack16x8<F16>(__ldg(reinterpret_cast<const uint4*>(dO) + gid), a);
    unpack16x8<F16>(__ldg(reinterpret_cast<const uint4*>(O) + gid), c);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] * c[i];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if ((threadIdx.x & 7) == 0 && head_row < rows) {
    const long long bl = head_row / H;
    const int h = (int)(head_row - bl * H);
    const long long b = bl / L;
    const int l = (int)(bl - b * L);
    delta[(b * H + h) * L + l] = s;
  }
}

// dq_acc fp32 [B, L, H*64] -> bf16 / fp16 view with token / batch strides
template <bool F16>
__global__ void __launch_bounds__(256)
    k_attn_dq_store(const float* __restrict__ acc, bf16* __restrict__ dq, long long n_vec, int E8, int L,
                    long long tok, long long batch) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_vec) return;
  const long long bl = gid / E8;
  const int c = (int)(gid - bl * E8);
  const long long b = bl / L;
  const int l = (int)(bl - b * L);
  const float4 lo = __ldg(reinterpret_cast<const float4*>(acc) + gid * 2);
  const float4 hi = __ldg(reinterpret_cast<const float4*>(acc) + gid * 2 + 1);
  uint4 v;
  v.x = pack16<F16>(lo.x, lo.y);
  v.y = pack16<F16>(lo.z, lo.w);
  v.z = pack16<F16>(hi.x, hi.y);
  v.w = pack16<F16>(hi.z, hi.w);
  *reinterpret_cast<uint4*>(dq + b * batch + (long long)l * tok + (long long)c * 8) = v;
}

void attention_bwd_bf16(const void* q, const void* k, const void* v, long long qkv_tok, long long qkv_batch,
                        const void* dout, const void* out, const float* lse, float* delta, float* dq_acc, void* dq,
                        void* dk, void* dv, long long dqkv_tok, long long dqkv_batch, int B, int H, int L, int causal,
                        cudaStream_t stream, int f16) {
  if (B <= 0 || H <= 0 || L <= 0) return;
  if ((qkv_tok % 8) || (qkv_batch % 8) || (dqkv_tok % 8) || (dqkv_batch % 8) ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dq) |
        reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(dq_acc)) & 15)) {
    fprintf(stderr, "[tnn_b200] attention_bwd_bf16: operands must be 16-byte aligned (strides %% 8 == 0)\n");
    abort();
  }
  const long long E = (long long)H * 64;
  {
    const long long rows = (long long)B * L * H;
    const long long threads = rows * 8;
    if (f16)
      k_attn_delta<true><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dout),
                                                                                reinterpret_cast<const bf16*>(out), delta, rows, H, L);
    else
      k_attn_delta<false><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(dout),
                                                                                 reinterpret_cast<const bf16*>(out), delta, rows, H, L);
    TNN_CHECK_LAUNCH();
  }
  cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)B * L * E, stream);
  CUtensorMap tmQ = make_head_map(q, B, L, H, qkv_tok, qkv_batch, f16);
  CUtensorMap tmK = make_head_map(k, B, L, H, qkv_tok, qkv_batch, f16);
  CUtensorMap tmV = make_head_map(v, B, L, H, qkv_tok, qkv_batch, f16);
  CUtensorMap tmdO = make_head_map(dout, B, L, H, E, (long long)L * E, f16);
  AttnBwdParams p{};
  p.B = B;
  p.H = H;
  p.L = L;
  p.scale = 1.f / sqrtf(64.f);
  p.scale_log2 = 1.4426950408889634f * p.scale;
  p.lse = lse;
  p.delta = delta;
  p.dq_acc = dq_acc;
  p.dk = reinterpret_cast<bf16*>(dk);
  p.dv = reinterpret_cast<bf16*>(dv);
  p.dkv_tok = dqkv_tok;
  p.dkv_batch = dqkv_batch;
  static PerDeviceOnce configured;
  if (configured.begin()) {
    cudaFuncSetAttribute(attn_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    cudaFuncSetAttribute(attn_bwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    cudaFuncSetAttribute(attn_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    cudaFuncSetAttribute(attn_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    configured.done();
  }
  dim3 grid((L + 127) / 128, H, B);
  if (f16) {
    if (causal) attn_bwd_kernel<true, true><<<grid, AB_THREADS, AB_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
    else attn_bwd_kernel<false, true><<<grid, AB_THREADS, AB_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
  } else if (causal)
    attn_bwd_kernel<true, false><<<grid, AB_THREADS, AB_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
  else
    attn_bwd_kernel<false, false><<<grid, AB_THREADS, AB_SMEM, stream>>>(tmQ, tmK, tmV, tmdO, p);
  TNN_CHECK_LAUNCH();
  {
    const long long n_vec = (long long)B * L * E / 8;
    if (f16)
      k_attn_dq_store<true><<<(unsigned)((n_vec + 255) / 256), 256, 0, stream>>>(dq_acc, reinterpret_cast<bf16*>(dq), n_vec, (int)(E / 8), L,
                                                                                dqkv_tok, dqkv_batch);
    else
      k_attn_dq_store<false><<<(unsigned)((n_vec + 255) / 256), 256, 0, stream>>>(dq_acc, reinterpret_cast<bf16*>(dq), n_vec, (int)(E / 8), L,
                                                                                 dqkv_tok, dqkv_batch);
    TNN_CHECK_LAUNCH();
  }
}

}  // namespace tnn
