// Persistent whole-model decode step: see decode_mega.cuh.  The GEMM inner loop is the weight-streaming scheme of
// skinny_gemm.cu (every lane's 16-byte weight load IS its mma.sync B fragment for two MMAs; x staged in padded shared
// memory); what is new is everything around it: grid-wide arrive/wait barriers, weight prefetch across the barrier
// and across work units, LayerNorm folded into weights + producer-side row statistics, split-sequence attention with
// a last-arriver merge, fused arg-max / token bookkeeping.
#include <cuda_fp8.h>

#include "common.cuh"
#include "decode_mega.cuh"

namespace tnn {

static constexpr int DM_THREADS = 256;                   // 2 CTAs per SM.  (512 threads x 1 CTA with K <= 3072 staged in one pass measured
                                                         //  slower: 215 KB of shared memory leaves ~13 KB of L1 for the loads in flight)
static constexpr int DM_WARPS = DM_THREADS / 32;
static constexpr int DM_KC = 1024;                       // K chunk staged in shared memory
static constexpr int DM_PAD = 32;                        // +64 bytes per row: conflict-free 16-byte fragment loads
static constexpr int DM_LDS = DM_KC + DM_PAD;
static constexpr int DM_X_BYTES = 32 * DM_LDS * 2;
static constexpr int DM_RED_BYTES = DM_WARPS * 32 * 8 * 4;
static constexpr int DM_SMEM = DM_X_BYTES + DM_RED_BYTES;
static constexpr int DM_UNROLL = 8;                      // 32-wide k steps in flight per lane (one round)
static constexpr int DM_MAX_CHUNKS = 16;                 // sequence splits per (batch, head)
static constexpr int DM_PART = 72;                       // floats per partial: 64 acc + m + l (+ pad)
static constexpr int DM_D = 64;                          // head dim

__device__ __forceinline__ void dm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned int dm_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long dm_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Device-wide barrier split into arrive (end of a phase) and wait (as late as possible in the next phase, after its
// weight loads were issued).  `expected` is the cumulative arrival count; the counter is zeroed before every launch.
// Data that crosses the barrier is written with plain stores (made visible by the release) and read with
// ld.global.cg / cp.async.cg (L2), so no L1 invalidation is needed on the consumer side.
struct DmBarrier {
  unsigned int* ctr;
  unsigned int expected;
  unsigned int nblk;
  bool pending;                                          // an arrive without its wait yet
};
__device__ __forceinline__ void dm_arrive(DmBarrier& b) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(b.ctr) : "memory");
  b.expected += b.nblk;
  b.pending = true;
}
__device__ __forceinline__ void dm_wait(DmBarrier& b) {
  if (!b.pending) return;
  if (threadIdx.x == 0) {
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (dm_ld_acquire(b.ctr) < b.expected) {
      if ((++spins & 0xFFFFu) == 0) {                    // watchdog: a CTA that never arrives would hang the GPU
        const unsigned long long now = dm_timer();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 5000000000ull) {
          printf("[tnn_b200] decode_mega: barrier timeout (cta %d, expected %u, have %u)\n", blockIdx.x, b.expected, dm_ld_acquire(b.ctr));
          __trap();
        }
      }
    }
  }
  __syncthreads();
  b.pending = false;
}

__device__ __noinline__ float dm_gelu(float v) {
  const float k0c = 0.7978845608028654f, k1c = 0.044715f;
  return 0.5f * v * (1.f + tanhf(k0c * (v + k1c * v * v * v)));
}

struct DmGemm {
  const bf16* x; int ldx;            // activations [M, K] written by other CTAs earlier in this launch
  const bf16* W;                     // [N, K] bf16 (LayerNorm-folded when csum != null)
  const float* bias;                 // [N] (folded) or null
  const float* csum;                 // [N] column sums of the folded W: the epilogue applies the LayerNorm of x
  const float* stats_in; float eps;  // [32][2] (sum, sum of squares) of the rows of x (with csum)
  const bf16* residual; int ldr;
  bf16* out; int ldo;
  float* stats_out;                  // [32][2]: += (sum, sum of squares) of the stored (bf16-rounded) rows, or null
  int M, N, K, act;
};

// out[M, N] = act(LN?(x) . W^T + bias) + residual.  Units of NT*8 output columns are dealt round-robin to the CTAs;
// inside a CTA the 8 warps are NT column tiles x (8 / NT) K-slices, reduced through shared memory.
__device__ __noinline__ void dm_gemm_phase(DmBarrier& bar, uint8_t* smem, const DmGemm& g, const int NT) {
  const int KS = DM_WARPS / NT;                          // NT (1 or 2) is a run-time value on purpose: ONE copy of this code (see dm_gemm)
  bf16* sx = reinterpret_cast<bf16*>(smem);
  float* red = reinterpret_cast<float*>(smem + DM_X_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = warp % NT, ks = warp / NT;
  const int kq = (lane & 3) * 8;
  const int M = g.M, N = g.N, K = g.K;
  const int units = (N + NT * 8 - 1) / (NT * 8);
  const bool single = K <= DM_KC;                        // the whole K range is staged once and reused across this CTA's units
  bool staged = false;
  uint4 wq[DM_UNROLL];
  // this lane's 16-byte weight loads of (unit u, K chunk k0): issued one (unit, chunk) AHEAD of the MMAs that use them
  auto load_w = [&](int u, int k0) {
    const int nrow = u * NT * 8 + cg * 8 + (lane >> 2);
    const bool n_ok = nrow < N;
    const bf16* wrow = g.W + (long long)(n_ok ? nrow : 0) * K;
    const int kc = min(DM_KC, K - k0);
    const int sub = ((kc + KS - 1) / KS + 31) / 32 * 32;           // <= 256 = DM_UNROLL steps (K chunk 1024, KS >= 4)
    const int kb = ks * sub;
#pragma unroll
    for (int q = 0; q < DM_UNROLL; ++q) {
      const int kk = kb + q * 32 + kq;
      wq[q] = make_uint4(0, 0, 0, 0);
      if (q * 32 < sub && n_ok && kk < kc) wq[q] = __ldg(reinterpret_cast<const uint4*>(wrow + k0 + kk));
    }
  };
  if ((int)blockIdx.x < units) load_w(blockIdx.x, 0);
  dm_wait(bar);
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int n_base = u * NT * 8 + cg * 8;
    float acc[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += DM_KC) {                         // (K chunks as separate units + global fp32 atomics measured slower)
      const int kc = min(DM_KC, K - k0);
      const int sub = ((kc + KS - 1) / KS + 31) / 32 * 32;          // per-warp K slice of this chunk (multiple of 32)
      const int kb = ks * sub;
      if (!(single && staged)) {
        __syncthreads();                                              // everybody is done with the previous contents
        const int kfill = min(DM_KC, KS * sub);
        const int vec = kfill / 8, total = 32 * vec;
        // cp.async.cg: L2 -> shared, no registers, the whole chunk in flight.  Every CTA reads the same few hundred
        // lines: start at a CTA-specific offset so that they are not all asking the same L2 slice at the same time.
        const int rot = (int)((blockIdx.x * 37u) % (unsigned)total);
        for (int i0 = threadIdx.x; i0 < total; i0 += DM_THREADS) {
          int i = i0 + rot;
          if (i >= total) i -= total;
          const int r = i / vec, c8 = (i - r * vec) * 8;
          const bool ok = r < M && c8 < kc;
          const bf16* src = g.x + (long long)(ok ? r : 0) * g.ldx + k0 + (ok ? c8 : 0);
          const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(sx + r * DM_LDS + c8));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        staged = true;
      }
#pragma unroll
      for (int q = 0; q < DM_UNROLL; ++q) {
        const int kk = kb + q * 32 + kq;
        if (q * 32 >= sub || kb + q * 32 >= kc) break;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int r0 = t * 16 + (lane >> 2);
          const uint4 xa = *reinterpret_cast<const uint4*>(sx + r0 * DM_LDS + kk);
          const uint4 xb = *reinterpret_cast<const uint4*>(sx + (r0 + 8) * DM_LDS + kk);
          const uint32_t a1[4] = {xa.x, xb.x, xa.y, xb.y};
          const uint32_t a2[4] = {xa.z, xb.z, xa.w, xb.w};
          dm_mma(acc[t], a1, wq[q].x, wq[q].y);
          dm_mma(acc[t], a2, wq[q].z, wq[q].w);
        }
      }
      // the registers are free again: put the next (unit, chunk)'s weights in flight under the reduction / epilogue
      int nu = u, nk = k0 + DM_KC;
      if (nk >= K) {
        nk = 0;
        nu = u + gridDim.x;
      }
      if (nu < units) load_w(nu, nk);
    }
    // K-slice reduction (KS partial sums per column tile)
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[((ks * NT + cg) * 32 + lane) * 8 + t * 4 + j] = acc[t][j];
    __syncthreads();
    if (ks == 0) {
      // C fragment: rows t*16 + hh*8 + lane/4, columns n_base + (lane%4)*2 + {0,1}
      const int col = n_base + (lane & 3) * 2;
      const bool c0ok = col < N, two = col + 1 < N;
      {
      float b0 = 0.f, b1 = 0.f, cs0 = 0.f, cs1 = 0.f;
      if (g.bias && c0ok) {
        b0 = __ldg(g.bias + col);
        if (two) b1 = __ldg(g.bias + col + 1);
      }
      if (g.csum && c0ok) {
        cs0 = __ldg(g.csum + col);
        if (two) cs1 = __ldg(g.csum + col + 1);
      }
      // a ROLLED loop over the lane's four (row, column pair) slots, summing the K-slice partials straight from shared
      // memory: a quarter of the epilogue code (this kernel is instruction-fetch bound)
#pragma unroll 1
      for (int sl = 0; sl < 4; ++sl) {
        {
          const int t = sl >> 1, hh = sl & 1;
          const int row = t * 16 + hh * 8 + (lane >> 2);
          const bool rok = row < M;
          float v0 = 0.f, v1 = 0.f;
#pragma unroll 4
          for (int q = 0; q < KS; ++q) {
            const float2 pr = *reinterpret_cast<const float2*>(red + ((q * NT + cg) * 32 + lane) * 8 + t * 4 + hh * 2);
            v0 += pr.x;
            v1 += pr.y;
          }
          if (g.csum) {                                               // LayerNorm of the input row, applied to the product
            float mean = 0.f, rstd = 0.f;
            if (rok) {
              const float2 st = __ldcg(reinterpret_cast<const float2*>(g.stats_in + row * 2));
              mean = st.x / (float)K;
              rstd = rsqrtf(fmaxf(st.y / (float)K - mean * mean, 0.f) + g.eps);
            }
            v0 = rstd * (v0 - mean * cs0);
            v1 = rstd * (v1 - mean * cs1);
          }
          v0 += b0;
          v1 += b1;
          if (g.act == 2) {
            v0 = dm_gelu(v0);
            v1 = dm_gelu(v1);
          } else if (g.act == 1) {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
          }
          if (g.residual && rok && c0ok) {
            if (two) {
              const unsigned int rr = __ldcg(reinterpret_cast<const unsigned int*>(g.residual + (long long)row * g.ldr + col));
              const float2 rf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr));
              v0 += rf.x;
              v1 += rf.y;
            } else {
              const unsigned short rr = __ldcg(reinterpret_cast<const unsigned short*>(g.residual + (long long)row * g.ldr + col));
              v0 += __bfloat162float(*reinterpret_cast<const bf16*>(&rr));
            }
          }
          const __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
          if (rok && c0ok) {
            if (two) *reinterpret_cast<__nv_bfloat162*>(g.out + (long long)row * g.ldo + col) = pk;
            else g.out[(long long)row * g.ldo + col] = pk.x;
          }
          if (g.stats_out) {                                          // statistics of what the consumer will read (rounded)
            const float2 rf = __bfloat1622float2(pk);
            float s = (rok && c0ok) ? rf.x + (two ? rf.y : 0.f) : 0.f;
            float q = (rok && c0ok) ? rf.x * rf.x + (two ? rf.y * rf.y : 0.f) : 0.f;
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            q += __shfl_xor_sync(0xffffffffu, q, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            q += __shfl_xor_sync(0xffffffffu, q, 2);
            if ((lane & 3) == 0 && rok) {
              asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g.stats_out + row * 2), "f"(s) : "memory");
              asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g.stats_out + row * 2 + 1), "f"(q) : "memory");
            }
          }
        }
      }
      }
    }
    __syncthreads();                                                  // red is reused by the next unit
  }
  dm_arrive(bar);
}

__device__ __forceinline__ void dm_gemm(DmBarrier& bar, uint8_t* smem, const DmGemm& g) {
  // one column tile per CTA while that still gives every CTA at most one unit; two tiles (half the x staging traffic)
  // otherwise.  The kernel is INSTRUCTION-FETCH bound: every phase runs its code once per layer, and with two template
  // instances of the GEMM, the attention phase and the main loop the per-layer code was ~95 KB (161 KB with a third GEMM
  // variant: 0.89 -> 1.07 ms per token).  Hence one run-time-parameterised copy of each phase and no inlined GELU.
  dm_gemm_phase(bar, smem, g, (g.N + 7) / 8 <= (int)gridDim.x ? 1 : 2);
}

// ---- fp8 KV cache helpers (MX rule of mxfp8.cu: e4m3 elements, one ue8m0 scale byte per 32 elements of a head row)
__device__ __forceinline__ void dm_e4m3x8_unpack(const uint2& q, float (&f)[8]) {
  const uint32_t w[2] = {q.x, q.y};
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const __half2_raw r = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((w[i] >> (16 * hf)) & 0xFFFFu), __NV_E4M3);
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&r));
      f[i * 4 + hf * 2] = t.x;
      f[i * 4 + hf * 2 + 1] = t.y;
    }
}
__device__ __forceinline__ float dm_ue8m0(unsigned int byte) { return __uint_as_float(byte << 23); }
__device__ __forceinline__ uint32_t dm_scale_byte(float amax) {       // 2^(floor(log2 amax) - 8), clamped
  int b = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 8;
  return (uint32_t)(b < 0 ? 0 : (b > 254 ? 254 : b));
}
__device__ __forceinline__ uint2 dm_pack_e4m3x8(const float (&f)[8], float inv) {
  uint2 w;
  w.x = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[0] * inv, f[1] * inv), __NV_SATFINITE, __NV_E4M3) |
        ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[2] * inv, f[3] * inv), __NV_SATFINITE, __NV_E4M3) << 16);
  w.y = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[4] * inv, f[5] * inv), __NV_SATFINITE, __NV_E4M3) |
        ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[6] * inv, f[7] * inv), __NV_SATFINITE, __NV_E4M3) << 16);
  return w;
}

// one online-softmax update of (m, l, acc) with up to DM_KEYS keys
template <int DM_KEYS>
__device__ __forceinline__ void dm_softmax_update(float& m, float& l, float (&acc)[8], const float (&s)[DM_KEYS], const uint4 (&vq)[DM_KEYS]) {
  float mn = m;
#pragma unroll
  for (int i = 0; i < DM_KEYS; ++i) mn = fmaxf(mn, s[i]);
  if (mn == -INFINITY) return;
  const float a = exp2f(m - mn);
  l *= a;
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] *= a;
#pragma unroll
  for (int i = 0; i < DM_KEYS; ++i) {
    const float e = exp2f(s[i] - mn);                    // exp2(-inf) = 0 for masked keys
    l += e;
    float vf[8];
    bf8_unpack(vq[i], vf);
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] += e * vf[d];
  }
  m = mn;
}

// Attention of the current token over the cache [0, pos) plus itself.  Work item = (batch, head, sequence chunk),
// one warp each (8 lanes per key: 4 keys per load instruction, 16 keys per step).  The warp that owns the last chunk
// also appends this token's K / V rows to the cache.  The last warp to finish a (batch, head) merges the partials.
template <int DM_KEYS, bool KV8>
__device__ __noinline__ void dm_attention_phase(DmBarrier& bar, const DecodeParams& p, const DecodeLayerParams& L, int pos) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g4 = lane >> 3, sub = lane & 7;
  const int E = p.E, H = p.H, BH = p.B * p.H;
  const int total_warps = gridDim.x * DM_WARPS;
  int nch = total_warps / BH;
  nch = nch < 1 ? 1 : (nch > DM_MAX_CHUNKS ? DM_MAX_CHUNKS : nch);
  if (nch > 1 && pos / nch < 16) nch = pos / 16 < 1 ? 1 : pos / 16;      // short contexts: at least 16 keys per chunk
  const float scale_log2 = 1.4426950408889634f * rsqrtf((float)DM_D);
  const bf16* qkv = reinterpret_cast<const bf16*>(p.qkv);
  bf16* kc = reinterpret_cast<bf16*>(L.kc);
  bf16* vc = reinterpret_cast<bf16*>(L.vc);
  dm_wait(bar);
  const int gw = warp * gridDim.x + blockIdx.x;                        // spread consecutive items over the SMs
  for (int item = gw; item < BH * nch; item += total_warps) {
    const int bh = item / nch, c = item % nch;
    const int b = bh / H, h = bh % H;
    const int start = (int)((long long)pos * c / nch), end = (int)((long long)pos * (c + 1) / nch);
    const bf16* row = qkv + (long long)b * 3 * E + h * DM_D + sub * 8;
    float qf[8];
    bf8_unpack(__ldcg(reinterpret_cast<const uint4*>(row)), qf);
#pragma unroll
    for (int d = 0; d < 8; ++d) qf[d] *= scale_log2;
    const bf16* kb = kc + (long long)b * p.Lmax * E + h * DM_D + sub * 8;
    const bf16* vb = vc + (long long)b * p.Lmax * E + h * DM_D + sub * 8;
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
    if constexpr (KV8) {
      // fp8 KV cache: 8 bytes of K and of V per lane and key (half the bytes of the bf16 cache), one scale byte per
      // 32 elements; the K scale multiplies the partial dot product, the V scale the softmax weight
      const int E32 = E >> 5;
      const uint8_t* kb8 = reinterpret_cast<const uint8_t*>(L.k8) + (long long)b * p.Lmax * E + h * DM_D + sub * 8;
      const uint8_t* vb8 = reinterpret_cast<const uint8_t*>(L.v8) + (long long)b * p.Lmax * E + h * DM_D + sub * 8;
      const uint8_t* ks8 = reinterpret_cast<const uint8_t*>(L.ks8) + (long long)b * p.Lmax * E32 + h * 2 + (sub >> 2);
      const uint8_t* vs8 = reinterpret_cast<const uint8_t*>(L.vs8) + (long long)b * p.Lmax * E32 + h * 2 + (sub >> 2);
      for (int j0 = start; j0 < end; j0 += 4 * DM_KEYS) {
        uint2 kq[DM_KEYS], vq[DM_KEYS];
        unsigned int ksb[DM_KEYS], vsb[DM_KEYS];
#pragma unroll
        for (int i = 0; i < DM_KEYS; ++i) {
          const int j = j0 + i * 4 + g4;
          kq[i] = make_uint2(0, 0);
          vq[i] = make_uint2(0, 0);
          ksb[i] = 0;
          vsb[i] = 0;
          if (j < end) {
            kq[i] = __ldcs(reinterpret_cast<const uint2*>(kb8 + (long long)j * E));
            vq[i] = __ldcs(reinterpret_cast<const uint2*>(vb8 + (long long)j * E));
            ksb[i] = __ldcs(ks8 + (long long)j * E32);
            vsb[i] = __ldcs(vs8 + (long long)j * E32);
          }
        }
        float s[DM_KEYS];
#pragma unroll
        for (int i = 0; i < DM_KEYS; ++i) {
          float kf[8];
          dm_e4m3x8_unpack(kq[i], kf);
          float d0 = 0.f;
#pragma unroll
          for (int d = 0; d < 8; ++d) d0 += qf[d] * kf[d];
          d0 *= dm_ue8m0(ksb[i]);
          d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
          d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
          d0 += __shfl_xor_sync(0xffffffffu, d0, 4);
          s[i] = (j0 + i * 4 + g4 < end) ? d0 : -INFINITY;
        }
        float mn = m;
#pragma unroll
        for (int i = 0; i < DM_KEYS; ++i) mn = fmaxf(mn, s[i]);
        if (mn != -INFINITY) {
          const float a = exp2f(m - mn);
          l *= a;
#pragma unroll
          for (int d = 0; d < 8; ++d) acc[d] *= a;
#pragma unroll
          for (int i = 0; i < DM_KEYS; ++i) {
            const float e = exp2f(s[i] - mn);
            l += e;
            const float ev = e * dm_ue8m0(vsb[i]);
            float vf[8];
            dm_e4m3x8_unpack(vq[i], vf);
#pragma unroll
            for (int d = 0; d < 8; ++d) acc[d] += ev * vf[d];
          }
          m = mn;
        }
      }
    } else {
    for (int j0 = start; j0 < end; j0 += 4 * DM_KEYS) {             // (a second step in flight measured slower: 29 vs 25 us)
        uint4 kq[DM_KEYS], vq[DM_KEYS];
#pragma unroll
        for (int i = 0; i < DM_KEYS; ++i) {
          const int j = j0 + i * 4 + g4;
          kq[i] = make_uint4(0, 0, 0, 0);
          vq[i] = make_uint4(0, 0, 0, 0);
          if (j < end) {
            kq[i] = __ldcs(reinterpret_cast<const uint4*>(kb + (long long)j * E));
            vq[i] = __ldcs(reinterpret_cast<const uint4*>(vb + (long long)j * E));
          }
        }
        float s[DM_KEYS];
#pragma unroll
        for (int i = 0; i < DM_KEYS; ++i) {
          float kf[8];
          bf8_unpack(kq[i], kf);
          float d0 = 0.f;
#pragma unroll
          for (int d = 0; d < 8; ++d) d0 += qf[d] * kf[d];
          d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
          d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
          d0 += __shfl_xor_sync(0xffffffffu, d0, 4);
          s[i] = (j0 + i * 4 + g4 < end) ? d0 : -INFINITY;
        }
        dm_softmax_update<DM_KEYS>(m, l, acc, s, vq);
      }
    }
    if (c == nch - 1) {                                                // the current token: from the projection, and into the cache
      const uint4 kcur = __ldcg(reinterpret_cast<const uint4*>(row + E));
      const uint4 vcur = __ldcg(reinterpret_cast<const uint4*>(row + 2 * E));
      if (g4 == 0) {
        *reinterpret_cast<uint4*>(kc + (long long)b * p.Lmax * E + (long long)pos * E + h * DM_D + sub * 8) = kcur;
        *reinterpret_cast<uint4*>(vc + (long long)b * p.Lmax * E + (long long)pos * E + h * DM_D + sub * 8) = vcur;
      }
      float kf[8];
      bf8_unpack(kcur, kf);
      if constexpr (KV8) {                                             // the fp8 cache gets this token's rows too
        float vf8[8];
        bf8_unpack(vcur, vf8);
        float ka = 0.f, va = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          ka = fmaxf(ka, fabsf(kf[d]));
          va = fmaxf(va, fabsf(vf8[d]));
        }
        ka = fmaxf(ka, __shfl_xor_sync(0xffffffffu, ka, 1));
        va = fmaxf(va, __shfl_xor_sync(0xffffffffu, va, 1));
        ka = fmaxf(ka, __shfl_xor_sync(0xffffffffu, ka, 2));
        va = fmaxf(va, __shfl_xor_sync(0xffffffffu, va, 2));
        const uint32_t kb_ = dm_scale_byte(ka), vb_ = dm_scale_byte(va);
        if (g4 == 0) {
          const long long tokoff = (long long)b * p.Lmax + pos;
          *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(L.k8) + tokoff * E + h * DM_D + sub * 8) =
              dm_pack_e4m3x8(kf, __uint_as_float((uint32_t)(254 - (int)kb_) << 23));
          *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(L.v8) + tokoff * E + h * DM_D + sub * 8) =
              dm_pack_e4m3x8(vf8, __uint_as_float((uint32_t)(254 - (int)vb_) << 23));
          if ((sub & 3) == 0) {
            reinterpret_cast<uint8_t*>(L.ks8)[tokoff * (E >> 5) + h * 2 + (sub >> 2)] = (uint8_t)kb_;
            reinterpret_cast<uint8_t*>(L.vs8)[tokoff * (E >> 5) + h * 2 + (sub >> 2)] = (uint8_t)vb_;
          }
        }
      }
      float d0 = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) d0 += qf[d] * kf[d];
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 4);
      if (g4 == 0) {                                                   // single-key online-softmax update
        const float mn = fmaxf(m, d0);
        const float a = exp2f(m - mn), e = exp2f(d0 - mn);
        float vf[8];
        bf8_unpack(vcur, vf);
        l = l * a + e;
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = acc[d] * a + e * vf[d];
        m = mn;
      }
    }
    // merge the 4 key groups of the warp
#pragma unroll 1
    for (int off = 8; off <= 16; off <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, off), lo = __shfl_xor_sync(0xffffffffu, l, off);
      const float M2 = fmaxf(m, mo);
      const float wa = (m == -INFINITY) ? 0.f : exp2f(m - M2), wb = (mo == -INFINITY) ? 0.f : exp2f(mo - M2);
      l = l * wa + lo * wb;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const float ao = __shfl_xor_sync(0xffffffffu, acc[d], off);
        acc[d] = acc[d] * wa + ao * wb;
      }
      m = M2;
    }
    bf16* orow = reinterpret_cast<bf16*>(p.o) + (long long)b * E + h * DM_D;
    if (nch == 1) {                                                    // nothing to merge: normalise and store
      if (g4 == 0) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        float f[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) f[d] = acc[d] * inv;
        *reinterpret_cast<uint4*>(orow + sub * 8) = bf8_pack(f);
      }
      continue;
    }
    float* part = p.attn_part + ((long long)bh * DM_MAX_CHUNKS + c) * DM_PART;
    if (g4 == 0) {
      *reinterpret_cast<float4*>(part + sub * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(part + sub * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      if (sub == 0) {
        part[64] = m;
        part[65] = l;
      }
    }
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      __threadfence();
      last = atomicAdd(&p.attn_cnt[bh], 1) == nch - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      const float* base = p.attn_part + (long long)bh * DM_MAX_CHUNKS * DM_PART;
      float Mx = -INFINITY;
      for (int c2 = 0; c2 < nch; ++c2) Mx = fmaxf(Mx, __ldcg(base + c2 * DM_PART + 64));
      float Ls = 0.f, o0 = 0.f, o1 = 0.f;
      for (int c2 = 0; c2 < nch; ++c2) {
        const float mc = __ldcg(base + c2 * DM_PART + 64);
        const float w = (mc == -INFINITY) ? 0.f : exp2f(mc - Mx);
        Ls += w * __ldcg(base + c2 * DM_PART + 65);
        const float2 a = __ldcg(reinterpret_cast<const float2*>(base + c2 * DM_PART + lane * 2));
        o0 += w * a.x;
        o1 += w * a.y;
      }
      const float inv = Ls > 0.f ? 1.f / Ls : 0.f;
      *reinterpret_cast<__nv_bfloat162*>(orow + lane * 2) = __floats2bfloat162_rn(o0 * inv, o1 * inv);
      if (lane == 0) p.attn_cnt[bh] = 0;                               // ready for the next layer / step
    }
  }
  dm_arrive(bar);
}

// LM head: logits = LN_f(x) Whead^T + b (folded), one WARP per 8-column tile over the full K (no K split, no block
// synchronisation per tile: 16 independent weight streams per SM), running arg-max per row in registers, merged with
// one 64-bit atomicMax per row and warp: key = (order-preserving logit bits, ~column) -> ties resolve to the smallest
// column.  x is staged once per CTA.
__device__ __forceinline__ unsigned long long dm_amax_key(float v, int col) {
  const unsigned int u = __float_as_uint(v);
  const unsigned int ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)ord << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)col);
}
__device__ __noinline__ void dm_head_phase(DmBarrier& bar, uint8_t* smem, const DecodeParams& p, const float* stats_x) {
  bf16* sx = reinterpret_cast<bf16*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kq = (lane & 3) * 8;
  const int M = p.B, N = p.V, K = p.E;
  const bf16* W = reinterpret_cast<const bf16*>(p.w_head);
  bf16* out = reinterpret_cast<bf16*>(p.logits);
  const int tiles = (N + 7) / 8, total_warps = gridDim.x * DM_WARPS;
  const int gw = blockIdx.x * DM_WARPS + warp;                         // consecutive warps stream consecutive weight rows
  const int rounds = (K + 255) / 256;                                  // 8 k-steps of 32 per round
  uint4 wq[DM_UNROLL];
  auto load_w = [&](int tile, int rd) {
    const int nrow = tile * 8 + (lane >> 2);
    const bool n_ok = nrow < N;
    const bf16* wrow = W + (long long)(n_ok ? nrow : 0) * K;
#pragma unroll
    for (int q = 0; q < DM_UNROLL; ++q) {
      const int kk = rd * 256 + q * 32 + kq;
      wq[q] = make_uint4(0, 0, 0, 0);
      if (n_ok && kk < K) wq[q] = __ldg(reinterpret_cast<const uint4*>(wrow + kk));
    }
  };
  if (gw < tiles) load_w(gw, 0);
  dm_wait(bar);
  {
    const int vec = K / 8, total = 32 * vec;
    const int rot = (int)((blockIdx.x * 37u) % (unsigned)total);
    for (int i0 = threadIdx.x; i0 < total; i0 += DM_THREADS) {
      int i = i0 + rot;
      if (i >= total) i -= total;
      const int r = i / vec, c8 = (i - r * vec) * 8;
      const bool ok = r < M;
      const bf16* src = reinterpret_cast<const bf16*>(p.x) + (long long)(ok ? r : 0) * K + c8;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(sx + r * DM_LDS + c8));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  // LayerNorm statistics of this lane's four rows
  float mean[4], rstd[4];
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    const int row = (sl >> 1) * 16 + (sl & 1) * 8 + (lane >> 2);
    mean[sl] = 0.f;
    rstd[sl] = 0.f;
    if (row < M) {
      const float2 st = __ldcg(reinterpret_cast<const float2*>(stats_x + row * 2));
      mean[sl] = st.x / (float)K;
      rstd[sl] = rsqrtf(fmaxf(st.y / (float)K - mean[sl] * mean[sl], 0.f) + p.epsf);
    }
  }
  float bestv[4];
  int bestc[4];
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    bestv[sl] = -INFINITY;
    bestc[sl] = 0x7fffffff;
  }
  for (int tile = gw; tile < tiles; tile += total_warps) {
    float acc[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
    for (int rd = 0; rd < rounds; ++rd) {
#pragma unroll
      for (int q = 0; q < DM_UNROLL; ++q) {
        const int kk = rd * 256 + q * 32 + kq;
        if (rd * 256 + q * 32 >= K) break;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int r0 = t * 16 + (lane >> 2);
          const uint4 xa = *reinterpret_cast<const uint4*>(sx + r0 * DM_LDS + kk);
          const uint4 xb = *reinterpret_cast<const uint4*>(sx + (r0 + 8) * DM_LDS + kk);
          const uint32_t a1[4] = {xa.x, xb.x, xa.y, xb.y};
          const uint32_t a2[4] = {xa.z, xb.z, xa.w, xb.w};
          dm_mma(acc[t], a1, wq[q].x, wq[q].y);
          dm_mma(acc[t], a2, wq[q].z, wq[q].w);
        }
      }
      if (rd + 1 < rounds) load_w(tile, rd + 1);
      else if (tile + total_warps < tiles) load_w(tile + total_warps, 0);
    }
    const int col = tile * 8 + (lane & 3) * 2;
    const bool c0ok = col < N, two = col + 1 < N;
    float b0 = 0.f, b1 = 0.f, cs0 = 0.f, cs1 = 0.f;
    if (c0ok) {
      if (p.b_head) b0 = __ldg(p.b_head + col);
      cs0 = __ldg(p.c_head + col);
      if (two) {
        if (p.b_head) b1 = __ldg(p.b_head + col + 1);
        cs1 = __ldg(p.c_head + col + 1);
      }
    }
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      const int t = sl >> 1, hh = sl & 1;
      const int row = t * 16 + hh * 8 + (lane >> 2);
      if (row >= M || !c0ok) continue;
      const float v0 = rstd[sl] * (acc[t][hh * 2] - mean[sl] * cs0) + b0;
      const float v1 = rstd[sl] * (acc[t][hh * 2 + 1] - mean[sl] * cs1) + b1;
      const __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
      if (two) *reinterpret_cast<__nv_bfloat162*>(out + (long long)row * p.vpad + col) = pk;
      else out[(long long)row * p.vpad + col] = pk.x;
      const float2 rf = __bfloat1622float2(pk);                         // arg-max over what is stored
      if (rf.x > bestv[sl]) {
        bestv[sl] = rf.x;
        bestc[sl] = col;
      }
      if (two && rf.y > bestv[sl]) {
        bestv[sl] = rf.y;
        bestc[sl] = col + 1;
      }
    }
  }
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    const int row = (sl >> 1) * 16 + (sl & 1) * 8 + (lane >> 2);
    unsigned long long key = bestc[sl] == 0x7fffffff ? 0ull : dm_amax_key(bestv[sl], bestc[sl]);
    const unsigned long long k1 = __shfl_xor_sync(0xffffffffu, key, 1);
    key = k1 > key ? k1 : key;
    const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, 2);
    key = k2 > key ? k2 : key;
    if ((lane & 3) == 0 && row < M && key) atomicMax(p.amax + row, key);
  }
  dm_arrive(bar);
}

template <bool KV8>
__global__ void __launch_bounds__(DM_THREADS, 2) k_decode_mega(const DecodeParams p) {
  extern __shared__ __align__(16) uint8_t dm_smem[];
  __shared__ float rv[DM_WARPS];
  __shared__ float rq[DM_WARPS];
  DmBarrier bar{p.bar, 0u, gridDim.x, false};
  const int pos = *p.pos, step = *p.step;
  const int B = p.B, E = p.E, F = p.F;
  bf16* x = reinterpret_cast<bf16*>(p.x);
  bf16* x2 = reinterpret_cast<bf16*>(p.x2);
  bf16* qkv = reinterpret_cast<bf16*>(p.qkv);
  bf16* o = reinterpret_cast<bf16*>(p.o);
  bf16* hbuf = reinterpret_cast<bf16*>(p.h);
  float* stats_x = p.stats;                                            // rows of x  (input of LN1 / LN_f)
  float* stats_x2 = p.stats + 64;                                      // rows of x2 (input of LN2)
  int stamp = 0;
#define DM_STAMP()                                                                   \
  do {                                                                               \
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) p.prof[stamp] = dm_timer();   \
    ++stamp;                                                                         \
  } while (0)
  DM_STAMP();
  // ---- phase 0: token + position embedding of the current token (+ the row statistics LN1 of layer 0 needs)
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    int id = p.tok[b];
    id = id < 0 ? 0 : (id >= p.V ? p.V - 1 : id);
    float s = 0.f, q = 0.f;
    for (int d = threadIdx.x; d < E; d += DM_THREADS) {
      const bf16 v = __float2bfloat16(__ldg(p.wte + (long long)id * E + d) + __ldg(p.wpe + (long long)pos * E + d));
      x[(long long)b * E + d] = v;
      const float f = __bfloat162float(v);
      s += f;
      q += f * f;
    }
    s = warp_sum(s);
    q = warp_sum(q);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      rv[threadIdx.x >> 5] = s;
      rq[threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float S = 0.f, Q = 0.f;
      for (int i = 0; i < DM_WARPS; ++i) {
        S += rv[i];
        Q += rq[i];
      }
      stats_x[b * 2] = S;
      stats_x[b * 2 + 1] = Q;
    }
  }
  dm_arrive(bar);
  DM_STAMP();
  for (int li = 0; li < p.n_layers; ++li) {
    const DecodeLayerParams L = p.layers[li];
    // A: qkv = LN1(x) Wqkv^T + b   (LN1 folded: raw x, statistics from the producer of x)
    {
      const DmGemm g{x, E, reinterpret_cast<const bf16*>(L.w_qkv), L.b_qkv, L.c_qkv, stats_x, L.eps1, nullptr, 0, qkv, 3 * E, nullptr, B, 3 * E, E, 0};
      dm_gemm(bar, dm_smem, g);
    }
    DM_STAMP();
    // B: o = attention(q, cache + current token); cache append
    dm_attention_phase<4, KV8>(bar, p, L, pos);          // 16 keys per warp step (2 and 8 per group measured: 1.13 / 1.70 ms vs 0.89 ms per token)
    DM_STAMP();
    // C: x2 = x + o Wout^T + b  (+ row statistics of x2).  stats_x was consumed in A (every CTA passed B's barrier
    // since) and is refilled in E: clear it here.
    if (blockIdx.x == 0 && threadIdx.x < 64) stats_x[threadIdx.x] = 0.f;
    {
      const DmGemm g{o, E, reinterpret_cast<const bf16*>(L.w_out), L.b_out, nullptr, nullptr, 0.f, x, E, x2, E, stats_x2, B, E, E, 0};
      dm_gemm(bar, dm_smem, g);
    }
    DM_STAMP();
    // D: h = gelu(LN2(x2) W1^T + b1)
    {
      const DmGemm g{x2, E, reinterpret_cast<const bf16*>(L.w_fc1), L.b_fc1, L.c_fc1, stats_x2, L.eps2, nullptr, 0, hbuf, F, nullptr, B, F, E, 2};
      dm_gemm(bar, dm_smem, g);
    }
    DM_STAMP();
    // E: x = x2 + h W2^T + b2  (+ row statistics of x)
    {
      const DmGemm g{hbuf, F, reinterpret_cast<const bf16*>(L.w_fc2), L.b_fc2, nullptr, nullptr, 0.f, x2, E, x, E, stats_x, B, E, F, 0};
      dm_gemm(bar, dm_smem, g);
    }
    // stats_x2 was consumed in D; this CTA passed E's barrier wait, i.e. every CTA is done with D: clear it for the
    // next layer's C (two barriers away)
    if (blockIdx.x == 0 && threadIdx.x < 64) stats_x2[threadIdx.x] = 0.f;
    DM_STAMP();
  }
  // ---- LM head on LN_f(x) (folded) with the running arg-max
  dm_head_phase(bar, dm_smem, p, stats_x);
  DM_STAMP();
  // ---- token bookkeeping (ties -> smallest index through the key order)
  if (blockIdx.x != 0) return;
  dm_wait(bar);
  if ((int)threadIdx.x < B) {
    const int r = threadIdx.x;
    const unsigned long long key = __ldcg(p.amax + r);
    const int t = key ? (int)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull)) : 0;
    p.amax[r] = 0ull;                                                  // ready for the next step
    p.tok[r] = t;
    if (step < p.hist_cap) p.hist[(long long)step * B + r] = t;
  }
  __syncthreads();
  DM_STAMP();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *p.pos = pos + 1;
    *p.step = step + 1;
  }
}

bool decode_mega_supported(int B, int E, int H, int F) {
  return B >= 1 && B <= 32 && H >= 1 && E == H * DM_D && E % 32 == 0 && E <= DM_KC && F % 32 == 0 && F >= 32;
}
size_t decode_mega_part_bytes(int B, int H) { return (size_t)B * H * DM_MAX_CHUNKS * DM_PART * sizeof(float); }

// one warp per output row n: Wf = bf16(W * gamma), csum = sum of the ROUNDED row, bias_f = bias + W . beta
__global__ void k_fold_ln(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                          int N, int K, bf16* __restrict__ Wf, float* __restrict__ csum, float* __restrict__ bias_f) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[(long long)n * K + k];
    const bf16 r = __float2bfloat16(w * gamma[k]);
    Wf[(long long)n * K + k] = r;
    cs += __bfloat162float(r);
    bs += w * beta[k];
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    csum[n] = cs;
    bias_f[n] = (bias ? bias[n] : 0.f) + bs;
  }
}
void decode_fold_ln(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K, void* Wf, float* csum, float* bias_f,
                    cudaStream_t stream) {
  if (N <= 0 || K <= 0) return;
  k_fold_ln<<<(N + 7) / 8, 256, 0, stream>>>(W, bias, gamma, beta, N, K, reinterpret_cast<bf16*>(Wf), csum, bias_f);
  TNN_CHECK_LAUNCH();
}

void decode_step_mega(const DecodeParams& p, cudaStream_t stream) {
  static int grids[64] = {};   // per device: the smem attribute and the occupancy are device properties
  static PerDeviceOnce once;
  int& grid = grids[PerDeviceOnce::device()];
  if (once.begin()) {
    cudaFuncSetAttribute(k_decode_mega<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM);
    cudaFuncSetAttribute(k_decode_mega<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM);
    int dev = 0, sms = 0, per_sm = 0, per_sm8 = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_decode_mega<false>, DM_THREADS, DM_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm8, k_decode_mega<true>, DM_THREADS, DM_SMEM);
    per_sm = per_sm < per_sm8 ? per_sm : per_sm8;
    if (per_sm < 1) {
      fprintf(stderr, "[tnn_b200] decode_step_mega: the kernel does not fit on an SM\n");
      abort();
    }
    // every CTA must be resident (the phases are separated by spin barriers): never more than the device holds
    grid = sms * (per_sm > 2 ? 2 : per_sm);
    const char* e = getenv("TNN_DECODE_CTAS");
    if (e && atoi(e) > 0 && atoi(e) < grid) grid = atoi(e);
    once.done();
  }
  cudaMemsetAsync(p.bar, 0, sizeof(unsigned int), stream);
  if (p.kv_fp8) k_decode_mega<true><<<grid, DM_THREADS, DM_SMEM, stream>>>(p);
  else k_decode_mega<false><<<grid, DM_THREADS, DM_SMEM, stream>>>(p);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
