// Skinny GEMM for decode: out[M <= 32, N] = act(x[M, K] . W[N, K]^T + bias (+ residual)), bf16.
//
// With one token per sequence the M dimension is the batch (32): the tcgen05 tile kernel would
// run N/128 CTAs, each wasting 3/4 of its 128-row tile, and the step is bound by how fast the
// WEIGHTS stream from HBM.  This kernel is weight-streaming: every warp owns 8 output columns and
// the full K range, lanes fetch 16 contiguous bytes of their weight row (UNROLL loads in flight
// per lane, no shared-memory staging of W), x (tiny, L2 resident) is staged once per K chunk in
// padded shared memory, and the products run on mma.sync m16n8k16 (two m16 tiles = 32 rows).
// The k index inside each 32-wide step is permuted consistently for both operands so that a
// lane's 16-byte weight load IS its B fragment for two MMAs -- no shuffles, no transposes.
// (tcgen05 needs M = 64/128 rows per instruction; at M = 32 the legacy warp-level MMA is the
// right tool and the kernel is HBM-bound, not tensor-bound.)
#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

static constexpr int SK_WARPS = 8;             // 2 column groups (8 columns each) x 4 K-sub-slices
static constexpr int SK_COLS = 16;             // output columns per CTA
static constexpr int SK_KC = 1024;             // K chunk staged in smem
static constexpr int SK_PAD = 32;              // +64 bytes per row: conflict-free 16-byte fragment loads
static constexpr int SK_UNROLL = 8;            // 32-wide k steps (16-byte weight loads) in flight per lane
static constexpr int SK_MAX_SPLIT = 8;         // thread-block cluster size limit (portable)
static constexpr int SK_X_BYTES = 32 * (SK_KC + SK_PAD) * 2;
static constexpr int SK_PART_BYTES = SK_MAX_SPLIT * 2 * 32 * 8 * 4;   // cross-CTA partial sums [rank][cg][lane][8]

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t sk_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void sk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// "every CTA of the cluster has started" barrier, split so that it costs nothing: arrive at kernel entry, wait right
// before the first distributed-shared-memory store (compute-sanitizer racecheck: a peer's shared memory must not be
// written before that CTA is known to have entered the kernel).
__device__ __forceinline__ void sk_cluster_arrive_entered() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void sk_cluster_wait_entered() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
// store 4 floats into the shared memory of cluster CTA `rank` at the address that `local` has here
__device__ __forceinline__ void sk_st_remote_v4(void* local, uint32_t rank, float a, float b, float c, float d) {
  uint32_t laddr = static_cast<uint32_t>(__cvta_generic_to_shared(local)), raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Split-K over a thread-block CLUSTER: the `ksplit` CTAs of a cluster own the same 16 output columns
// and a 1/ksplit slice of K each, so even N = 768 (48 column groups) puts 192..384 CTAs' worth of
// 16-byte weight loads in flight (the weights ARE the traffic: 0.15 TB/s with 48 CTAs).  Partial
// sums travel through distributed shared memory into cluster rank 0, which applies the epilogue.
__global__ void __launch_bounds__(SK_WARPS * 32)
    k_skinny_gemm(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w, long long ldw,
                  const float* __restrict__ bias, const bf16* __restrict__ residual, long long ldr,
                  bf16* __restrict__ out, long long ldo, int M, int N, int K, int act, int ksplit) {
  // programmatic dependent launch: let the NEXT kernel of the decode chain start its own prologue now
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (ksplit > 1) sk_cluster_arrive_entered();
  extern __shared__ __align__(16) uint8_t sk_smem[];
  bf16* sx = reinterpret_cast<bf16*>(sk_smem);                                // [32][SK_KC + SK_PAD]
  float* part = reinterpret_cast<float*>(sk_smem + SK_X_BYTES);               // [ksplit][2][32][8], written by peers
  constexpr int LDS = SK_KC + SK_PAD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = warp & 1;                                 // column group inside the CTA
  const int ks = warp >> 1;                                // K-sub-slice 0..3 of every staged chunk
  const uint32_t rank = ksplit > 1 ? sk_cluster_rank() : 0u;
  const int col_cta = blockIdx.x / ksplit;
  const int n_base = col_cta * SK_COLS + cg * 8;
  const int nrow = n_base + (lane >> 2);                   // this lane's weight row (B column)
  const int kq = (lane & 3) * 8;                           // this lane's 8-element slice of a 32-wide k step
  const bool n_ok = nrow < N;
  const bf16* wrow = w + (long long)(n_ok ? nrow : 0) * ldw;
  // this CTA's K range (multiples of 32 so the 16-byte loads stay aligned)
  const int k_len = ((K + ksplit - 1) / ksplit + 31) / 32 * 32;
  const int k_beg = min(K, (int)rank * k_len);
  const int k_end = min(K, k_beg + k_len);
  float acc[2][4];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;

  for (int k0 = k_beg; k0 < k_end; k0 += SK_KC) {
    const int kc = min(SK_KC, k_end - k0);
    // every K-sub-slice warp gets an equal share of the chunk (a multiple of 32)
    const int sub = ((kc + 3) / 4 + 31) / 32 * 32;
    const int kb = ks * sub;
    // 1) weight loads first: they do not depend on the staged activations, so the HBM latency of W
    //    overlaps the L2 latency of staging x
    uint4 wq[SK_UNROLL];
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      const int kk = kb + u * 32 + kq;
      wq[u] = make_uint4(0, 0, 0, 0);
      if (u * 32 < sub) {
        if (n_ok && kk + 8 <= kc) {
          wq[u] = __ldg(reinterpret_cast<const uint4*>(wrow + k0 + kk));
        } else if (n_ok && kk < kc) {
          bf16 tmp[8];
          for (int j = 0; j < 8; ++j) tmp[j] = (kk + j < kc) ? wrow[k0 + kk + j] : __float2bfloat16(0.f);
          wq[u] = *reinterpret_cast<uint4*>(tmp);
        }
      }
    }
    // (launched with the PDL attribute when the weights are static: everything above overlapped the
    //  previous kernel of the stream; its output x is only valid after this wait.  No-op otherwise.)
    if (k0 == k_beg) asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    // 2) stage x[:, k0:k0+kc] (rows >= M and the k tail up to 4*sub are zero filled: no masking below)
    const int kfill = min(SK_KC, 4 * sub);
    for (int i = threadIdx.x; i < 32 * (kfill / 8); i += blockDim.x) {
      const int r = i / (kfill / 8), c8 = (i % (kfill / 8)) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r < M && c8 < kc) {
        if (c8 + 8 <= kc) {
          v = *reinterpret_cast<const uint4*>(x + (long long)r * ldx + k0 + c8);
        } else {
          bf16 tmp[8];
          for (int j = 0; j < 8; ++j) tmp[j] = (c8 + j < kc) ? x[(long long)r * ldx + k0 + c8 + j] : __float2bfloat16(0.f);
          v = *reinterpret_cast<uint4*>(tmp);
        }
      }
      *reinterpret_cast<uint4*>(sx + r * LDS + c8) = v;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      const int kk = kb + u * 32 + kq;
      if (u * 32 >= sub || kb + u * 32 >= kc) break;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int r0 = t * 16 + (lane >> 2);
        const uint4 xa = *reinterpret_cast<const uint4*>(sx + r0 * LDS + kk);
        const uint4 xb = *reinterpret_cast<const uint4*>(sx + (r0 + 8) * LDS + kk);
        // logical k (lane%4)*2+{0,1} <-> elements 0,1 ; logical k+8 <-> elements 2,3 (first MMA);
        // elements 4..7 feed the second MMA the same way
        const uint32_t a1[4] = {xa.x, xb.x, xa.y, xb.y};
        const uint32_t a2[4] = {xa.z, xb.z, xa.w, xb.w};
        mma_bf16_16816(acc[t], a1, wq[u].x, wq[u].y);
        mma_bf16_16816(acc[t], a2, wq[u].z, wq[u].w);
      }
    }
  }
  // reduce the 4 K-sub-slices of each column group through shared memory (reuse the x staging area)
  __syncthreads();
  float* red = reinterpret_cast<float*>(sk_smem);          // [4 ks][2 cg][32 lanes][8]
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(((ks * 2 + cg) * 32) + lane) * 8 + t * 4 + j] = acc[t][j];
  __syncthreads();
  if (ksplit > 1) sk_cluster_wait_entered();               // all peers are running: their shared memory may be written
  if (ks == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) v += red[(((q * 2 + cg) * 32) + lane) * 8 + t * 4 + j];
        acc[t][j] = v;
      }
    if (ksplit > 1 && rank != 0) {
      // this CTA's partial sums -> rank 0's `part[rank]` slot (distributed shared memory)
      float* slot = part + ((rank * 2 + cg) * 32 + lane) * 8;
      sk_st_remote_v4(slot, 0, acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
      sk_st_remote_v4(slot + 4, 0, acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
    }
  }
  if (ksplit > 1) {
    sk_cluster_sync();                 // release my remote stores / acquire everybody else's
    if (rank != 0) return;
    if (ks == 0) {
      for (int r = 1; r < ksplit; ++r) {
        const float4 p0 = *reinterpret_cast<const float4*>(part + ((r * 2 + cg) * 32 + lane) * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(part + ((r * 2 + cg) * 32 + lane) * 8 + 4);
        acc[0][0] += p0.x; acc[0][1] += p0.y; acc[0][2] += p0.z; acc[0][3] += p0.w;
        acc[1][0] += p1.x; acc[1][1] += p1.y; acc[1][2] += p1.z; acc[1][3] += p1.w;
      }
    }
  }
  if (ks != 0) return;
  // C fragment: rows lane/4 (+8), columns n_base + (lane%4)*2 + {0,1}
#pragma unroll
  for (int t = 0; t < 2; ++t) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = t * 16 + h * 8 + (lane >> 2);
      if (row >= M) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = n_base + (lane & 3) * 2 + j;
        if (col >= N) continue;
        float v = acc[t][h * 2 + j];
        if (bias) v += bias[col];
        if (residual) v += __bfloat162float(residual[(long long)row * ldr + col]);
        if (act == 1) v = fmaxf(v, 0.f);
        else if (act == 2) {
          const float k0c = 0.7978845608028654f, k1c = 0.044715f;
          v = 0.5f * v * (1.f + tanhf(k0c * (v + k1c * v * v * v)));
        }
        out[(long long)row * ldo + col] = __float2bfloat16(v);
      }
    }
  }
}

void skinny_gemm_bf16(const void* x, long long ldx, const void* w, long long ldw, const float* bias,
                      const void* residual, long long ldr, void* out, long long ldo, int M, int N, int K, int act,
                      int static_weights, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  if (M > 32 || (ldx % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) {
    fprintf(stderr, "[tnn_b200] skinny_gemm_bf16: needs M <= 32 and 16-byte aligned rows\n");
    abort();
  }
  constexpr int smem = SK_X_BYTES + SK_PART_BYTES;
  static PerDeviceOnce configured;
  static int forced = -1, pdl = 1;
  if (configured.begin()) {
    cudaFuncSetAttribute(k_skinny_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* e = getenv("TNN_SKINNY_KSPLIT");
    forced = (e != nullptr) ? atoi(e) : 0;
    const char* e2 = getenv("TNN_PDL");
    pdl = (e2 != nullptr) ? atoi(e2) : 1;
    configured.done();
  }
  const int col_ctas = (N + SK_COLS - 1) / SK_COLS;
  // cluster size: enough CTAs for >= ~2 per SM, every rank keeps >= 128 of K
  int ksplit = 1;
  while (ksplit < SK_MAX_SPLIT && col_ctas * ksplit < 2 * 148 && K / (ksplit * 2) >= 128) ksplit *= 2;
  if (forced > 0) ksplit = forced > SK_MAX_SPLIT ? SK_MAX_SPLIT : forced;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(col_ctas * ksplit), 1, 1);
  cfg.blockDim = dim3(SK_WARPS * 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)ksplit;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (static_weights && pdl) {
    // the weight stream does not depend on the previous kernel: start early, wait inside the kernel
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t err = cudaLaunchKernelEx(&cfg, k_skinny_gemm, (const bf16*)x, ldx, (const bf16*)w, ldw, bias,
                                       (const bf16*)residual, ldr, (bf16*)out, ldo, M, N, K, act, ksplit);
  if (err != cudaSuccess) {
    fprintf(stderr, "[tnn_b200] skinny_gemm_bf16 launch failed: %s\n", cudaGetErrorString(err));
    abort();
  }
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
