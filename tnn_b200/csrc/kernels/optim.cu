// Optimizer kernels.
//
// (1) Flat single-GPU Adam/AdamW/SGD: ONE launch over the whole parameter slab (the
//     reference launches one kernel per parameter tensor, src/nn/optimizers_impl/cuda/
//     adam_kernels.cu:19-130), fused with the bf16 compute-copy refresh and grad zeroing.
//
// (2) The data-parallel product path: a single kernel per step that
//        barrier(all ranks' grads complete)                      [peer flags, .sys scope]
//     -> reduce my 1/W slice of the gradient across all W GPUs   [NVLink peer loads, or one
//                                                                  multimem.ld_reduce via NVLS]
//     -> scale (1/W, 1/accum) -> Adam/SGD update of my slice     [optimizer state is sharded:
//                                                                  m, v exist only for my slice]
//     -> broadcast new fp32 + bf16 params to every replica       [peer stores / multimem.st]
//     -> barrier(all ranks done) -> zero my gradient slab.
//     No NCCL call, no scale/cast/optimizer/memset kernels on that path.
#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

__device__ __forceinline__ float adam_update(float& p, float g, float& m, float& v, const AdamHyper& h) {
  if (h.weight_decay != 0.f && h.decoupled) p -= h.weight_decay * h.lr * p;
  m = h.beta1 * m + (1.f - h.beta1) * g;
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  float mh = m / h.bias_c1, vh = v / h.bias_c2;
  float upd = h.lr * mh / (sqrtf(vh) + h.eps);
  if (h.weight_decay != 0.f && !h.decoupled) upd += h.weight_decay * h.lr * p;
  p -= upd;
  return p;
}

__global__ void __launch_bounds__(256)
    k_adam_flat(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                float* __restrict__ v, bf16* __restrict__ pb, long long n, AdamHyper h, int zero_grad) {
  const long long nv = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    adam_update(P.x, G.x * h.grad_scale, M.x, V.x, h);
    adam_update(P.y, G.y * h.grad_scale, M.y, V.y, h);
    adam_update(P.z, G.z * h.grad_scale, M.z, V.z, h);
    adam_update(P.w, G.w * h.grad_scale, M.w, V.w, h);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pb) {
      uint2 o;
      __nv_bfloat162 a = __floats2bfloat162_rn(P.x, P.y), b = __floats2bfloat162_rn(P.z, P.w);
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(pb)[i] = o;
    }
  }
  for (long long i = nv * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float P = p[i], M = m[i], V = v[i];
    adam_update(P, g[i] * h.grad_scale, M, V, h);
    p[i] = P;
    m[i] = M;
    v[i] = V;
    if (zero_grad) g[i] = 0.f;
    if (pb) pb[i] = __float2bfloat16(P);
  }
}
// Step state in device memory (written by the host once per step OUTSIDE any CUDA graph):
//   [0] lr [1] beta1 [2] beta2 [3] eps [4] weight_decay [5] decoupled [6] grad_scale [7] momentum
//   [8] 1-beta1^t [9] 1-beta2^t   [12..13] uint64 seed counter
__device__ __forceinline__ AdamHyper hyper_from_state(const float* st) {
  AdamHyper h;
  h.lr = st[0]; h.beta1 = st[1]; h.beta2 = st[2]; h.eps = st[3]; h.weight_decay = st[4];
  h.decoupled = st[5] != 0.f; h.grad_scale = st[6]; h.bias_c1 = st[8]; h.bias_c2 = st[9];
  return h;
}
__global__ void __launch_bounds__(256)
    k_adam_flat_dev(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                    float* __restrict__ v, bf16* __restrict__ pb, long long n, const float* __restrict__ st,
                    int zero_grad) {
  const AdamHyper h = hyper_from_state(st);
  const long long nv = (n + 3) / 4;   // slabs are padded to a multiple of 64 elements
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    adam_update(P.x, G.x * h.grad_scale, M.x, V.x, h);
    adam_update(P.y, G.y * h.grad_scale, M.y, V.y, h);
    adam_update(P.z, G.z * h.grad_scale, M.z, V.z, h);
    adam_update(P.w, G.w * h.grad_scale, M.w, V.w, h);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pb) {
      uint2 o;
      __nv_bfloat162 a = __floats2bfloat162_rn(P.x, P.y), b = __floats2bfloat162_rn(P.z, P.w);
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(pb)[i] = o;
    }
  }
}
void adam_step_flat_dev(float* p, float* g, float* m, float* v, void* p_bf16, long long n,
                        const float* state, int zero_grad, cudaStream_t s) {
  if (n <= 0) return;
  if (n % 4 != 0) {
    fprintf(stderr, "[tnn_b200] adam_step_flat_dev: slab length must be a multiple of 4\n");
    abort();
  }
  k_adam_flat_dev<<<grid_for(n / 4 + 1, 256, kSMs * 8), 256, 0, s>>>(p, g, m, v, (bf16*)p_bf16, n, state, zero_grad);
  TNN_CHECK_LAUNCH();
}
__global__ void __launch_bounds__(256)
    k_sgd_flat_dev(float* __restrict__ p, float* __restrict__ g, float* __restrict__ mom,
                   bf16* __restrict__ pb, long long n, const float* __restrict__ st, int zero_grad) {
  const float lr = st[0], wd = st[4], gs = st[6], momentum = st[7];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float P = p[i], G = g[i] * gs + wd * P;
    if (mom) {
      float V = momentum * mom[i] - lr * G;
      mom[i] = V;
      P += V;
    } else {
      P -= lr * G;
    }
    p[i] = P;
    if (zero_grad) g[i] = 0.f;
    if (pb) pb[i] = __float2bfloat16(P);
  }
}
void sgd_step_flat_dev(float* p, float* g, float* mom, void* p_bf16, long long n, const float* state,
                       int zero_grad, cudaStream_t s) {
  if (n <= 0) return;
  k_sgd_flat_dev<<<grid_for(n, 256, kSMs * 8), 256, 0, s>>>(p, g, mom, (bf16*)p_bf16, n, state, zero_grad);
  TNN_CHECK_LAUNCH();
}

void adam_step_flat(float* p, float* g, float* m, float* v, void* p_bf16, long long n,
                    const AdamHyper& h, int zero_grad, cudaStream_t s) {
  if (n <= 0) return;
  k_adam_flat<<<grid_for(n / 4 + 1, 256, kSMs * 8), 256, 0, s>>>(p, g, m, v, (bf16*)p_bf16, n, h, zero_grad);
  TNN_CHECK_LAUNCH();
}

__global__ void __launch_bounds__(256)
    k_sgd_flat(float* __restrict__ p, float* __restrict__ g, float* __restrict__ mom,
               bf16* __restrict__ pb, long long n, float lr, float momentum, float wd, float gs,
               int zero_grad) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float P = p[i], G = g[i] * gs + wd * P;
    if (mom) {
      float V = momentum * mom[i] - lr * G;
      mom[i] = V;
      P += V;
    } else {
      P -= lr * G;
    }
    p[i] = P;
    if (zero_grad) g[i] = 0.f;
    if (pb) pb[i] = __float2bfloat16(P);
  }
}
void sgd_step_flat(float* p, float* g, float* mom, void* p_bf16, long long n, float lr,
                   float momentum, float weight_decay, float grad_scale, int zero_grad,
                   cudaStream_t s) {
  if (n <= 0) return;
  k_sgd_flat<<<grid_for(n, 256, kSMs * 8), 256, 0, s>>>(p, g, mom, (bf16*)p_bf16, n, lr, momentum,
                                                      weight_decay, grad_scale, zero_grad);
  TNN_CHECK_LAUNCH();
}

// =====================================================================================
// Fused data-parallel step over NVLink peer memory
// =====================================================================================
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st_bf16x4(bf16* mc, uint32_t lo, uint32_t hi) {
  asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(mc), "r"(lo), "r"(hi)
               : "memory");
}

static constexpr int kMaxWorld = 16;
struct DpArgs {
  int world, rank;
  long long n;          // total parameters
  long long chunk;      // per-rank slice length (multiple of 4)
  long long lo0, hi0;   // element range [lo0, hi0) handled by this launch (a gradient bucket); the range is
                        // split W ways; optimizer state is indexed by (element - rank slice start) + st_off
  long long st_off;
  float* g_peers[kMaxWorld];    // every rank's gradient slab (peer mapped), [rank] is local
  float* p_peers[kMaxWorld];    // every rank's fp32 parameter slab
  bf16* pb_peers[kMaxWorld];    // every rank's bf16 parameter copy
  uint32_t* flag_peers[kMaxWorld];  // every rank's flag array (world words)
  const float* g_mc;    // multicast views (NVLS) or null
  float* p_mc;
  bf16* pb_mc;
  float* m;             // optimizer state for MY slice only (length chunk)
  float* v;
  uint32_t* grid_ctr;   // sync[0]: local grid barrier counter
  uint32_t* go_flag;    // sync[1]: local release flag
  uint32_t* epoch_ctr;  // sync[2]: launches so far (device resident so the kernel is graph-replayable)
  const float* state;   // device step state (hyper-parameters, see hyper_from_state)
  int use_adam;
};

// Watchdog for the cross-GPU spins: a rank that never arrives (crashed peer, diverged schedule, one rank
// skipped a bucket) would otherwise hang every GPU of the node silently.  After kDpWatchdogNs of spinning the waiting
// thread prints who waits for whom and traps, which surfaces on the host as a launch failure of this rank.
__device__ unsigned long long g_dp_watchdog_ns = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s; dp_set_watchdog_ms() overrides
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __noinline__ void dp_watchdog_fire(const DpArgs& a, uint32_t epoch, int peer, uint32_t seen, const char* what) {
  printf("[tnn_b200] k_dp_step watchdog: rank %d of %d waited > %llu ms on %s (peer %d shows %u, want >= %u, bucket [%lld, %lld))\n", a.rank, a.world,
         g_dp_watchdog_ns / 1000000ull, what, peer, seen, epoch, a.lo0, a.hi0);
  __trap();
}
__device__ __forceinline__ void dp_barrier(const DpArgs& a, uint32_t epoch, uint32_t grid_target) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(a.grid_ctr, 1u);
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      const unsigned long long t0 = globaltimer_ns();
      uint32_t seen;
      while ((seen = ld_acquire_gpu(a.grid_ctr)) < grid_target)
        if (globaltimer_ns() - t0 > g_dp_watchdog_ns) dp_watchdog_fire(a, grid_target, -1, seen, "the local grid barrier");
    }
    __syncthreads();
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(a.flag_peers[threadIdx.x] + a.rank, epoch);
      const unsigned long long t0 = globaltimer_ns();
      uint32_t seen;
      while ((seen = ld_acquire_sys(a.flag_peers[a.rank] + threadIdx.x)) < epoch)
        if (globaltimer_ns() - t0 > g_dp_watchdog_ns) dp_watchdog_fire(a, epoch, (int)threadIdx.x, seen, "a peer's barrier flag");
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(a.go_flag, epoch);
  }
  if (threadIdx.x == 0) {
    const unsigned long long t0 = globaltimer_ns();
    uint32_t seen;
    while ((seen = ld_acquire_gpu(a.go_flag)) < epoch)
      if (globaltimer_ns() - t0 > g_dp_watchdog_ns + 2000000000ull) dp_watchdog_fire(a, epoch, -2, seen, "the release of block 0");
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_dp_step(const DpArgs a) {
  const uint32_t G = gridDim.x;
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(a.epoch_ctr);
  const uint32_t grid_base = epoch * 2u * G;
  dp_barrier(a, epoch * 2 + 1, grid_base + G);

  const long long lo = a.lo0 + (long long)a.rank * a.chunk;
  const long long hi = min(a.hi0, lo + a.chunk);
  const long long nv = hi > lo ? (hi - lo + 3) / 4 : 0;   // slab is padded to a multiple of 4
  AdamHyper h = hyper_from_state(a.state);
  h.grad_scale /= (float)a.world;
  const float sgd_momentum = a.state[7];
  // The loop is LATENCY bound, not bandwidth bound: a multimem.ld_reduce through the switch (or a peer load) takes
  // microseconds, and with one float4 in flight per thread a 3 M-element slice needed ~22 dependent round trips per
  // thread (measured 120 us per bucket at 64 x 512 threads, inversely proportional to the CTA count).  Four
  // independent vectors per thread are issued before any of them is consumed.
  constexpr int U = 4;
  const long long T = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < nv; i0 += T * U) {
    float4 g[U], P[U], M[U], V[U];
    long long idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      idx[u] = i0 + u * T;
      if (idx[u] >= nv) continue;
      const long long e = lo + idx[u] * 4;
      if (a.g_mc) {
        g[u] = multimem_ld_reduce_v4(a.g_mc + e);
      } else {
        g[u] = ld_peer_v4(a.g_peers[0] + e);
        for (int r = 1; r < a.world; ++r) {
          float4 t = ld_peer_v4(a.g_peers[r] + e);
          g[u].x += t.x; g[u].y += t.y; g[u].z += t.z; g[u].w += t.w;
        }
      }
      P[u] = *reinterpret_cast<const float4*>(a.p_peers[a.rank] + e);
      M[u] = reinterpret_cast<float4*>(a.m + a.st_off)[idx[u]];
      if (a.use_adam) V[u] = reinterpret_cast<float4*>(a.v + a.st_off)[idx[u]];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (idx[u] >= nv) continue;
      const long long e = lo + idx[u] * 4;
      float4 gg = g[u];
      gg.x *= h.grad_scale; gg.y *= h.grad_scale; gg.z *= h.grad_scale; gg.w *= h.grad_scale;
      float4 Pu = P[u], Mu = M[u];
      if (a.use_adam) {
        float4 Vu = V[u];
        adam_update(Pu.x, gg.x, Mu.x, Vu.x, h);
        adam_update(Pu.y, gg.y, Mu.y, Vu.y, h);
        adam_update(Pu.z, gg.z, Mu.z, Vu.z, h);
        adam_update(Pu.w, gg.w, Mu.w, Vu.w, h);
        reinterpret_cast<float4*>(a.v + a.st_off)[idx[u]] = Vu;
      } else {
        const float mu = sgd_momentum, lr = h.lr, wd = h.weight_decay;
        gg.x += wd * Pu.x; gg.y += wd * Pu.y; gg.z += wd * Pu.z; gg.w += wd * Pu.w;
        Mu.x = mu * Mu.x - lr * gg.x; Mu.y = mu * Mu.y - lr * gg.y;
        Mu.z = mu * Mu.z - lr * gg.z; Mu.w = mu * Mu.w - lr * gg.w;
        Pu.x += Mu.x; Pu.y += Mu.y; Pu.z += Mu.z; Pu.w += Mu.w;
      }
      reinterpret_cast<float4*>(a.m + a.st_off)[idx[u]] = Mu;
      __nv_bfloat162 b0 = __floats2bfloat162_rn(Pu.x, Pu.y), b1 = __floats2bfloat162_rn(Pu.z, Pu.w);
      const uint32_t u0 = *reinterpret_cast<uint32_t*>(&b0), u1 = *reinterpret_cast<uint32_t*>(&b1);
      if (a.p_mc) {
        multimem_st_v4(a.p_mc + e, Pu);
        multimem_st_bf16x4(a.pb_mc + e, u0, u1);
      } else {
        for (int r = 0; r < a.world; ++r) {
          *reinterpret_cast<float4*>(a.p_peers[r] + e) = Pu;
          *reinterpret_cast<uint2*>(a.pb_peers[r] + e) = make_uint2(u0, u1);
        }
      }
    }
  }

  dp_barrier(a, epoch * 2 + 2, grid_base + 2 * G);

  // every rank is done reading my gradients: clear them for the next accumulation
  float4* gz = reinterpret_cast<float4*>(a.g_peers[a.rank] + a.lo0);
  const long long nz = (a.hi0 - a.lo0 + 3) / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nz;
       i += (long long)gridDim.x * blockDim.x)
    gz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // every thread of this launch read `epoch` before the first barrier; safe to advance it now
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.epoch_ctr = epoch + 1;
}

// Raw-argument entry used by the binding layer (keeps DpArgs private to this TU).
void dp_fused_step(int world, int rank, long long n, long long chunk, float* const* g_peers,
                   float* const* p_peers, void* const* pb_peers, uint32_t* const* flag_peers,
                   const float* g_mc, float* p_mc, void* pb_mc, float* m, float* v,
                   uint32_t* sync, const float* state, int use_adam, int ctas, long long lo0, long long hi0,
                   long long st_off, cudaStream_t s) {
  if (world > kMaxWorld) {
    fprintf(stderr, "[tnn_b200] dp_fused_step: world %d > %d\n", world, kMaxWorld);
    abort();
  }
  DpArgs a{};
  a.world = world;
  a.rank = rank;
  a.n = n;
  a.chunk = chunk;
  a.lo0 = lo0;
  a.hi0 = hi0 < 0 ? n : hi0;
  a.st_off = st_off;
  for (int r = 0; r < world; ++r) {
    a.g_peers[r] = g_peers[r];
    a.p_peers[r] = p_peers[r];
    a.pb_peers[r] = (bf16*)pb_peers[r];
    a.flag_peers[r] = flag_peers[r];
  }
  a.g_mc = g_mc;
  a.p_mc = p_mc;
  a.pb_mc = (bf16*)pb_mc;
  a.m = m;
  a.v = v;
  a.grid_ctr = sync;
  a.go_flag = sync + 1;
  a.epoch_ctr = sync + 2;
  a.state = state;
  a.use_adam = use_adam;
  k_dp_step<<<ctas, 256, 0, s>>>(a);
  TNN_CHECK_LAUNCH();
}

void dp_set_watchdog_ms(unsigned long long ms) {
  const unsigned long long ns = ms * 1000000ull;
  cudaMemcpyToSymbol(g_dp_watchdog_ns, &ns, sizeof(ns));
}

// Stand-alone cross-rank barrier (used by tests and by the pipeline runtime).
__global__ void k_p2p_barrier(uint32_t* const* flags, int world, int rank, uint32_t epoch) {
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(flags[threadIdx.x] + rank, epoch);
    while (ld_acquire_sys(flags[rank] + threadIdx.x) < epoch) {
    }
  }
}
void p2p_barrier(uint32_t* const* flags, int world, int rank, uint32_t epoch, cudaStream_t s) {
  k_p2p_barrier<<<1, 32, 0, s>>>(flags, world, rank, epoch);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
