// Token + positional embedding (fused gather) and row argmax (greedy sampling on device).
// Reference: src/nn/layers_impl/cuda/embedding_ops.cu:17-75 (float token ids, separate
// positional layer, host argmax over a 103 MB D2H copy in examples/gpt2_inference.cpp:94-121).
#include "common.cuh"
#include "simt_ops.cuh"

namespace tnn {

__global__ void __launch_bounds__(256)
    k_embed_fwd(const int* __restrict__ ids, const float* __restrict__ table,
                const float* __restrict__ pos, bf16* __restrict__ y, long long T, int L, int D, int V) {
  const long long total = T * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i / D;
    int d = (int)(i - t * D);
    int id = ids[t];
    float v = (id >= 0 && id < V) ? table[(long long)id * D + d] : 0.f;
    if (pos) v += pos[(t % L) * D + d];
    y[i] = __float2bfloat16(v);
  }
}
void embedding_fwd(const int* ids, const float* table, const float* pos_table, void* y, long long T,
                   int L, int D, int V, cudaStream_t s) {
  if (T <= 0) return;
  k_embed_fwd<<<grid_for(T * D, 256), 256, 0, s>>>(ids, table, pos_table, (bf16*)y, T, L, D, V);
  TNN_CHECK_LAUNCH();
}
__global__ void __launch_bounds__(256)
    k_embed_bwd(const int* __restrict__ ids, const bf16* __restrict__ dy, float* __restrict__ dtable,
                float* __restrict__ dpos, long long T, int L, int D, int V) {
  const long long total = T * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i / D;
    int d = (int)(i - t * D);
    float g = __bfloat162float(dy[i]);
    int id = ids[t];
    if (dtable && id >= 0 && id < V) atomicAdd(dtable + (long long)id * D + d, g);
    if (dpos) atomicAdd(dpos + (t % L) * D + d, g);
  }
}
void embedding_bwd(const int* ids, const void* dy, float* dtable, float* dpos, long long T, int L,
                   int D, int V, cudaStream_t s) {
  if (T <= 0) return;
  k_embed_bwd<<<grid_for(T * D, 256), 256, 0, s>>>(ids, (const bf16*)dy, dtable, dpos, T, L, D, V);
  TNN_CHECK_LAUNCH();
}

// Row arg-max (greedy decoding over [B, V] logits, V = 50257).  One CTA of 1024 threads per row;
// bf16 rows whose stride allows it are read as 16-byte vectors (8 logits per load, ~6 loads per thread
// at GPT-2's vocabulary instead of ~200 dependent 2-byte loads: 97 us -> a few us in the decode graph).
// Ties resolve to the smallest index.
template <typename T>
__global__ void __launch_bounds__(1024)
    k_argmax_rows(const T* __restrict__ x, long long ld, long long rows, int V, int* __restrict__ out, int vec_ok) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: dependents may start their prologue
  __shared__ float rv[32];
  __shared__ int ri[32];
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    float mx = -INFINITY;
    int am = 0x7fffffff;
    const T* row = x + r * ld;
    if (sizeof(T) == 2 && vec_ok) {
      const int nv = V >> 3;
      for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        float f[8];
        bf8_unpack(__ldg(reinterpret_cast<const uint4*>(row) + i), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > mx) {
            mx = f[j];
            am = i * 8 + j;
          }
      }
      for (int c = (nv << 3) + threadIdx.x; c < V; c += blockDim.x) {
        const float v = (float)row[c];
        if (v > mx || (v == mx && c < am)) {
          mx = v;
          am = c;
        }
      }
    } else {
      for (int c = threadIdx.x; c < V; c += blockDim.x) {
        const float v = (float)row[c];
        if (v > mx) {
          mx = v;
          am = c;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, mx, o);
      int oi = __shfl_xor_sync(0xffffffffu, am, o);
      if (ov > mx || (ov == mx && oi < am)) {
        mx = ov;
        am = oi;
      }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) {
      rv[w] = mx;
      ri[w] = am;
    }
    __syncthreads();
    if (w == 0) {
      mx = lane < nw ? rv[lane] : -INFINITY;
      am = lane < nw ? ri[lane] : 0x7fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, mx, o);
        int oi = __shfl_xor_sync(0xffffffffu, am, o);
        if (ov > mx || (ov == mx && oi < am)) {
          mx = ov;
          am = oi;
        }
      }
      if (lane == 0) out[r] = (am == 0x7fffffff) ? 0 : am;
    }
  }
}
void argmax_rows(const void* x, int is_f32, long long ld, long long rows, int V, int* out,
                 cudaStream_t s) {
  if (rows <= 0) return;
  int blocks = (int)(rows < kSMs * 2 ? rows : kSMs * 2);
  const int vec_ok = (!is_f32 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
  if (is_f32) k_argmax_rows<float><<<blocks, 1024, 0, s>>>((const float*)x, ld, rows, V, out, 0);
  else k_argmax_rows<bf16><<<blocks, 1024, 0, s>>>((const bf16*)x, ld, rows, V, out, vec_ok);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
