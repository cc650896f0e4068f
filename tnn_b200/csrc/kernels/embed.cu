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

template <typename T>
__global__ void __launch_bounds__(256)
    k_argmax_rows(const T* __restrict__ x, long long ld, long long rows, int V, int* __restrict__ out) {
  __shared__ float rv[32];
  __shared__ int ri[32];
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    float mx = -INFINITY;
    int am = 0;
    for (int c = threadIdx.x; c < V; c += blockDim.x) {
      float v = (float)x[r * ld + c];
      if (v > mx) {
        mx = v;
        am = c;
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
      if (lane == 0) out[r] = am;
    }
  }
}
void argmax_rows(const void* x, int is_f32, long long ld, long long rows, int V, int* out,
                 cudaStream_t s) {
  if (rows <= 0) return;
  int blocks = (int)(rows < kSMs * 4 ? rows : kSMs * 4);
  if (is_f32) k_argmax_rows<float><<<blocks, 256, 0, s>>>((const float*)x, ld, rows, V, out);
  else k_argmax_rows<bf16><<<blocks, 256, 0, s>>>((const bf16*)x, ld, rows, V, out);
  TNN_CHECK_LAUNCH();
}

}  // namespace tnn
