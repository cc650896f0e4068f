// Probe for NEXT_STEPS item 1 (conv main loop: one halo'd input window in shared memory, the 9 filter taps
// addressed by offsetting the A-operand descriptor).  Question it answers on a real B200:
//
//   Does tcgen05.mma read a K-major SWIZZLE_128B operand correctly when the descriptor START ADDRESS is not
//   1024-byte aligned (start = window + dh*1280 + dw*128) and the 8-row group stride (SBO) is 1280 bytes,
//   i.e. is the 128B swizzle a function of the absolute shared-memory address bits (mode 0), or does the
//   descriptor need its base_offset field (bits 49..51) set to (start >> 7) & 7 (mode 1)?
//
// Layout under test: window = 18 image rows x 10 pixels, each pixel = 64 bf16 channels = one 128-byte row,
// stored with the swizzle TMA would apply (16-byte chunk index ^= address bits [7,10)).  The MMA tile is
// M = 128 output pixels (16 rows x 8 pixels) x N = 64 x K = 64; A row m = (r, c) reads window pixel
// (r + dh, c + dw).  Every 8-row core-matrix group is 8 consecutive pixels of one image row, group stride
// 10 pixels = 1280 B.
//
// Build + run (single GPU, a few milliseconds):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tnn_b200/csrc/kernels \
//        tools/experiments/umma_window_probe.cu -o gpurun_out/umma_window_probe && gpurun_out/umma_window_probe
// Prints one line per (mode, tap) and a summary; exit code 0 iff at least one mode is correct for all 9 taps.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100.cuh"

using namespace tnn::sm100;

static constexpr int WR = 18, WC = 10, CH = 64;          // window rows, pixels per row, channels (K)
static constexpr int NPIX = WR * WC;                     // 180 pixel rows of 128 B
static constexpr int TM = 128, TN = 64;
static constexpr uint32_t kWindowBytes = 24 * 1024;      // 180 * 128 = 23040, rounded up
static constexpr uint32_t kBBytes = TN * 128;

__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* __restrict__ win /*[180][64]*/,
                                             const __nv_bfloat16* __restrict__ bmat /*[64 n][64 k]*/, float* __restrict__ out /*[128][64]*/,
                                             int dh, int dw, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = base;
  uint8_t* sB = base + kWindowBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kBBytes);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  // stage operands with the 128B swizzle keyed on the offset from the 1024-aligned base (== absolute address bits)
  for (int i = tid; i < NPIX * 8; i += 128) {            // 16-byte chunks
    const uint32_t off = i * 16;
    *reinterpret_cast<uint4*>(sW + swz(off)) = reinterpret_cast<const uint4*>(win)[i];
  }
  for (int i = tid; i < TN * 8; i += 128) {
    const uint32_t off = i * 16;
    *reinterpret_cast<uint4*>(sB + swz(off)) = reinterpret_cast<const uint4*>(bmat)[i];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0) tmem_alloc<64>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sW) + dh * (WC * 128) + dw * 128;
    const uint32_t b0 = smem_u32(sB);
    const uint32_t idesc = make_idesc_bf16(TM, TN, 0, 0);
    for (int k = 0; k < CH / 16; ++k) {
      uint64_t ad = make_smem_desc_sw128(a0 + k * 32, 0, WC * 128);
      if (mode == 1) ad |= static_cast<uint64_t>((a0 >> 7) & 7u) << 49;        // matrix base offset
      const uint64_t bd = make_smem_desc_sw128(b0 + k * 32, 0, 1024);
      umma_f16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c = 0; c < TN; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + (tid & 31)) * TN + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
  std::vector<__nv_bfloat16> hw(NPIX * CH), hb(TN * CH);
  std::vector<float> fw(NPIX * CH), fb(TN * CH);
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 9) & 0xFF) / 64.0f - 2.0f; };   // exact in bf16
  for (size_t i = 0; i < hw.size(); ++i) { fw[i] = rnd(); hw[i] = __float2bfloat16(fw[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = rnd(); hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *dw_, *db_;
  float* dout;
  cudaMalloc(&dw_, hw.size() * 2);
  cudaMalloc(&db_, hb.size() * 2);
  cudaMalloc(&dout, TM * TN * 4);
  cudaMemcpy(dw_, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db_, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = kWindowBytes + kBBytes + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ho(TM * TN);
  int ok_taps[2] = {0, 0};
  for (int mode = 0; mode < 2; ++mode)
    for (int dh = 0; dh < 3; ++dh)
      for (int dwv = 0; dwv < 3; ++dwv) {
        cudaMemset(dout, 0, TM * TN * 4);
        probe<<<1, 128, smem>>>(dw_, db_, dout, dh, dwv, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d tap (%d,%d): CUDA error %s\n", mode, dh, dwv, cudaGetErrorString(e));
          return 2;
        }
        cudaMemcpy(ho.data(), dout, TM * TN * 4, cudaMemcpyDeviceToHost);
        double maxdiff = 0;
        for (int m = 0; m < TM; ++m) {
          const int pix = (m / 8 + dh) * WC + (m % 8 + dwv);
          for (int n = 0; n < TN; ++n) {
            double ref = 0;
            for (int k = 0; k < CH; ++k) ref += (double)fw[pix * CH + k] * fb[n * CH + k];
            maxdiff = std::fmax(maxdiff, std::fabs(ref - ho[m * TN + n]));
          }
        }
        const bool ok = maxdiff < 1e-3;
        ok_taps[mode] += ok;
        printf("mode %d (%s) tap (%d,%d) start%%1024=%4d : max|diff| %.4g %s\n", mode, mode ? "base_offset field" : "absolute-address swizzle", dh,
               dwv, (dh * WC * 128 + dwv * 128) % 1024, maxdiff, ok ? "OK" : "MISMATCH");
      }
  printf("SUMMARY absolute-address swizzle: %d/9 taps correct; base_offset field: %d/9 taps correct\n", ok_taps[0], ok_taps[1]);
  return (ok_taps[0] == 9 || ok_taps[1] == 9) ? 0 : 1;
}
