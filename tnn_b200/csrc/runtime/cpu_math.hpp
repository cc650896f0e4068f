// CPU math for the plumbing path (BASELINE config 1: mnist_cnn fp32 on CPU) and the CPU oracles:
// a cache-blocked, register-tiled SGEMM (NN/NT/TN/TT through strides) parallelised over row
// panels with std::thread, plus a thread-pool parallel_for.  The reference ships hand-written
// AVX2 kernels / MKL (src/math/cpu/sgemm.cpp:15-656, include/threading/thread_handler.hpp:37-140);
// here: Goto-style packing with a 6x16 AVX2+FMA intrinsics micro-kernel selected at run time (scalar tile otherwise).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace tnn {

class ThreadPool {
 public:
  explicit ThreadPool(int n) : stop_(false) {
    n = std::max(1, n);
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }

  // fn(begin, end) over [0, n) in `grain`-sized chunks; returns when all chunks ran
  void parallel_for(int64_t n, int64_t grain, const std::function<void(int64_t, int64_t)>& fn) {
    if (n <= 0) return;
    grain = std::max<int64_t>(1, grain);
    const int64_t chunks = (n + grain - 1) / grain;
    if (chunks == 1 || workers_.size() == 1) {
      fn(0, n);
      return;
    }
    // The bookkeeping lives on the heap and is shared with the helpers: a helper that only gets scheduled
    // after every chunk is done (and this call has returned) must still find valid counters to see that.
    struct State {
      std::atomic<int64_t> next{0}, done{0};
      std::mutex m;
      std::condition_variable cv;
      int64_t chunks = 0, grain = 0, n = 0;
      const std::function<void(int64_t, int64_t)>* fn = nullptr;   // only dereferenced while chunks remain
    };
    auto st = std::make_shared<State>();
    st->chunks = chunks;
    st->grain = grain;
    st->n = n;
    st->fn = &fn;
    auto body = [st] {
      for (;;) {
        const int64_t c = st->next.fetch_add(1);
        if (c >= st->chunks) break;
        (*st->fn)(c * st->grain, std::min(st->n, (c + 1) * st->grain));
        if (st->done.fetch_add(1) + 1 == st->chunks) {
          std::lock_guard<std::mutex> g(st->m);
          st->cv.notify_all();
        }
      }
    };
    const int helpers = (int)std::min<int64_t>(chunks - 1, (int64_t)workers_.size());
    {
      std::lock_guard<std::mutex> g(mu_);
      for (int i = 0; i < helpers; ++i) tasks_.push_back(body);
    }
    cv_.notify_all();
    body();
    std::unique_lock<std::mutex> lk(st->m);
    st->cv.wait(lk, [&] { return st->done.load() == st->chunks; });
  }

 private:
  void loop() {
    for (;;) {
      std::function<void()> t;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !tasks_.empty(); });
        if (stop_ && tasks_.empty()) return;
        t = std::move(tasks_.back());
        tasks_.pop_back();
      }
      t();
    }
  }
  std::vector<std::thread> workers_;
  std::vector<std::function<void()>> tasks_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_;
};

inline ThreadPool& global_pool(int threads = 0) {
  static std::unique_ptr<ThreadPool> pool;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (!pool || (threads > 0 && pool->size() != threads))
    pool.reset(new ThreadPool(threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency())));
  return *pool;
}

// ---------------------------------------------------------------------------------------------- SGEMM
// C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C.  Element (i,k) of op(A) is A[i*sa_m + k*sa_k], so the four
// NN/NT/TN/TT cases are one code path (reference: src/math/cpu/sgemm.cpp:15-656, four hand-written variants).
//
// Goto-style blocking: B is packed once per (KC x NC) block into 16-wide column panels, each task packs its own
// (MC x KC) block of A into 6-tall row panels, and a 6x16 register-tiled micro-kernel (12 ymm accumulators, two
// B loads + six broadcasts + twelve FMAs per k) runs over the panels.  Packing zero-pads, so the micro-kernel
// has no edge cases; partial tiles go through a small stack buffer.  Rows of C are split over the thread pool.
// The AVX2+FMA micro-kernel is chosen at run time; other CPUs get the same blocking with a scalar tile.
namespace sgemm_detail {
constexpr int64_t MR = 6, NR = 16, MC = 96, KC = 256, NC = 4096;

inline void micro_scalar(int64_t kc, const float* ap, const float* bp, float* acc /*[MR*NR]*/) {
  for (int64_t i = 0; i < MR * NR; ++i) acc[i] = 0.f;
  for (int64_t k = 0; k < kc; ++k) {
    const float* a = ap + k * MR;
    const float* b = bp + k * NR;
    for (int64_t i = 0; i < MR; ++i)
      for (int64_t j = 0; j < NR; ++j) acc[i * NR + j] += a[i] * b[j];
  }
}

#if defined(__x86_64__)
__attribute__((target("avx2,fma"))) inline void micro_avx2(int64_t kc, const float* ap, const float* bp, float* acc) {
  __m256 c00 = _mm256_setzero_ps(), c01 = _mm256_setzero_ps(), c10 = _mm256_setzero_ps(), c11 = _mm256_setzero_ps();
  __m256 c20 = _mm256_setzero_ps(), c21 = _mm256_setzero_ps(), c30 = _mm256_setzero_ps(), c31 = _mm256_setzero_ps();
  __m256 c40 = _mm256_setzero_ps(), c41 = _mm256_setzero_ps(), c50 = _mm256_setzero_ps(), c51 = _mm256_setzero_ps();
  for (int64_t k = 0; k < kc; ++k) {
    const __m256 b0 = _mm256_loadu_ps(bp), b1 = _mm256_loadu_ps(bp + 8);
    __m256 a;
    a = _mm256_broadcast_ss(ap + 0); c00 = _mm256_fmadd_ps(a, b0, c00); c01 = _mm256_fmadd_ps(a, b1, c01);
    a = _mm256_broadcast_ss(ap + 1); c10 = _mm256_fmadd_ps(a, b0, c10); c11 = _mm256_fmadd_ps(a, b1, c11);
    a = _mm256_broadcast_ss(ap + 2); c20 = _mm256_fmadd_ps(a, b0, c20); c21 = _mm256_fmadd_ps(a, b1, c21);
    a = _mm256_broadcast_ss(ap + 3); c30 = _mm256_fmadd_ps(a, b0, c30); c31 = _mm256_fmadd_ps(a, b1, c31);
    a = _mm256_broadcast_ss(ap + 4); c40 = _mm256_fmadd_ps(a, b0, c40); c41 = _mm256_fmadd_ps(a, b1, c41);
    a = _mm256_broadcast_ss(ap + 5); c50 = _mm256_fmadd_ps(a, b0, c50); c51 = _mm256_fmadd_ps(a, b1, c51);
    ap += MR;
    bp += NR;
  }
  _mm256_storeu_ps(acc + 0 * NR, c00); _mm256_storeu_ps(acc + 0 * NR + 8, c01);
  _mm256_storeu_ps(acc + 1 * NR, c10); _mm256_storeu_ps(acc + 1 * NR + 8, c11);
  _mm256_storeu_ps(acc + 2 * NR, c20); _mm256_storeu_ps(acc + 2 * NR + 8, c21);
  _mm256_storeu_ps(acc + 3 * NR, c30); _mm256_storeu_ps(acc + 3 * NR + 8, c31);
  _mm256_storeu_ps(acc + 4 * NR, c40); _mm256_storeu_ps(acc + 4 * NR + 8, c41);
  _mm256_storeu_ps(acc + 5 * NR, c50); _mm256_storeu_ps(acc + 5 * NR + 8, c51);
}
#endif

inline bool have_avx2() {
#if defined(__x86_64__)
  static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  return ok;
#else
  return false;
#endif
}
inline void micro(int64_t kc, const float* ap, const float* bp, float* acc) {
#if defined(__x86_64__)
  if (have_avx2()) return micro_avx2(kc, ap, bp, acc);
#endif
  micro_scalar(kc, ap, bp, acc);
}
}  // namespace sgemm_detail

inline void cpu_sgemm(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, bool transA, bool transB,
                      float alpha, float beta, int64_t lda, int64_t ldb, int64_t ldc, int threads = 0) {
  using namespace sgemm_detail;
  if (M <= 0 || N <= 0) return;
  const int64_t sa_m = transA ? 1 : lda, sa_k = transA ? lda : 1;
  const int64_t sb_k = transB ? 1 : ldb, sb_n = transB ? ldb : 1;
  if (K <= 0 || alpha == 0.f) {                       // C = beta * C
    for (int64_t i = 0; i < M; ++i)
      for (int64_t j = 0; j < N; ++j) C[i * ldc + j] = beta == 0.f ? 0.f : beta * C[i * ldc + j];
    return;
  }
  ThreadPool& pool = global_pool(threads);
  std::vector<float> bpack;
  for (int64_t jc = 0; jc < N; jc += NC) {
    const int64_t nc = std::min(NC, N - jc);
    const int64_t npanels = (nc + NR - 1) / NR;
    for (int64_t pc = 0; pc < K; pc += KC) {
      const int64_t kc = std::min(KC, K - pc);
      const float beta_eff = pc == 0 ? beta : 1.f;    // later K blocks accumulate onto the first one's result
      bpack.resize((size_t)npanels * kc * NR);
      float* bpk = bpack.data();
      pool.parallel_for(npanels, 4, [&](int64_t p0, int64_t p1) {
        for (int64_t p = p0; p < p1; ++p) {
          float* dst = bpk + p * kc * NR;
          const int64_t j0 = jc + p * NR, nr = std::min(NR, jc + nc - j0);
          for (int64_t k = 0; k < kc; ++k) {
            const float* src = B + (pc + k) * sb_k + j0 * sb_n;
            int64_t j = 0;
            for (; j < nr; ++j) dst[k * NR + j] = src[j * sb_n];
            for (; j < NR; ++j) dst[k * NR + j] = 0.f;
          }
        }
      });
      pool.parallel_for(M, MC, [&](int64_t m0, int64_t m1) {
        // parallel_for may hand over several MC blocks at once (single-thread / single-chunk case)
        std::vector<float> apack((size_t)((MC + MR - 1) / MR) * MR * kc);
        alignas(32) float acc[MR * NR];
        for (int64_t ic = m0; ic < m1; ic += MC) {
          const int64_t mc = std::min(MC, m1 - ic);
          const int64_t mpanels = (mc + MR - 1) / MR;
          for (int64_t q = 0; q < mpanels; ++q) {
            float* dst = apack.data() + q * kc * MR;
            const int64_t i0 = ic + q * MR, mr = std::min(MR, ic + mc - i0);
            for (int64_t k = 0; k < kc; ++k) {
              const float* src = A + i0 * sa_m + (pc + k) * sa_k;
              int64_t i = 0;
              for (; i < mr; ++i) dst[k * MR + i] = src[i * sa_m];
              for (; i < MR; ++i) dst[k * MR + i] = 0.f;
            }
          }
          for (int64_t p = 0; p < npanels; ++p) {
            const int64_t j0 = jc + p * NR, nr = std::min(NR, jc + nc - j0);
            for (int64_t q = 0; q < mpanels; ++q) {
              const int64_t i0 = ic + q * MR, mr = std::min(MR, ic + mc - i0);
              micro(kc, apack.data() + q * kc * MR, bpk + p * kc * NR, acc);
              for (int64_t i = 0; i < mr; ++i) {
                float* c = C + (i0 + i) * ldc + j0;
                const float* r = acc + i * NR;
                if (beta_eff == 0.f) {
                  for (int64_t j = 0; j < nr; ++j) c[j] = alpha * r[j];
                } else {
                  for (int64_t j = 0; j < nr; ++j) c[j] = alpha * r[j] + beta_eff * c[j];
                }
              }
            }
          }
        }
      });
    }
  }
}


// ---- SIMD dot product (reference: include/utils/simd_asm.hpp:13-15, hand-written AVX2 assembly).  Here: AVX2 +
// FMA intrinsics in a function compiled for that target and selected at run time, four independent
// accumulators (32 floats per iteration), scalar fallback on CPUs without AVX2.
#if defined(__x86_64__)
__attribute__((target("avx2,fma"))) inline float cpu_dot_avx2(const float* a, const float* b, int64_t n) {
  __m256 s0 = _mm256_setzero_ps(), s1 = _mm256_setzero_ps(), s2 = _mm256_setzero_ps(), s3 = _mm256_setzero_ps();
  int64_t i = 0;
  for (; i + 32 <= n; i += 32) {
    s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s0);
    s1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8), s1);
    s2 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 16), _mm256_loadu_ps(b + i + 16), s2);
    s3 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 24), _mm256_loadu_ps(b + i + 24), s3);
  }
  for (; i + 8 <= n; i += 8) s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s0);
  __m256 s = _mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3));
  __m128 lo = _mm256_castps256_ps128(s), hi = _mm256_extractf128_ps(s, 1);
  lo = _mm_add_ps(lo, hi);
  lo = _mm_hadd_ps(lo, lo);
  lo = _mm_hadd_ps(lo, lo);
  float r = _mm_cvtss_f32(lo);
  for (; i < n; ++i) r += a[i] * b[i];
  return r;
}
#endif
inline float cpu_dot(const float* a, const float* b, int64_t n) {
#if defined(__x86_64__)
  static const bool has_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  if (has_avx2) return cpu_dot_avx2(a, b, n);
#endif
  double r = 0.0;
  for (int64_t i = 0; i < n; ++i) r += (double)a[i] * b[i];
  return (float)r;
}

}  // namespace tnn
