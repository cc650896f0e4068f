// CPU math for the plumbing path (BASELINE config 1: mnist_cnn fp32 on CPU) and the CPU oracles:
// a cache-blocked, register-tiled SGEMM (NN/NT/TN/TT through strides) parallelised over row
// panels with std::thread, plus a thread-pool parallel_for.  The reference ships hand-written
// AVX2 kernels / MKL (src/math/cpu/sgemm.cpp:15-656, include/threading/thread_handler.hpp:37-140);
// here the inner 4x16 micro-kernel is plain C++ that GCC vectorises (-O2 -ftree-vectorize).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

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
    std::atomic<int64_t> next{0}, done{0};
    std::mutex m;
    std::condition_variable cv;
    auto body = [&] {
      for (;;) {
        const int64_t c = next.fetch_add(1);
        if (c >= chunks) break;
        fn(c * grain, std::min(n, (c + 1) * grain));
        if (done.fetch_add(1) + 1 == chunks) {
          std::lock_guard<std::mutex> g(m);
          cv.notify_all();
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
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return done.load() == chunks; });
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

// C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C.  Element (i,k) of op(A) is A[i*sa_m + k*sa_k].
inline void cpu_sgemm(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, bool transA, bool transB,
                      float alpha, float beta, int64_t lda, int64_t ldb, int64_t ldc, int threads = 0) {
  const int64_t sa_m = transA ? 1 : lda, sa_k = transA ? lda : 1;
  const int64_t sb_k = transB ? 1 : ldb, sb_n = transB ? ldb : 1;
  constexpr int64_t MB = 64, KB = 256, MR = 4, NR = 16;
  auto panel = [&](int64_t m_begin, int64_t m_end) {
    std::vector<float> bpack((size_t)KB * NR);
    for (int64_t i0 = m_begin; i0 < m_end; i0 += MR) {
      const int64_t mr = std::min(MR, m_end - i0);
      for (int64_t j0 = 0; j0 < N; j0 += NR) {
        const int64_t nr = std::min(NR, N - j0);
        float acc[MR][NR] = {};
        for (int64_t k0 = 0; k0 < K; k0 += KB) {
          const int64_t kb = std::min(KB, K - k0);
          for (int64_t k = 0; k < kb; ++k)
            for (int64_t j = 0; j < NR; ++j) bpack[k * NR + j] = j < nr ? B[(k0 + k) * sb_k + (j0 + j) * sb_n] : 0.f;
          for (int64_t k = 0; k < kb; ++k) {
            const float* bp = &bpack[k * NR];
            for (int64_t i = 0; i < mr; ++i) {
              const float a = A[(i0 + i) * sa_m + (k0 + k) * sa_k];
              for (int64_t j = 0; j < NR; ++j) acc[i][j] += a * bp[j];
            }
          }
        }
        for (int64_t i = 0; i < mr; ++i)
          for (int64_t j = 0; j < nr; ++j) {
            float* c = &C[(i0 + i) * ldc + j0 + j];
            *c = alpha * acc[i][j] + (beta == 0.f ? 0.f : beta * *c);
          }
      }
    }
  };
  global_pool(threads).parallel_for(M, MB, panel);
}


// ---- SIMD dot product (reference: include/utils/simd_asm.hpp:13-15, hand-written AVX2 assembly).  Here: AVX2 +
// FMA intrinsics in a function compiled for that target and selected at run time, four independent
// accumulators (32 floats per iteration), scalar fallback on CPUs without AVX2.
#if defined(__x86_64__)
#include <immintrin.h>
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
