// CPU elementwise / reduction family for fp32 and fp64 (reference: include/ops/cpu/kernels.hpp:16-493,
// src/ops/cpu/skernels.cpp, dkernels.cpp -- hand-written AVX2 loops per op and per dtype).
//
// Here every op is ONE templated loop that the compiler vectorises; the loops are compiled twice
// (AVX2+FMA and baseline x86-64) through GCC function multi-versioning, so the wide version is
// (the loops carry optimize("O3"): the tree's -O2 uses GCC's very-cheap vectoriser cost model, which skips
// loops with a run-time trip count) picked at load time on CPUs that have it -- the same run-time dispatch the dot product uses.
// Work is split over the shared thread pool in cache-sized grains; reductions accumulate per
// chunk in double and are combined in chunk order, so results do not depend on the thread count.
//
// Op codes match the CUDA family (csrc/kernels/tensor_ops.cu) so ops/tensor_ops.py dispatches both
// with the same tables.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "cpu_math.hpp"

namespace tnn {
namespace cpuops {

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define TNN_MV __attribute__((target_clones("arch=haswell", "default"), optimize("O3")))
#else
#define TNN_MV
#endif

enum BinOp { ADD = 0, SUB, MUL, DIV, MIN, MAX, EQ, GT };
enum UnOp { SQRT = 0, RSQRT, RCP, ABS, CLAMP, NEG, EXP, LOG, SQUARE };
enum TerOp { FMADD = 0, FMSUB, FNMADD };
enum RedOp { SUM = 0, DOT, SQDIFF };

constexpr int64_t kGrain = 1 << 15;   // elements per task: 128 KiB of fp32, fits L2 with both operands

template <class T, int OP>
static inline T bin1(T a, T b) {
  if (OP == ADD) return a + b;
  if (OP == SUB) return a - b;
  if (OP == MUL) return a * b;
  if (OP == DIV) return a / b;
  if (OP == MIN) return a < b ? a : b;
  if (OP == MAX) return a > b ? a : b;
  if (OP == EQ) return a == b ? T(1) : T(0);
  return a > b ? T(1) : T(0);
}

// b == nullptr: the right operand is the scalar s
template <class T, int OP>
TNN_MV static void bin_loop(const T* __restrict__ a, const T* __restrict__ b, T s, T* __restrict__ o, int64_t n) {
  if (b) {
    for (int64_t i = 0; i < n; ++i) o[i] = bin1<T, OP>(a[i], b[i]);
  } else {
    for (int64_t i = 0; i < n; ++i) o[i] = bin1<T, OP>(a[i], s);
  }
}
// in-place variants (o aliases a): no __restrict__ promise
template <class T, int OP>
TNN_MV static void bin_loop_inplace(T* a, const T* b, T s, int64_t n) {
  if (b) {
    for (int64_t i = 0; i < n; ++i) a[i] = bin1<T, OP>(a[i], b[i]);
  } else {
    for (int64_t i = 0; i < n; ++i) a[i] = bin1<T, OP>(a[i], s);
  }
}

template <class T, int OP>
static inline T un1(T a, T lo, T hi) {
  if (OP == SQRT) return std::sqrt(a);
  if (OP == RSQRT) return T(1) / std::sqrt(a);
  if (OP == RCP) return T(1) / a;
  if (OP == ABS) return a < T(0) ? -a : a;
  if (OP == CLAMP) return a < lo ? lo : (a > hi ? hi : a);
  if (OP == NEG) return -a;
  if (OP == EXP) return std::exp(a);
  if (OP == LOG) return std::log(a);
  return a * a;
}
template <class T, int OP>
TNN_MV static void un_loop(const T* a, T* o, T lo, T hi, int64_t n) {
  for (int64_t i = 0; i < n; ++i) o[i] = un1<T, OP>(a[i], lo, hi);
}

template <class T, int OP>
TNN_MV static void ter_loop(const T* a, const T* b, const T* c, T* o, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const T p = a[i] * b[i];
    o[i] = OP == FMADD ? p + c[i] : (OP == FMSUB ? p - c[i] : c[i] - p);
  }
}

template <class T>
TNN_MV static void axpy_loop(T* y, const T* x, T alpha, int64_t n) {
  for (int64_t i = 0; i < n; ++i) y[i] += alpha * x[i];
}
template <class T>
TNN_MV static void fill_loop(T* y, T v, int64_t n) {
  for (int64_t i = 0; i < n; ++i) y[i] = v;
}

// 8 independent double lanes: a fixed-width inner loop the vectoriser turns into two 4-wide FMA chains without
// needing -ffast-math to reassociate one serial chain
template <class T, int OP>
TNN_MV static double red_loop(const T* a, const T* b, double mu, int64_t n) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    for (int k = 0; k < 8; ++k) {
      const double x = (double)a[i + k];
      if (OP == SUM) acc[k] += x;
      else if (OP == DOT) acc[k] += x * (double)b[i + k];
      else acc[k] += (x - mu) * (x - mu);
    }
  }
  double tail = 0.0;
  for (; i < n; ++i) {
    const double x = (double)a[i];
    tail += OP == SUM ? x : (OP == DOT ? x * (double)b[i] : (x - mu) * (x - mu));
  }
  return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7])) + tail;
}

// ---- fused optimizer steps over the flat slabs (reference: include/nn/optimizers_impl/cpu/adam_kernels.hpp,
// sgd_kernels.hpp).  One pass over p, g (and m, v / velocity) instead of one pass per arithmetic op.
struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, grad_scale, bias1, bias2;   // bias_k = 1 - beta_k^t
  bool decoupled, zero_grad;
};
TNN_MV static void adam_loop(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             int64_t n, AdamArgs a) {
  const float inv_b1 = 1.f / a.bias1, inv_b2 = 1.f / a.bias2;
  const float decay = (a.weight_decay != 0.f && a.decoupled) ? 1.f - a.weight_decay * a.lr : 1.f;
  const float l2 = (a.weight_decay != 0.f && !a.decoupled) ? a.weight_decay * a.lr : 0.f;
  for (int64_t i = 0; i < n; ++i) {
    const float gg = g[i] * a.grad_scale;
    const float pi = p[i] * decay;
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * gg;
    const float vi = a.beta2 * v[i] + (1.f - a.beta2) * gg * gg;
    m[i] = mi;
    v[i] = vi;
    const float upd = a.lr * (mi * inv_b1) / (std::sqrt(vi * inv_b2) + a.eps) + l2 * pi;
    p[i] = pi - upd;
    if (a.zero_grad) g[i] = 0.f;
  }
}
struct SgdArgs {
  float lr, momentum, weight_decay, grad_scale;
  bool zero_grad;
};
TNN_MV static void sgd_loop(float* __restrict__ p, float* __restrict__ g, float* vel /*may be null*/, int64_t n, SgdArgs a) {
  if (vel) {
    for (int64_t i = 0; i < n; ++i) {
      const float gg = g[i] * a.grad_scale + a.weight_decay * p[i];
      const float vi = a.momentum * vel[i] - a.lr * gg;
      vel[i] = vi;
      p[i] += vi;
      if (a.zero_grad) g[i] = 0.f;
    }
  } else {
    for (int64_t i = 0; i < n; ++i) {
      const float gg = g[i] * a.grad_scale + a.weight_decay * p[i];
      p[i] -= a.lr * gg;
      if (a.zero_grad) g[i] = 0.f;
    }
  }
}

template <class F>
static void for_chunks(int64_t n, int threads, F&& f) {
  if (n <= 0) return;
  if (n <= kGrain) {
    f(0, n);
    return;
  }
  global_pool(threads).parallel_for(n, kGrain, [&](int64_t b, int64_t e) { f(b, e); });
}

template <class T>
void binary(const T* a, const T* b, T s, T* o, int64_t n, int op, int threads = 0) {
  const bool inplace = (o == a);
#define TNN_CASE(OPC)                                                                                   \
  case OPC:                                                                                             \
    for_chunks(n, threads, [&](int64_t lo, int64_t hi) {                                                \
      if (inplace) bin_loop_inplace<T, OPC>(o + lo, b ? b + lo : nullptr, s, hi - lo);                  \
      else bin_loop<T, OPC>(a + lo, b ? b + lo : nullptr, s, o + lo, hi - lo);                          \
    });                                                                                                 \
    break;
  switch (op) {
    TNN_CASE(ADD) TNN_CASE(SUB) TNN_CASE(MUL) TNN_CASE(DIV) TNN_CASE(MIN) TNN_CASE(MAX) TNN_CASE(EQ) TNN_CASE(GT)
    default: throw std::invalid_argument("cpu binary: unknown op");
  }
#undef TNN_CASE
}

template <class T>
void unary(const T* a, T* o, int64_t n, int op, T lo_v, T hi_v, int threads = 0) {
#define TNN_CASE(OPC)                                                                                   \
  case OPC:                                                                                             \
    for_chunks(n, threads, [&](int64_t lo, int64_t hi) { un_loop<T, OPC>(a + lo, o + lo, lo_v, hi_v, hi - lo); }); \
    break;
  switch (op) {
    TNN_CASE(SQRT) TNN_CASE(RSQRT) TNN_CASE(RCP) TNN_CASE(ABS) TNN_CASE(CLAMP) TNN_CASE(NEG) TNN_CASE(EXP) TNN_CASE(LOG)
    TNN_CASE(SQUARE)
    default: throw std::invalid_argument("cpu unary: unknown op");
  }
#undef TNN_CASE
}

template <class T>
void ternary(const T* a, const T* b, const T* c, T* o, int64_t n, int op, int threads = 0) {
  switch (op) {
    case FMADD: for_chunks(n, threads, [&](int64_t l, int64_t h) { ter_loop<T, FMADD>(a + l, b + l, c + l, o + l, h - l); }); break;
    case FMSUB: for_chunks(n, threads, [&](int64_t l, int64_t h) { ter_loop<T, FMSUB>(a + l, b + l, c + l, o + l, h - l); }); break;
    case FNMADD: for_chunks(n, threads, [&](int64_t l, int64_t h) { ter_loop<T, FNMADD>(a + l, b + l, c + l, o + l, h - l); }); break;
    default: throw std::invalid_argument("cpu ternary: unknown op");
  }
}

template <class T>
void axpy(T* y, const T* x, T alpha, int64_t n, int threads = 0) {
  for_chunks(n, threads, [&](int64_t l, int64_t h) { axpy_loop<T>(y + l, x + l, alpha, h - l); });
}
template <class T>
void fill(T* y, T v, int64_t n, int threads = 0) {
  for_chunks(n, threads, [&](int64_t l, int64_t h) { fill_loop<T>(y + l, v, h - l); });
}

inline void adam_step(float* p, float* g, float* m, float* v, int64_t n, AdamArgs a, int threads = 0) {
  for_chunks(n, threads, [&](int64_t l, int64_t h) { adam_loop(p + l, g + l, m + l, v + l, h - l, a); });
}
inline void sgd_step(float* p, float* g, float* vel, int64_t n, SgdArgs a, int threads = 0) {
  for_chunks(n, threads, [&](int64_t l, int64_t h) { sgd_loop(p + l, g + l, vel ? vel + l : nullptr, h - l, a); });
}

// sum / dot / sum of squared differences, double accumulation, deterministic in the thread count
template <class T>
double reduce(const T* a, const T* b, double mu, int64_t n, int op, int threads = 0) {
  if (n <= 0) return 0.0;
  const int64_t chunks = (n + kGrain - 1) / kGrain;
  std::vector<double> part((size_t)chunks, 0.0);
  auto body = [&](int64_t lo, int64_t hi) {
    // the pool hands out whole grains, so lo is a multiple of kGrain
    for (int64_t c = lo / kGrain; c * kGrain < hi; ++c) {
      const int64_t s = c * kGrain, e = std::min(n, s + kGrain);
      double r;
      if (op == SUM) r = red_loop<T, SUM>(a + s, nullptr, 0.0, e - s);
      else if (op == DOT) r = red_loop<T, DOT>(a + s, b + s, 0.0, e - s);
      else r = red_loop<T, SQDIFF>(a + s, nullptr, mu, e - s);
      part[(size_t)c] = r;
    }
  };
  if (chunks == 1) body(0, n);
  else global_pool(threads).parallel_for(n, kGrain, body);
  double total = 0.0;
  for (double p : part) total += p;
  return total;
}

}  // namespace cpuops
}  // namespace tnn
