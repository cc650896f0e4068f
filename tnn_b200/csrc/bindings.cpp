// Python bindings (pybind11 through torch/extension.h).  Every function takes torch tensors
// only as typed pointer carriers; all work happens in the hand-written kernels on the
// CURRENT torch CUDA stream, so they compose with torch streams / events / CUDA graphs.
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include <optional>
#include <vector>

#include "kernels/gemm_xt.cuh"
#include "kernels/simt_ops.cuh"
#include "kernels/umma_gemm.cuh"
namespace tnn {
void runtime_bind(pybind11::module_& m);   // csrc/runtime/runtime.cpp: archiver / wire format, allocators, data loaders, tokenizer, profiler
}

namespace py = pybind11;
using torch::Tensor;
using OptT = std::optional<Tensor>;

namespace tnn {
void attention_bind(py::module_& m);  // attention.cpp
void mxfp8_bind(py::module_& m);      // mxfp8.cpp
}

static int dt_code(const Tensor& t);
// element type of an im2col / col2im pair: bf16 (default), fp32 or fp16 -- both tensors must agree
static int conv_dt(const Tensor& a, const Tensor& b) {
  TORCH_CHECK(a.scalar_type() == b.scalar_type(), "im2col / col2im: source and destination dtypes differ");
  const int dt = dt_code(a);
  TORCH_CHECK(dt != tnn::XT_F64, "im2col / col2im: bf16 / fp32 / fp16");
  return dt;
}
static cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
static void* vp(const Tensor& t) { return t.data_ptr(); }
// element type code of the generic tensor ops / casts / extra-dtype GEMMs (tnn::XtDType)
static int dt_code(const Tensor& t) {
  TORCH_CHECK(t.is_cuda(), "CUDA tensor expected");
  switch (t.scalar_type()) {
    case torch::kBFloat16: return tnn::XT_BF16;
    case torch::kFloat32: return tnn::XT_F32;
    case torch::kFloat16: return tnn::XT_F16;
    case torch::kFloat64: return tnn::XT_F64;
    default: TORCH_CHECK(false, "tensor ops take bf16 / fp32 / fp16 / fp64 tensors"); return 0;
  }
}
static void* vp(const OptT& t) { return t.has_value() ? t->data_ptr() : nullptr; }
static float* fp(const Tensor& t) {
  TORCH_CHECK(t.scalar_type() == torch::kFloat32, "expected fp32 tensor");
  return t.data_ptr<float>();
}
static float* fp(const OptT& t) { return t.has_value() ? fp(*t) : nullptr; }
static void check_cuda(const Tensor& t) { TORCH_CHECK(t.is_cuda(), "expected a CUDA tensor"); }

static tnn::Epilogue make_epi(const Tensor& out, int out_mode, int act, double alpha, const OptT& bias,
                              const OptT& residual, int64_t ldc, const std::vector<int64_t>& grid) {
  tnn::Epilogue e;
  e.out = vp(out);
  e.out_mode = out_mode;
  e.act = act;
  e.alpha = (float)alpha;
  e.bias = fp(bias);
  e.residual = vp(residual);
  e.ldc = ldc;
  if (!grid.empty()) {
    // {OH, OW, sN, sH, sW, base}
    TORCH_CHECK(grid.size() == 6, "grid spec = {OH, OW, sN, sH, sW, base}");
    e.grid_mode = 1;
    e.OH = (int)grid[0];
    e.OW = (int)grid[1];
    e.sN = grid[2];
    e.sH = grid[3];
    e.sW = grid[4];
    e.base = grid[5];
  }
  return e;
}

static tnn::ConvGeom make_geom(const std::vector<int64_t>& g, const std::vector<int64_t>& taps) {
  TORCH_CHECK(g.size() == 7, "geom = {IN, IH, IW, C, ON, OH, OW}");
  TORCH_CHECK(taps.size() % 4 == 0 && taps.size() / 4 <= tnn::kMaxTaps, "taps = flat {dh,dw,nb,koff}*");
  tnn::ConvGeom cg{};
  cg.IN = (int)g[0];
  cg.IH = (int)g[1];
  cg.IW = (int)g[2];
  cg.C = (int)g[3];
  cg.ON = (int)g[4];
  cg.OH = (int)g[5];
  cg.OW = (int)g[6];
  cg.ntaps = (int)(taps.size() / 4);
  for (int t = 0; t < cg.ntaps; ++t)
    cg.taps[t] = tnn::ConvTap{(int)taps[4 * t], (int)taps[4 * t + 1], (int)taps[4 * t + 2], (int)taps[4 * t + 3]};
  return cg;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "tnn_b200 native kernels (sm_100a) and C++ runtime";

  // ------------------------------------------------------------------ tcgen05 GEMM / conv
  m.def("gemm", [](const Tensor& A, const Tensor& B, int64_t M, int64_t N, int64_t K, int64_t lda,
                   int64_t ldb, bool a_mn, bool b_mn, const Tensor& out, int64_t ldc, int out_mode,
                   int act, double alpha, const OptT& bias, const OptT& residual, int splits) {
    check_cuda(A);
    TORCH_CHECK(A.scalar_type() == torch::kBFloat16 && B.scalar_type() == torch::kBFloat16, "bf16 operands");
    tnn::Epilogue e = make_epi(out, out_mode, act, alpha, bias, residual, ldc, {});
    tnn::gemm_bf16(vp(A), vp(B), (int)M, (int)N, (int)K, lda, ldb, a_mn, b_mn, e, splits, cur_stream());
  });
  m.def("conv_fprop", [](const Tensor& X, const Tensor& Wt, int64_t ldw, int64_t N,
                         const std::vector<int64_t>& geom, const std::vector<int64_t>& taps,
                         const Tensor& out, int64_t ldc, const std::vector<int64_t>& grid, int out_mode,
                         int act, double alpha, const OptT& bias, const OptT& residual, const OptT& col_stats,
                         const OptT& bnb_x, const OptT& bnb_pack, bool bnb_relu, double bnb_drop, uint64_t bnb_seed,
                         const OptT& bnb_state) {
    check_cuda(X);
    tnn::ConvGeom g = make_geom(geom, taps);
    tnn::Epilogue e = make_epi(out, out_mode, act, alpha, bias, residual, ldc, grid);
    if (col_stats.has_value()) {
      TORCH_CHECK(col_stats->numel() >= (int64_t)tnn::bn_stat_replicas() * 2 * N, "col_stats workspace too small");
      e.col_sum = fp(*col_stats);
      e.col_repl = tnn::bn_stat_replicas();
    }
    if (bnb_x.has_value()) {
      // BatchNorm-backward fusion: this launch is the dgrad of the conv that consumes BN(x)
      TORCH_CHECK(col_stats.has_value() && bnb_pack.has_value(), "bnb_x needs col_stats and bnb_pack");
      TORCH_CHECK(bnb_x->scalar_type() == torch::kBFloat16 && bnb_x->is_contiguous() && bnb_x->numel() == out.numel(),
                  "bnb_x must be a contiguous bf16 tensor shaped like out");
      TORCH_CHECK(bnb_pack->scalar_type() == torch::kFloat32 && bnb_pack->numel() >= 4 * N &&
                      (reinterpret_cast<uintptr_t>(bnb_pack->data_ptr()) & 15) == 0, "bnb_pack: 16-byte aligned fp32 [N][4]");
      TORCH_CHECK(out_mode == tnn::OUT_BF16 && N % 8 == 0 && ldc == N, "BN-backward fusion needs a dense bf16 [M, N] output");
      e.bnb_x = vp(*bnb_x);
      e.bnb_pack = fp(*bnb_pack);
      e.bnb_relu = bnb_relu ? 1 : 0;
      e.bnb_drop = (float)bnb_drop;
      e.bnb_dscale = bnb_drop > 0.0 ? (float)(1.0 / (1.0 - bnb_drop)) : 1.f;
      e.bnb_seed = bnb_seed;
      e.bnb_seed_ptr = bnb_state.has_value()
          ? reinterpret_cast<const unsigned long long*>(bnb_state->data_ptr<float>() + 12) : nullptr;
    }
    tnn::conv_fprop_bf16(vp(X), vp(Wt), ldw, (int)N, g, e, cur_stream());
  }, py::arg("X"), py::arg("Wt"), py::arg("ldw"), py::arg("N"), py::arg("geom"), py::arg("taps"), py::arg("out"),
     py::arg("ldc"), py::arg("grid"), py::arg("out_mode"), py::arg("act"), py::arg("alpha"), py::arg("bias"),
     py::arg("residual"), py::arg("col_stats") = py::none(), py::arg("bnb_x") = py::none(),
     py::arg("bnb_pack") = py::none(), py::arg("bnb_relu") = false, py::arg("bnb_drop") = 0.0, py::arg("bnb_seed") = 0,
     py::arg("bnb_state") = py::none());
  m.def("bn_finalize_stats", [](const Tensor& ws, int64_t M, int C, const OptT& gamma, const OptT& beta, double eps,
                                double momentum, const OptT& rmean, const OptT& rvar, const Tensor& smean,
                                const Tensor& sinvstd, const Tensor& ss, const OptT& pack) {
    tnn::bn_finalize_stats(fp(ws), M, C, fp(gamma), fp(beta), (float)eps, (float)momentum, fp(rmean), fp(rvar),
                           fp(smean), fp(sinvstd), fp(ss), fp(pack), cur_stream());
  }, py::arg("ws"), py::arg("M"), py::arg("C"), py::arg("gamma"), py::arg("beta"), py::arg("eps"), py::arg("momentum"),
     py::arg("rmean"), py::arg("rvar"), py::arg("smean"), py::arg("sinvstd"), py::arg("ss"), py::arg("pack") = py::none());
  m.def("bn_bwd_coef", [](const Tensor& ws, int64_t M, int C, const OptT& gamma, const Tensor& sinvstd,
                          const OptT& dgamma, const OptT& dbeta, const Tensor& coef) {
    tnn::bn_bwd_coef(fp(ws), M, C, fp(gamma), fp(sinvstd), fp(dgamma), fp(dbeta), fp(coef), cur_stream());
  });
  m.def("bn_stat_replicas", &tnn::bn_stat_replicas);
  m.def("conv_wgrad", [](const Tensor& dY, int64_t ldy, int64_t Kout, const Tensor& X,
                         const std::vector<int64_t>& geom, const std::vector<int64_t>& taps,
                         const Tensor& dW, int64_t ldw, double alpha) {
    check_cuda(X);
    tnn::ConvGeom g = make_geom(geom, taps);
    tnn::conv_wgrad_bf16(vp(dY), ldy, (int)Kout, vp(X), g, fp(dW), ldw, (float)alpha, cur_stream());
  });
  m.def("conv_geom_supported", [](const std::vector<int64_t>& geom, const std::vector<int64_t>& taps) {
    return tnn::conv_geom_supported(make_geom(geom, taps));
  });

  // ------------------------------------------------------------------ elementwise
  m.def("cast_f32_to_bf16", [](const Tensor& x, const Tensor& y) {
    tnn::cast_f32_to_bf16(fp(x), vp(y), x.numel(), cur_stream());
  });
  m.def("cast_bf16_to_f32", [](const Tensor& x, const Tensor& y) {
    tnn::cast_bf16_to_f32(vp(x), fp(y), x.numel(), cur_stream());
  });
  m.def("add_bf16", [](const Tensor& a, const Tensor& b, const Tensor& o, bool relu) {
    tnn::add_bf16(vp(a), vp(b), vp(o), a.numel(), relu, cur_stream());
  });
  m.def("binary_bf16", [](const Tensor& a, const Tensor& b, const Tensor& o, int op) {
    tnn::binary_bf16(vp(a), vp(b), vp(o), a.numel(), op, cur_stream());
  });
  m.def("act_fwd", [](const Tensor& x, const Tensor& y, int kind, double slope) {
    tnn::act_fwd(vp(x), vp(y), x.numel(), kind, (float)slope, dt_code(x), cur_stream());
  });
  m.def("act_bwd", [](const Tensor& x, const Tensor& dy, const Tensor& dx, int kind, double slope) {
    tnn::act_bwd(vp(x), vp(dy), vp(dx), x.numel(), kind, (float)slope, dt_code(x), cur_stream());
  });
  m.def("dropout_fwd", [](const Tensor& x, const Tensor& y, double p, uint64_t seed, const OptT& state,
                          const OptT& residual) {
    // `state` = the fp32[16] device step-state tensor; its uint64 seed counter lives at floats 12..13
    const uint64_t* sp = state.has_value() ? reinterpret_cast<const uint64_t*>(state->data_ptr<float>() + 12) : nullptr;
    if (residual.has_value())
      TORCH_CHECK(residual->scalar_type() == torch::kBFloat16 && residual->is_contiguous() && residual->numel() == x.numel(),
                  "dropout_fwd: residual must be a contiguous bf16 tensor shaped like x");
    tnn::dropout_fwd(vp(x), vp(y), x.numel(), (float)p, seed, sp, vp(residual), cur_stream());
  }, py::arg("x"), py::arg("y"), py::arg("p"), py::arg("seed"), py::arg("state"), py::arg("residual") = py::none());
  m.def("fill_f32", [](const Tensor& x, double v) { tnn::fill_f32(fp(x), x.numel(), (float)v, cur_stream()); });
  m.def("axpy_f32", [](const Tensor& y, const Tensor& x, double a) {
    tnn::axpy_f32(fp(y), fp(x), (float)a, x.numel(), cur_stream());
  });
  m.def("scale_f32", [](const Tensor& x, double a) { tnn::scale_f32(fp(x), (float)a, x.numel(), cur_stream()); });
  m.def("colsum_bf16", [](const Tensor& x, int64_t M, int64_t N, int64_t ld, double alpha, const Tensor& out) {
    tnn::colsum_bf16(vp(x), M, (int)N, ld, (float)alpha, fp(out), cur_stream());
  });
  m.def("phase_split", [](const Tensor& x, const Tensor& xp, int N, int H, int W, int C) {
    tnn::phase_split(vp(x), vp(xp), N, H, W, C, cur_stream());
  });
  m.def("phase_merge", [](const Tensor& xp, const Tensor& x, int N, int H, int W, int C) {
    tnn::phase_merge(vp(xp), vp(x), N, H, W, C, cur_stream());
  });
  m.def("nchw_to_nhwc", [](const Tensor& x, const Tensor& y, int N, int C, int H, int W) {
    tnn::nchw_to_nhwc(vp(x), vp(y), N, C, H, W, x.scalar_type() == torch::kFloat32, cur_stream());
  });
  m.def("nhwc_to_nchw", [](const Tensor& x, const Tensor& y, int N, int C, int H, int W) {
    tnn::nhwc_to_nchw(vp(x), vp(y), N, C, H, W, x.scalar_type() == torch::kFloat32, cur_stream());
  });
  m.def("transpose_bf16", [](const Tensor& x, const Tensor& y, int rows, int cols) {
    tnn::transpose_bf16(vp(x), vp(y), rows, cols, cur_stream());
  });
  m.def("conv_weight_prep", [](const Tensor& w, const OptT& wb, const OptT& wt, int K, int R, int S, int C) {
    tnn::conv_weight_prep(fp(w), vp(wb), vp(wt), K, R, S, C, cur_stream());
  });
  m.def("philox_uniform", [](const Tensor& x, double lo, double hi, uint64_t seed) {
    tnn::philox_uniform_f32(fp(x), x.numel(), (float)lo, (float)hi, seed, cur_stream());
  });
  m.def("philox_normal", [](const Tensor& x, double mean, double std, uint64_t seed) {
    tnn::philox_normal_f32(fp(x), x.numel(), (float)mean, (float)std, seed, cur_stream());
  });

  // ------------------------------------------------------------------ normalisation
  m.def("bn_train_stats", [](const Tensor& x, int64_t M, int C, const OptT& gamma, const OptT& beta,
                             double eps, double momentum, const OptT& rmean, const OptT& rvar,
                             const Tensor& smean, const Tensor& sinvstd, const Tensor& ss, const Tensor& ws,
                             const OptT& pack) {
    tnn::bn_train_stats(vp(x), M, C, fp(gamma), fp(beta), (float)eps, (float)momentum, fp(rmean), fp(rvar),
                        fp(smean), fp(sinvstd), fp(ss), fp(ws), fp(pack), cur_stream());
  }, py::arg("x"), py::arg("M"), py::arg("C"), py::arg("gamma"), py::arg("beta"), py::arg("eps"), py::arg("momentum"),
     py::arg("rmean"), py::arg("rvar"), py::arg("smean"), py::arg("sinvstd"), py::arg("ss"), py::arg("ws"),
     py::arg("pack") = py::none());
  m.def("bn_eval_scale_shift", [](int C, const OptT& gamma, const OptT& beta, const Tensor& rmean,
                                  const Tensor& rvar, double eps, const Tensor& ss) {
    tnn::bn_eval_scale_shift(C, fp(gamma), fp(beta), fp(rmean), fp(rvar), (float)eps, fp(ss), cur_stream());
  });
  m.def("bn_apply", [](const Tensor& x, const Tensor& y, int64_t M, int C, const Tensor& ss, bool relu,
                       double drop_p, uint64_t seed, const OptT& state, const OptT& mask) {
    const uint64_t* sp = state.has_value() ? reinterpret_cast<const uint64_t*>(state->data_ptr<float>() + 12) : nullptr;
    uint8_t* mp = nullptr;
    if (mask.has_value()) {
      TORCH_CHECK(mask->scalar_type() == torch::kUInt8 && mask->is_contiguous() && C % 8 == 0 &&
                      mask->numel() >= M * (C / 8), "bn_apply: mask must be uint8 [M * C / 8] and C % 8 == 0");
      mp = mask->data_ptr<uint8_t>();
    }
    tnn::bn_apply(vp(x), vp(y), M, C, fp(ss), relu, (float)drop_p, seed, sp, mp, cur_stream());
  }, py::arg("x"), py::arg("y"), py::arg("M"), py::arg("C"), py::arg("ss"), py::arg("relu"), py::arg("drop_p"),
     py::arg("seed"), py::arg("state"), py::arg("mask") = py::none());
  m.def("bn_bwd_reduce", [](const Tensor& dy, const Tensor& x, const OptT& y, int64_t M, int C,
                            const OptT& gamma, const Tensor& smean, const Tensor& sinvstd, bool masked,
                            double drop_p, const OptT& dgamma, const OptT& dbeta, const Tensor& coef,
                            const Tensor& ws, const OptT& beta, bool relu, uint64_t seed, const OptT& state,
                            const OptT& bits) {
    const uint64_t* sp = state.has_value() ? reinterpret_cast<const uint64_t*>(state->data_ptr<float>() + 12) : nullptr;
    const uint8_t* bp = bits.has_value() ? bits->data_ptr<uint8_t>() : nullptr;
    tnn::bn_bwd_reduce(vp(dy), vp(x), vp(y), M, C, fp(gamma), fp(smean), fp(sinvstd), masked, (float)drop_p,
                       fp(dgamma), fp(dbeta), fp(coef), fp(ws), fp(beta), relu ? 1 : 0, seed, sp, bp, cur_stream());
  }, py::arg("dy"), py::arg("x"), py::arg("y"), py::arg("M"), py::arg("C"), py::arg("gamma"), py::arg("smean"),
     py::arg("sinvstd"), py::arg("masked"), py::arg("drop_p"), py::arg("dgamma"), py::arg("dbeta"), py::arg("coef"),
     py::arg("ws"), py::arg("beta") = py::none(), py::arg("relu") = false, py::arg("seed") = 0, py::arg("state") = py::none(),
     py::arg("bits") = py::none());
  m.def("bn_bwd_apply", [](const Tensor& dy, const Tensor& x, const OptT& y, const OptT& extra,
                           const Tensor& dx, int64_t M, int C, const Tensor& smean, const Tensor& sinvstd,
                           const Tensor& coef, bool masked, double drop_p, const OptT& colsum_out,
                           const OptT& colsum_ws, const OptT& gamma, const OptT& beta, bool relu, uint64_t seed,
                           const OptT& state, const OptT& bits) {
    const uint64_t* sp = state.has_value() ? reinterpret_cast<const uint64_t*>(state->data_ptr<float>() + 12) : nullptr;
    const uint8_t* bp = bits.has_value() ? bits->data_ptr<uint8_t>() : nullptr;
    tnn::bn_bwd_apply(vp(dy), vp(x), vp(y), vp(extra), vp(dx), M, C, fp(smean), fp(sinvstd), fp(coef), masked,
                      (float)drop_p, fp(colsum_out), fp(colsum_ws), fp(gamma), fp(beta), relu ? 1 : 0, seed, sp, bp, cur_stream());
  }, py::arg("dy"), py::arg("x"), py::arg("y"), py::arg("extra"), py::arg("dx"), py::arg("M"), py::arg("C"),
     py::arg("smean"), py::arg("sinvstd"), py::arg("coef"), py::arg("masked"), py::arg("drop_p"), py::arg("colsum_out"),
     py::arg("colsum_ws"), py::arg("gamma") = py::none(), py::arg("beta") = py::none(), py::arg("relu") = false,
     py::arg("seed") = 0, py::arg("state") = py::none(), py::arg("bits") = py::none());
  m.def("bn_ws_floats", &tnn::bn_ws_floats);
  m.def("layernorm_fwd", [](const Tensor& x, const OptT& res, const Tensor& y, const OptT& sum_out,
                            int64_t rows, int C, const OptT& gamma, const OptT& beta, double eps,
                            const OptT& mean, const OptT& rstd) {
    tnn::layernorm_fwd(vp(x), vp(res), vp(y), vp(sum_out), rows, C, fp(gamma), fp(beta), (float)eps, fp(mean),
                       fp(rstd), cur_stream());
  });
  m.def("layernorm_bwd", [](const Tensor& dy, const Tensor& x, int64_t rows, int C, const OptT& gamma,
                            const Tensor& mean, const Tensor& rstd, const Tensor& dx, const OptT& dgamma,
                            const OptT& dbeta, const OptT& extra) {
    tnn::layernorm_bwd(vp(dy), vp(x), rows, C, fp(gamma), fp(mean), fp(rstd), vp(dx), fp(dgamma), fp(dbeta),
                       vp(extra), cur_stream());
  }, py::arg("dy"), py::arg("x"), py::arg("rows"), py::arg("C"), py::arg("gamma"), py::arg("mean"), py::arg("rstd"),
     py::arg("dx"), py::arg("dgamma"), py::arg("dbeta"), py::arg("extra") = py::none());
  m.def("groupnorm_fwd", [](const Tensor& x, const Tensor& y, int N, int HW, int C, int G, const OptT& gamma,
                            const OptT& beta, double eps, const Tensor& mean, const Tensor& rstd) {
    tnn::groupnorm_fwd(vp(x), vp(y), N, HW, C, G, fp(gamma), fp(beta), (float)eps, fp(mean), fp(rstd),
                       cur_stream());
  });
  m.def("groupnorm_bwd", [](const Tensor& dy, const Tensor& x, int N, int HW, int C, int G, const OptT& gamma,
                            const Tensor& mean, const Tensor& rstd, const Tensor& dx, const OptT& dgamma,
                            const OptT& dbeta) {
    tnn::groupnorm_bwd(vp(dy), vp(x), N, HW, C, G, fp(gamma), fp(mean), fp(rstd), vp(dx), fp(dgamma),
                       fp(dbeta), cur_stream());
  });

  // ------------------------------------------------------------------ pooling
  m.def("maxpool_fwd", [](const Tensor& x, const Tensor& y, const OptT& mask, std::vector<int> d) {
    TORCH_CHECK(d.size() == 12, "dims = {N,H,W,C,kh,kw,sh,sw,ph,pw,OH,OW}");
    tnn::maxpool_fwd(vp(x), vp(y), (uint8_t*)vp(mask), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8],
                     d[9], d[10], d[11], cur_stream());
  });
  m.def("maxpool_bwd", [](const Tensor& dy, const Tensor& mask, const Tensor& dx, std::vector<int> d) {
    tnn::maxpool_bwd(vp(dy), (const uint8_t*)vp(mask), vp(dx), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7],
                     d[8], d[9], d[10], d[11], cur_stream());
  });
  m.def("avgpool_fwd", [](const Tensor& x, const Tensor& y, std::vector<int> d) {
    tnn::avgpool_fwd(vp(x), vp(y), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9], d[10], d[11],
                     cur_stream());
  });
  m.def("avgpool_bwd", [](const Tensor& dy, const Tensor& dx, std::vector<int> d) {
    tnn::avgpool_bwd(vp(dy), vp(dx), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9], d[10], d[11],
                     cur_stream());
  });

  // ------------------------------------------------------------------ loss
  m.def("softmax_ce", [](const Tensor& logits, int64_t ld, const Tensor& labels, int64_t rows, int V,
                         double loss_scale, double grad_scale, const OptT& grad, const Tensor& out) {
    TORCH_CHECK(labels.scalar_type() == torch::kInt32, "labels must be int32");
    tnn::softmax_ce(vp(logits), logits.scalar_type() == torch::kFloat32, ld, labels.data_ptr<int>(), rows, V,
                    (float)loss_scale, (float)grad_scale, vp(grad), fp(out), cur_stream());
  });
  m.def("regression_loss", [](const Tensor& pred, const Tensor& target, int kind, double delta,
                              double grad_scale, const OptT& grad, const Tensor& out) {
    tnn::regression_loss(vp(pred), vp(target), pred.scalar_type() == torch::kFloat32, pred.numel(), kind,
                         (float)delta, (float)grad_scale, vp(grad), fp(out), cur_stream());
  });

  // ------------------------------------------------------------------ optimizers
  m.def("adam_step_flat", [](const Tensor& p, const Tensor& g, const Tensor& mm, const Tensor& v,
                             const OptT& pb, double lr, double b1, double b2, double eps, double wd,
                             bool decoupled, int64_t step, double grad_scale, bool zero_grad) {
    tnn::AdamHyper h{(float)lr, (float)b1, (float)b2, (float)eps, (float)wd, decoupled ? 1 : 0,
                     (float)(1.0 - std::pow(b1, (double)step)), (float)(1.0 - std::pow(b2, (double)step)),
                     (float)grad_scale};
    tnn::adam_step_flat(fp(p), fp(g), fp(mm), fp(v), vp(pb), p.numel(), h, zero_grad, cur_stream());
  });
  m.def("sgd_step_flat", [](const Tensor& p, const Tensor& g, const OptT& mom, const OptT& pb, double lr,
                            double momentum, double wd, double grad_scale, bool zero_grad) {
    tnn::sgd_step_flat(fp(p), fp(g), fp(mom), vp(pb), p.numel(), (float)lr, (float)momentum, (float)wd,
                       (float)grad_scale, zero_grad, cur_stream());
  });
  m.def("adam_step_flat_dev", [](const Tensor& p, const Tensor& g, const Tensor& mm, const Tensor& v,
                                 const OptT& pb, const Tensor& state, bool zero_grad) {
    tnn::adam_step_flat_dev(fp(p), fp(g), fp(mm), fp(v), vp(pb), p.numel(), fp(state), zero_grad, cur_stream());
  });
  m.def("sgd_step_flat_dev", [](const Tensor& p, const Tensor& g, const OptT& mom, const OptT& pb,
                                const Tensor& state, bool zero_grad) {
    tnn::sgd_step_flat_dev(fp(p), fp(g), fp(mom), vp(pb), p.numel(), fp(state), zero_grad, cur_stream());
  });
  m.def("dp_set_watchdog_ms", [](uint64_t ms) { tnn::dp_set_watchdog_ms(ms); });
  m.def("dp_fused_step", [](int world, int rank, int64_t n, int64_t chunk, std::vector<int64_t> g_peers,
                            std::vector<int64_t> p_peers, std::vector<int64_t> pb_peers,
                            std::vector<int64_t> flag_peers, int64_t g_mc, int64_t p_mc, int64_t pb_mc,
                            const Tensor& mm, const OptT& v, const Tensor& sync, const Tensor& state,
                            bool use_adam, int ctas, int64_t lo0, int64_t hi0, int64_t st_off) {
    TORCH_CHECK((int)g_peers.size() == world && (int)p_peers.size() == world &&
                (int)pb_peers.size() == world && (int)flag_peers.size() == world, "peer lists");
    std::vector<float*> g(world), p(world);
    std::vector<void*> pb(world);
    std::vector<uint32_t*> fl(world);
    for (int r = 0; r < world; ++r) {
      g[r] = reinterpret_cast<float*>(g_peers[r]);
      p[r] = reinterpret_cast<float*>(p_peers[r]);
      pb[r] = reinterpret_cast<void*>(pb_peers[r]);
      fl[r] = reinterpret_cast<uint32_t*>(flag_peers[r]);
    }
    tnn::dp_fused_step(world, rank, n, chunk, g.data(), p.data(), pb.data(), fl.data(),
                       reinterpret_cast<const float*>(g_mc), reinterpret_cast<float*>(p_mc),
                       reinterpret_cast<void*>(pb_mc), fp(mm), fp(v), reinterpret_cast<uint32_t*>(sync.data_ptr()),
                       fp(state), use_adam ? 1 : 0, ctas, lo0, hi0, st_off, cur_stream());
  }, py::arg("world"), py::arg("rank"), py::arg("n"), py::arg("chunk"), py::arg("g_peers"), py::arg("p_peers"),
     py::arg("pb_peers"), py::arg("flag_peers"), py::arg("g_mc"), py::arg("p_mc"), py::arg("pb_mc"), py::arg("m"),
     py::arg("v"), py::arg("sync"), py::arg("state"), py::arg("use_adam"), py::arg("ctas"), py::arg("lo0") = 0,
     py::arg("hi0") = -1, py::arg("st_off") = 0);
  m.def("p2p_barrier", [](const Tensor& flag_ptrs, int world, int rank, uint32_t epoch) {
    tnn::p2p_barrier(reinterpret_cast<uint32_t* const*>(flag_ptrs.data_ptr()), world, rank, epoch, cur_stream());
  });

  // ------------------------------------------------------------------ conv helpers
  m.def("conv_direct_fwd", [](const Tensor& x, const Tensor& w, const OptT& bias, const Tensor& y,
                              std::vector<int> d) {
    TORCH_CHECK(d.size() == 11, "dims = {N,H,W,C,K,R,S,stride,pad,OH,OW}");
    tnn::conv_direct_fwd(vp(x), fp(w), fp(bias), vp(y), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8],
                         d[9], d[10], cur_stream());
  });
  m.def("conv_direct_wgrad", [](const Tensor& x, const Tensor& dy, const Tensor& dw, const OptT& dbias,
                                std::vector<int> d) {
    tnn::conv_direct_wgrad(vp(x), vp(dy), fp(dw), fp(dbias), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7],
                           d[8], d[9], d[10], cur_stream());
  });
  m.def("im2col_nhwc", [](const Tensor& x, const Tensor& col, int64_t ldcol, std::vector<int> d) {
    TORCH_CHECK(d.size() == 10 || d.size() == 12, "dims = {N,H,W,C,R,S,stride,pad,OH,OW} or {N,H,W,C,R,S,sh,sw,ph,pw,OH,OW}");
    if (d.size() == 12)
      tnn::im2col_nhwc2(vp(x), vp(col), ldcol, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9], d[10], d[11], cur_stream(),
                        conv_dt(x, col));
    else
      tnn::im2col_nhwc(vp(x), vp(col), ldcol, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9], cur_stream());
  });
  m.def("col2im_nhwc", [](const Tensor& col, int64_t ldcol, const Tensor& dx, std::vector<int> d) {
    TORCH_CHECK(d.size() == 10 || d.size() == 12, "dims = {N,H,W,C,R,S,stride,pad,OH,OW} or {N,H,W,C,R,S,sh,sw,ph,pw,OH,OW}");
    if (d.size() == 12)
      tnn::col2im_nhwc2(vp(col), ldcol, vp(dx), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9], d[10], d[11], cur_stream(),
                        conv_dt(col, dx));
    else
      tnn::col2im_nhwc(vp(col), ldcol, vp(dx), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9], cur_stream());
  });

  // ------------------------------------------------------------------ embedding
  m.def("embedding_fwd", [](const Tensor& ids, const Tensor& table, const OptT& pos, const Tensor& y,
                            int64_t T, int L, int D, int V) {
    tnn::embedding_fwd(ids.data_ptr<int>(), fp(table), fp(pos), vp(y), T, L, D, V, cur_stream());
  });
  m.def("embedding_bwd", [](const Tensor& ids, const Tensor& dy, const OptT& dtable, const OptT& dpos,
                            int64_t T, int L, int D, int V) {
    tnn::embedding_bwd(ids.data_ptr<int>(), vp(dy), fp(dtable), fp(dpos), T, L, D, V, cur_stream());
  });
  m.def("argmax_rows", [](const Tensor& x, int64_t ld, int64_t rows, int V, const Tensor& out) {
    tnn::argmax_rows(vp(x), x.scalar_type() == torch::kFloat32, ld, rows, V, out.data_ptr<int>(), cur_stream());
  });

  // ---- generic tensor op family (tensor_ops.cu)
  auto is32 = [](const Tensor& t) { return dt_code(t); };   // XtDType code: 0 bf16, 1 fp32, 2 fp16, 3 fp64
  m.def("tensor_binary", [is32](const Tensor& a, const OptT& b, double scalar, const Tensor& out, int op) {
    tnn::tensor_binary(vp(a), vp(b), (float)scalar, vp(out), a.numel(), op, is32(a), cur_stream());
  });
  m.def("tensor_unary", [is32](const Tensor& a, const Tensor& out, int op, double lo, double hi) {
    tnn::tensor_unary(vp(a), vp(out), a.numel(), op, (float)lo, (float)hi, is32(a), cur_stream());
  });
  m.def("tensor_ternary", [is32](const Tensor& a, const Tensor& b, const Tensor& c, const Tensor& out, int op) {
    tnn::tensor_ternary(vp(a), vp(b), vp(c), vp(out), a.numel(), op, is32(a), cur_stream());
  });
  m.def("tensor_reduce", [is32](const Tensor& a, const OptT& b, double mu, int mode, const Tensor& out) {
    tnn::tensor_reduce(vp(a), vp(b), (float)mu, a.numel(), mode, is32(a), fp(out), cur_stream());
  });
  m.def("tensor_window_copy", [is32](const Tensor& src, const Tensor& dst, int N, int SH, int SW, int DH, int DW, int C,
                                     int off_h, int off_w, double fill) {
    tnn::tensor_window_copy(vp(src), vp(dst), N, SH, SW, DH, DW, C, off_h, off_w, (float)fill, is32(src), cur_stream());
  });
  m.def("softmax_rows", [is32](const Tensor& x, const Tensor& y, int64_t rows, int C) {
    tnn::softmax_rows(vp(x), vp(y), rows, C, is32(x), cur_stream());
  });
  m.def("softmax_rows_bwd", [is32](const Tensor& y, const Tensor& dy, const Tensor& dx, int64_t rows, int C) {
    tnn::softmax_rows_bwd(vp(y), vp(dy), vp(dx), rows, C, is32(y), cur_stream());
  });
  m.def("tensor_bswap", [](const Tensor& t) {
    tnn::tensor_bswap(t.data_ptr(), t.numel(), (int)t.element_size(), cur_stream());
  });
  m.def("skinny_gemm", [](const Tensor& x, const Tensor& w, const OptT& bias, const OptT& residual, const Tensor& out,
                          int act, bool static_weights) {
    TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kBFloat16 && w.scalar_type() == torch::kBFloat16 &&
                    out.scalar_type() == torch::kBFloat16, "bf16 CUDA tensors");
    TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.size(1) == w.size(1) && x.stride(1) == 1 && w.stride(1) == 1 &&
                    out.stride(1) == 1, "x [M,K], w [N,K] row-major");
    tnn::skinny_gemm_bf16(vp(x), x.stride(0), vp(w), w.stride(0), fp(bias), vp(residual),
                          residual.has_value() ? residual->stride(0) : 0, vp(out), out.stride(0), (int)x.size(0),
                          (int)w.size(0), (int)x.size(1), act, static_weights ? 1 : 0, cur_stream());
  }, py::arg("x"), py::arg("w"), py::arg("bias"), py::arg("residual"), py::arg("out"), py::arg("act"),
     py::arg("static_weights") = false);
  // ---- GEMMs / casts for fp16, fp32 (tf32 tensor cores) and fp64 (gemm_xt.cu).  a: [M,K] or (a_mn) [K,M]; b: [N,K] or (b_mn) [K,N]
  m.def("gemm_xt", [](const Tensor& a, const Tensor& b, bool a_mn, bool b_mn, const Tensor& out, const OptT& bias, const OptT& residual,
                      int act, double alpha, bool accumulate) {
    TORCH_CHECK(a.is_cuda() && a.dim() == 2 && b.dim() == 2 && out.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1 && out.stride(1) == 1,
                "gemm_xt: row-major 2-D CUDA tensors");
    TORCH_CHECK(a.scalar_type() == b.scalar_type(), "gemm_xt: operand dtypes differ");
    const int M = (int)(a_mn ? a.size(1) : a.size(0)), K = (int)(a_mn ? a.size(0) : a.size(1));
    const int N = (int)(b_mn ? b.size(1) : b.size(0)), Kb = (int)(b_mn ? b.size(0) : b.size(1));
    TORCH_CHECK(K == Kb && out.size(0) == M && out.size(1) == N, "gemm_xt: shape mismatch");
    const int dt = dt_code(a);
    if (dt == tnn::XT_F64) {
      TORCH_CHECK(out.scalar_type() == torch::kFloat64 && !residual.has_value() && act == 0, "fp64 GEMM: fp64 out, no residual / activation");
      TORCH_CHECK(!bias.has_value() || bias->scalar_type() == torch::kFloat64, "fp64 GEMM: fp64 bias");
      tnn::gemm_f64(a.data_ptr<double>(), b.data_ptr<double>(), M, N, K, a.stride(0), b.stride(0), a_mn, b_mn, alpha,
                    bias.has_value() ? bias->data_ptr<double>() : nullptr, out.data_ptr<double>(), out.stride(0), accumulate, cur_stream());
      return;
    }
    TORCH_CHECK(dt == tnn::XT_F16 || dt == tnn::XT_F32, "gemm_xt: fp16 / fp32 / fp64 operands (bf16 goes through gemm_bf16)");
    tnn::XtEpilogue e;
    e.out = vp(out);
    e.out_f32 = out.scalar_type() == torch::kFloat32;
    TORCH_CHECK(e.out_f32 || (dt == tnn::XT_F16 && out.scalar_type() == torch::kFloat16), "gemm_xt: out is the operand type or fp32");
    e.ldc = out.stride(0);
    e.bias = fp(bias);
    if (residual.has_value()) {
      TORCH_CHECK(residual->scalar_type() == a.scalar_type() && residual->stride(1) == 1, "gemm_xt: residual has the operand type");
      e.residual = vp(residual);
      e.ldr = residual->stride(0);
    }
    e.act = act;
    e.alpha = (float)alpha;
    e.accumulate = accumulate ? 1 : 0;
    if (dt == tnn::XT_F16) tnn::gemm_f16(vp(a), vp(b), M, N, K, a.stride(0), b.stride(0), a_mn, b_mn, e, cur_stream());
    else tnn::gemm_tf32(vp(a), vp(b), M, N, K, a.stride(0), b.stride(0), a_mn, b_mn, e, cur_stream());
  }, py::arg("a"), py::arg("b"), py::arg("a_mn"), py::arg("b_mn"), py::arg("out"), py::arg("bias") = py::none(),
     py::arg("residual") = py::none(), py::arg("act") = 0, py::arg("alpha") = 1.0, py::arg("accumulate") = false);
  // ---- native NCHW BatchNorm (norm_nchw.cu)
  m.def("bn_nchw_ws_bytes", [](int C) { return (int64_t)tnn::bn_nchw_ws_bytes(C); });
  m.def("bn_nchw_train_stats", [](const Tensor& x, const OptT& gamma, const OptT& beta, double eps, double momentum, const OptT& rmean,
                                  const OptT& rvar, const Tensor& mean, const Tensor& invstd, const Tensor& ss, const Tensor& ws) {
    TORCH_CHECK(x.is_cuda() && x.dim() == 4 && x.is_contiguous(), "bn_nchw: contiguous [N,C,H,W] CUDA tensor");
    TORCH_CHECK(x.scalar_type() == torch::kFloat32 || x.scalar_type() == torch::kBFloat16, "bn_nchw: bf16 / fp32 activations");
    const int N = (int)x.size(0), C = (int)x.size(1), HW = (int)(x.size(2) * x.size(3));
    tnn::bn_nchw_train_stats(vp(x), x.scalar_type() == torch::kFloat32, N, C, HW, fp(gamma), fp(beta), (float)eps, (float)momentum, fp(rmean),
                             fp(rvar), fp(mean), fp(invstd), fp(ss), ws.data_ptr(), cur_stream());
  });
  m.def("bn_nchw_eval_scale_shift", [](int C, const OptT& gamma, const OptT& beta, const Tensor& rmean, const Tensor& rvar, double eps,
                                       const Tensor& ss) {
    tnn::bn_nchw_eval_scale_shift(C, fp(gamma), fp(beta), fp(rmean), fp(rvar), (float)eps, fp(ss), cur_stream());
  });
  m.def("bn_nchw_apply", [](const Tensor& x, const Tensor& y, const Tensor& ss, bool relu) {
    TORCH_CHECK(x.is_cuda() && x.dim() == 4 && x.is_contiguous() && y.is_contiguous() && x.scalar_type() == y.scalar_type(), "bn_nchw_apply");
    tnn::bn_nchw_apply(vp(x), vp(y), x.scalar_type() == torch::kFloat32, (int)x.size(0), (int)x.size(1), (int)(x.size(2) * x.size(3)), fp(ss),
                       relu ? 1 : 0, cur_stream());
  });
  m.def("bn_nchw_backward", [](const Tensor& dy, const Tensor& x, const Tensor& dx, const OptT& gamma, const Tensor& mean, const Tensor& invstd,
                               const Tensor& ss, bool relu, const OptT& dgamma, const OptT& dbeta, const Tensor& coef, const Tensor& ws) {
    TORCH_CHECK(x.is_cuda() && x.dim() == 4 && x.is_contiguous() && dy.is_contiguous() && dx.is_contiguous() &&
                    dy.scalar_type() == x.scalar_type() && dx.scalar_type() == x.scalar_type(), "bn_nchw_backward");
    tnn::bn_nchw_backward(vp(dy), vp(x), vp(dx), x.scalar_type() == torch::kFloat32, (int)x.size(0), (int)x.size(1),
                          (int)(x.size(2) * x.size(3)), fp(gamma), fp(mean), fp(invstd), fp(ss), relu ? 1 : 0, fp(dgamma), fp(dbeta), fp(coef),
                          ws.data_ptr(), cur_stream());
  });
  m.def("layernorm_fwd_any", [](const Tensor& x, const OptT& residual, const Tensor& y, const OptT& sum_out, int64_t rows, int C,
                                const OptT& gamma, const OptT& beta, double eps, const Tensor& mean, const Tensor& rstd) {
    TORCH_CHECK(x.is_contiguous() && y.is_contiguous() && y.scalar_type() == x.scalar_type(), "layernorm_fwd_any: contiguous tensors of one dtype");
    tnn::layernorm_fwd_any(vp(x), vp(residual), vp(y), vp(sum_out), dt_code(x), rows, C, fp(gamma), fp(beta), (float)eps, fp(mean), fp(rstd),
                           cur_stream());
  });
  m.def("layernorm_bwd_any", [](const Tensor& dy, const Tensor& x, int64_t rows, int C, const OptT& gamma, const Tensor& mean, const Tensor& rstd,
                                const Tensor& dx, const OptT& dgamma, const OptT& dbeta, const OptT& extra) {
    TORCH_CHECK(dy.is_contiguous() && x.is_contiguous() && dx.is_contiguous() && dy.scalar_type() == x.scalar_type() &&
                    dx.scalar_type() == x.scalar_type(), "layernorm_bwd_any: contiguous tensors of one dtype");
    tnn::layernorm_bwd_any(vp(dy), vp(x), dt_code(x), rows, C, fp(gamma), fp(mean), fp(rstd), vp(dx), fp(dgamma), fp(dbeta), vp(extra),
                           cur_stream());
  });
  m.def("colsum_any", [](const Tensor& x, const Tensor& out, double alpha, bool accumulate) {
    TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && out.numel() == x.size(1) && out.is_contiguous(), "colsum_any: x [R,C], out [C]");
    const int dt = dt_code(x);
    TORCH_CHECK(out.scalar_type() == (dt == tnn::XT_F64 ? torch::kFloat64 : torch::kFloat32), "colsum_any: out is fp32 (fp64 for fp64 input)");
    tnn::colsum_any(vp(x), dt, x.size(0), (int)x.size(1), x.stride(0), alpha, vp(out), accumulate, cur_stream());
  });
  m.def("cast_any", [](const Tensor& src, const Tensor& dst) {
    TORCH_CHECK(src.numel() == dst.numel() && src.is_contiguous() && dst.is_contiguous(), "cast_any: contiguous tensors of one size");
    tnn::tensor_cast_any(vp(src), dt_code(src), vp(dst), dt_code(dst), src.numel(), cur_stream());
  });
  m.def("transpose_any", [](const Tensor& src, const Tensor& dst) {
    TORCH_CHECK(src.dim() == 2 && dst.dim() == 2 && src.is_contiguous() && dst.is_contiguous() && src.size(0) == dst.size(1) &&
                    src.size(1) == dst.size(0) && src.scalar_type() == dst.scalar_type(), "transpose_any: [R,C] -> [C,R]");
    tnn::tensor_transpose_any(vp(src), vp(dst), dt_code(src), (int)src.size(0), (int)src.size(1), cur_stream());
  });
  m.def("launch_count", []() { return (unsigned long long)tnn::g_launch_count.load(); });
  tnn::attention_bind(m);
  tnn::mxfp8_bind(m);
  tnn::runtime_bind(m);
}
