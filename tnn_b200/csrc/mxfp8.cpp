// Python bindings of the block-scaled fp8 family (csrc/kernels/mxfp8.cu).  Tensors are pointer carriers only.
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include <optional>

#include "kernels/mxfp8.cuh"

namespace py = pybind11;
using torch::Tensor;
using OptT = std::optional<Tensor>;

namespace tnn {

static cudaStream_t cur() { return at::cuda::getCurrentCUDAStream().stream(); }
static void* vp(const OptT& t) { return t.has_value() ? t->data_ptr() : nullptr; }
static const float* fpo(const OptT& t) {
  if (!t.has_value()) return nullptr;
  TORCH_CHECK(t->scalar_type() == torch::kFloat32, "expected fp32");
  return t->data_ptr<float>();
}

void mxfp8_bind(py::module_& m) {
  m.def("mxfp8_sf_bytes", [](int64_t rows, int64_t K) { return (int64_t)mxfp8_sf_bytes(rows, K); });
  // x [rows, K] bf16 or fp32 (row stride in elements) -> q uint8 [rows, K], sf uint8 [mxfp8_sf_bytes]
  m.def("mxfp8_quantize", [](const Tensor& x, const Tensor& q, const Tensor& sf) {
    TORCH_CHECK(x.is_cuda() && x.dim() == 2 && x.stride(1) == 1 && q.dim() == 2 && q.stride(1) == 1 &&
                    q.scalar_type() == torch::kUInt8 && sf.scalar_type() == torch::kUInt8, "x [rows,K], q uint8 [rows,K]");
    const int64_t rows = x.size(0), K = x.size(1);
    TORCH_CHECK(q.size(0) == rows && q.size(1) == K && sf.numel() >= mxfp8_sf_bytes(rows, K), "q / sf too small");
    if (x.scalar_type() == torch::kBFloat16)
      mxfp8_quantize_bf16(x.data_ptr(), x.stride(0), rows, (int)K, q.data_ptr(), q.stride(0), sf.data_ptr(), cur());
    else if (x.scalar_type() == torch::kFloat32)
      mxfp8_quantize_f32(x.data_ptr<float>(), x.stride(0), rows, (int)K, q.data_ptr(), q.stride(0), sf.data_ptr(), cur());
    else
      TORCH_CHECK(false, "mxfp8_quantize: bf16 or fp32 source");
  });
  m.def("mxfp8_dequantize", [](const Tensor& q, const Tensor& sf, const Tensor& out) {
    TORCH_CHECK(q.is_cuda() && q.dim() == 2 && q.stride(1) == 1 && out.scalar_type() == torch::kFloat32 && out.is_contiguous());
    mxfp8_dequantize(q.data_ptr(), q.stride(0), sf.data_ptr(), q.size(0), (int)q.size(1), out.data_ptr<float>(), cur());
  });
  m.def("layernorm_mxfp8", [](const Tensor& x, const OptT& residual, const OptT& sum_out, int64_t rows, int C, const Tensor& gamma,
                              const Tensor& beta, double eps, const Tensor& q, const Tensor& sf, const OptT& y_bf16) {
    TORCH_CHECK(x.is_cuda() && x.scalar_type() == torch::kBFloat16 && x.is_contiguous(), "x: contiguous bf16");
    TORCH_CHECK(q.scalar_type() == torch::kUInt8 && q.is_contiguous() && q.numel() >= rows * C && sf.numel() >= mxfp8_sf_bytes(rows, C));
    layernorm_fwd_mxfp8(x.data_ptr(), vp(residual), vp(sum_out), rows, C, gamma.data_ptr<float>(), beta.data_ptr<float>(),
                        (float)eps, q.data_ptr(), sf.data_ptr(), vp(y_bf16), cur());
  }, py::arg("x"), py::arg("residual"), py::arg("sum_out"), py::arg("rows"), py::arg("C"), py::arg("gamma"), py::arg("beta"),
     py::arg("eps"), py::arg("q"), py::arg("sf"), py::arg("y_bf16") = py::none());
  // out = act(alpha * A B^T + bias + residual);  A [M,K] / B [N,K] uint8 e4m3 with scales
  m.def("gemm_mxfp8", [](const Tensor& A, const Tensor& sfa, const Tensor& B, const Tensor& sfb, const OptT& out,
                         const OptT& bias, const OptT& residual, int act, double alpha, const OptT& out_q, const OptT& out_sf) {
    TORCH_CHECK(A.is_cuda() && A.scalar_type() == torch::kUInt8 && B.scalar_type() == torch::kUInt8 && A.dim() == 2 &&
                    B.dim() == 2 && A.stride(1) == 1 && B.stride(1) == 1 && A.size(1) == B.size(1), "A [M,K], B [N,K] uint8");
    const int64_t M = A.size(0), N = B.size(0), K = A.size(1);
    TORCH_CHECK(sfa.numel() >= mxfp8_sf_bytes(M, K) && sfb.numel() >= mxfp8_sf_bytes(N, K), "scale buffers too small");
    Fp8Epilogue e;
    if (out.has_value()) {
      TORCH_CHECK(out->scalar_type() == torch::kBFloat16 && out->dim() == 2 && out->stride(1) == 1 && out->size(0) == M &&
                      out->size(1) >= N, "out bf16 [M, >=N]");
      e.out = out->data_ptr();
      e.ldc = out->stride(0);
    }
    if (residual.has_value()) {
      TORCH_CHECK(out.has_value() && residual->scalar_type() == torch::kBFloat16 && residual->stride(0) == out->stride(0),
                  "residual must share out's layout");
      e.residual = residual->data_ptr();
    }
    e.bias = fpo(bias);
    e.act = act;
    e.alpha = (float)alpha;
    if (out_q.has_value()) {
      TORCH_CHECK(out_sf.has_value() && out_q->scalar_type() == torch::kUInt8 && out_q->dim() == 2 && out_q->stride(1) == 1 &&
                      out_sf->numel() >= mxfp8_sf_bytes(M, N), "out_q uint8 [M,N] + out_sf");
      e.out_q = out_q->data_ptr();
      e.ldq = out_q->stride(0);
      e.out_sf = out_sf->data_ptr();
    }
    TORCH_CHECK(e.out != nullptr || e.out_q != nullptr, "gemm_mxfp8: no output");
    gemm_mxfp8(A.data_ptr(), A.stride(0), sfa.data_ptr(), B.data_ptr(), B.stride(0), sfb.data_ptr(), (int)M, (int)N, (int)K, e, cur());
  }, py::arg("A"), py::arg("sfa"), py::arg("B"), py::arg("sfb"), py::arg("out"), py::arg("bias") = py::none(),
     py::arg("residual") = py::none(), py::arg("act") = 0, py::arg("alpha") = 1.0, py::arg("out_q") = py::none(),
     py::arg("out_sf") = py::none());
}

}  // namespace tnn
