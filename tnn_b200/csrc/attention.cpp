// Python bindings of the fused attention kernels.
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include "kernels/attention.cuh"

namespace py = pybind11;
using torch::Tensor;

namespace tnn {

static void check_head_view(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && (t.scalar_type() == torch::kBFloat16 || t.scalar_type() == torch::kFloat16), name,
              ": expected a CUDA bf16 / fp16 tensor");
  TORCH_CHECK(t.dim() == 4 && t.size(3) == 64 && t.stride(3) == 1 && t.stride(2) == 64, name,
              ": expected a [B, L, H, 64] view with contiguous heads");
}

void attention_bind(py::module_& m) {
  m.def("attention_fwd", [](const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& out,
                            const std::optional<Tensor>& lse, bool causal, int64_t q_offset) {
    check_head_view(q, "q");
    check_head_view(k, "k");
    check_head_view(v, "v");
    const int B = (int)q.size(0), Lq = (int)q.size(1), H = (int)q.size(2), Lk = (int)k.size(1);
    TORCH_CHECK(k.size(0) == B && v.size(0) == B && k.size(2) == H && v.size(2) == H && v.size(1) == Lk, "q/k/v shape mismatch");
    const bool f16 = q.scalar_type() == torch::kFloat16;       // fp16 q / k / v / out: the kind::f16 fp16 operand format
    TORCH_CHECK(k.scalar_type() == q.scalar_type() && v.scalar_type() == q.scalar_type(), "q / k / v dtypes differ");
    TORCH_CHECK(out.is_cuda() && out.scalar_type() == q.scalar_type() && out.is_contiguous() &&
                    out.numel() == (int64_t)B * Lq * H * 64, "out must be contiguous [B, Lq, H*64] of the q dtype");
    float* lp = nullptr;
    if (lse.has_value()) {
      TORCH_CHECK(lse->scalar_type() == torch::kFloat32 && lse->is_contiguous() && lse->numel() == (int64_t)B * H * Lq,
                  "lse must be contiguous fp32 [B, H, Lq]");
      lp = lse->data_ptr<float>();
    }
    attention_fwd_bf16(q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0), v.data_ptr(),
                       v.stride(1), v.stride(0), out.data_ptr(), lp, B, H, Lq, Lk, causal ? 1 : 0, (int)q_offset,
                       at::cuda::getCurrentCUDAStream().stream(), f16 ? 1 : 0);
  });
  m.def("attention_bwd", [](const Tensor& qkv, const Tensor& dout, const Tensor& out, const Tensor& lse,
                            const Tensor& dqkv, const Tensor& delta, const Tensor& dq_acc, bool causal) {
    const bool f16 = qkv.scalar_type() == torch::kFloat16;
    const auto dt16 = qkv.scalar_type();
    TORCH_CHECK(qkv.is_cuda() && (f16 || dt16 == torch::kBFloat16) && qkv.dim() == 5 && qkv.size(2) == 3 &&
                    qkv.size(4) == 64 && qkv.is_contiguous(), "qkv must be contiguous bf16 / fp16 [B, L, 3, H, 64]");
    TORCH_CHECK(dqkv.is_cuda() && dqkv.scalar_type() == dt16 && dqkv.is_contiguous() &&
                    dqkv.sizes() == qkv.sizes(), "dqkv must match qkv");
    const int B = (int)qkv.size(0), L = (int)qkv.size(1), H = (int)qkv.size(3);
    const int64_t E = (int64_t)H * 64, n = (int64_t)B * L * E;
    TORCH_CHECK(dout.scalar_type() == dt16 && dout.is_contiguous() && dout.numel() == n, "dout: contiguous [B, L, H*64] of the qkv dtype");
    TORCH_CHECK(out.scalar_type() == dt16 && out.is_contiguous() && out.numel() == n, "out: contiguous [B, L, H*64] of the qkv dtype");
    TORCH_CHECK(lse.scalar_type() == torch::kFloat32 && lse.is_contiguous() && lse.numel() == (int64_t)B * H * L, "lse: fp32 [B, H, L]");
    TORCH_CHECK(delta.scalar_type() == torch::kFloat32 && delta.is_contiguous() && delta.numel() == (int64_t)B * H * L, "delta: fp32 [B, H, L]");
    TORCH_CHECK(dq_acc.scalar_type() == torch::kFloat32 && dq_acc.is_contiguous() && dq_acc.numel() == n, "dq_acc: fp32 [B, L, H*64]");
    auto* base = reinterpret_cast<at::BFloat16*>(qkv.data_ptr());
    auto* dbase = reinterpret_cast<at::BFloat16*>(dqkv.data_ptr());
    attention_bwd_bf16(base, base + E, base + 2 * E, 3 * E, (int64_t)L * 3 * E, dout.data_ptr(), out.data_ptr(),
                       lse.data_ptr<float>(), delta.data_ptr<float>(), dq_acc.data_ptr<float>(), dbase, dbase + E,
                       dbase + 2 * E, 3 * E, (int64_t)L * 3 * E, B, H, L, causal ? 1 : 0,
                       at::cuda::getCurrentCUDAStream().stream(), f16 ? 1 : 0);
  });
  m.def("attention_decode", [](const Tensor& q, const Tensor& kc, const Tensor& vc, const Tensor& out, int64_t len) {
    TORCH_CHECK(q.is_cuda() && q.scalar_type() == torch::kBFloat16, "q: CUDA bf16");
    check_head_view(kc, "k cache");
    check_head_view(vc, "v cache");
    const int B = (int)kc.size(0), H = (int)kc.size(2);
    TORCH_CHECK(q.numel() == (int64_t)B * H * 64, "q must hold one token per batch row");
    TORCH_CHECK(kc.stride(0) == vc.stride(0) && kc.stride(1) == vc.stride(1), "k/v cache strides differ");
    const int64_t q_batch = q.dim() >= 2 ? q.stride(0) : (int64_t)H * 64;
    attention_decode_bf16(q.data_ptr(), q_batch, kc.data_ptr(), vc.data_ptr(), kc.stride(1), kc.stride(0), out.data_ptr(),
                          B, H, (int)len, at::cuda::getCurrentCUDAStream().stream());
  });
  m.def("kv_cache_store", [](const Tensor& qkv, const Tensor& kc, const Tensor& vc, int64_t pos) {
    check_head_view(kc, "k cache");
    check_head_view(vc, "v cache");
    TORCH_CHECK(qkv.is_cuda() && qkv.scalar_type() == torch::kBFloat16 && qkv.dim() == 5 && qkv.size(2) == 3 &&
                    qkv.size(4) == 64 && qkv.is_contiguous(), "qkv must be contiguous bf16 [B, L, 3, H, 64]");
    const int B = (int)qkv.size(0), L = (int)qkv.size(1), H = (int)qkv.size(3);
    TORCH_CHECK(kc.size(0) >= B && kc.size(2) == H && pos >= 0 && pos + L <= kc.size(1), "KV cache too small");
    TORCH_CHECK(kc.stride(0) == vc.stride(0) && kc.stride(1) == vc.stride(1) && kc.stride(1) % 8 == 0 && kc.stride(0) % 8 == 0,
                "k/v cache strides differ or are not 16-byte multiples");
    kv_cache_store_bf16(qkv.data_ptr(), kc.data_ptr(), vc.data_ptr(), kc.stride(1), kc.stride(0), B, L, H * 64, (int)pos,
                        at::cuda::getCurrentCUDAStream().stream());
  });
  m.def("attention_decode_append", [](const Tensor& qkv, const Tensor& kc, const Tensor& vc, const Tensor& out,
                                      const Tensor& pos) {
    check_head_view(kc, "k cache");
    check_head_view(vc, "v cache");
    const int B = (int)kc.size(0), H = (int)kc.size(2);
    TORCH_CHECK(qkv.is_cuda() && qkv.scalar_type() == torch::kBFloat16 && qkv.is_contiguous() &&
                    qkv.numel() == (int64_t)B * 3 * H * 64, "qkv must be contiguous bf16 [B, 3*H*64]");
    TORCH_CHECK(pos.is_cuda() && pos.scalar_type() == torch::kInt32, "pos: CUDA int32 scalar");
    attention_decode_append_bf16(qkv.data_ptr(), kc.data_ptr(), vc.data_ptr(), kc.stride(1), kc.stride(0), out.data_ptr(), B, H,
                                 pos.data_ptr<int>(), at::cuda::getCurrentCUDAStream().stream());
  });
  m.def("embed_decode", [](const Tensor& tok, const Tensor& wte, const Tensor& wpe, const Tensor& x, const Tensor& pos) {
    TORCH_CHECK(tok.scalar_type() == torch::kInt32 && pos.scalar_type() == torch::kInt32, "int32 ids / pos");
    TORCH_CHECK(wte.scalar_type() == torch::kFloat32 && wpe.scalar_type() == torch::kFloat32, "fp32 tables");
    embed_decode(tok.data_ptr<int>(), wte.data_ptr<float>(), wpe.data_ptr<float>(), x.data_ptr(), (int)tok.numel(),
                 (int)wte.size(1), (int)wte.size(0), pos.data_ptr<int>(), at::cuda::getCurrentCUDAStream().stream());
  });
  m.def("decode_epilogue", [](const Tensor& next, const Tensor& tok, const Tensor& history, const Tensor& pos,
                              const Tensor& step) {
    const int B = (int)tok.numel();
    decode_epilogue(next.data_ptr<int>(), tok.data_ptr<int>(), history.data_ptr<int>(), B, (int)(history.numel() / B),
                    pos.data_ptr<int>(), step.data_ptr<int>(), at::cuda::getCurrentCUDAStream().stream());
  });
}

}  // namespace tnn
