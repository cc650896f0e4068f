"""Dense GEMM family on the tcgen05 kernel (csrc/kernels/umma_gemm.cu).

Reference call sites: cuBLAS ``GemmEx`` (src/math/cuda/gemm.cu:74) and the cuDNN-frontend
matmul graphs (src/math/cuda/cudnn_gemm.cu:108,170,228) behind DenseLayer
(src/nn/layers_impl/dense_layer.cpp:267-333).  Here fwd / dgrad / wgrad are the same kernel
with different operand major-ness, bias/activation/residual fused into the epilogue and
wgrad accumulating (beta = 1) through fp32 atomics.
"""
from __future__ import annotations

import torch

from . import native

import os

_SKINNY = os.environ.get("TNN_SKINNY_GEMM", "1") == "1"
# Set (through ``static_weights()``) while an inference engine runs: the decode GEMMs may then be
# launched as programmatic dependents that stream their weights before the previous kernel has finished.
_STATIC_WEIGHTS = False


class static_weights:
    """Context manager: weights are not written by any kernel in flight (inference)."""

    def __enter__(self):
        global _STATIC_WEIGHTS
        self.prev, _STATIC_WEIGHTS = _STATIC_WEIGHTS, True
        return self

    def __exit__(self, *exc):
        global _STATIC_WEIGHTS
        _STATIC_WEIGHTS = self.prev
        return False

ACT = {"none": 0, "linear": 0, None: 0, "relu": 1, "gelu": 2}
OUT_BF16, OUT_F32, OUT_F32_ATOMIC = 0, 1, 2


def _act_ref(y: torch.Tensor, act) -> torch.Tensor:
    if act in (None, "none", "linear"):
        return y
    if act == "relu":
        return torch.relu(y)
    if act == "gelu":
        return torch.nn.functional.gelu(y, approximate="tanh")
    raise ValueError(act)


# ------------------------------------------------------------------------------------------- fp16 / fp32 / fp64
# The reference's DenseLayer is instantiated per IO dtype (dense_layer.cpp:336-340: fp32, fp16, bf16, fp64) and
# picks cublasHgemm / Sgemm / Dgemm / GemmEx (src/math/cuda/cublas_gemm.cu:20-164).  Here: fp16 -> tcgen05 kind::f16
# with fp16 operands, fp32 -> tcgen05 kind::tf32 on the untouched fp32 tensors, fp64 -> register-tiled DFMA kernel
# (csrc/kernels/gemm_xt.cu).  bf16 (the default GPU activation type) keeps the tuned kernels below.
XT_DTYPES = (torch.float16, torch.float32, torch.float64)


def _xt_rows_ok(t: torch.Tensor) -> bool:
    """TMA operands: unit inner stride, 16-byte aligned base and row pitch."""
    return t.stride(1) == 1 and (t.stride(0) * t.element_size()) % 16 == 0 and t.data_ptr() % 16 == 0


def _xt_operand(t: torch.Tensor) -> torch.Tensor:
    """A 2-D view the TMA descriptors accept (copy into 16-byte padded rows only when the pitch is odd)."""
    if t.dtype == torch.float64 or _xt_rows_ok(t):
        return t if t.stride(1) == 1 else t.contiguous()
    per16 = 16 // t.element_size()
    ld = (t.shape[1] + per16 - 1) // per16 * per16
    buf = torch.zeros((t.shape[0], ld), dtype=t.dtype, device=t.device)
    buf[:, :t.shape[1]].copy_(t)
    return buf[:, :t.shape[1]]


def gemm_xt(a: torch.Tensor, b: torch.Tensor, a_mn: bool = False, b_mn: bool = False, *, out: torch.Tensor | None = None,
            bias: torch.Tensor | None = None, residual: torch.Tensor | None = None, act=None, alpha: float = 1.0,
            accumulate: bool = False, out_f32: bool = False) -> torch.Tensor:
    """out[M,N] (+)= act(alpha * A B^T-ish + bias + residual) for fp16 / fp32 (tf32 tensor cores) / fp64 CUDA tensors.
    a is [M,K] (or [K,M] with a_mn), b is [N,K] (or [K,N] with b_mn)."""
    C = native.require()
    a, b = _xt_operand(a), _xt_operand(b)
    M = a.shape[1] if a_mn else a.shape[0]
    N = b.shape[1] if b_mn else b.shape[0]
    if out is None:
        odt = torch.float32 if (out_f32 and a.dtype == torch.float16) else a.dtype
        out = torch.empty((M, N), dtype=odt, device=a.device)
    if bias is not None:
        want = torch.float64 if a.dtype == torch.float64 else torch.float32
        bias = bias if bias.dtype == want else bias.to(want)
    C.gemm_xt(a, b, a_mn, b_mn, out, bias, residual, ACT[act], float(alpha), accumulate)
    return out


def _xt(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype in XT_DTYPES and native.use_native(x)


def _linear_fwd_xt(x, w, bias, act, residual, out, out_f32):
    assert w.dtype == x.dtype, f"weight dtype {w.dtype} != activation dtype {x.dtype}"
    if x.dtype == torch.float64 and (act not in (None, "none", "linear") or residual is not None):
        y = gemm_xt(x, w, bias=bias)
        if residual is not None:
            y = y + residual
        y = _act_ref(y, act)
        if out is not None:
            out.copy_(y)
            return out
        return y
    return gemm_xt(x, w, out=out, bias=bias, residual=residual, act=act, out_f32=out_f32)


def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, act=None,
               residual: torch.Tensor | None = None, out: torch.Tensor | None = None,
               out_f32: bool = False) -> torch.Tensor:
    """y[M,N] = act(x[M,K] @ w[N,K]^T + bias (+ residual)).  GPU: x, w bf16; bias fp32."""
    M, K = x.shape
    N = w.shape[0]
    if _xt(x):
        return _linear_fwd_xt(x, w, bias, act, residual, out, out_f32)
    if not native.use_native(x):
        y = x.float() @ w.float().t()
        if bias is not None:
            y = y + bias.float()
        if residual is not None:
            y = y + residual.float()
        y = _act_ref(y, act)
        return y.to(torch.float32 if out_f32 else x.dtype)
    C = native.require()
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert x.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    if M <= 32 and out.dtype == torch.bfloat16 and _SKINNY and x.stride(0) % 8 == 0 and w.stride(0) % 8 == 0 and \
            (residual is None or (residual.dtype == torch.bfloat16 and residual.stride(-1) == 1)):
        # decode-sized batch: weight-streaming kernel (HBM bound) instead of a 3/4-empty 128-row tile
        C.skinny_gemm(x, w, bias, residual, out, ACT[act], _STATIC_WEIGHTS)
        return out
    C.gemm(x, w, M, N, K, x.stride(0), w.stride(0), False, False, out, out.stride(0),
           OUT_F32 if out.dtype == torch.float32 else OUT_BF16, ACT[act], 1.0, bias, residual, 1)
    return out


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """dx[M,K] = dy[M,N] @ w[N,K]  (w consumed MN-major, no transpose copy)."""
    M, N = dy.shape
    K = w.shape[1]
    if _xt(dy):
        return gemm_xt(dy, w, False, True, out=out)          # w [N,K] read MN-major: no transpose copy
    if not native.use_native(dy):
        return (dy.float() @ w.float()).to(dy.dtype)
    C = native.require()
    if out is None:
        out = torch.empty((M, K), device=dy.device, dtype=torch.bfloat16)
    C.gemm(dy, w, M, K, N, dy.stride(0), w.stride(0), False, True, out, out.stride(0), OUT_BF16, 0, 1.0,
           None, None, 1)
    return out


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, alpha: float = 1.0) -> None:
    """dw[N,K] += alpha * dy[M,N]^T @ x[M,K]  (dw fp32; both operands MN-major; split-K)."""
    M, N = dy.shape
    K = x.shape[1]
    if _xt(dy):
        if dy.dtype == torch.float64 and dw.dtype != torch.float64:
            # fp64 activations over fp32 master gradients: exact fp64 product, one rounding into the fp32 accumulator
            t = gemm_xt(dy, x, True, True, alpha=alpha)
            t32 = torch.empty(t.shape, dtype=dw.dtype, device=t.device)
            native.require().cast_any(t, t32)
            native.require().tensor_binary(dw, t32, 0.0, dw, 0)
            return
        gemm_xt(dy, x, True, True, out=dw, alpha=alpha, accumulate=True)   # dw fp32 (fp64 for fp64 masters), both operands MN-major
        return
    if not native.use_native(dy):
        dw.add_(alpha * (dy.float().t() @ x.float()))
        return
    C = native.require()
    tiles = ((N + 127) // 128) * ((K + 127) // 128)
    # split-K so that tiles * splits FITS the 2 x 148 resident persistent CTAs (floor, not ceil: 108 tiles
    # x 3 splits = 324 work items run as two rounds of K/3 each, 108 x 2 = 216 as one round of K/2)
    splits = max(1, min((M + 63) // 64, (2 * 148) // tiles))
    C.gemm(dy, x, N, K, M, dy.stride(0), x.stride(0), True, True, dw, dw.stride(0), OUT_F32_ATOMIC, 0,
           float(alpha), None, None, splits)


def bias_grad(dy: torch.Tensor, db: torch.Tensor, alpha: float = 1.0) -> None:
    """db[N] += alpha * column sums of dy[M,N]."""
    if _xt(dy):
        if dy.dtype == torch.float64 and db.dtype != torch.float64:
            t = torch.empty(dy.shape[1], dtype=torch.float64, device=dy.device)
            native.require().colsum_any(dy, t, float(alpha), False)
            t32 = torch.empty(t.shape, dtype=db.dtype, device=t.device)
            native.require().cast_any(t, t32)
            native.require().tensor_binary(db, t32, 0.0, db, 0)
            return
        native.require().colsum_any(dy, db, float(alpha), True)
        return
    if not native.use_native(dy):
        db.add_(alpha * dy.float().sum(0))
        return
    assert dy.stride(1) == 1
    native.require().colsum_bf16(dy, dy.shape[0], dy.shape[1], dy.stride(0), float(alpha), db)


def matmul_nt(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """a[M,K] @ b[N,K]^T * alpha (bf16 out) -- utility for attention fallbacks/tests."""
    if not native.use_native(a):
        return (alpha * (a.float() @ b.float().t())).to(a.dtype)
    C = native.require()
    out = torch.empty((a.shape[0], b.shape[0]), device=a.device, dtype=torch.bfloat16)
    C.gemm(a, b, a.shape[0], b.shape[0], a.shape[1], a.stride(0), b.stride(0), False, False, out,
           out.stride(0), OUT_BF16, 0, float(alpha), None, None, 1)
    return out
