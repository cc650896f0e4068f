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


def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, act=None,
               residual: torch.Tensor | None = None, out: torch.Tensor | None = None,
               out_f32: bool = False) -> torch.Tensor:
    """y[M,N] = act(x[M,K] @ w[N,K]^T + bias (+ residual)).  GPU: x, w bf16; bias fp32."""
    M, K = x.shape
    N = w.shape[0]
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
