"""Block-scaled fp8 (MXFP8) operators: e4m3 elements + one ue8m0 scale per 32 elements along K.

Device side: ``csrc/kernels/mxfp8.cu`` (tcgen05 ``kind::mxf8f6f4.block_scale`` GEMM, scale factors through
``tcgen05.cp`` into TMEM).  The reference has no fp8 path (its GPT-2 runs through cuBLAS GemmEx,
/root/reference/src/math/cuda/gemm.cu:74); BASELINE config 4 asks for one.

``sf`` tensors use the tcgen05 scale-factor atom layout (see mxfp8.cuh); ``sf_index`` gives the flat
byte offset of (row, k-group) so tests and tools can decode it without the kernels.
"""
from __future__ import annotations

import torch

from . import native

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


def sf_bytes(rows: int, K: int) -> int:
    return ((rows + 127) // 128 + 1) * (K // 128) * 512


def sf_index(rows: torch.Tensor, groups: torch.Tensor, K: int) -> torch.Tensor:
    """Flat offset of the scale of (row, 32-element group) -- mirrors sf_offset() in mxfp8.cu."""
    kb = K // 128
    r = rows % 128
    return ((rows // 128) * kb + groups // 4) * 512 + (r % 32) * 16 + (r // 32) * 4 + (groups % 4)


def quantize(x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """x [rows, K] (bf16 or fp32, unit inner stride) -> (q uint8 [rows, K], sf uint8 [sf_bytes])."""
    assert x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 128 == 0, "x: [rows, K], K % 128 == 0"
    C = native.require()
    q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    sf = torch.empty(sf_bytes(*x.shape), dtype=torch.uint8, device=x.device)
    C.mxfp8_quantize(x, q, sf)
    return q, sf


def dequantize(q: torch.Tensor, sf: torch.Tensor) -> torch.Tensor:
    out = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    native.require().mxfp8_dequantize(q, sf, out)
    return out


def dequantize_torch(q: torch.Tensor, sf: torch.Tensor) -> torch.Tensor:
    """Pure-torch decoder (oracle): e4m3 bytes through torch.float8_e4m3fn, scales gathered by sf_index."""
    rows, K = q.shape
    vals = q.view(torch.float8_e4m3fn).float()
    r = torch.arange(rows, device=q.device).view(-1, 1).expand(rows, K // 32)
    g = torch.arange(K // 32, device=q.device).view(1, -1).expand(rows, K // 32)
    e = sf[sf_index(r, g, K)].to(torch.int32) - 127
    scale = torch.exp2(e.float())
    return vals * scale.repeat_interleave(32, dim=1)


def gemm(a_q, a_sf, b_q, b_sf, out=None, bias=None, residual=None, act=ACT_NONE, alpha=1.0, out_q=None, out_sf=None,
         want_bf16=True, want_fp8=False):
    """act(alpha * A @ B.T + bias + residual).  Returns (out_bf16 | None, out_q | None, out_sf | None)."""
    M, N = a_q.shape[0], b_q.shape[0]
    dev = a_q.device
    if out is None and want_bf16:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    if want_fp8 and out_q is None:
        out_q = torch.empty((M, N), dtype=torch.uint8, device=dev)
        out_sf = torch.empty(sf_bytes(M, N), dtype=torch.uint8, device=dev)
    native.require().gemm_mxfp8(a_q, a_sf, b_q, b_sf, out, bias, residual, act, alpha, out_q, out_sf)
    return out, out_q, out_sf


def layernorm_quant(x, gamma, beta, eps, residual=None, sum_out=None, q=None, sf=None, y_bf16=None):
    """LayerNorm(x [+ residual]) written directly as MXFP8.  x: bf16 [rows, C]."""
    rows, Cdim = x.shape[0], x.shape[-1]
    if q is None:
        q = torch.empty((rows, Cdim), dtype=torch.uint8, device=x.device)
        sf = torch.empty(sf_bytes(rows, Cdim), dtype=torch.uint8, device=x.device)
    native.require().layernorm_mxfp8(x, residual, sum_out, rows, Cdim, gamma, beta, eps, q, sf, y_bf16)
    return q, sf
