"""Fused multi-head attention.

Input is the packed projection ``qkv`` [B, L, 3, H, D] straight out of the fused QKV GEMM (no
permute_heads kernel: the attention kernel addresses heads through TMA strides; reference:
src/nn/blocks_impl/cuda/permute_heads.cu, causal_mask.cu, softmax.cu + 2 batched cuBLAS GEMMs,
or cuDNN SDPA in cudnn_flash_attention_ops.cu:123,209).

GPU: tcgen05 flash-attention forward (csrc/kernels/attention.cu): S = QK^T and O += PV on the
tensor cores with S / PV tiles in TMEM, one softmax thread per query row, online softmax,
causal tile skipping; decode (L_q = 1) uses a split-KV SIMT kernel over the KV cache.
Backward: tcgen05 flash backward (attention_bwd.cu) recomputing P tile by tile from the saved
log-sum-exp; dQ accumulates in fp32 through vector reductions.
CPU: fp32 reference.
"""
from __future__ import annotations

import math
import os

import torch

from . import native


_KERNEL_DTYPES = (torch.bfloat16, torch.float16)      # element types of the fused tcgen05 attention kernels


class KVCache:
    """Per-layer key/value cache [B, Lmax, H, D] (bf16) for incremental decoding."""

    def __init__(self, B, Lmax, H, D, device, dtype):
        self.k = torch.zeros((B, Lmax, H, D), device=device, dtype=dtype)
        self.v = torch.zeros((B, Lmax, H, D), device=device, dtype=dtype)
        self.len = 0


def attention_ref(q, k, v, causal: bool, q_offset: int = 0):
    """q [B, Lq, H, D], k/v [B, Lk, H, D] -> o [B, Lq, H, D], lse [B, H, Lq] (fp32 math)."""
    D = q.shape[-1]
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) / math.sqrt(D)
    if causal:
        Lq, Lk = s.shape[-2], s.shape[-1]
        i = torch.arange(Lq, device=s.device)[:, None] + q_offset
        j = torch.arange(Lk, device=s.device)[None, :]
        s = s.masked_fill(j > i, float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.exp(s - lse[..., None])
    o = (p @ vf).permute(0, 2, 1, 3)
    return o, lse


def attention_fwd(qkv: torch.Tensor, causal: bool, kv_cache: KVCache | None = None, pos: int = 0):
    """Returns (o [B, L, H*D], lse [B, H, L])."""
    B, L, _, H, D = qkv.shape
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    if kv_cache is not None:
        Cx = native.require() if qkv.is_cuda else None
        if Cx is not None and hasattr(Cx, "kv_cache_store") and D == 64 and qkv.dtype == torch.bfloat16 and \
                qkv.is_contiguous() and kv_cache.k.dtype == torch.bfloat16:
            Cx.kv_cache_store(qkv, kv_cache.k, kv_cache.v, int(pos))      # one vectorised kernel for K and V
        else:
            kv_cache.k[:, pos:pos + L] = k
            kv_cache.v[:, pos:pos + L] = v
        kv_cache.len = pos + L
        k, v = kv_cache.k[:, :pos + L], kv_cache.v[:, :pos + L]
    if qkv.is_cuda:
        Cx = native.require()
        if hasattr(Cx, "attention_fwd") and D == 64 and qkv.dtype in _KERNEL_DTYPES and \
                (qkv.dtype == torch.bfloat16 or not (L == 1 and kv_cache is not None)):
            # bf16, or fp16 (the reference's flash attention is fp16-only, cudnn_flash_attention_ops.cu:123): the same
            # tcgen05 kernel with the fp16 operand format; the decode / KV-cache kernels are bf16
            o = torch.empty((B, L, H * D), device=qkv.device, dtype=qkv.dtype)
            lse = torch.empty((B, H, L), device=qkv.device, dtype=torch.float32)
            if L == 1 and kv_cache is not None:
                Cx.attention_decode(q.contiguous(), kv_cache.k, kv_cache.v, o, pos + 1)
            else:
                Cx.attention_fwd(q, k, v, o, lse, bool(causal), int(pos))
            return o, lse
    o, lse = attention_ref(q, k, v, causal, q_offset=pos)
    return o.reshape(B, L, H * D).to(qkv.dtype), lse


def attention_bwd(qkv: torch.Tensor, do: torch.Tensor, o: torch.Tensor, lse: torch.Tensor, causal: bool):
    """dqkv [B, L, 3, H, D] from do [B, L, H*D]; P is recomputed from the saved lse."""
    B, L, _, H, D = qkv.shape
    if qkv.is_cuda and D == 64 and qkv.dtype in _KERNEL_DTYPES and os.environ.get("TNN_ATTN_BWD_KERNEL", "1") != "0":
        # tcgen05 flash backward (csrc/kernels/attention_bwd.cu): five MMAs per (key tile, query tile)
        Cx = native.require()
        dqkv = torch.empty_like(qkv, memory_format=torch.contiguous_format)
        delta = torch.empty((B, H, L), device=qkv.device, dtype=torch.float32)
        dq_acc = torch.empty((B, L, H * D), device=qkv.device, dtype=torch.float32)
        Cx.attention_bwd(qkv.contiguous(), do.reshape(B, L, H * D).to(qkv.dtype).contiguous(),
                         o.reshape(B, L, H * D).contiguous(), lse.contiguous(), dqkv, delta, dq_acc, bool(causal))
        return dqkv
    q, k, v = (qkv[:, :, i].float().permute(0, 2, 1, 3) for i in range(3))   # [B,H,L,D]
    dof = do.float().reshape(B, L, H, D).permute(0, 2, 1, 3)
    of = o.float().reshape(B, L, H, D).permute(0, 2, 1, 3)
    scale = 1.0 / math.sqrt(D)
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        i = torch.arange(L, device=s.device)[:, None]
        j = torch.arange(L, device=s.device)[None, :]
        s = s.masked_fill(j > i, float("-inf"))
    p = torch.exp(s - lse.float()[..., None])
    dv = p.transpose(-1, -2) @ dof
    dp = dof @ v.transpose(-1, -2)
    delta = (dof * of).sum(-1, keepdim=True)
    ds = p * (dp - delta) * scale
    dq = ds @ k
    dk = ds.transpose(-1, -2) @ q
    dqkv = torch.stack([dq, dk, dv], 0).permute(1, 3, 0, 2, 4)   # [B, L, 3, H, D]
    return dqkv.contiguous().to(qkv.dtype)
