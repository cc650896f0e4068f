#!/usr/bin/env python
"""Attention micro-benchmark (reference: benchmarks/attention_benchmark.cpp, vanilla vs flash at
B=16 L=512 E=768): the tcgen05 flash-attention forward vs torch SDPA, plus the GPT-2 inference
shape (B=32, L=1024, 12 heads) and the decode kernel over a full KV cache."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit, peaks, time_cuda  # noqa: E402

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from tnn_b200.ops import attention as A  # noqa: E402

if __name__ == "__main__":
    peak, hbm, how = peaks()
    dev = torch.device("cuda", 0)
    for name, B, L, H, causal in [("ref_bench", 16, 512, 12, True), ("gpt2_prefill", 32, 1024, 12, True),
                                  ("vit", 64, 257, 4, False), ("long", 4, 4096, 12, True)]:
        qkv = (torch.randn(B, L, 3, H, 64, device=dev) * 0.5).bfloat16()
        flops = 4.0 * B * H * L * L * 64 * (0.5 if causal else 1.0)
        t = time_cuda(lambda: A.attention_fwd(qkv, causal), 20)
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
        t_ref = time_cuda(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=causal), 20)
        emit(bench="attention_fwd", shape=name, B=B, L=L, H=H, causal=causal, ms=t, tflops=flops / t / 1e9, sdpa_ms=t_ref,
             sdpa_tflops=flops / t_ref / 1e9, frac_of_peak=flops / t / 1e9 / peak, peak=how)
        # backward: tcgen05 flash backward vs the op-by-op recompute it replaces vs SDPA autograd
        o, lse = A.attention_fwd(qkv, causal)
        do = torch.randn_like(o)
        tb = time_cuda(lambda: A.attention_bwd(qkv, do, o, lse, causal), 10)
        row = dict(bench="attention_bwd", shape=name, B=B, L=L, H=H, causal=causal, ms=tb, tflops=2.5 * flops / tb / 1e9,
                   frac_of_peak=2.5 * flops / tb / 1e9 / peak)
        if B * H * L * L * 4 * 6 < (40 << 30):
            os.environ["TNN_ATTN_BWD_KERNEL"] = "0"
            row["fallback_ms"] = time_cuda(lambda: A.attention_bwd(qkv, do, o, lse, causal), 3)
            os.environ["TNN_ATTN_BWD_KERNEL"] = "1"
        qg, kg, vg = (t_.detach().clone().requires_grad_(True) for t_ in (q, k, v))
        og = F.scaled_dot_product_attention(qg, kg, vg, is_causal=causal)
        dog = torch.randn_like(og)
        row["sdpa_bwd_ms"] = time_cuda(lambda: torch.autograd.grad(og, (qg, kg, vg), dog, retain_graph=True), 10)
        emit(**row)
    B, H, Lmax = 32, 12, 1024
    cache = A.KVCache(B, Lmax, H, 64, dev, torch.bfloat16)
    cache.k.normal_()
    cache.v.normal_()
    step = torch.randn(B, 1, 3, H, 64, device=dev).bfloat16()
    t = time_cuda(lambda: A.attention_fwd(step, True, kv_cache=cache, pos=Lmax - 1), 20)
    nbytes = 2 * B * Lmax * H * 64 * 2
    emit(bench="attention_decode", B=B, L=Lmax, H=H, ms=t, gbs=nbytes / t / 1e6, frac_of_hbm=nbytes / t / 1e6 / hbm, peak=how)
