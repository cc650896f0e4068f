#!/usr/bin/env python
"""GEMM / Dense micro-benchmark (reference: benchmarks/gemm_benchmark.cpp, dense_benchmark.cpp):
the hand-written tcgen05 kernel vs cuBLAS (torch.matmul) at the Dense shapes of the model zoo
(SURVEY 2.7), fwd / dgrad / wgrad, TFLOP/s and fraction of the measured cuBLAS peak."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit, peaks, time_cuda  # noqa: E402

import torch  # noqa: E402

from tnn_b200.ops import gemm  # noqa: E402

SHAPES = [  # (name, M, N, K)
    ("gpt2_qkv", 32768, 2304, 768), ("gpt2_proj", 32768, 768, 768), ("gpt2_fc1", 32768, 3072, 768),
    ("gpt2_fc2", 32768, 768, 3072), ("gpt2_head", 32768, 50264, 768), ("square_8k", 8192, 8192, 8192),
    ("square_4k", 4096, 4096, 4096), ("wrn_fc", 256, 104, 512), ("decode_qkv", 32, 2304, 768),
]

if __name__ == "__main__":
    peak, _, how = peaks()
    dev = torch.device("cuda", 0)
    only = sys.argv[1:] or None
    for name, M, N, K in SHAPES:
        if only and name not in only:
            continue
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        dy = torch.randn(M, N, device=dev).bfloat16()
        dw = torch.zeros(N, K, device=dev)
        flops = 2.0 * M * N * K
        big = M * N * 2 > (1 << 28)
        it = 5 if big else 20
        t_f = time_cuda(lambda: gemm.linear_fwd(x, w), it)
        t_ref = time_cuda(lambda: torch.matmul(x, w.t()), it)
        row = dict(bench="gemm", shape=name, M=M, N=N, K=K, fwd_ms=t_f, fwd_tflops=flops / t_f / 1e9,
                   cublas_ms=t_ref, cublas_tflops=flops / t_ref / 1e9, frac_of_peak=flops / t_f / 1e9 / peak, peak=how)
        if not big:
            t_d = time_cuda(lambda: gemm.linear_dgrad(dy, w), it)
            t_w = time_cuda(lambda: gemm.linear_wgrad(dy, x, dw), it)
            row.update(dgrad_ms=t_d, dgrad_tflops=flops / t_d / 1e9, wgrad_ms=t_w, wgrad_tflops=flops / t_w / 1e9)
        emit(**row)
