#!/usr/bin/env python
"""fp16 / fp32 (tf32 tensor cores) / fp64 GEMM micro-benchmark (csrc/kernels/gemm_xt.cu) next to the library routine of the
same dtype (torch.matmul -> cuBLAS Hgemm-class / TF32 Sgemm / Dgemm): forward, dgrad (MN-major B) and wgrad (both operands
MN-major, fp32 accumulate) of a Dense layer.  These are the plain library-slot GEMMs of the non-bf16 element types, not the
tuned bf16 / MXFP8 kernels; the numbers say how far the simple one-tile-per-CTA kernel is from cuBLAS."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit, time_cuda  # noqa: E402

import torch  # noqa: E402

from tnn_b200.ops import gemm  # noqa: E402

SHAPES = [("gpt2_fc1", 8192, 3072, 768), ("gpt2_fc2", 8192, 768, 3072), ("square_4k", 4096, 4096, 4096), ("wrn_fc", 256, 104, 512)]

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True            # the library arm of the fp32 row uses TF32 tensor cores too
    for dtype, tag in ((torch.float16, "fp16"), (torch.float32, "tf32"), (torch.float64, "fp64")):
        for name, M, N, K in SHAPES:
            if dtype == torch.float64 and M * N * K > 8192 * 3072 * 768:
                continue
            x = torch.randn(M, K, device=dev).to(dtype)
            w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dtype)
            dy = torch.randn(M, N, device=dev).to(dtype)
            dw = torch.zeros(N, K, device=dev, dtype=torch.float64 if dtype == torch.float64 else torch.float32)
            flops = 2.0 * M * N * K
            it = 5 if dtype == torch.float64 else 20
            t_f = time_cuda(lambda: gemm.linear_fwd(x, w), it)
            t_d = time_cuda(lambda: gemm.linear_dgrad(dy, w), it)
            t_w = time_cuda(lambda: gemm.linear_wgrad(dy, x, dw), it)
            t_ref = time_cuda(lambda: torch.matmul(x, w.t()), it)
            emit(bench="gemm_xt", dtype=tag, shape=name, M=M, N=N, K=K, fwd_ms=t_f, fwd_tflops=flops / t_f / 1e9, dgrad_ms=t_d,
                 dgrad_tflops=flops / t_d / 1e9, wgrad_ms=t_w, wgrad_tflops=flops / t_w / 1e9, lib_ms=t_ref, lib_tflops=flops / t_ref / 1e9)
