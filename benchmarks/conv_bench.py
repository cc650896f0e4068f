#!/usr/bin/env python
"""Conv2d micro-benchmark (reference: benchmarks/conv2d_benchmark.cpp): implicit-GEMM tcgen05
fprop / dgrad / wgrad vs cuDNN (torch channels_last bf16) at the WRN-16-8 / ResNet-9 shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit, peaks, time_cuda  # noqa: E402

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from tnn_b200.nn.layers import Conv2DLayer  # noqa: E402

SHAPES = [  # (name, N, H, C, K, ksize, stride)
    ("wrn_g1_128", 256, 32, 128, 128, 3, 1), ("wrn_g2_256", 256, 16, 256, 256, 3, 1), ("wrn_g3_512", 256, 8, 512, 512, 3, 1),
    ("wrn_g2_down", 256, 32, 128, 256, 3, 2), ("wrn_g3_down", 256, 16, 256, 512, 3, 2), ("wrn_first", 256, 32, 16, 128, 3, 1),
    ("wrn_sc_1x1", 256, 32, 128, 256, 1, 2), ("r9_64_128", 256, 32, 64, 128, 3, 1),
]

if __name__ == "__main__":
    peak, _, how = peaks()
    dev = torch.device("cuda", 0)
    for name, N, H, C, K, ks, st in SHAPES:
        pad = ks // 2
        l = Conv2DLayer(C, K, ks, ks, st, st, pad, pad, True, name).init(dev, 0)
        x = torch.randn(N, H, H, C, device=dev).bfloat16()
        y = l.forward(x, 0)
        dy = torch.randn_like(y)
        flops = 2.0 * y.numel() * ks * ks * C

        def fwd():
            return l.forward(x, 0)

        def bwd():
            l.forward_impl  # noqa: B018
            return l.backward_impl(dy, 0)

        t_f = time_cuda(fwd, 20)
        l.forward(x, 0)
        t_b = time_cuda(lambda: (l.cache(0).update(x=x, plan=l._plan(x)), l.backward_impl(dy, 0)), 20)
        xc = x.permute(0, 3, 1, 2)
        wc = l.params["weight"].permute(0, 3, 1, 2).bfloat16().contiguous(memory_format=torch.channels_last)
        t_ref = time_cuda(lambda: F.conv2d(xc, wc, None, st, pad), 20)
        emit(bench="conv2d", shape=name, N=N, HW=H, C=C, K=K, k=ks, stride=st, fwd_ms=t_f, fwd_tflops=flops / t_f / 1e9,
             bwd_ms=t_b, bwd_tflops=2 * flops / t_b / 1e9, cudnn_fwd_ms=t_ref, cudnn_fwd_tflops=flops / t_ref / 1e9,
             frac_of_peak=flops / t_f / 1e9 / peak, peak=how)
