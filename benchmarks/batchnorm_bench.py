#!/usr/bin/env python
"""BatchNorm micro-benchmark (reference: benchmarks/batchnorm_benchmark.cpp): train-forward
(stats + apply with fused ReLU) and backward (reduce + apply) vs torch/cuDNN at WRN shapes;
reports achieved GB/s against the measured copy bandwidth."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit, peaks, time_cuda  # noqa: E402

import torch  # noqa: E402

from tnn_b200.nn.layers import BatchNormLayer  # noqa: E402

if __name__ == "__main__":
    _, hbm, how = peaks()
    dev = torch.device("cuda", 0)
    only = sys.argv[1:]
    for name, N, H, C in [("wrn_g1", 256, 32, 128), ("wrn_g2", 256, 16, 256), ("wrn_g3", 256, 8, 512), ("r9_64", 256, 32, 64),
                          ("pipeline_224", 32, 224, 64)]:
        if only and name not in only:
            continue
        l = BatchNormLayer(C, 1e-2, 0.1, True, True, name).init(dev, 0)
        x = torch.randn(N, H, H, C, device=dev).bfloat16()
        dy = torch.randn_like(x)
        nbytes = x.numel() * 2
        t_f = time_cuda(lambda: l.forward(x, 0), 20)
        l.forward(x, 0)
        keep = dict(l.cache(0))

        def bwd():
            l._cache[0] = dict(keep)
            return l.backward_impl(dy, 0)

        t_b = time_cuda(bwd, 20)
        bn = torch.nn.BatchNorm2d(C, eps=1e-2).to(dev).to(memory_format=torch.channels_last)
        xc = x.permute(0, 3, 1, 2)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            t_ref = time_cuda(lambda: torch.relu(bn(xc)), 20)
        # fwd: read x twice (stats, apply) + write y ; bwd: reduce reads dy,x,y ; apply reads dy,x,y writes dx
        emit(bench="batchnorm", shape=name, N=N, HW=H, C=C, fwd_ms=t_f, fwd_gbs=3 * nbytes / t_f / 1e6, bwd_ms=t_b,
             bwd_gbs=7 * nbytes / t_b / 1e6, torch_fwd_ms=t_ref, frac_of_hbm_fwd=3 * nbytes / t_f / 1e6 / hbm, peak=how)
