#!/usr/bin/env python
"""Layer-chain micro-benchmark (reference: benchmarks/pipeline_benchmark.cpp: BatchNorm -> MaxPool
on [128, 224, 224, 64]) through a Sequential, forward and backward."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit, time_cuda  # noqa: E402

import torch  # noqa: E402

from tnn_b200.nn import LayerBuilder, Sequential  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    seq = Sequential(LayerBuilder([224, 224, 64]).batchnorm(1e-5, 0.1, True, True, "bn").maxpool2d(2, 2, name="pool").build(), "chain")
    seq.init(dev, 0)
    x = torch.randn(128, 224, 224, 64, device=dev).bfloat16()
    y = seq.forward(x, 0)
    dy = torch.randn_like(y)
    t_f = time_cuda(lambda: seq.forward(x, 0), 10)

    def fb():
        seq.forward(x, 0)
        return seq.backward(dy, 0)

    t_fb = time_cuda(fb, 10)
    emit(bench="bn_maxpool_chain", shape=list(x.shape), fwd_ms=t_f, fwd_bwd_ms=t_fb, fwd_gbs=(2 * x.numel() * 2 + x.numel() * 2 + y.numel() * 3) / t_f / 1e6)
