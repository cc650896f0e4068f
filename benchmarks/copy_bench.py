#!/usr/bin/env python
"""Copy / grad-clear micro-benchmark (reference: benchmarks/copy_benchmark.cpp,
grad_clear_benchmark.cpp): H2D from pinned and pageable memory, D2D, D2H, peer copy when two GPUs
are visible, and clearing the contiguous gradient slab (one fill vs one fill per tensor)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit, time_cuda  # noqa: E402

import torch  # noqa: E402

from tnn_b200.models import load_or_create_model  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    for mb in (1, 16, 256, 1024):
        n = mb << 20
        d = torch.empty(n, dtype=torch.uint8, device=dev)
        d2 = torch.empty_like(d)
        hp = torch.empty(n, dtype=torch.uint8).pin_memory()
        hu = torch.empty(n, dtype=torch.uint8)
        row = dict(bench="copy", MB=mb)
        row["h2d_pinned_gbs"] = n / time_cuda(lambda: d.copy_(hp, non_blocking=True), 10, flush=False) / 1e6
        row["h2d_pageable_gbs"] = n / time_cuda(lambda: d.copy_(hu), 5, flush=False) / 1e6
        row["d2h_pinned_gbs"] = n / time_cuda(lambda: hp.copy_(d, non_blocking=True), 10, flush=False) / 1e6
        row["d2d_gbs"] = 2 * n / time_cuda(lambda: d2.copy_(d), 10, flush=False) / 1e6
        if torch.cuda.device_count() > 1:
            p = torch.empty(n, dtype=torch.uint8, device="cuda:1")
            row["peer_gbs"] = n / time_cuda(lambda: p.copy_(d, non_blocking=True), 10, flush=False) / 1e6
        emit(**row)
    g = load_or_create_model("cifar100_wrn16_8", device=dev, seed=0)
    ctx = g.context()
    t_slab = time_cuda(ctx.zero_grads, 20)
    grads = [l.grads[d.name] for l, d, _, _ in ctx.entries]
    t_each = time_cuda(lambda: [x.zero_() for x in grads], 20)
    emit(bench="grad_clear", params=ctx.num_parameters(), tensors=len(grads), slab_ms=t_slab, per_tensor_ms=t_each)
