#!/usr/bin/env python
"""Gradient-clear benchmark (reference: benchmarks/grad_clear_benchmark.cpp:10-59): zero every
gradient tensor of GPT-2 small one by one ("manual") vs one fill of the contiguous gradient slab
(``GraphContext.zero_grads``, "bulk").  Device-timed on a GPU, wall-clock on the host."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit  # noqa: E402

import torch  # noqa: E402

from tnn_b200.models import load_or_create_model  # noqa: E402


def timed(fn, dev, passes):
    ts = []
    for _ in range(passes):
        if dev.type == "cuda":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        else:
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
    return ts


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "gpt2_small"
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    graph = load_or_create_model(model, device=dev, seed=123456)
    ctx = graph.context()
    grads = [l.grads[d.name] for l, d, _, _ in ctx.entries if d.name in l.grads]
    passes = 10

    def manual():
        for g in grads:
            g.zero_()

    manual(), ctx.zero_grads()                                    # warm-up
    t_manual = timed(manual, dev, passes)
    t_bulk = timed(ctx.zero_grads, dev, passes)
    for i, t in enumerate(t_manual):
        print(f"Manual Clear: Pass {i + 1} took {t:.3f} ms")
    for i, t in enumerate(t_bulk):
        print(f"Bulk Clear: Pass {i + 1} took {t:.3f} ms")
    nbytes = ctx.grad_slab.numel() * ctx.grad_slab.element_size()
    m, b = sorted(t_manual)[passes // 2], sorted(t_bulk)[passes // 2]
    emit(bench="grad_clear", model=model, device=str(dev), tensors=len(grads), grad_MB=round(nbytes / 2 ** 20, 1),
         manual_ms=round(m, 4), bulk_ms=round(b, 4), bulk_GBps=round(nbytes / b / 1e6, 1), speedup=round(m / b, 2))
    return 0


if __name__ == "__main__":
    sys.exit(main())
