#!/usr/bin/env python
"""Dense-layer benchmark with a parity check between the two layer generations (reference:
benchmarks/dense_benchmark.cpp:12-183: DenseLayer vs LegacyDenseLayer at 262144 -> 1024 features,
batch 128; forward and backward timed over 10 passes, then max |diff| of outputs, input gradients
and weight gradients).

  python benchmarks/dense_bench.py [--in 262144 --out 1024 --batch 128 --passes 10]

On a GPU both layers run the tcgen05 GEMM (the legacy class only differs in its registered type
name), so the diffs must be exactly zero and the times equal within noise; the torch.nn.Linear
(cuBLAS) row is the library baseline.  Runs on the CPU fp32 path at small sizes too."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit  # noqa: E402

import torch  # noqa: E402

from tnn_b200.nn.layers import DenseLayer, LegacyDenseLayer  # noqa: E402


def _time(fn, dev, passes):
    for _ in range(3):
        fn()
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
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--in", dest="fin", type=int, default=262144)
    ap.add_argument("--out", dest="fout", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--passes", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    dt = torch.bfloat16 if dev.type == "cuda" else torch.float32
    cur = DenseLayer(a.fin, a.fout, True, "dense").init(dev, seed=123456).set_training(True)
    leg = LegacyDenseLayer(a.fin, a.fout, True, "legacy_dense").init(dev, seed=654321).set_training(True)
    for k in cur.params:                                   # same weights in both (dense_benchmark.cpp:31)
        leg.params[k].copy_(cur.params[k])
    for l in (cur, leg):
        if hasattr(l, "refresh_lp"):
            l.refresh_lp()
    x = torch.randn(a.batch, a.fin, device=dev).to(dt)
    dy = torch.randn(a.batch, a.fout, device=dev).to(dt)
    flops = 2.0 * a.batch * a.fin * a.fout
    res = dict(bench="dense", batch=a.batch, in_features=a.fin, out_features=a.fout, dtype=str(dt).split(".")[1], device=str(dev))
    outs, gins = {}, {}
    for name, l in (("dense", cur), ("legacy_dense", leg)):
        t_f = _time(lambda: l.forward(x, 0), dev, a.passes)
        outs[name] = l.forward(x, 0)
        l.zero_grads()
        gins[name] = l.backward(dy, 0)
        keep = {k: v.clone() for k, v in l.grads.items()}

        def bwd():
            l.forward(x, 0)
            return l.backward(dy, 0)

        t_fb = _time(bwd, dev, a.passes)
        l.grads_snapshot = keep
        print(f"{name}: forward {t_f:.3f} ms, forward+backward {t_fb:.3f} ms")
        res[f"{name}_fwd_ms"] = round(t_f, 4)
        res[f"{name}_fwd_bwd_ms"] = round(t_fb, 4)
        res[f"{name}_fwd_tflops"] = round(flops / t_f / 1e9, 2)
    res["max_diff"] = float((outs["dense"].float() - outs["legacy_dense"].float()).abs().max())
    res["max_grad_diff"] = float((gins["dense"].float() - gins["legacy_dense"].float()).abs().max())
    res["max_weight_grad_diff"] = max(float((cur.grads_snapshot[k].float() - leg.grads_snapshot[k].float()).abs().max())
                                      for k in cur.grads_snapshot)
    print(f"Max diff: {res['max_diff']}\nMax grad diff: {res['max_grad_diff']}\nMax weight grad diff: {res['max_weight_grad_diff']}")
    lin = torch.nn.Linear(a.fin, a.fout, device=dev, dtype=dt)
    res["torch_fwd_ms"] = round(_time(lambda: lin(x), dev, a.passes), 4)
    emit(**res)
    return 0 if res["max_diff"] == 0.0 else 1


if __name__ == "__main__":
    sys.exit(main())
