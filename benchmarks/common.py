"""Shared timing helpers for the micro-benchmarks: CUDA events on the launching stream, warm-up,
an L2 flush (a 256 MB write) between timed iterations, result as a JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

_FLUSH = None


def flush_l2():
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _FLUSH.zero_()


def time_cuda(fn, iters=20, warm=3, flush=True):
    """Median device time (ms) of fn() over `iters` runs, each preceded by an L2 flush."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 6650.0, "fallback"


def emit(**kv):
    print(json.dumps(kv), flush=True)
