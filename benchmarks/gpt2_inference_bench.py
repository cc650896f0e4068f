#!/usr/bin/env python
"""GPT-2 small inference, bs=32, seq=1024 (BASELINE.json config 4): prefill tokens/s over the
full 1024-token window and KV-cached greedy decode tokens/s, tnn_b200 engine vs a torch
(cuBLAS + flash SDPA) baseline.  Random-init weights, synthetic token ids, device-event timing.

  python benchmarks/gpt2_inference_bench.py [--impl tnn|torch|both] [--batch 32] [--seq 1024] [--decode 64]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))

import torch


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_tnn(a, dev):
    from tnn_b200.inference import GPT2InferenceEngine
    from tnn_b200.models import load_or_create_model
    from tnn_b200.ops import native
    Cx = native.require()
    g = load_or_create_model(a.model, device=dev, seed=1)
    eng = GPT2InferenceEngine(g, max_batch=a.batch, max_len=a.seq)
    gen = torch.Generator().manual_seed(0)
    prompt = torch.randint(0, 50257, (a.batch, a.seq), generator=gen, dtype=torch.int32).to(dev)
    l0 = Cx.launch_count()
    eng.prefill(prompt)
    launches_prefill = Cx.launch_count() - l0
    ms_prefill = timed(lambda: eng.prefill(prompt), a.iters)
    # decode from a half-full cache up to the full window
    start = a.seq - a.decode
    eng.prefill(prompt[:, :start])
    tok = torch.randint(0, 50257, (a.batch,), generator=gen, dtype=torch.int32).to(dev)
    torch.cuda.synchronize()
    l0 = Cx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.decode):
        tok = eng.decode(tok)
    e1.record()
    torch.cuda.synchronize()
    ms_decode = e0.elapsed_time(e1) / a.decode
    # the same decode as replays of one captured CUDA graph (tokens + position stay on the device)
    eng.prefill(prompt[:, :start])
    eng.decode_graph(tok, 2)
    eng.prefill(prompt[:, :start])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.decode_graph(tok, a.decode)
    e1.record()
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / a.decode
    return {"impl": "tnn_b200", "decode_graph_ms_per_step": ms_graph, "decode_graph_tok_s": a.batch / ms_graph * 1e3,
            "prefill_ms": ms_prefill, "prefill_tok_s": a.batch * a.seq / ms_prefill * 1e3,
            "decode_ms_per_step": ms_decode, "decode_tok_s": a.batch / ms_decode * 1e3,
            "launches_prefill": int(launches_prefill), "launches_per_decode_step": (Cx.launch_count() - l0) / a.decode}


def run_torch(a, dev):
    from torch_gpt2 import GPT2
    m = GPT2(Lmax=a.seq).to(dev).to(torch.bfloat16).eval()
    gen = torch.Generator().manual_seed(0)
    prompt = torch.randint(0, 50257, (a.batch, a.seq), generator=gen).to(dev)
    caches = m.new_cache(a.batch, dev, torch.bfloat16)
    ms_prefill = timed(lambda: m.step(prompt, caches, 0), a.iters)
    start = a.seq - a.decode
    m.step(prompt[:, :start], caches, 0)
    tok = torch.randint(0, 50257, (a.batch, 1), generator=gen).to(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.decode):
        tok = m.step(tok, caches, start + i).reshape(-1, 1)
    e1.record()
    torch.cuda.synchronize()
    ms_decode = e0.elapsed_time(e1) / a.decode
    return {"impl": "torch", "prefill_ms": ms_prefill, "prefill_tok_s": a.batch * a.seq / ms_prefill * 1e3,
            "decode_ms_per_step": ms_decode, "decode_tok_s": a.batch / ms_decode * 1e3}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="both")
    ap.add_argument("--model", default="gpt2_small")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--decode", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = {"model": a.model, "batch": a.batch, "seq": a.seq, "decode_steps": a.decode, "dtype": "bf16", "data": "synthetic"}
    for impl in (["tnn", "torch"] if a.impl == "both" else [a.impl]):
        r = run_tnn(a, dev) if impl == "tnn" else run_torch(a, dev)
        r["config"] = cfg
        print(json.dumps(r), flush=True)
