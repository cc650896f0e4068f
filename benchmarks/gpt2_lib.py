"""GPT-2 small inference measurement (BASELINE.json config 4: bs=32, seq=1024, fp8, single B200) shared by
``bench.py`` (the driver's JSON line carries a ``gpt2`` block) and ``benchmarks/gpt2_inference_bench.py``.

Random-init weights, synthetic token ids, CUDA-event timing after warm-up.  Per timed iteration the working set
(163 M parameters + [32768, 768..3072] activations, ~1 GB) is far larger than the 126 MB L2, so no explicit flush.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _timed(fn, iters, warm=2):
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


def tnn_gpt2(dev, batch=32, seq=1024, decode_steps=64, iters=5, model="gpt2_small") -> dict:
    """Prefill tokens/s (bf16 and block-scaled fp8) and KV-cached greedy decode tokens/s through the public
    ``GPT2InferenceEngine`` API, plus the fp8-vs-bf16 logits error of the same model on the same prompt."""
    from tnn_b200.inference import GPT2InferenceEngine
    from tnn_b200.models import load_or_create_model
    from tnn_b200.ops import native
    Cx = native.require()
    g = load_or_create_model(model, device=dev, seed=1)
    eng = GPT2InferenceEngine(g, max_batch=batch, max_len=seq)
    gen = torch.Generator().manual_seed(0)
    prompt = torch.randint(0, 50257, (batch, seq), generator=gen, dtype=torch.int32).to(dev)
    out = {"model": model, "batch": batch, "seq": seq, "decode_steps": decode_steps}
    # ---- bf16 prefill
    _, logits_bf16 = eng.prefill(prompt, return_logits=True)
    logits_bf16 = logits_bf16.float().clone()
    ms = _timed(lambda: eng.prefill(prompt), iters)
    out["prefill_bf16_ms"] = ms
    out["prefill_bf16_tok_s"] = batch * seq / ms * 1e3
    # ---- decode (CUDA-graph replays, tokens + position stay on the device)
    start = seq - decode_steps
    tok = torch.randint(0, 50257, (batch,), generator=gen, dtype=torch.int32).to(dev)
    eng.prefill(prompt[:, :start])
    eng.decode_graph(tok, 2)
    eng.prefill(prompt[:, :start])
    torch.cuda.synchronize()
    l0 = Cx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.decode_graph(tok, decode_steps)
    e1.record()
    torch.cuda.synchronize()
    msd = e0.elapsed_time(e1) / decode_steps
    out["decode_ms_per_step"] = msd
    out["decode_tok_s"] = batch / msd * 1e3
    out["decode_dtype"] = "bf16 (cluster split-K skinny GEMMs, CUDA-graph replay)"
    # ---- fp8 prefill
    eng.enable_fp8()
    l0 = Cx.launch_count()
    _, logits_fp8 = eng.prefill(prompt, return_logits=True)
    out["launches_prefill"] = int(Cx.launch_count() - l0)
    logits_fp8 = logits_fp8.float()
    ms8 = _timed(lambda: eng.prefill(prompt), iters)
    out["prefill_ms"] = ms8
    out["prefill_tok_s"] = batch * seq / ms8 * 1e3
    out["dtype"] = "fp8 (mxfp8 e4m3 + ue8m0/32 block scales on the qkv/proj/fc1/fc2 GEMMs; attention, residual stream, LM head bf16)"
    d = (logits_fp8 - logits_bf16).abs()
    out["logits_max_abs_err_vs_bf16"] = float(d.max())
    out["logits_rms_err_vs_bf16"] = float(d.pow(2).mean().sqrt())
    out["logits_rms_bf16"] = float(logits_bf16.pow(2).mean().sqrt())
    out["argmax_agreement"] = float((logits_fp8.argmax(-1) == logits_bf16.argmax(-1)).float().mean())
    return out


def tnn_gpt2_cpp(gpu: int = 0, batch=32, seq=1024, decode_steps=64, iters=5, model="gpt2_small") -> dict:
    """The same measurement through the C++ host's ``GPT2Engine`` (cpp/src/inference.cpp, bound as
    ``tnn_b200._host.GPT2Engine``): model, KV caches, buffers, CUDA-graph decode and the CUDA-event timing all live in
    libtnn_b200; Python only passes token arrays."""
    import numpy as np
    from tnn_b200 import _host
    dev = f"GPU:{gpu}"
    eng = _host.GPT2Engine(model, dev, 1, batch, seq, "bf16")
    rs = np.random.RandomState(0)
    prompt = rs.randint(0, 50257, (batch, seq)).astype(np.int32)
    out = {"model": model, "batch": batch, "seq": seq, "decode_steps": decode_steps, "host": "cpp (libtnn_b200 GPT2Engine)"}
    _, logits_bf16 = eng.prefill(prompt, True)
    ms = eng.time_prefill_ms(batch, seq, iters, 2)
    out["prefill_bf16_ms"] = ms
    out["prefill_bf16_tok_s"] = batch * seq / ms * 1e3
    msd = eng.time_decode_ms(batch, seq - decode_steps, decode_steps, True)
    out["decode_ms_per_step"] = msd
    out["decode_tok_s"] = batch / msd * 1e3
    out["decode_dtype"] = "bf16 (persistent whole-model decode kernel, one launch per token, CUDA-graph replay)" if eng.uses_persistent_decode() \
        else "bf16 (cluster split-K skinny GEMMs, CUDA-graph replay)"
    eng.enable_fp8()
    l0 = _host.launch_count()
    _, logits_fp8 = eng.prefill(prompt, True)
    out["launches_prefill"] = int(_host.launch_count() - l0)
    ms8 = eng.time_prefill_ms(batch, seq, iters, 2)
    out["prefill_ms"] = ms8
    out["prefill_tok_s"] = batch * seq / ms8 * 1e3
    out["dtype"] = "fp8 (mxfp8 e4m3 + ue8m0/32 block scales on the qkv/proj/fc1/fc2 GEMMs; attention, residual stream, LM head bf16)"
    d = np.abs(logits_fp8 - logits_bf16)
    out["logits_max_abs_err_vs_bf16"] = float(d.max())
    out["logits_rms_err_vs_bf16"] = float(np.sqrt((d ** 2).mean()))
    out["logits_rms_bf16"] = float(np.sqrt((logits_bf16 ** 2).mean()))
    out["argmax_agreement"] = float((logits_fp8.argmax(-1) == logits_bf16.argmax(-1)).mean())
    # decode again with the fp8 engine: the persistent kernel then attends over the fp8 KV cache the fp8 prefill wrote
    msd8 = eng.time_decode_ms(batch, seq - decode_steps, decode_steps, True)
    if eng.uses_fp8_kv_cache():
        out["decode_bf16_kv_ms_per_step"] = out["decode_ms_per_step"]
        out["decode_ms_per_step"] = msd8
        out["decode_tok_s"] = batch / msd8 * 1e3
        out["decode_dtype"] = "bf16 weights, fp8 (e4m3 + ue8m0 per 32) KV cache; persistent whole-model decode kernel, CUDA-graph replay"
    return out


def torch_gpt2(dev, batch=32, seq=1024, decode_steps=64, iters=5) -> dict:
    """Library arm: torch + cuBLAS + flash SDPA in bf16; the decode step is captured in a CUDA graph
    (static shapes: the position is a device tensor, attention runs over the full cache with a length mask)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import torch.nn.functional as F
    from torch_gpt2 import GPT2
    m = GPT2(Lmax=seq).to(dev).to(torch.bfloat16).eval()
    gen = torch.Generator().manual_seed(0)
    prompt = torch.randint(0, 50257, (batch, seq), generator=gen).to(dev)
    caches = m.new_cache(batch, dev, torch.bfloat16)
    out = {"batch": batch, "seq": seq, "decode_steps": decode_steps, "dtype": "bf16"}
    ms = _timed(lambda: m.step(prompt, caches, 0), iters)
    out["prefill_ms"] = ms
    out["prefill_tok_s"] = batch * seq / ms * 1e3
    start = seq - decode_steps
    m.step(prompt[:, :start], caches, 0)
    tok = torch.randint(0, 50257, (batch,), generator=gen).to(dev)
    pos_t = torch.full((1,), start, dtype=torch.int64, device=dev)
    ar = torch.arange(seq, device=dev)

    @torch.no_grad()
    def static_step():
        x = m.wte(tok).unsqueeze(1) + m.wpe(pos_t).unsqueeze(0)
        mask = (ar <= pos_t).view(1, 1, 1, seq)
        for b, c in zip(m.blocks, caches):
            B, L, E = x.shape
            q, k, v = b.qkv(b.ln1(x)).view(B, 1, 3, b.H, E // b.H).unbind(2)
            c[0].index_copy_(1, pos_t, k)
            c[1].index_copy_(1, pos_t, v)
            o = F.scaled_dot_product_attention(q.transpose(1, 2), c[0].transpose(1, 2), c[1].transpose(1, 2), attn_mask=mask)
            x = x + b.proj(o.transpose(1, 2).reshape(B, 1, E))
            x = x + b.fc2(F.gelu(b.fc1(b.ln2(x)), approximate="tanh"))
        nxt = m.head(m.ln_f(x[:, -1:])).argmax(-1).reshape(batch)
        tok.copy_(nxt)
        pos_t.add_(1)

    mode = "cuda-graph"
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                static_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        pos_t.fill_(start)
        with torch.cuda.graph(graph):
            static_step()
        run = graph.replay
    except Exception as e:  # noqa: BLE001
        mode, run = f"eager ({type(e).__name__})", static_step
        torch.cuda.synchronize()
    pos_t.fill_(start)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(decode_steps - 1):
        run()
    e1.record()
    torch.cuda.synchronize()
    msd = e0.elapsed_time(e1) / (decode_steps - 1)
    out["decode_ms_per_step"] = msd
    out["decode_tok_s"] = batch / msd * 1e3
    out["decode_execution"] = mode
    return out
