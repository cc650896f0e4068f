#!/usr/bin/env python
"""GPT-2 small TRAINING step benchmark (reference: examples/trainer.cpp with MODEL_NAME=gpt2_small on
open-webtext, bs x 1024 tokens): one full step = forward (fused QKV GEMM, tcgen05 flash attention,
GELU/residual epilogues) + CE loss + backward (tcgen05 flash-attention backward, dgrad / wgrad GEMMs)
+ Adam, through tnn_b200's TrainStep, against a torch (cuBLAS + flash SDPA + fused AdamW, bf16
autocast) re-creation of the same model.  Synthetic token ids, random-init weights.

  python benchmarks/gpt2_train_bench.py [--impl tnn|torch|both] [--batch 8] [--seq 1024] [--steps 10]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit  # noqa: E402

import torch  # noqa: E402


def flops_per_token(E=768, layers=12, V=50257, L=1024):
    dense = layers * (2 * E * 3 * E + 2 * E * E + 2 * 2 * E * 4 * E) + 2 * E * V
    attn = layers * 4 * L * E // 2
    return 3 * (dense + attn)


def run_tnn(a, dev):
    from tnn_b200.models import load_or_create_model
    from tnn_b200.nn import LossFactory, OptimizerFactory
    from tnn_b200.nn.train import TrainStep
    graph = load_or_create_model(a.model, device=dev, seed=1234)
    opt = OptimizerFactory.create_adam(3e-4, 0.9, 0.999, 1e-8, 0.0)
    opt.attach(graph.context())
    step = TrainStep(graph, LossFactory.create_logsoftmax_crossentropy(), opt, None, 1, None)
    g = torch.Generator().manual_seed(7)
    batches = [(torch.randint(0, 50257, (a.batch, a.seq), generator=g, dtype=torch.int32).to(dev),
                torch.randint(0, 50257, (a.batch, a.seq), generator=g, dtype=torch.int32).to(dev)) for _ in range(2)]
    for i in range(a.warmup):
        st = step(*batches[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        st = step(*batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    return dict(impl="tnn_b200", ms_per_step=ms, tok_s=a.batch * a.seq / ms * 1e3, loss=float(st[0]), cuda_graph=bool(step.use_graph))


def run_torch(a, dev):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))
    from torch_gpt2 import GPT2
    import torch.nn.functional as F
    torch.manual_seed(0)
    m = GPT2().to(dev)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4, weight_decay=0.0, fused=True)
    ids = torch.randint(0, 50257, (a.batch, a.seq), device=dev)
    tgt = torch.randint(0, 50257, (a.batch, a.seq), device=dev)

    def one():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = m.forward_train(ids)
        loss = F.cross_entropy(logits.float().view(-1, logits.shape[-1]), tgt.view(-1))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss
    for _ in range(a.warmup):
        loss = one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    return dict(impl="torch", ms_per_step=ms, tok_s=a.batch * a.seq / ms * 1e3, loss=float(loss))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="both")
    ap.add_argument("--model", default="gpt2_small")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    fl = flops_per_token(L=a.seq)
    for impl in (["tnn", "torch"] if a.impl == "both" else [a.impl]):
        r = run_tnn(a, dev) if impl == "tnn" else run_torch(a, dev)
        r.update(bench="gpt2_train", model=a.model, batch=a.batch, seq=a.seq, tflops=fl * r["tok_s"] / 1e12)
        emit(**r)
        torch.cuda.empty_cache()
