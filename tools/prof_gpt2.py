#!/usr/bin/env python
"""One GPT-2 small prefill (fp8 or bf16) + a few decode steps, for `ncu --metrics gpu__time_duration.sum` launch lists.
  ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <N> --csv --log-file out.csv python tools/prof_gpt2.py --mode prefill-fp8"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tnn_b200.inference import GPT2InferenceEngine
from tnn_b200.models import load_or_create_model

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="prefill-fp8", choices=["prefill-fp8", "prefill-bf16", "decode"])
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seq", type=int, default=1024)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = load_or_create_model("gpt2_small", device=dev, seed=1)
eng = GPT2InferenceEngine(g, max_batch=a.batch, max_len=a.seq, precision="fp8" if a.mode == "prefill-fp8" else "bf16")
prompt = torch.randint(0, 50257, (a.batch, a.seq), dtype=torch.int32).to(dev)
if a.mode.startswith("prefill"):
    for _ in range(a.reps):
        eng.prefill(prompt)
else:
    tok = eng.prefill(prompt[:, : a.seq - 16])
    for _ in range(a.reps + 2):
        tok = eng.decode(tok)
torch.cuda.synchronize()
print("done")
