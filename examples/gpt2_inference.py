#!/usr/bin/env python
"""GPT-2 text generation (reference: examples/gpt2_inference.cpp:19-127): KV-cached greedy
decoding of a zoo GPT-2 model.  ``--vocab vocab.bin`` is the reference's decode table
(python/export_vocab.py); without it a byte-level fallback vocabulary is used.  ``--prompt`` is
text (encoded by the built-in BPE/byte encoder) or ``--prompt-tokens`` a comma list of ids.

  python examples/gpt2_inference.py --model gpt2_small --path model_snapshots/gpt2 --prompt "Hello" --tokens 50
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.data import Tokenizer  # noqa: E402
from tnn_b200.inference import GPT2InferenceEngine  # noqa: E402
from tnn_b200.models import load_or_create_model  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gpt2_small")
    ap.add_argument("--path", default="")
    ap.add_argument("--vocab", default="")
    ap.add_argument("--prompt", default="The quick brown fox")
    ap.add_argument("--prompt-tokens", default="")
    ap.add_argument("--tokens", type=int, default=50)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    a = ap.parse_args()
    tok = Tokenizer(a.vocab) if a.vocab else Tokenizer.byte_fallback()
    ids = [int(t) for t in a.prompt_tokens.split(",")] if a.prompt_tokens else tok.encode(a.prompt)
    g = load_or_create_model(a.model, a.path or None, device=a.device, seed=1)
    eng = GPT2InferenceEngine(g, max_batch=a.batch)
    prompt = torch.tensor([ids] * a.batch, dtype=torch.int32)
    t0 = time.time()
    out = eng.generate(prompt, a.tokens, eos_token=50256)
    if a.device.startswith("cuda"):
        torch.cuda.synchronize()
    dt = time.time() - t0
    print(a.prompt + tok.decode([t for t in out[0].tolist() if t < tok.vocab_size()]))
    print(f"[gpt2_inference] {out.numel()} tokens in {dt:.3f}s ({out.numel() / dt:.1f} tok/s, batch {a.batch})", file=sys.stderr)
