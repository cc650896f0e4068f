#!/usr/bin/env python
"""Tokenise a text corpus into the uint16 token streams ``train.bin`` / ``val.bin`` the OpenWebText loaders of both hosts
read (reference: python/openwebtext.py -- Hugging Face ``datasets`` + tiktoken, which need the network; this tool works
from local files with the repository's own byte-level BPE encoder over ``vocab.bin``, or tiktoken when it is installed).

Inputs: any mix of ``.txt`` files (one document per file), ``.jsonl`` files (one document per line, ``--text-field``), or
directories of them.  Every document is followed by the end-of-text token.  Documents are split ``--val-fraction`` by a hash
of their index, so the split is reproducible.

  python tools/export_vocab.py --from vocab-json gpt2/vocab.json --out data/open-web-text/vocab.bin
  python tools/prepare_openwebtext.py corpus/ --vocab data/open-web-text/vocab.bin --out data/open-web-text
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def documents(paths, text_field):
    for p in paths:
        if os.path.isdir(p):
            files = sorted(os.path.join(r, f) for r, _, fs in os.walk(p) for f in fs if f.endswith((".txt", ".jsonl")))
        else:
            files = [p]
        for f in files:
            if f.endswith(".jsonl"):
                with open(f, encoding="utf-8") as fh:
                    for line in fh:
                        line = line.strip()
                        if line:
                            yield json.loads(line)[text_field]
            else:
                with open(f, encoding="utf-8", errors="replace") as fh:
                    yield fh.read()


def make_encoder(vocab, use_tiktoken):
    if use_tiktoken:
        import tiktoken
        enc = tiktoken.get_encoding("gpt2")
        return enc.encode_ordinary, enc.eot_token, enc.n_vocab
    from tnn_b200.data import Tokenizer
    tok = Tokenizer(vocab)
    n = tok.vocab_size()
    eot = None
    for i in range(n - 1, max(-1, n - 8), -1):                 # GPT-2: the last id; byte-fallback export: the 257th
        if tok.decode_bytes([i]) == b"<|endoftext|>":
            eot = i
            break
    return tok.encode, eot, n


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("inputs", nargs="+")
    ap.add_argument("--vocab", default="data/open-web-text/vocab.bin")
    ap.add_argument("--out", default="data/open-web-text")
    ap.add_argument("--text-field", default="text")
    ap.add_argument("--val-fraction", type=float, default=0.0005, help="share of documents that goes to val.bin (reference split: 0.0005)")
    ap.add_argument("--tiktoken", action="store_true", help="encode with tiktoken's gpt2 encoding instead of vocab.bin")
    a = ap.parse_args(argv)
    encode, eot, n_vocab = make_encoder(a.vocab, a.tiktoken)
    if n_vocab > 65536:
        raise SystemExit(f"vocabulary of {n_vocab} tokens does not fit the uint16 stream format")
    os.makedirs(a.out, exist_ok=True)
    outs = {s: open(os.path.join(a.out, f"{s}.bin"), "wb") for s in ("train", "val")}
    counts = {"train": 0, "val": 0}
    docs = 0
    period = max(1, round(1.0 / a.val_fraction)) if a.val_fraction > 0 else 0
    for i, text in enumerate(documents(a.inputs, a.text_field)):
        ids = list(encode(text))
        if eot is not None:
            ids.append(eot)
        split = "val" if period and (i * 2654435761 % 2 ** 32) % period == period - 1 else "train"
        np.asarray(ids, dtype=np.uint16).tofile(outs[split])
        counts[split] += len(ids)
        docs += 1
    for f in outs.values():
        f.close()
    print(json.dumps({"documents": docs, "train_tokens": counts["train"], "val_tokens": counts["val"], "eot": eot, "vocab": n_vocab, "out": a.out}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
