#!/usr/bin/env python
"""Write the GPT-2 decode table ``vocab.bin`` (``u32 n; {u32 len; bytes} x n``; reference: python/export_vocab.py,
include/tokenizer/tokenizer.hpp:15-53) that both hosts' tokenizers load.

The reference exports it from ``tiktoken``.  There is no network on the build / GPU boxes, so three sources are supported:

  --from tiktoken [NAME]          tiktoken.get_encoding(NAME)           (the reference's route; needs the package + its cache)
  --from vocab-json FILE          a Hugging Face ``vocab.json`` of a byte-level BPE tokenizer (token string -> id); the
                                  printable-unicode alphabet is mapped back to raw bytes (the GPT-2 bytes<->unicode table)
  --from byte-fallback            256 single-byte tokens (+ ``--eot``): a self-contained vocabulary for smoke runs

  python tools/export_vocab.py --from vocab-json gpt2/vocab.json --out data/open-web-text/vocab.bin
"""
import argparse
import json
import os
import struct
import sys


def bytes_to_unicode() -> dict[int, str]:
    """GPT-2's reversible byte -> printable unicode character table (public algorithm of the GPT-2 encoder)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    chars = keep[:]
    n = 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(keep, chars)}


def tokens_from_vocab_json(path: str) -> list[bytes]:
    with open(path, encoding="utf-8") as f:
        vocab = json.load(f)
    back = {c: b for b, c in bytes_to_unicode().items()}
    n = max(vocab.values()) + 1
    table: list[bytes | None] = [None] * n
    for tok, i in vocab.items():
        try:
            table[i] = bytes(back[ch] for ch in tok)
        except KeyError:                      # special tokens such as <|endoftext|> are stored verbatim
            table[i] = tok.encode("utf-8")
    missing = [i for i, t in enumerate(table) if t is None]
    if missing:
        raise SystemExit(f"{path}: ids without a token: {missing[:8]}{' ...' if len(missing) > 8 else ''}")
    return table  # type: ignore[return-value]


def tokens_from_tiktoken(name: str) -> list[bytes]:
    try:
        import tiktoken
    except ImportError as e:
        raise SystemExit(f"tiktoken is not installed ({e}); use --from vocab-json or --from byte-fallback") from e
    enc = tiktoken.get_encoding(name)
    return [enc.decode_bytes([i]) for i in range(enc.n_vocab)]


def write_vocab(tokens: list[bytes], path: str) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(tokens)))
        for t in tokens:
            f.write(struct.pack("<I", len(t)))
            f.write(t)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--from", dest="source", nargs="+", required=True, metavar="KIND",
                    help="tiktoken [NAME] | vocab-json FILE | byte-fallback")
    ap.add_argument("--out", default="data/open-web-text/vocab.bin")
    ap.add_argument("--eot", default="<|endoftext|>", help="byte-fallback: text of the extra end-of-text token ('' for none)")
    a = ap.parse_args(argv)
    kind, rest = a.source[0], a.source[1:]
    if kind == "tiktoken":
        tokens = tokens_from_tiktoken(rest[0] if rest else "gpt2")
    elif kind == "vocab-json":
        if not rest:
            ap.error("--from vocab-json needs the file")
        tokens = tokens_from_vocab_json(rest[0])
    elif kind == "byte-fallback":
        tokens = [bytes([i]) for i in range(256)] + ([a.eot.encode()] if a.eot else [])
    else:
        ap.error(f"unknown source {kind!r}")
    write_vocab(tokens, a.out)
    print(f"wrote {len(tokens)} tokens to {a.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
