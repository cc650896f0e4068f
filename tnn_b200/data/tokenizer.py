"""GPT-2 tokenizer: ``vocab.bin`` decode table of the reference (include/tokenizer/tokenizer.hpp:11-68,
python/export_vocab.py: ``u32 n; {u32 len; bytes} x n``) plus a byte-level BPE ENCODER over the
same table (native, csrc/runtime/tokenizer.hpp) so prompts can be text."""
from __future__ import annotations

from ..ops import native


class Tokenizer:
    def __init__(self, vocab_path: str | None = None):
        self._t = native.require().Tokenizer()
        if vocab_path:
            self.load(vocab_path)

    def load(self, path: str):
        self._t.load(path)
        return self

    def save(self, path: str):
        self._t.save(path)

    def set_vocab(self, tokens: list[bytes]):
        self._t.set_vocab([bytes(t) for t in tokens])
        return self

    @staticmethod
    def byte_fallback(extra_merges: list[bytes] | None = None) -> "Tokenizer":
        """256 single-byte tokens (+ optional merged tokens): a self-contained vocabulary for tests
        and for running without the tiktoken export."""
        t = Tokenizer()
        t.set_vocab([bytes([i]) for i in range(256)] + list(extra_merges or []))
        return t

    def vocab_size(self) -> int:
        return self._t.vocab_size()

    def decode(self, ids) -> str:
        if isinstance(ids, int):
            return self._t.decode_one(ids).decode("utf-8", errors="replace")
        return self._t.decode([int(i) for i in ids]).decode("utf-8", errors="replace")

    def decode_bytes(self, ids) -> bytes:
        return self._t.decode([int(i) for i in ids])

    def encode(self, text: str) -> list[int]:
        return list(self._t.encode(text))
