"""GPT-2 inference engine: batched prefill + incremental decode over a real KV cache.

The reference's ``gpt2_inference`` (examples/gpt2_inference.cpp:44-121) has no KV
cache: every generated token re-runs the whole model on a zero-padded [1, 512] float tensor,
copies the [1, 512, 50257] logits (~103 MB) to the host and arg-maxes there.  This engine

  * walks a zoo GPT graph (``gpt2_small|medium|large`` / ``flash_*``) and runs it with
    per-layer KV caches [B, Lmax, H, 64] (bf16 on the GPU),
  * prefill: one pass over the prompt -- fused token+position embedding, LayerNorm, fused QKV
    GEMM, the tcgen05 flash-attention kernel reading Q/K/V in place from the packed projection,
    output projection and MLP GEMMs with bias / GELU / residual epilogues -- and the LM head is
    evaluated for the LAST position only,
  * decode: one token per sequence per step, the split-lane decode-attention kernel over the
    cache, greedy arg-max on the device (4 bytes per sequence cross PCIe, not 103 MB),
  * integer token ids end to end.

It uses the layer objects' own forward paths, so on a CPU it runs the fp32 oracle math and is
testable against a full recompute; on a GPU every op is a hand-written kernel.
"""
from __future__ import annotations

import torch

from ..nn.blocks import AttentionBlock, ResidualBlock, Sequential
from ..nn.graph import Graph
from ..nn.layers import (ActivationLayer, DenseLayer, DropoutLayer, EmbeddingLayer, LayerNormLayer,
                         PositionalEmbeddingLayer)
from ..ops import attention as attn_ops
from ..ops import gemm as gemm_ops
from ..ops import native


class _Block:
    __slots__ = ("ln1", "attn", "ln2", "fc1", "act", "fc2")


class GPT2InferenceEngine:
    """``precision="fp8"`` (GPU only) runs every token GEMM of the transformer blocks in block-scaled fp8
    (MXFP8: e4m3 + one ue8m0 scale per 32 elements, ``ops/mxfp8.py`` / ``csrc/kernels/mxfp8.cu``): weights are
    quantised once, LayerNorm writes its output directly as MXFP8, the fc1 epilogue applies GELU and
    re-quantises for fc2; attention, the residual stream and the LM head stay bf16."""

    def __init__(self, model: Graph | Sequential, max_batch: int = 1, max_len: int | None = None, precision: str = "bf16"):
        seq = model.get_layers()[0] if isinstance(model, Graph) else model
        self.model = seq
        self.device = seq.device
        seq.set_training(False)
        top = list(seq.layers)
        self.embed = next(l for l in top if isinstance(l, EmbeddingLayer))
        self.pos = next(l for l in top if isinstance(l, PositionalEmbeddingLayer))
        self.blocks: list[_Block] = []
        res = [l for l in top if isinstance(l, ResidualBlock)]
        if len(res) % 2:
            raise ValueError("expected (attention, mlp) residual pairs")
        for a, f in zip(res[0::2], res[1::2]):
            b = _Block()
            al, fl = a.main.layers, f.main.layers
            b.ln1 = next(l for l in al if isinstance(l, LayerNormLayer))
            b.attn = next(l for l in al if isinstance(l, AttentionBlock))
            b.ln2 = next(l for l in fl if isinstance(l, LayerNormLayer))
            dense = [l for l in fl if isinstance(l, DenseLayer)]
            b.fc1, b.fc2 = dense[0], dense[1]
            b.act = next((l.activation for l in fl if isinstance(l, ActivationLayer)), "gelu")
            self.blocks.append(b)
        tail = top[top.index(res[-1]) + 1:]
        self.ln_f = next(l for l in tail if isinstance(l, LayerNormLayer))
        self.head = next(l for l in tail if isinstance(l, DenseLayer))
        self.E = self.embed.embed_dim
        self.H = self.blocks[0].attn.num_heads
        self.D = self.E // self.H
        self.vocab = self.embed.vocab_size
        self.max_len = int(max_len or self.pos.seq_len)
        self.max_batch = int(max_batch)
        self.cuda = self.device.type == "cuda"
        self.dtype = torch.bfloat16 if self.cuda else torch.float32
        self.caches = [attn_ops.KVCache(self.max_batch, self.max_len, self.H, self.D, self.device, self.dtype)
                       for _ in self.blocks]
        self.cur_len = 0
        self.cur_batch = 0
        self._dev = None
        self._graph = None
        self.precision = "bf16"
        self._fp8 = None
        if precision == "fp8":
            self.enable_fp8()
        elif precision != "bf16":
            raise ValueError(f"precision must be bf16 or fp8, not {precision!r}")

    # ------------------------------------------------------------------ block-scaled fp8
    def enable_fp8(self):
        """Quantise the block weights to MXFP8 once (offline weight quantiser) and switch prefill to the fp8 path."""
        if not self.cuda:
            raise RuntimeError("the fp8 path needs a GPU (tcgen05 block-scaled MMA)")
        from ..ops import mxfp8
        self._fp8 = []
        for b in self.blocks:
            w = {}
            for key, layer, wn, bn in (("qkv", b.attn, "qkv_weight", "qkv_bias"), ("out", b.attn, "out_weight", "out_bias"),
                                       ("fc1", b.fc1, "weight", "bias"), ("fc2", b.fc2, "weight", "bias")):
                W = layer.params[wn].detach().float().contiguous()
                q, sf = mxfp8.quantize(W)
                bias = layer.params[bn].detach().float().contiguous() if bn in layer.params else None
                w[key] = (q, sf, bias)
            if b.act != "gelu":
                raise ValueError("the fp8 MLP epilogue implements GELU (GPT-2); got " + str(b.act))
            self._fp8.append(w)
        self.precision = "fp8"
        return self

    def _layers_fp8(self, x, pos: int):
        from ..ops import mxfp8
        B, L, E = x.shape
        H, D = self.H, self.D
        x2 = x.reshape(B * L, E).contiguous()
        for b, w, cache in zip(self.blocks, self._fp8, self.caches):
            view = attn_ops.KVCache.__new__(attn_ops.KVCache)
            view.k, view.v, view.len = cache.k[:B], cache.v[:B], cache.len
            q, sf = mxfp8.layernorm_quant(x2, b.ln1.params["gamma"], b.ln1.params["beta"], b.ln1.epsilon)
            qkv, _, _ = mxfp8.gemm(q, sf, w["qkv"][0], w["qkv"][1], bias=w["qkv"][2])
            o, _ = attn_ops.attention_fwd(qkv.reshape(B, L, 3, H, D), b.attn.is_causal, kv_cache=view, pos=pos)
            cache.len = view.len
            oq, osf = mxfp8.quantize(o.reshape(B * L, E))
            x2, _, _ = mxfp8.gemm(oq, osf, w["out"][0], w["out"][1], bias=w["out"][2], residual=x2)
            q, sf = mxfp8.layernorm_quant(x2, b.ln2.params["gamma"], b.ln2.params["beta"], b.ln2.epsilon)
            _, hq, hsf = mxfp8.gemm(q, sf, w["fc1"][0], w["fc1"][1], bias=w["fc1"][2], act=mxfp8.ACT_GELU,
                                    want_bf16=False, want_fp8=True)
            x2, _, _ = mxfp8.gemm(hq, hsf, w["fc2"][0], w["fc2"][1], bias=w["fc2"][2], residual=x2)
        return x2.reshape(B, L, E)

    # ------------------------------------------------------------------ building blocks
    def _embed(self, ids: torch.Tensor, pos: int) -> torch.Tensor:
        B, L = ids.shape
        table = self.pos.params["weight"][pos:pos + L]
        if self.cuda:
            return self.embed.forward_impl(ids, 0, pos_table=table.contiguous())
        return self.embed.params["weight"][ids.long()] + table

    def _mlp(self, b: _Block, x):
        y = b.ln2.forward_impl(x, 0)
        if self.cuda and b.act in ("relu", "gelu"):
            saved, b.fc1.activation = b.fc1.activation, b.act      # GELU in the fc1 epilogue
            try:
                h = b.fc1.forward_impl(y, 0)
            finally:
                b.fc1.activation = saved
        else:
            h = b.fc1.forward_impl(y, 0)
            h = ActivationLayer(b.act).init(self.device).set_training(False).forward(h, 0)
        return b.fc2.forward_impl(h, 0, residual=x)

    def _layers(self, x, pos: int):
        B = x.shape[0]
        for b, cache in zip(self.blocks, self.caches):
            view = attn_ops.KVCache.__new__(attn_ops.KVCache)
            view.k, view.v, view.len = cache.k[:B], cache.v[:B], cache.len
            y = b.ln1.forward_impl(x, 0)
            x = b.attn.forward_impl(y, 0, residual=x, kv_cache=view, pos=pos)
            cache.len = view.len
            x = self._mlp(b, x)
        return x

    def _logits_last(self, x_last: torch.Tensor) -> torch.Tensor:
        y = self.ln_f.forward_impl(x_last, 0)
        return self.head.forward_impl(y, 0)

    def _argmax(self, logits: torch.Tensor) -> torch.Tensor:
        rows = logits.reshape(-1, logits.shape[-1])
        if self.cuda:
            out = torch.empty(rows.shape[0], dtype=torch.int32, device=self.device)
            native.require().argmax_rows(rows, rows.stride(0), rows.shape[0], self.vocab, out)
            return out
        return rows[:, :self.vocab].argmax(-1).to(torch.int32)

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def prefill(self, ids: torch.Tensor, return_logits: bool = False):
        """ids [B, L] integer tokens.  Fills the KV caches; returns the greedy next token [B]
        (and the last-position logits [B, V] when asked)."""
        ids = ids.to(self.device).to(torch.int32)
        B, L = ids.shape
        if B > self.max_batch or L > self.max_len:
            raise ValueError(f"prompt [{B}, {L}] exceeds engine capacity [{self.max_batch}, {self.max_len}]")
        self.cur_batch, self.cur_len = B, 0
        with gemm_ops.static_weights():
            x = self._embed(ids, 0)
            x = self._layers_fp8(x, 0) if (self.precision == "fp8" and B * L >= 128) else self._layers(x, 0)
            self.cur_len = L
            logits = self._logits_last(x[:, -1:, :].contiguous()).reshape(B, -1)
            nxt = self._argmax(logits)
        return (nxt, logits) if return_logits else nxt

    @torch.no_grad()
    def decode(self, tokens: torch.Tensor, return_logits: bool = False):
        """tokens [B] = the tokens at position ``cur_len``; returns the next greedy token [B]."""
        B = self.cur_batch
        if self.cur_len >= self.max_len:
            raise ValueError("KV cache is full")
        ids = tokens.to(self.device).to(torch.int32).reshape(B, 1)
        with gemm_ops.static_weights():
            x = self._embed(ids, self.cur_len)
            x = self._layers(x, self.cur_len)
            self.cur_len += 1
            logits = self._logits_last(x).reshape(B, -1)
            nxt = self._argmax(logits)
        return (nxt, logits) if return_logits else nxt

    @torch.no_grad()
    def generate(self, prompt_ids: torch.Tensor, max_new_tokens: int = 50, eos_token: int | None = 50256) -> torch.Tensor:
        """Greedy generation (examples/gpt2_inference.cpp:71-121).  Returns [B, <= max_new_tokens] ids."""
        tok = self.prefill(prompt_ids)
        out = [tok]
        done = torch.zeros_like(tok, dtype=torch.bool)
        for _ in range(max_new_tokens - 1):
            if self.cur_len >= self.max_len:
                break
            if eos_token is not None:
                done |= tok == eos_token
                if bool(done.all()):
                    break
            tok = self.decode(tok)
            out.append(tok)
        return torch.stack(out, 1)

    # ------------------------------------------------------------------ CUDA-graph decode (GPU)
    def _decode_step_device(self):
        """One decode step whose position / tokens live in device memory (graph replayable)."""
        Cx = native.require()
        B, E = self.cur_batch, self.E
        st = self._dev
        x = torch.empty((B, E), dtype=torch.bfloat16, device=self.device)
        Cx.embed_decode(st["tok"], self.embed.params["weight"], self.pos.params["weight"], x, st["pos"])
        with gemm_ops.static_weights():
            for b, cache in zip(self.blocks, self.caches):
                y = b.ln1.forward_impl(x, 0)
                qkv = b.attn._qkv(y)                                         # [B, 3E]
                o = torch.empty((B, E), dtype=torch.bfloat16, device=self.device)
                Cx.attention_decode_append(qkv, cache.k[:B], cache.v[:B], o, st["pos"])
                x = b.attn._out(o, x)
                x = self._mlp(b, x)
            logits = self._logits_last(x).reshape(B, -1)
        nxt = self._argmax(logits)
        Cx.decode_epilogue(nxt, st["tok"], st["history"], st["pos"], st["step"])

    @torch.no_grad()
    def decode_graph(self, tokens: torch.Tensor, steps: int) -> torch.Tensor:
        """``steps`` greedy decode steps as replays of ONE captured CUDA graph: token ids and the
        position stay on the device between steps (no host round trip, no per-step launches from
        Python).  Returns the generated ids [B, steps]."""
        if not self.cuda:
            out, tok = [], tokens
            for _ in range(steps):
                tok = self.decode(tok)
                out.append(tok)
            return torch.stack(out, 1)
        B = self.cur_batch
        steps = min(int(steps), self.max_len - self.cur_len)
        dev = self.device
        if getattr(self, "_dev", None) is None or self._dev["tok"].numel() != B or self._dev["history"].shape[0] < steps:
            self._dev = dict(tok=torch.zeros(B, dtype=torch.int32, device=dev), pos=torch.zeros(1, dtype=torch.int32, device=dev),
                             step=torch.zeros(1, dtype=torch.int32, device=dev),
                             history=torch.zeros((max(steps, 64), B), dtype=torch.int32, device=dev))
            self._graph = None
        st = self._dev
        st["tok"].copy_(tokens.to(dev).to(torch.int32).reshape(B))
        st["pos"].fill_(self.cur_len)
        st["step"].zero_()
        if self._graph is None:
            self._decode_step_device()                                   # warm-up (allocator, plans) ...
            st["tok"].copy_(tokens.to(dev).to(torch.int32).reshape(B))   # ... then rewind its side effects
            st["pos"].fill_(self.cur_len)
            st["step"].zero_()
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._decode_step_device()
            self._graph = g
        for _ in range(steps):
            self._graph.replay()
        self.cur_len += steps
        for c in self.caches:
            c.len = self.cur_len
        return st["history"][:steps].t().contiguous()

    def full_forward_logits(self, ids: torch.Tensor) -> torch.Tensor:
        """Reference path: the whole model over the whole sequence, no cache (what the reference
        does per generated token).  Used by tests as the oracle for the cached path."""
        self.model.set_training(False)
        return self.model.forward(ids.to(self.device).to(torch.int32), 0)
