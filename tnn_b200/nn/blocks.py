"""Composite blocks: Sequential, MSequential, ResidualBlock, AttentionBlock, FlashAttentionBlock.

Reference: src/nn/blocks_impl/{sequential,msequential,residual_block,attention_block,
flash_attention_block}.cpp.  Differences by design: no stream synchronisation at
the end of a Sequential pass; BatchNorm+ReLU+Dropout, conv+residual-add(+ReLU) and
BN-backward+shortcut-gradient-add run as single kernels; Q/K/V are one fused GEMM.
"""
from __future__ import annotations

import os

import math

import torch

from ..core.config import TConfig
from ..ops import attention as attn_ops
from ..ops import gemm as gemm_ops
from ..ops import native
from ..ops import tensor_ops
from .layer import Layer, LayerFactory, ParamDescriptor, register
from .layers import (ActivationLayer, BatchNormLayer, Conv2DLayer, DenseLayer, DropoutLayer, LayerNormLayer, _aligned_rows,
                     _normal_init, _uniform_init, to_bf16, to_dtype)


# TNN_NVTX=1: one NVTX range per layer forward / backward (shows up in Nsight Systems and `ncu --nvtx`)
_NVTX = os.environ.get("TNN_NVTX", "0") not in ("", "0")

class Block(Layer):
    """A layer that owns child layers and forwards every setter to them (block.hpp:6-71)."""

    def set_training(self, flag):
        self.training = flag
        for c in self.children():
            c.set_training(flag)
        return self

    def set_seed(self, seed):
        self.seed = int(seed)
        for i, c in enumerate(self.children()):
            c.set_seed(seed)
        return self

    def set_io_dtype(self, dt):
        self.io_dtype = dt
        for c in self.children():
            c.set_io_dtype(dt)
        return self

    def init(self, device="cpu", seed=None):
        self.device = torch.device(device)
        if seed is not None:
            self.seed = int(seed)
        Layer.init(self, device, seed)
        return self

    def forward_flops(self, s):
        return 0


@register()
class Sequential(Block):
    TYPE_NAME = "sequential"

    def __init__(self, layers=None, name="sequential"):
        super().__init__(name)
        self.layers: list[Layer] = list(layers or [])
        self._analyze()

    def _analyze(self):
        # BatchNorm(+ReLU) directly followed by Dropout -> one fused apply kernel
        for a, b in zip(self.layers, self.layers[1:]):
            if isinstance(a, BatchNormLayer) and isinstance(b, DropoutLayer) and b.dropout_rate > 0:
                a.fused_dropout_p = b.dropout_rate
                b.fused_into_prev = True
        # BatchNorm(+ReLU) [+ fused Dropout] -> Conv2D: the conv's dgrad epilogue performs the BatchNorm's
        # backward reduction (the BN output has no other consumer inside a Sequential)
        for i, a in enumerate(self.layers):
            if not isinstance(a, BatchNormLayer):
                continue
            j = i + 1
            if j < len(self.layers) and isinstance(self.layers[j], DropoutLayer) and getattr(self.layers[j], "fused_into_prev", False):
                j += 1
            if j < len(self.layers) and isinstance(self.layers[j], Conv2DLayer) and not self.layers[j]._generic:
                self.layers[j]._bn_prev = a
                a._bwd_consumer = self.layers[j]

    def add(self, layer: Layer):
        self.layers.append(layer)
        self._analyze()
        return self

    def children(self):
        return self.layers

    def __len__(self):
        return len(self.layers)

    def __getitem__(self, i):
        return self.layers[i]

    def forward_impl(self, x, mb_id, last_residual=None, last_act=None):
        n = len(self.layers)
        for i, layer in enumerate(self.layers):
            if _NVTX and x.is_cuda:
                with torch.cuda.nvtx.range("fwd:" + layer.name):
                    x = self._forward_one(i, n, layer, x, mb_id, last_residual, last_act)
                continue
            x = self._forward_one(i, n, layer, x, mb_id, last_residual, last_act)
        return x

    @staticmethod
    def _forward_one(i, n, layer, x, mb_id, last_residual, last_act):
        if True:
            if i == n - 1 and (last_residual is not None or last_act is not None):
                x = layer.forward_impl(layer._migrate(x), mb_id, residual=last_residual, act=last_act) \
                    if isinstance(layer, Conv2DLayer) else layer.forward_impl(layer._migrate(x), mb_id, residual=last_residual)
                # (Dense / Conv2D add it in the GEMM epilogue, Dropout in its own kernel)
            else:
                x = layer.forward(x, mb_id)
        return x

    backward_hook = None        # optional callable(layer_index), fired after each layer's backward (DP overlap)

    def backward_impl(self, dy, mb_id, first_extra=None):
        hook = self.backward_hook
        for i in range(len(self.layers) - 1, -1, -1):
            layer = self.layers[i]
            if dy is None:
                if hook is not None:
                    for j in range(i, -1, -1):
                        hook(j)
                break
            if i == 0 and first_extra is not None:
                if isinstance(layer, (BatchNormLayer, LayerNormLayer)):
                    dy = layer.backward_impl(layer._migrate(dy), mb_id, extra=first_extra)
                    layer._cache.pop(mb_id, None)
                    first_extra = None
                    if hook is not None:
                        hook(i)
                    continue
            if _NVTX and dy.is_cuda:
                with torch.cuda.nvtx.range("bwd:" + layer.name):
                    dy = layer.backward(dy, mb_id)
            else:
                dy = layer.backward(dy, mb_id)
            if hook is not None:
                hook(i)
        if first_extra is not None and dy is not None:
            dy = _add(dy, first_extra)
        return dy

    def fuses_first_extra(self) -> bool:
        return len(self.layers) > 0 and isinstance(self.layers[0], (BatchNormLayer, LayerNormLayer))

    def fuses_last_residual(self) -> bool:
        return len(self.layers) > 0 and isinstance(self.layers[-1], (Conv2DLayer, DenseLayer, DropoutLayer)) and \
            not getattr(self.layers[-1], "_generic", False) and not getattr(self.layers[-1], "fused_into_prev", False)

    def output_shape(self, s):
        for layer in self.layers:
            s = layer.output_shape(s)
        return s

    def forward_flops(self, s):
        total = 0
        for layer in self.layers:
            total += layer.forward_flops(s)
            s = layer.output_shape(s)
        return total

    def get_config(self):
        c = TConfig(self.name, self.TYPE_NAME, layers=[l.get_config() for l in self.layers])
        return c

    @classmethod
    def create_from_config(cls, c):
        return cls([LayerFactory.create(lc) for lc in c.get("layers", [])], c.name)


def _add(a, b, relu=False):
    if a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.shape == b.shape:
        o = torch.empty_like(a)
        native.require().add_bf16(a.contiguous(), b.contiguous(), o, relu)
        return o
    if a.is_cuda and a.shape == b.shape and a.dtype in (torch.float16, torch.float32, torch.float64, torch.bfloat16):
        # other float dtypes (explicit io dtype): the templated tensor-op kernels, not a library add
        o = tensor_ops.add(a.contiguous(), to_dtype(b, a.dtype).contiguous())
        if relu:
            y = torch.empty_like(o)
            native.require().act_fwd(o, y, 1, 0.0)
            return y
        return o
    o = a + b
    return torch.relu(o) if relu else o


@register()
class ResidualBlock(Block):
    """act(F(x) + shortcut(x)) (residual_block.cpp:63-142)."""
    TYPE_NAME = "residual_block"

    def __init__(self, main_path: Sequential, shortcut_path: Sequential | None = None, activation="relu",
                 name="residual_block"):
        super().__init__(name)
        self.main = main_path
        self.shortcut = shortcut_path
        self.activation = (activation or "linear").lower()

    def children(self):
        return [self.main] + ([self.shortcut] if self.shortcut is not None else [])

    def forward_impl(self, x, mb_id):
        short = self.shortcut.forward(x, mb_id) if self.shortcut is not None else x
        act = self.activation
        last = self.main.layers[-1] if len(self.main.layers) else None
        fuse = self.main.fuses_last_residual() and (
            (isinstance(last, Conv2DLayer) and act in ("linear", "none", "relu")) or
            (isinstance(last, (DenseLayer, DropoutLayer)) and act in ("linear", "none") and x.is_cuda and
             short.shape == self.main.output_shape(tuple(x.shape))))
        if fuse:
            # Conv2D / Dense add the shortcut in the GEMM epilogue, Dropout in its own kernel
            y = self.main.forward_impl(x, mb_id, last_residual=short,
                                       last_act="relu" if act == "relu" else None)
        else:
            y = _add(self.main.forward(x, mb_id), short, relu=(act == "relu"))
            if act not in ("linear", "none", "relu"):
                self._act = getattr(self, "_act", None) or ActivationLayer(act, self.name + "_act")
                self._act.training = self.training
                y = self._act.forward(y, mb_id)
        if self.training and act == "relu":
            self.cache(mb_id)["y"] = y
        return y

    def backward_impl(self, dy, mb_id):
        act = self.activation
        if act == "relu":
            y = self.cache(mb_id)["y"]
            if dy.is_cuda:
                g = torch.empty_like(y)
                native.require().act_bwd(y, dy.to(y.dtype).contiguous(), g, 1, 0.0)
                dy = g
            else:
                dy = dy * (y > 0).to(dy.dtype)
        elif act not in ("linear", "none"):
            dy = self._act.backward(dy, mb_id)
        d_short = self.shortcut.backward(dy, mb_id) if self.shortcut is not None else dy
        if d_short is None:
            return self.main.backward(dy, mb_id)
        return self.main.backward_impl(dy, mb_id, first_extra=d_short)

    def output_shape(self, s):
        return self.main.output_shape(s)

    def forward_flops(self, s):
        return self.main.forward_flops(s) + (self.shortcut.forward_flops(s) if self.shortcut is not None else 0)

    def get_config(self):
        c = TConfig(self.name, self.TYPE_NAME, activation=self.activation, main_path=self.main.get_config())
        # the reference writes an EMPTY json object for "no shortcut" (residual_block.cpp:156-161)
        c.set("shortcut_path", self.shortcut.get_config() if self.shortcut is not None else {})
        return c

    @classmethod
    def create_from_config(cls, c):
        main = LayerFactory.create(c["main_path"])
        sp = c.get("shortcut_path")
        sc = LayerFactory.create(sp) if isinstance(sp, TConfig) else None
        return cls(main, sc, c.get("activation", "relu"), c.name)


@register()
class MSequential(Block):
    """M parallel Sequential branches joined by an N-ary layer (msequential.cpp:40-163).  On the
    first call each branch's peak/retained activation memory is measured and branches then run in
    descending (peak - retained) order, which minimises the allocator high-water mark."""
    TYPE_NAME = "msequential"

    def __init__(self, sequences: list[Sequential], join_layer: Layer, name="msequential"):
        super().__init__(name)
        self.sequences = list(sequences)
        self.join = join_layer
        self._order = None

    def children(self):
        return self.sequences + [self.join]

    def _measure_order(self, x, mb_id):
        stats = []
        cuda = x.is_cuda
        for i, s in enumerate(self.sequences):
            if cuda:
                torch.cuda.synchronize()
                torch.cuda.reset_peak_memory_stats()
                base = torch.cuda.memory_allocated()
            out = s.forward(x, mb_id)
            if cuda:
                peak = torch.cuda.max_memory_allocated() - base
                retained = torch.cuda.memory_allocated() - base
            else:
                retained = out.numel() * out.element_size()
                peak = retained
            stats.append((peak - retained, i, out))
        self._order = [i for _, i, _ in sorted(stats, key=lambda t: -t[0])]
        outs = [None] * len(self.sequences)
        for _, i, o in stats:
            outs[i] = o
        return outs

    def forward_impl(self, x, mb_id):
        if self._order is None:
            outs = self._measure_order(x, mb_id)
        else:
            outs = [None] * len(self.sequences)
            for i in self._order:
                outs[i] = self.sequences[i].forward(x, mb_id)
        return self.join.forward(outs, mb_id)

    def backward_impl(self, dy, mb_id):
        gs = self.join.backward(dy, mb_id)
        if not isinstance(gs, (list, tuple)):
            gs = [gs] * len(self.sequences)
        total = None
        for s, g in zip(self.sequences, gs):
            d = s.backward(g, mb_id)
            if d is not None:
                total = d if total is None else _add(total, d)
        return total

    def output_shape(self, s):
        return self.sequences[0].output_shape(s)

    def get_config(self):
        return TConfig(self.name, self.TYPE_NAME, sequences=[s.get_config() for s in self.sequences],
                       join_layer=self.join.get_config())

    @classmethod
    def create_from_config(cls, c):
        return cls([LayerFactory.create(s) for s in c["sequences"]], LayerFactory.create(c["join_layer"]), c.name)


# =============================================================================== attention
@register()
class AttentionBlock(Block):
    """Multi-head self-attention [B, L, E] -> [B, L, E] (attention_block.cpp:42-297).
    One fused QKV GEMM, fused attention kernel (online softmax, in-register causal mask), output
    projection with bias.  Training backward recomputes the probabilities (as the reference does)."""
    TYPE_NAME = "attention_block"
    FLASH = False

    def __init__(self, embed_dim, num_heads, is_causal=True, name="attention_block"):
        super().__init__(name)
        self.embed_dim, self.num_heads, self.is_causal = int(embed_dim), int(num_heads), bool(is_causal)
        assert self.embed_dim % self.num_heads == 0
        self.head_dim = self.embed_dim // self.num_heads

    def param_descriptors(self):
        E = self.embed_dim
        b = 1.0 / math.sqrt(E)
        return [ParamDescriptor("qkv_weight", (3 * E, E), _uniform_init(b)),
                ParamDescriptor("qkv_bias", (3 * E,), _uniform_init(b)),
                ParamDescriptor("out_weight", (E, E), _uniform_init(b)),
                ParamDescriptor("out_bias", (E,), _uniform_init(b))]

    def forward_impl(self, x, mb_id, residual=None, kv_cache=None, pos: int = 0):
        B, L, E = x.shape
        H, D = self.num_heads, self.head_dim
        x2 = x.reshape(B * L, E)
        # io_dtype fp16 (what the reference's flash attention requires, flash_attention_block.cpp:96-97): projections on the
        # fp16 tcgen05 GEMM, the fused attention kernel with fp16 operands; default = bf16, the tuned path
        f16 = x.is_cuda and self.io_dtype == torch.float16 and D == 64 and kv_cache is None
        if f16:
            x2 = to_dtype(x2, torch.float16)
            qkv = gemm_ops.linear_fwd(x2, self.lp_as("qkv_weight", torch.float16), self.params["qkv_bias"])
        elif x.is_cuda:
            x2 = _aligned_rows(to_bf16(x2))
            qkv = gemm_ops.linear_fwd(x2, self.lp("qkv_weight"), self.params["qkv_bias"])
        else:
            qkv = gemm_ops.linear_fwd(x2.float(), self.params["qkv_weight"], self.params["qkv_bias"])
        qkv = qkv.reshape(B, L, 3, H, D)
        o, lse = attn_ops.attention_fwd(qkv, self.is_causal, kv_cache=kv_cache, pos=pos)   # [B, L, H*D]
        o2 = o.reshape(B * L, E)
        res2 = residual.reshape(B * L, E) if residual is not None else None
        if f16:
            y = gemm_ops.linear_fwd(o2, self.lp_as("out_weight", torch.float16), self.params["out_bias"],
                                    residual=to_dtype(res2, torch.float16) if res2 is not None else None)
        elif x.is_cuda:
            y = gemm_ops.linear_fwd(o2, self.lp("out_weight"), self.params["out_bias"], residual=res2)
        else:
            y = gemm_ops.linear_fwd(o2, self.params["out_weight"], self.params["out_bias"], residual=res2)
        if self.training:
            self.cache(mb_id).update(x=x2, qkv=qkv, o=o2, lse=lse)
        return y.reshape(B, L, E)

    # inference helpers (inference/gpt2.py): the two projections without the attention in between
    def _qkv(self, x2: torch.Tensor) -> torch.Tensor:
        if x2.is_cuda:
            return gemm_ops.linear_fwd(_aligned_rows(to_bf16(x2)), self.lp("qkv_weight"), self.params["qkv_bias"])
        return gemm_ops.linear_fwd(x2.float(), self.params["qkv_weight"], self.params["qkv_bias"])

    def _out(self, o2: torch.Tensor, residual: torch.Tensor | None = None) -> torch.Tensor:
        w = self.lp("out_weight") if o2.is_cuda else self.params["out_weight"]
        return gemm_ops.linear_fwd(o2, w, self.params["out_bias"], residual=residual)

    def backward_impl(self, dy, mb_id):
        c = self.cache(mb_id)
        x2, qkv, o2 = c["x"], c["qkv"], c["o"]
        B, L = qkv.shape[0], qkv.shape[1]
        E = self.embed_dim
        dy2 = dy.reshape(B * L, E)
        cuda = dy.is_cuda
        f16 = cuda and qkv.dtype == torch.float16
        if f16:
            dy2 = to_dtype(dy2, torch.float16).contiguous()
        elif cuda:
            dy2 = _aligned_rows(to_bf16(dy2))
        else:
            dy2 = dy2.float()

        def w(name):
            return self.lp_as(name, torch.float16) if f16 else (self.lp(name) if cuda else self.params[name])

        gemm_ops.linear_wgrad(dy2, o2, self.grads["out_weight"])
        gemm_ops.bias_grad(dy2, self.grads["out_bias"])
        do = gemm_ops.linear_dgrad(dy2, w("out_weight"))
        dqkv = attn_ops.attention_bwd(qkv, do.reshape(B, L, E), c["o"].reshape(B, L, E), c["lse"], self.is_causal)
        dqkv2 = dqkv.reshape(B * L, 3 * E)
        if cuda and not f16:
            dqkv2 = _aligned_rows(dqkv2)
        gemm_ops.linear_wgrad(dqkv2, x2, self.grads["qkv_weight"])
        gemm_ops.bias_grad(dqkv2, self.grads["qkv_bias"])
        dx = gemm_ops.linear_dgrad(dqkv2, w("qkv_weight"))
        return dx.reshape(B, L, E)

    def forward_flops(self, s):
        B, L, E = s
        return 2 * B * L * E * 4 * E + 4 * B * L * L * E // (2 if self.is_causal else 1)

    def config_params(self):
        return dict(embed_dim=self.embed_dim, num_heads=self.num_heads, is_causal=self.is_causal)

    def get_config(self):
        return Layer.get_config(self)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["embed_dim"], c["num_heads"], c.get("is_causal", True), c.name)


@register()
class FlashAttentionBlock(AttentionBlock):
    """Same math; the reference's flash variant forces fp16 through cuDNN SDPA (flash_attention_block.cpp:96-97).  Here both
    names run the same fused tcgen05 kernel: bf16 by default, fp16 operands end to end when ``set_io_dtype(torch.float16)``."""
    TYPE_NAME = "flash_attention_block"
    FLASH = True


@register()
class SDPALayer(Layer):
    """Scaled dot-product attention over pre-projected [B, L, 3, H, D] (sdpa_layer.cpp)."""
    TYPE_NAME = "sdpa"

    def __init__(self, is_causal=False, name="sdpa"):
        super().__init__(name)
        self.is_causal = bool(is_causal)

    def forward_impl(self, qkv, mb_id):
        o, lse = attn_ops.attention_fwd(qkv, self.is_causal)
        if self.training:
            self.cache(mb_id).update(qkv=qkv, o=o, lse=lse)
        return o

    def backward_impl(self, dy, mb_id):
        c = self.cache(mb_id)
        return attn_ops.attention_bwd(c["qkv"], dy, c["o"], c["lse"], self.is_causal)

    def config_params(self):
        return dict(is_causal=self.is_causal)

    @classmethod
    def create_from_config(cls, c):
        return cls(c.get("is_causal", False), c.name)
