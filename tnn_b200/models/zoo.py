"""ExampleModels zoo: the 17 named models of the reference (src/nn/example_models.cpp:21-559) built
with the LayerBuilder DSL, plus ``load_or_create_model`` (include/nn/example_models.hpp:49-84)."""
from __future__ import annotations

import os
from typing import Callable

import torch

from ..nn.blocks import Sequential
from ..nn.builder import LayerBuilder
from ..nn.graph import Graph, graph_from_sequential


def _b(shape, io_dtype):
    b = LayerBuilder([list(shape)])
    if io_dtype is not None:
        b.dtype(io_dtype)
    return b


def mnist_cnn(io_dtype=None):
    l = (_b((28, 28, 1), io_dtype).conv2d(8, 5, 5, 1, 1, 0, 0, False, "conv1").batchnorm(None, 0.1, True, True, "bn1")
         .maxpool2d(3, 3, 3, 3, 0, 0, "pool1").conv2d(16, 1, 1, 1, 1, 0, 0, False, "conv2_1x1")
         .batchnorm(None, 0.1, True, True, "bn2_1x1").activation("relu", "relu2")
         .conv2d(48, 5, 5, 1, 1, 0, 0, False, "conv3").batchnorm(None, 0.1, True, True, "bn3")
         .maxpool2d(2, 2, 2, 2, 0, 0, "pool2").flatten(1, -1, "flatten").dense(10, False, "output").build())
    return Sequential(l, "mnist_cnn")


def cifar10_vgg(io_dtype=None):
    b = _b((32, 32, 3), io_dtype)
    plan = [(64, "0", True), (64, "1", True), "p0", (128, "2", True), (128, "3", True), "p1", (256, "4", True),
            (256, "5", False), (256, "6", True), "p2", (512, "7", True), (512, "8", True), (512, "9", True), "p3"]
    for it in plan:
        if isinstance(it, str):
            b.maxpool2d(2, 2, 2, 2, 0, 0, "pool" + it[1:])
        else:
            ch, i, bn = it
            b.conv2d(ch, 3, 3, 1, 1, 1, 1, False, "conv" + i)
            if bn:
                b.batchnorm(None, 0.1, True, True, "bn" + i)
            else:
                b.activation("relu", "relu" + i)
    b.flatten(1, -1, "flatten").dense(512, True, "fc0").activation("relu", "relu10").dense(10, True, "fc1")
    return Sequential(b.build(), "cifar10_vgg")


def cifar10_resnet9(io_dtype=None):
    b = _b((32, 32, 3), io_dtype)
    (b.conv2d(64, 3, 3, 1, 1, 1, 1, False, "conv1").batchnorm(None, 0.1, True, True, "bn1")
     .conv2d(128, 3, 3, 1, 1, 1, 1, False, "conv2").batchnorm(None, 0.1, True, True, "bn2")
     .maxpool2d(2, 2, 2, 2, 0, 0, "pool1")
     .basic_residual_block(128, 128, 1, "res_block1").basic_residual_block(128, 128, 1, "res_block2")
     .conv2d(256, 3, 3, 1, 1, 1, 1, False, "conv3").batchnorm(None, 0.1, True, True, "bn3")
     .maxpool2d(2, 2, 2, 2, 0, 0, "pool2")
     .basic_residual_block(256, 256, 1, "res_block3").basic_residual_block(256, 256, 1, "res_block4")
     .conv2d(512, 3, 3, 1, 1, 1, 1, False, "conv4").batchnorm(None, 0.1, True, True, "bn4")
     .maxpool2d(2, 2, 2, 2, 0, 0, "pool3").basic_residual_block(512, 512, 1, "res_block5")
     .avgpool2d(4, 4, 1, 1, 0, 0, "avgpool").flatten(1, -1, "flatten").dense(10, True, "output"))
    return Sequential(b.build(), "cifar10_resnet9")


def _resnet18(shape, classes, pool, name, io_dtype):
    b = _b(shape, io_dtype)
    b.conv2d(32, 3, 3, 1, 1, 1, 1, False, "conv1").batchnorm(None, 0.1, True, True, "bn1").maxpool2d(2, 2, 2, 2, 0, 0, "maxpool")
    spec = [(32, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2), (512, 512, 1)]
    for i, (ci, co, s) in enumerate(spec):
        b.basic_residual_block(ci, co, s, f"layer{i // 2 + 1}_block{i % 2 + 1}")
    b.avgpool2d(pool, pool, 1, 1, 0, 0, "avgpool").flatten(1, -1, "flatten").dense(classes, True, "fc")
    return Sequential(b.build(), name)


def cifar100_resnet18(io_dtype=None):
    return _resnet18((32, 32, 3), 100, 2, "cifar100_resnet18", io_dtype)


def tiny_imagenet_resnet18(io_dtype=None):
    return _resnet18((64, 64, 3), 200, 4, "tiny_imagenet_resnet18", io_dtype)


def _wrn16_8(shape, classes, name, io_dtype, dropout=0.3):
    c1, c2, c3 = 128, 256, 512
    b = _b(shape, io_dtype)
    b.conv2d(16, 3, 3, 1, 1, 1, 1, True, "conv1")
    for i, (ci, co, s) in enumerate([(16, c1, 1), (c1, c1, 1), (c1, c2, 2), (c2, c2, 1), (c2, c3, 2), (c3, c3, 1)]):
        b.wide_residual_block(ci, co, s, dropout, f"group{i // 2 + 1}_block{i % 2 + 1}")
    b.batchnorm(None, 0.1, True, True, "bn_final").avgpool2d(8, 8, 1, 1, 0, 0, "avgpool").flatten(1, -1, "flatten")
    b.dense(classes, True, "fc")
    return Sequential(b.build(), name)


def cifar100_wrn16_8(io_dtype=None):
    return _wrn16_8((32, 32, 3), 100, "cifar100_wrn16_8", io_dtype)


def tiny_imagenet_wrn16_8(io_dtype=None):
    return _wrn16_8((64, 64, 3), 200, "tiny_imagenet_wrn16_8", io_dtype)


def _resnet50(shape, classes, stem7, pool, name, io_dtype):
    b = _b(shape, io_dtype)
    if stem7:
        b.conv2d(64, 7, 7, 2, 2, 3, 3, True, "conv1")
    else:
        b.conv2d(64, 3, 3, 1, 1, 1, 1, True, "conv1")
    b.batchnorm(None, 0.1, True, True, "bn1").maxpool2d(3, 3, 2, 2, 1, 1, "maxpool")
    cfg = [(64, 64, 256, 1, 3), (256, 128, 512, 2, 4), (512, 256, 1024, 2, 6), (1024, 512, 2048, 2, 3)]
    for li, (cin, mid, cout, stride, n) in enumerate(cfg):
        for j in range(n):
            b.bottleneck_residual_block(cin if j == 0 else cout, mid, cout, stride if j == 0 else 1, f"layer{li + 1}_block{j + 1}")
    b.avgpool2d(pool, pool, 1, 1, 0, 0, "avgpool").flatten(1, -1, "flatten").dense(classes, True, "fc")
    return Sequential(b.build(), name)


def tiny_imagenet_resnet50(io_dtype=None):
    return _resnet50((64, 64, 3), 200, False, 4, "tiny_imagenet_resnet50", io_dtype)


def imagenet_resnet50(io_dtype=None):
    return _resnet50((224, 224, 3), 1000, True, 7, "imagenet_resnet50", io_dtype)


def _vit(flash, name, io_dtype):
    patch, E, H, ratio, depth, classes = 4, 256, 4, 4, 4, 200
    seq = (64 // patch) ** 2 + 1
    b = _b((64, 64, 3), io_dtype)
    (b.conv2d(E, patch, patch, patch, patch, 0, 0, True, "patch_embed").flatten(1, 2, "flatten_patches")
     .class_token(E, "cls").positional_embedding(E, seq, "pos_embed").dropout(0.1, "embed_dropout"))
    for i in range(depth):
        ab = _b((seq, E), io_dtype).layernorm(1e-5, True, f"enc{i}_ln_attn")
        (ab.flash_attention if flash else ab.attention)(E, H, False, f"enc{i}_attn")
        ab.dropout(0.1, f"enc{i}_attn_dropout")
        b.residual_block(ab.build(), None, "linear", f"encoder_{i}_attn")
        mb = (_b((seq, E), io_dtype).layernorm(1e-5, True, f"enc{i}_ln_mlp").dense(E * ratio, False, f"enc{i}_fc1")
              .activation("gelu", f"enc{i}_gelu").dropout(0.1, f"enc{i}_d1").dense(E, False, f"enc{i}_fc2").dropout(0.1, f"enc{i}_d2"))
        b.residual_block(mb.build(), None, "linear", f"encoder_{i}_mlp")
    b.layernorm(1e-5, True, "ln_final").slice(1, 0, 1, "extract_cls").flatten(1, -1, "flatten_cls").dense(classes, True, "head")
    return Sequential(b.build(), name)


def tiny_imagenet_vit(io_dtype=None):
    return _vit(False, "tiny_imagenet_vit", io_dtype)


def tiny_imagenet_flash_vit(io_dtype=None):
    return _vit(True, "tiny_imagenet_flash_vit", io_dtype)


def _gpt2(E, H, layers, flash, name, io_dtype, seq_len=1024, vocab=50257, dropout=0.1):
    b = _b((seq_len,), io_dtype)
    b.embedding(vocab, E, -1, "token_embed").positional_embedding(E, seq_len, "pos_embed")
    if dropout > 0:
        b.dropout(dropout, "embed_dropout")
    for i in range(layers):
        b.gpt_block(E, H, 4 * E, dropout, True, "gelu", f"h{i}", flash=flash)
    b.layernorm(1e-5, True, "ln_f").dense(vocab, True, "head")
    return Sequential(b.build(), name)


def gpt2_small(io_dtype=None, **kw):
    return _gpt2(768, 12, 12, False, "gpt2_small", io_dtype, **kw)


def flash_gpt2_small(io_dtype=None, **kw):
    return _gpt2(768, 12, 12, True, "flash_gpt2_small", io_dtype, **kw)


def gpt2_medium(io_dtype=None, **kw):
    return _gpt2(1024, 16, 24, False, "gpt2_medium", io_dtype, **kw)


def flash_gpt2_medium(io_dtype=None, **kw):
    return _gpt2(1024, 16, 24, True, "flash_gpt2_medium", io_dtype, **kw)


def gpt2_large(io_dtype=None, **kw):
    return _gpt2(1280, 20, 36, False, "gpt2_large", io_dtype, **kw)


def flash_gpt2_large(io_dtype=None, **kw):
    return _gpt2(1280, 20, 36, True, "flash_gpt2_large", io_dtype, **kw)


class ExampleModels:
    _creators: dict[str, Callable] = {}

    @classmethod
    def register_model(cls, name: str, fn: Callable):
        cls._creators[name] = fn

    @classmethod
    def register_defaults(cls):
        for fn in (mnist_cnn, cifar10_vgg, cifar10_resnet9, cifar100_resnet18, cifar100_wrn16_8, tiny_imagenet_resnet18,
                   tiny_imagenet_wrn16_8, tiny_imagenet_resnet50, tiny_imagenet_vit, tiny_imagenet_flash_vit,
                   imagenet_resnet50, gpt2_small, flash_gpt2_small, gpt2_medium, flash_gpt2_medium, gpt2_large,
                   flash_gpt2_large):
            cls._creators[fn.__name__] = fn
        cls._creators["gpt2"] = gpt2_small

    @classmethod
    def create(cls, name: str, io_dtype=None, **kw) -> Sequential:
        if not cls._creators:
            cls.register_defaults()
        if name not in cls._creators:
            raise KeyError(f"unknown model {name!r}; available: {cls.available_models()}")
        return cls._creators[name](io_dtype, **kw)

    @classmethod
    def available_models(cls) -> list[str]:
        if not cls._creators:
            cls.register_defaults()
        return sorted(cls._creators)


def load_or_create_model(name: str, path: str | None = None, device="cpu", seed: int | None = None, alloc_fn=None,
                         **kw) -> Graph:
    """Loads ``path`` when it exists, else builds the named zoo model as input->model->output."""
    if path and os.path.exists(path):
        return Graph.load_state(path, device, alloc_fn)
    seq = ExampleModels.create(name, **kw)
    return graph_from_sequential(seq, device, seed, name, alloc_fn)
