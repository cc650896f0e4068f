"""Legacy (NCHW) model zoo and loaders (reference: include/nn/legacy/example_models.hpp:14-81,
src/nn/legacy/example_models.cpp:21-510, include/data_loading/legacy/*.hpp).

The reference keeps an older generation of its zoo whose image models take ``[N, C, H, W]`` tensors and are built from the
``legacy_*`` layers.  Here every legacy model is DERIVED from the current NHWC definition: the layer
configs of ``models/zoo.py`` are walked and each spatial layer is replaced by its ``legacy_*`` twin
(an NCHW facade over the same kernels, ``nn/layers._NCHWFacade``), so the two zoos cannot drift
apart.  The only structural edit is ViT patch flattening: ``[N,E,h,w]`` is flattened over (h, w) and
transposed to ``[N, h*w, E]`` where the NHWC model flattens dims (1, 2).

``legacy.DataLoaderFactory`` wraps the NHWC loaders and hands out NCHW batches.
"""
from __future__ import annotations

import copy
import os

import torch

from ..core.config import TConfig
from ..nn.blocks import Sequential
from ..nn.graph import Graph, graph_from_sequential
from ..nn.layer import LayerFactory
from . import zoo

_TWIN = {"conv2d": "legacy_conv2d", "batchnorm": "legacy_batchnorm", "maxpool2d": "legacy_maxpool2d",
         "avgpool2d": "legacy_avgpool2d", "dense": "legacy_dense"}


def _val(d, key, default=None):
    v = d.get("parameters", {}).get(key, default)
    return v.get("value") if isinstance(v, dict) and "value" in v else v


def _convert(node):
    """Recursively renames layer types inside a TConfig JSON tree; returns the converted copy."""
    if isinstance(node, list):
        out = []
        for i, item in enumerate(node):
            conv = _convert(item)
            prev = out[-1] if out else None
            if (isinstance(conv, dict) and conv.get("type") == "flatten" and isinstance(prev, dict)
                    and prev.get("type") == "legacy_conv2d" and _val(conv, "start_dim") == 1 and _val(conv, "end_dim") == 2):
                # NHWC patch flatten (H,W) -> NCHW: flatten (H,W) = dims 2..3, then [N,E,T] -> [N,T,E]
                conv["parameters"]["start_dim"]["value"] = 2
                conv["parameters"]["end_dim"]["value"] = 3
                out.append(conv)
                out.append({"name": conv["name"] + "_to_tokens", "type": "transpose", "parameters": {}})
                continue
            out.append(conv)
        return out
    if isinstance(node, dict):
        node = {k: _convert(v) for k, v in node.items()}
        if "parameters" in node and node.get("type") in _TWIN:
            node["type"] = _TWIN[node["type"]]
        return node
    return node


def to_legacy(seq: Sequential) -> Sequential:
    """NHWC Sequential -> the same network on NCHW tensors built from legacy_* layers."""
    LayerFactory.register_defaults()
    cfg = _convert(copy.deepcopy(seq.get_config().to_json()))
    return LayerFactory.create(TConfig.from_json(cfg))


_NAMES = ["mnist_cnn", "cifar10_vgg", "cifar10_resnet9", "cifar100_resnet18", "cifar100_wrn16_8", "tiny_imagenet_resnet18",
          "tiny_imagenet_wrn16_8", "tiny_imagenet_resnet50", "tiny_imagenet_vit", "imagenet_resnet50", "gpt2_small",
          "flash_gpt2_small", "gpt2_medium", "flash_gpt2_medium", "gpt2_large", "flash_gpt2_large"]


class ExampleModels:
    """legacy::ExampleModels (example_models.hpp:16-44): same registry API, NCHW models."""
    _creators: dict = {}

    @classmethod
    def register_model(cls, name, fn):
        cls._creators[name] = fn

    @classmethod
    def register_defaults(cls):
        for n in _NAMES:
            cls._creators[n] = (lambda io_dtype=None, _n=n, **kw: to_legacy(zoo.ExampleModels.create(_n, io_dtype, **kw)))

    @classmethod
    def create(cls, name, io_dtype=None, **kw) -> Sequential:
        if not cls._creators:
            cls.register_defaults()
        if name not in cls._creators:
            raise KeyError(f"unknown legacy model {name!r}; available: {cls.available_models()}")
        return cls._creators[name](io_dtype, **kw)

    @classmethod
    def available_models(cls):
        if not cls._creators:
            cls.register_defaults()
        return sorted(cls._creators)


def load_or_create_model(name, path=None, device="cpu", seed=None, alloc_fn=None, **kw) -> Graph:
    if path and os.path.exists(path):
        return Graph.load_state(path, device, alloc_fn)
    return graph_from_sequential(ExampleModels.create(name, **kw), device, seed, name, alloc_fn)


class _NCHWLoader:
    """Presents an NHWC loader's image batches as NCHW (data_loading/legacy/*: the CHW record order
    of the CIFAR/MNIST files is what those loaders return unchanged)."""

    def __init__(self, inner):
        self.inner = inner

    def get_batch(self, batch_size):
        b = self.inner.get_batch(batch_size)
        if b is None:
            return None
        x, y = b
        return (x.permute(0, 3, 1, 2).contiguous() if x.dim() == 4 else x), y

    def __getattr__(self, k):
        return getattr(self.inner, k)

    def __len__(self):
        return len(self.inner)


class DataLoaderFactory:
    """legacy::DataLoaderFactory::create(name, path) -> (train, val) in NCHW."""

    @staticmethod
    def create(name, path, **kw):
        from ..data import DataLoaderFactory as F
        train, val = F.create(name, path, **kw)
        return _NCHWLoader(train), (_NCHWLoader(val) if val is not None else None)

    @staticmethod
    def wrap(loader):
        return _NCHWLoader(loader)
