"""Layer core (reference: include/nn/layer.hpp:23-144, src/nn/layer.cpp:25-230, siso_layer.hpp,
include/nn/layers.hpp:95-167 LayerFactory).

Canonical API generation (SURVEY F2 b+c): ``forward(inputs, mb_id) -> outputs``,
``backward(grads, mb_id) -> input grads``; activations needed by backward live in a
per-microbatch cache keyed by ``mb_id`` and are released when that microbatch's backward ran.
Parameters are fp32 (master) with fp32 gradients that ACCUMULATE until ``zero_grads``; on a GPU
every parameter also has a bf16 compute copy that the optimizer refreshes in the same kernel.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import torch

from ..core import side_stream
from ..core.config import TConfig
from ..core.dtype import DType, dtype_to_string, string_to_dtype, to_torch


@dataclass
class ParamDescriptor:
    name: str
    shape: tuple
    init: Callable[[torch.Tensor, torch.Generator], None] | None = None
    trainable: bool = True          # False: buffer (BN running stats) -- serialised, never optimised
    dtype: torch.dtype = torch.float32


class Layer:
    TYPE_NAME = "layer"

    def __init__(self, name: str = ""):
        self.name = name or self.TYPE_NAME
        self.training = True
        self.seed = 0
        self.io_dtype: torch.dtype | None = None      # None: fp32 on CPU, bf16 on CUDA
        self.param_dtype = torch.float32
        self.compute_dtype = torch.float32
        self.device = torch.device("cpu")
        self.params: dict[str, torch.Tensor] = {}
        self.grads: dict[str, torch.Tensor] = {}
        self.params_lp: dict[str, torch.Tensor] = {}   # bf16 compute copies (GPU)
        self._lp_version: dict[str, int] = {}
        self.param_version = 0                        # bumped whenever params change
        self._cache: dict[int, dict] = {}
        self._initialized = False
        self._step = 0

    # ------------------------------------------------------------------ configuration
    def set_training(self, flag: bool):
        self.training = flag
        return self

    def set_seed(self, seed: int):
        self.seed = int(seed)
        return self

    def set_io_dtype(self, dt):
        """Activation dtype contract of this layer and every child (layer.hpp:65-70).  None = fp32 on CPU, bf16 on CUDA."""
        self.io_dtype = dt
        for c in self.children():
            c.set_io_dtype(dt)
        return self

    def children(self) -> list["Layer"]:
        return []

    def all_layers(self) -> list["Layer"]:
        out = [self]
        for c in self.children():
            out += c.all_layers()
        return out

    def act_dtype(self, device=None) -> torch.dtype:
        device = device or self.device
        if self.io_dtype is not None:
            return self.io_dtype
        return torch.bfloat16 if device.type == "cuda" else torch.float32

    # ------------------------------------------------------------------ parameters
    def param_descriptors(self) -> list[ParamDescriptor]:
        return []

    def init(self, device: torch.device | str = "cpu", seed: int | None = None):
        """Allocate + initialise own parameters (children included)."""
        self.device = torch.device(device)
        if seed is not None:
            self.seed = int(seed)
        gen = torch.Generator(device="cpu")
        gen.manual_seed(self.seed)
        for d in self.param_descriptors():
            t = torch.zeros(d.shape, dtype=d.dtype)
            if d.init is not None:
                d.init(t, gen)
            self.params[d.name] = t.to(self.device)
            if d.trainable:
                self.grads[d.name] = torch.zeros(d.shape, dtype=torch.float32, device=self.device)
        for c in self.children():
            c.init(self.device, self.seed if seed is None else seed)
        self._initialized = True
        self.mark_params_dirty()
        return self

    def to(self, device):
        device = torch.device(device)
        self.device = device
        for k in list(self.params):
            self.params[k] = self.params[k].to(device)
        for k in list(self.grads):
            self.grads[k] = self.grads[k].to(device)
        self.params_lp.clear()
        self._lp_version.clear()
        for c in self.children():
            c.to(device)
        self.mark_params_dirty()
        return self

    def rebind(self, name: str, param: torch.Tensor, grad: torch.Tensor | None, lp: torch.Tensor | None):
        """Re-home a parameter into slab views (GraphContext)."""
        self.params[name] = param
        if grad is not None:
            self.grads[name] = grad
        if lp is not None:
            self.params_lp[name] = lp
            self._lp_version[name] = -1
        self._slab_lp = lp is not None or getattr(self, "_slab_lp", False)

    def mark_params_dirty(self):
        self.param_version += 1
        for c in self.children():
            c.mark_params_dirty()

    def mark_lp_fresh(self):
        """Called by the optimizer after its kernel refreshed the bf16 slab."""
        self.param_version += 1
        for k in self.params_lp:
            self._lp_version[k] = self.param_version
        for c in self.children():
            c.mark_lp_fresh()

    def lp(self, name: str) -> torch.Tensor:
        """bf16 compute copy of a parameter (refreshed lazily when stale)."""
        p = self.params[name]
        if not p.is_cuda:
            return p
        t = self.params_lp.get(name)
        if t is None:
            t = torch.empty(p.shape, dtype=torch.bfloat16, device=p.device)
            self.params_lp[name] = t
            self._lp_version[name] = -1
        if self._lp_version.get(name, -1) != self.param_version:
            from ..ops import native
            native.require().cast_f32_to_bf16(p, t)
            self._lp_version[name] = self.param_version
        return t

    def lp_as(self, name: str, dtype: torch.dtype) -> torch.Tensor:
        """Compute copy of a parameter in an explicit activation dtype (layers whose io_dtype is fp16 / fp64; fp32 is the
        master itself, bf16 the default copy).  Same lazy refresh as ``lp``."""
        p = self.params[name]
        if dtype == p.dtype or not p.is_cuda:
            return p
        if dtype == torch.bfloat16:
            return self.lp(name)
        key = f"{name}@{dtype}"
        t = self.params_lp.get(key)
        if t is None:
            t = torch.empty(p.shape, dtype=dtype, device=p.device)
            self.params_lp[key] = t
            self._lp_version[key] = -1
        if self._lp_version.get(key, -1) != self.param_version:
            from ..ops import native
            native.require().cast_any(p, t)
            self._lp_version[key] = self.param_version
        return t

    def parameters(self) -> list[tuple[str, torch.Tensor]]:
        out = [(f"{self.name}.{k}", v) for k, v in self.params.items()]
        for c in self.children():
            out += c.parameters()
        return out

    def zero_grads(self):
        for g in self.grads.values():
            g.zero_()
        for c in self.children():
            c.zero_grads()

    # ------------------------------------------------------------------ execution
    def _foreign_device(self) -> bool:
        """True when this layer lives on a CUDA device that is not the calling thread's current one (the hand-written kernels
        launch on the current device's stream, so the call is wrapped in a device guard)."""
        d = self.device
        return d.type == "cuda" and d.index is not None and torch.cuda.current_device() != d.index

    def forward(self, inputs, mb_id: int = 0):
        if self._foreign_device():
            with torch.cuda.device(self.device):
                return self.forward(inputs, mb_id)
        if isinstance(inputs, (list, tuple)):
            inputs = [self._migrate(t) for t in inputs]
            if len(inputs) == 1 and not self.MULTI_INPUT:
                return self.forward_impl(inputs[0], mb_id)
            return self.forward_impl(inputs, mb_id)
        return self.forward_impl(self._migrate(inputs), mb_id)

    def backward(self, grads, mb_id: int = 0):
        if self._foreign_device():
            with torch.cuda.device(self.device):
                return self.backward(grads, mb_id)
        if isinstance(grads, (list, tuple)) and len(grads) == 1 and not self.MULTI_OUTPUT:
            grads = grads[0]
        if isinstance(grads, (list, tuple)):
            grads = [self._migrate(g) for g in grads]
        else:
            grads = self._migrate(grads)
        side_stream.enter()
        try:
            out = self.backward_impl(grads, mb_id)
        finally:
            side_stream.exit_()          # outermost backward: join the side-stream weight gradients
        self._cache.pop(mb_id, None)
        return out

    MULTI_INPUT = False
    MULTI_OUTPUT = False

    def _migrate(self, t: torch.Tensor) -> torch.Tensor:
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        return t

    def cache(self, mb_id: int) -> dict:
        return self._cache.setdefault(mb_id, {})

    def clear_cache(self):
        self._cache.clear()
        for c in self.children():
            c.clear_cache()

    def forward_impl(self, x, mb_id):
        raise NotImplementedError

    def backward_impl(self, dy, mb_id):
        raise NotImplementedError

    def output_shape(self, input_shape: tuple) -> tuple:
        return tuple(input_shape)

    def forward_flops(self, input_shape: tuple) -> int:
        return 0

    # ------------------------------------------------------------------ config / (de)serialisation
    def config_params(self) -> dict:
        return {}

    def get_config(self) -> TConfig:
        c = TConfig(self.name, self.TYPE_NAME, **self.config_params())
        if self.io_dtype is not None:
            c.set("io_dtype", dtype_to_string(self.io_dtype))
        return c

    @classmethod
    def create_from_config(cls, cfg: TConfig) -> "Layer":
        raise NotImplementedError

    def __repr__(self):
        return f"{type(self).__name__}({self.name})"


class LayerFactory:
    """String-keyed registry (reference: include/nn/layers.hpp:95-167)."""
    _creators: dict[str, Callable[[TConfig], Layer]] = {}

    @classmethod
    def register(cls, type_name: str, creator: Callable[[TConfig], Layer]):
        cls._creators[type_name] = creator

    @classmethod
    def register_layer(cls, layer_cls, *aliases):
        cls._creators[layer_cls.TYPE_NAME] = layer_cls.create_from_config
        for a in aliases:
            cls._creators[a] = layer_cls.create_from_config
        return layer_cls

    @classmethod
    def create(cls, cfg: TConfig | dict | str) -> Layer:
        if not isinstance(cfg, TConfig):
            cfg = TConfig.from_json(cfg)
        if cfg.type not in cls._creators:
            cls.register_defaults()
        if cfg.type not in cls._creators:
            raise KeyError(f"unknown layer type {cfg.type!r}; known: {sorted(cls._creators)}")
        layer = cls._creators[cfg.type](cfg)
        if cfg.has("io_dtype"):
            layer.io_dtype = to_torch(string_to_dtype(cfg.get("io_dtype")))
        return layer

    @classmethod
    def register_defaults(cls):
        from . import blocks, layers  # noqa: F401  (registration side effects)

    @classmethod
    def available(cls) -> list[str]:
        cls.register_defaults()
        return sorted(cls._creators)


def register(*aliases):
    def deco(layer_cls):
        return LayerFactory.register_layer(layer_cls, *aliases)
    return deco


# ----------------------------------------------------------------------------------------------- extension bases
class SISOLayer(Layer):
    """Single-input / single-output layer (reference: include/nn/siso_layer.hpp:9-34).  The base ``Layer`` already
    unwraps one-element lists; this class makes the contract explicit for user-defined layers: a list of more than
    one tensor is an error, and the shape rule is ``compute_output_shape`` (``output_shapes`` is the list form)."""

    def forward(self, inputs, mb_id: int = 0):
        if isinstance(inputs, (list, tuple)) and len(inputs) != 1:
            raise RuntimeError("SISOLayer only supports single input and single output")
        return super().forward(inputs, mb_id)

    def backward(self, grads, mb_id: int = 0):
        if isinstance(grads, (list, tuple)) and len(grads) != 1:
            raise RuntimeError("SISOLayer only supports single grad output and single grad input")
        return super().backward(grads, mb_id)

    def compute_output_shape(self, input_shape):
        return tuple(input_shape)

    def output_shape(self, s):
        return tuple(self.compute_output_shape(tuple(s)))

    def output_shapes(self, input_shapes):
        if len(input_shapes) != 1:
            raise RuntimeError("SISOLayer expects exactly one input shape")
        return [self.output_shape(input_shapes[0])]


class StatelessLayer(SISOLayer):
    """A SISO layer without parameters (layers_impl/stateless_layer.hpp:16-22): activations, shape ops, ..."""

    def param_descriptors(self):
        return []


class ParameterizedLayer(SISOLayer):
    """A SISO layer that owns parameters (layers_impl/parameterized_layer.hpp:15-22): subclasses list them in
    ``param_descriptors`` (name, shape, initialiser) and the base allocates, initialises, slabs and serialises them."""

    def param_descriptors(self):
        raise NotImplementedError(f"{type(self).__name__} must describe its parameters")
