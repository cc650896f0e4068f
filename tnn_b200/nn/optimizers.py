"""Optimizers (reference: include/nn/optimizers.hpp:28-280, cuda/adam_kernels.cu, sgd_kernels.cu).

``attach(GraphContext)`` binds to the contiguous parameter / gradient slabs; ``update()`` is ONE
kernel launch over the whole slab that also refreshes the bf16 compute copies and zeroes the
gradients (the reference launches one kernel per parameter tensor plus a memset).
Semantics match the reference: Adam m,v EMA in fp32, bias correction, coupled weight decay adds
``wd*lr*p`` to the update, decoupled subtracts ``wd*lr*p`` first; SGD ``v = mu*v - lr*g; p += v``.
"""
from __future__ import annotations

import torch

from ..core import stepstate
from ..core.config import TConfig
from ..ops import native


class Optimizer:
    NAME = "optimizer"

    def __init__(self, lr: float):
        self.lr = float(lr)
        self.ctx = None
        self.step_count = 0
        self.grad_scale = 1.0
        self.fused_zero_grad = True
        self.external_state = False     # True: the caller (TrainStep) already pushed the step state

    def _push_state(self, device):
        if not self.external_state:
            stepstate.get(device).advance(self)
        return stepstate.get(device).dev

    def attach(self, ctx):
        self.ctx = ctx
        self._alloc_state()
        return self

    def _alloc_state(self):
        pass

    def set_learning_rate(self, lr: float):
        self.lr = float(lr)

    def get_learning_rate(self) -> float:
        return self.lr

    def zero_grads(self):
        if self.ctx is not None:
            self.ctx.zero_grads()

    def update(self):
        raise NotImplementedError

    def state_dict(self) -> dict:
        return {"step": self.step_count, "lr": self.lr}

    def load_state_dict(self, d: dict):
        self.step_count = int(d.get("step", 0))
        self.lr = float(d.get("lr", self.lr))

    def get_config(self) -> TConfig:
        raise NotImplementedError

    def clone(self) -> "Optimizer":
        return OptimizerFactory.create_from_config(self.get_config())

    @staticmethod
    def _cpu_native(*ts) -> bool:
        """Flat fp32 CPU slabs -> the fused one-pass native step (csrc/runtime/cpu_ops.hpp::adam_step / sgd_step);
        TNN_CPU_OPT_TORCH=1 keeps the op-by-op torch formulation (the oracle the tests compare against)."""
        import os
        if os.environ.get("TNN_CPU_OPT_TORCH", "0") == "1":
            return False
        C = native.load()
        return C is not None and hasattr(C, "cpu_adam_step") and all(
            t is None or ((not t.is_cuda) and t.dtype == torch.float32 and t.is_contiguous()) for t in ts)

    def _after_update(self):
        ctx = self.ctx
        if ctx.lp_slab is not None:
            for layer in ctx.layers:
                layer.mark_lp_fresh()
        else:
            for layer in ctx.layers:
                layer.mark_params_dirty()


class SGD(Optimizer):
    NAME = "sgd"

    def __init__(self, lr=0.01, momentum=0.0, weight_decay=0.0):
        super().__init__(lr)
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        self.velocity = None

    def _alloc_state(self):
        self.velocity = torch.zeros_like(self.ctx.param_slab) if self.momentum > 0 else None

    def update(self):
        ctx = self.ctx
        self.step_count += 1
        p, g = ctx.param_slab, ctx.grad_slab
        if p.is_cuda:
            self.step_count -= 1
            st = self._push_state(p.device)
            self.step_count += 1
            native.require().sgd_step_flat_dev(p, g, self.velocity, ctx.lp_slab, st, self.fused_zero_grad)
        elif self._cpu_native(p, g, self.velocity):
            native.load().cpu_sgd_step(p.data_ptr(), g.data_ptr(), self.velocity.data_ptr() if self.velocity is not None else 0,
                                       p.numel(), self.lr, self.momentum, self.weight_decay, self.grad_scale, self.fused_zero_grad)
        else:
            gg = g * self.grad_scale + self.weight_decay * p
            if self.velocity is not None:
                self.velocity.mul_(self.momentum).add_(gg, alpha=-self.lr)
                p.add_(self.velocity)
            else:
                p.add_(gg, alpha=-self.lr)
            if self.fused_zero_grad:
                g.zero_()
        self._after_update()

    def state_dict(self):
        d = super().state_dict()
        if self.velocity is not None:
            d["velocity"] = self.velocity
        return d

    def load_state_dict(self, d):
        super().load_state_dict(d)
        if "velocity" in d and self.velocity is not None:
            self.velocity.copy_(d["velocity"])

    def get_config(self):
        return TConfig("sgd", "sgd", learning_rate=self.lr, momentum=self.momentum, weight_decay=self.weight_decay)


class Adam(Optimizer):
    NAME = "adam"

    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, weight_decay=0.0, decouple_weight_decay=False):
        super().__init__(lr)
        self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)
        self.weight_decay, self.decoupled = float(weight_decay), bool(decouple_weight_decay)
        self.m = self.v = None

    def _alloc_state(self):
        self.m = torch.zeros_like(self.ctx.param_slab)
        self.v = torch.zeros_like(self.ctx.param_slab)

    def update(self):
        ctx = self.ctx
        self.step_count += 1
        t = self.step_count
        p, g = ctx.param_slab, ctx.grad_slab
        if p.is_cuda:
            self.step_count -= 1
            st = self._push_state(p.device)
            self.step_count += 1
            native.require().adam_step_flat_dev(p, g, self.m, self.v, ctx.lp_slab, st, self.fused_zero_grad)
        elif self._cpu_native(p, g, self.m, self.v):
            native.load().cpu_adam_step(p.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), p.numel(), self.lr,
                                        self.beta1, self.beta2, self.epsilon, self.weight_decay, self.decoupled, self.grad_scale,
                                        t, self.fused_zero_grad)
        else:
            gg = g * self.grad_scale
            if self.weight_decay != 0 and self.decoupled:
                p.mul_(1 - self.weight_decay * self.lr)
            self.m.mul_(self.beta1).add_(gg, alpha=1 - self.beta1)
            self.v.mul_(self.beta2).addcmul_(gg, gg, value=1 - self.beta2)
            mh = self.m / (1 - self.beta1 ** t)
            vh = self.v / (1 - self.beta2 ** t)
            upd = self.lr * mh / (vh.sqrt() + self.epsilon)
            if self.weight_decay != 0 and not self.decoupled:
                upd = upd + self.weight_decay * self.lr * p
            p.sub_(upd)
            if self.fused_zero_grad:
                g.zero_()
        self._after_update()

    def state_dict(self):
        d = super().state_dict()
        # under FusedDataParallel the moments are sharded in dp.m / dp.v (self.m is None): the DP wrapper
        # contributes the gathered tensors (FusedDataParallel.state_dict); never write a null moment
        if self.m is not None:
            d["m"] = self.m
        if self.v is not None:
            d["v"] = self.v
        return d

    def load_state_dict(self, d):
        super().load_state_dict(d)
        if d.get("m") is not None and self.m is not None:
            self.m.copy_(d["m"])
        if d.get("v") is not None and self.v is not None:
            self.v.copy_(d["v"])

    def get_config(self):
        return TConfig("adam", "adam", learning_rate=self.lr, beta1=self.beta1, beta2=self.beta2,
                       epsilon=self.epsilon, weight_decay=self.weight_decay, decouple_weight_decay=self.decoupled)


class OptimizerFactory:
    @staticmethod
    def create_sgd(lr=0.01, momentum=0.0, weight_decay=0.0):
        return SGD(lr, momentum, weight_decay)

    @staticmethod
    def create_adam(lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, weight_decay=0.0, decouple_weight_decay=False):
        return Adam(lr, beta1, beta2, epsilon, weight_decay, decouple_weight_decay)

    @staticmethod
    def create_from_config(cfg: TConfig | dict | str) -> Optimizer:
        if not isinstance(cfg, TConfig):
            cfg = TConfig.from_json(cfg)
        t = (cfg.type or cfg.name).lower()
        if t == "sgd":
            return SGD(cfg.get("learning_rate", 0.01), cfg.get("momentum", 0.0), cfg.get("weight_decay", 0.0))
        if t in ("adam", "adamw"):
            return Adam(cfg.get("learning_rate", 1e-3), cfg.get("beta1", 0.9), cfg.get("beta2", 0.999),
                        cfg.get("epsilon", 1e-8), cfg.get("weight_decay", 0.0),
                        cfg.get("decouple_weight_decay", t == "adamw"))
        raise ValueError(f"unknown optimizer {t}")
