"""Losses (reference: include/nn/loss.hpp:24-523, src/nn/loss_impl/cuda/loss_ops.cu).

CrossEntropy takes int class-index labels ([N] or [B, L]); loss = mean over rows of
(logsumexp - logit[label]); grad = (softmax - onehot)/rows.  On the GPU loss, gradient AND the
correct-prediction count come out of ONE fused kernel and stay on the device
(``last_stats`` = [loss_sum, n_correct]); ``compute_loss`` only syncs when the caller asks for a
Python float.
"""
from __future__ import annotations

import torch

from ..core.config import TConfig
from ..ops import native


class Loss:
    NAME = "loss"

    def __init__(self):
        self.last_stats: torch.Tensor | None = None   # device tensor [loss, n_correct]

    def compute_loss(self, pred, target) -> float:
        raise NotImplementedError

    def compute_gradient(self, pred, target, scale: float = 1.0) -> torch.Tensor:
        raise NotImplementedError

    def loss_and_grad(self, pred, target, scale: float = 1.0):
        """Returns (stats device tensor [loss, n_correct], grad).  No host sync."""
        raise NotImplementedError

    def config_params(self) -> dict:
        return {}

    def get_config(self) -> TConfig:
        """LossConfig (loss.hpp:75-83,376-382): type name + the constructor arguments."""
        return TConfig(self.NAME, self.NAME, **self.config_params())

    def name(self) -> str:
        return self.NAME

    def clone(self) -> "Loss":
        return LossFactory.create_from_config(self.get_config())


def _rows_view(pred):
    """[..., V] -> [rows, V] WITHOUT copying when the leading dims collapse (a Dense output whose rows are
    padded to a multiple of 8 elements is such a view: copying GPT-2's [8192, 50257] logits costs 0.9 ms)."""
    V = pred.shape[-1]
    if pred.dim() == 2:
        return pred, V
    if pred.stride(-1) == 1:
        try:
            return pred.view(-1, V), V
        except RuntimeError:
            pass
    return pred.contiguous().reshape(-1, V), V


class CrossEntropyLoss(Loss):
    NAME = "crossentropy"

    def __init__(self, use_logits: bool = True, eps: float = 1e-15):
        super().__init__()
        self.use_logits, self.eps = bool(use_logits), float(eps)

    def config_params(self):
        return dict(use_logits=self.use_logits, epsilon=self.eps)

    def loss_and_grad(self, pred, target, scale: float = 1.0, want_grad: bool = True):
        p2, V = _rows_view(pred)
        rows = p2.shape[0]
        labels = target.reshape(-1)
        if p2.is_cuda and self.use_logits:
            Cx = native.require()
            labels = labels.to(torch.int32)
            stats = torch.zeros(2, dtype=torch.float32, device=p2.device)
            grad = torch.empty_strided(p2.shape, p2.stride(), dtype=p2.dtype, device=p2.device) if want_grad else None
            Cx.softmax_ce(p2, p2.stride(0), labels, rows, V, 1.0 / rows, scale / rows, grad, stats)
            self.last_stats = stats
            return stats, (grad.reshape(pred.shape) if grad is not None else None)
        x = p2.float()
        labels = labels.long()
        if self.use_logits:
            logp = torch.log_softmax(x, -1)
            prob = logp.exp()
        else:
            prob = x.clamp(self.eps, 1.0 - self.eps)      # loss_ops.cpp:31-34
            logp = prob.log()
        loss = -(logp[torch.arange(rows), labels]).mean()
        correct = (x.argmax(-1) == labels).sum().float()
        stats = torch.stack([loss, correct])
        grad = None
        if want_grad:
            onehot = torch.zeros_like(x)
            onehot[torch.arange(rows), labels] = 1.0
            grad = ((prob - onehot) if self.use_logits else (-onehot / prob)) * (scale / rows)
            grad = grad.to(pred.dtype).reshape(pred.shape)
        self.last_stats = stats
        return stats, grad

    def compute_loss(self, pred, target) -> float:
        stats, _ = self.loss_and_grad(pred, target, want_grad=False)
        return float(stats[0])

    def compute_gradient(self, pred, target, scale: float = 1.0):
        return self.loss_and_grad(pred, target, scale)[1]


class LogSoftmaxCrossEntropyLoss(CrossEntropyLoss):
    NAME = "logsoftmax_crossentropy"

    def __init__(self):
        super().__init__(use_logits=True)

    def config_params(self):
        return {}


class _Regression(Loss):
    KIND = 0

    def __init__(self, delta: float = 1.0):
        super().__init__()
        self.delta = delta

    def loss_and_grad(self, pred, target, scale: float = 1.0, want_grad: bool = True):
        n = pred.numel()
        if pred.is_cuda and pred.dtype in (torch.float32, torch.bfloat16):
            Cx = native.require()
            p, t = pred.contiguous(), target.to(pred.dtype).contiguous()
            stats = torch.zeros(2, dtype=torch.float32, device=pred.device)
            grad = torch.empty_like(p) if want_grad else None
            Cx.regression_loss(p, t, self.KIND, self.delta, scale / n, grad, stats)
            self.last_stats = stats
            return stats, grad
        d = pred.float() - target.float()
        if self.KIND == 0:
            loss, g = (d * d).mean(), 2 * d
        elif self.KIND == 1:
            loss, g = d.abs().mean(), torch.sign(d)
        else:
            a = d.abs()
            loss = torch.where(a <= self.delta, 0.5 * d * d, self.delta * (a - 0.5 * self.delta)).mean()
            g = torch.where(a <= self.delta, d, self.delta * torch.sign(d))
        stats = torch.stack([loss, torch.zeros(())])
        self.last_stats = stats
        return stats, (g * (scale / n)).to(pred.dtype) if want_grad else None

    def compute_loss(self, pred, target) -> float:
        return float(self.loss_and_grad(pred, target, want_grad=False)[0][0])

    def compute_gradient(self, pred, target, scale: float = 1.0):
        return self.loss_and_grad(pred, target, scale)[1]


class MSELoss(_Regression):
    NAME, KIND = "mse", 0


class MAELoss(_Regression):
    NAME, KIND = "mae", 1


class HuberLoss(_Regression):
    NAME, KIND = "huber", 2

    def config_params(self):
        return dict(delta=float(self.delta))

    def set_delta(self, delta: float):
        self.delta = float(delta)

    def get_delta(self) -> float:
        return self.delta


class LossFactory:
    """loss.hpp:464-523."""

    @staticmethod
    def create(name: str, **kw) -> Loss:
        n = name.lower()
        if n in ("crossentropy", "ce"):
            return CrossEntropyLoss(kw.get("use_logits", True), kw.get("epsilon", kw.get("eps", 1e-15)))
        if n in ("logsoftmax_crossentropy", "logsoftmax_ce"):
            return LogSoftmaxCrossEntropyLoss()
        if n in ("mse", "mean_squared_error"):
            return MSELoss()
        if n in ("mae", "mean_absolute_error"):
            return MAELoss()
        if n == "huber":
            return HuberLoss(kw.get("delta", 1.0))
        raise ValueError(f"unknown loss {name}")

    @staticmethod
    def create_crossentropy(use_logits=True, eps=1e-15):
        return CrossEntropyLoss(use_logits, eps)

    @staticmethod
    def create_logsoftmax_crossentropy():
        return LogSoftmaxCrossEntropyLoss()

    @staticmethod
    def create_mse():
        return MSELoss()

    @staticmethod
    def create_mae():
        return MAELoss()

    @staticmethod
    def create_huber(delta: float = 1.0):
        return HuberLoss(delta)

    @staticmethod
    def create_from_config(cfg: TConfig) -> Loss:
        return LossFactory.create(cfg.type or cfg.name, **cfg.parameters)


def compute_class_corrects(pred: torch.Tensor, labels: torch.Tensor) -> int:
    """argmax == label count (accuracy.hpp:14-40)."""
    V = pred.shape[-1]
    p2 = pred.reshape(-1, V)
    if p2.is_cuda:
        Cx = native.require()
        out = torch.empty(p2.shape[0], dtype=torch.int32, device=p2.device)
        p2 = p2 if p2.stride(1) == 1 else p2.contiguous()
        Cx.argmax_rows(p2, p2.stride(0), p2.shape[0], V, out)
        return int((out == labels.reshape(-1).to(torch.int32)).sum())
    return int((p2.argmax(-1) == labels.reshape(-1).long()).sum())
