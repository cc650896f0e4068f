"""Learning-rate schedulers with config round-trip (reference: include/nn/schedulers.hpp:27-751).
``step()`` is called once per optimizer step.  Closed-form schedules (everything except
ReduceLROnPlateau) are pure functions of ``current_step`` so a resumed run is exact."""
from __future__ import annotations

import math

from ..core.config import TConfig


class Scheduler:
    TYPE = "scheduler"

    def __init__(self, optimizer):
        self.optimizer = optimizer
        self.base_lr = optimizer.get_learning_rate() if optimizer is not None else 0.0
        self.current_step = 0

    def get_lr(self) -> float:
        return self.optimizer.get_learning_rate() if self.optimizer is not None else self.base_lr

    def get_base_lr(self) -> float:
        return self.base_lr

    def get_current_step(self) -> int:
        return self.current_step

    def _set_lr(self, lr: float):
        if self.optimizer is not None:
            self.optimizer.set_learning_rate(lr)

    def reset(self):
        self.current_step = 0
        self._set_lr(self.base_lr)

    def lr_at(self, step: int) -> float | None:
        return None

    def step(self, metric: float | None = None):
        self.current_step += 1
        lr = self.lr_at(self.current_step)
        if lr is not None:
            self._set_lr(lr)

    def name(self) -> str:
        return type(self).__name__

    def params(self) -> dict:
        return {}

    def get_config(self) -> TConfig:
        return TConfig(self.name(), self.TYPE, base_lr=self.base_lr, **self.params())

    def clone(self, optimizer) -> "Scheduler":
        return SchedulerFactory.create_from_config(self.get_config(), optimizer)

    def state_dict(self) -> dict:
        return {"current_step": self.current_step}

    def load_state_dict(self, d: dict):
        self.current_step = int(d.get("current_step", 0))
        lr = self.lr_at(self.current_step)
        if lr is not None and self.current_step > 0:
            self._set_lr(lr)


class NoOpScheduler(Scheduler):
    TYPE = "no_op"


class StepLR(Scheduler):
    TYPE = "step_lr"

    def __init__(self, optimizer, step_size: int, gamma: float = 0.1):
        super().__init__(optimizer)
        self.step_size, self.gamma = max(1, int(step_size)), float(gamma)

    def lr_at(self, s):
        return self.base_lr * self.gamma ** (s // self.step_size)

    def params(self):
        return dict(step_size=self.step_size, gamma=self.gamma)


class MultiStepLR(Scheduler):
    TYPE = "multi_step_lr"

    def __init__(self, optimizer, milestones, gamma: float = 0.1):
        super().__init__(optimizer)
        self.milestones, self.gamma = sorted(int(m) for m in milestones), float(gamma)

    def lr_at(self, s):
        return self.base_lr * self.gamma ** sum(1 for m in self.milestones if s >= m)

    def params(self):
        return dict(milestones=list(self.milestones), gamma=self.gamma)


class ExponentialLR(Scheduler):
    TYPE = "exponential_lr"

    def __init__(self, optimizer, gamma: float = 0.95):
        super().__init__(optimizer)
        self.gamma = float(gamma)

    def lr_at(self, s):
        return self.base_lr * self.gamma ** s

    def params(self):
        return dict(gamma=self.gamma)


class CosineAnnealingLR(Scheduler):
    TYPE = "cosine_annealing_lr"

    def __init__(self, optimizer, T_max: int, eta_min: float = 0.0):
        super().__init__(optimizer)
        self.T_max, self.eta_min = max(1, int(T_max)), float(eta_min)

    def lr_at(self, s):
        s = min(s, self.T_max)
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * s / self.T_max)) / 2

    def params(self):
        return dict(T_max=self.T_max, eta_min=self.eta_min)


class CosineAnnealingWarmRestarts(Scheduler):
    TYPE = "cosine_annealing_warm_restarts"

    def __init__(self, optimizer, T_0: int, T_mult: int = 1, eta_min: float = 0.0):
        super().__init__(optimizer)
        self.T_0, self.T_mult, self.eta_min = max(1, int(T_0)), max(1, int(T_mult)), float(eta_min)

    def lr_at(self, s):
        T_i, t = self.T_0, s
        while t >= T_i:
            t -= T_i
            T_i *= self.T_mult
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * t / T_i)) / 2

    def params(self):
        return dict(T_0=self.T_0, T_mult=self.T_mult, eta_min=self.eta_min)


class LinearWarmup(Scheduler):
    TYPE = "linear_warmup"

    def __init__(self, optimizer, warmup_steps: int, start_lr: float = 0.0):
        super().__init__(optimizer)
        self.warmup_steps, self.start_lr = max(1, int(warmup_steps)), float(start_lr)

    def lr_at(self, s):
        if s >= self.warmup_steps:
            return self.base_lr
        return self.start_lr + (self.base_lr - self.start_lr) * s / self.warmup_steps

    def params(self):
        return dict(warmup_steps=self.warmup_steps, start_lr=self.start_lr)


class WarmupCosineAnnealing(Scheduler):
    TYPE = "warmup_cosine_annealing"

    def __init__(self, optimizer, warmup_steps: int, total_steps: int, start_lr: float = 0.0, eta_min: float = 0.0):
        super().__init__(optimizer)
        self.warmup_steps, self.total_steps = max(1, int(warmup_steps)), max(int(total_steps), int(warmup_steps) + 1)
        self.start_lr, self.eta_min = float(start_lr), float(eta_min)

    def lr_at(self, s):
        if s < self.warmup_steps:
            return self.start_lr + (self.base_lr - self.start_lr) * s / self.warmup_steps
        p = min(1.0, (s - self.warmup_steps) / (self.total_steps - self.warmup_steps))
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * p)) / 2

    def params(self):
        return dict(warmup_steps=self.warmup_steps, total_steps=self.total_steps, start_lr=self.start_lr,
                    eta_min=self.eta_min)


class ReduceLROnPlateau(Scheduler):
    TYPE = "reduce_lr_on_plateau"

    def __init__(self, optimizer, mode="min", factor=0.1, patience=10, threshold=1e-4, min_lr=0.0):
        super().__init__(optimizer)
        self.mode, self.factor, self.patience = mode, float(factor), int(patience)
        self.threshold, self.min_lr = float(threshold), float(min_lr)
        self.best = math.inf if mode == "min" else -math.inf
        self.bad = 0

    def step(self, metric: float | None = None):
        self.current_step += 1
        if metric is None:
            return
        better = metric < self.best - self.threshold if self.mode == "min" else metric > self.best + self.threshold
        if better:
            self.best, self.bad = metric, 0
        else:
            self.bad += 1
            if self.bad > self.patience:
                self._set_lr(max(self.get_lr() * self.factor, self.min_lr))
                self.bad = 0

    def params(self):
        return dict(mode=self.mode, factor=self.factor, patience=self.patience, threshold=self.threshold,
                    min_lr=self.min_lr)

    def state_dict(self):
        return {"current_step": self.current_step, "best": self.best, "bad": self.bad, "lr": self.get_lr()}

    def load_state_dict(self, d):
        self.current_step = int(d.get("current_step", 0))
        self.best, self.bad = d.get("best", self.best), int(d.get("bad", 0))
        if "lr" in d:
            self._set_lr(d["lr"])


class PolynomialLR(Scheduler):
    TYPE = "polynomial_lr"

    def __init__(self, optimizer, total_steps: int, power: float = 1.0, end_lr: float = 0.0):
        super().__init__(optimizer)
        self.total_steps, self.power, self.end_lr = max(1, int(total_steps)), float(power), float(end_lr)

    def lr_at(self, s):
        s = min(s, self.total_steps)
        return (self.base_lr - self.end_lr) * (1 - s / self.total_steps) ** self.power + self.end_lr

    def params(self):
        return dict(total_steps=self.total_steps, power=self.power, end_lr=self.end_lr)


class OneCycleLR(Scheduler):
    TYPE = "one_cycle_lr"

    def __init__(self, optimizer, max_lr=0.1, total_steps=100, pct_start=0.3, div_factor=25.0, final_div_factor=1e4):
        super().__init__(optimizer)
        self.max_lr, self.total_steps, self.pct_start = float(max_lr), max(2, int(total_steps)), float(pct_start)
        self.div_factor, self.final_div_factor = float(div_factor), float(final_div_factor)
        self.initial_lr = self.max_lr / self.div_factor
        self.min_lr = self.initial_lr / self.final_div_factor
        self._set_lr(self.initial_lr)

    @staticmethod
    def _cos(a, b, p):
        return b + (a - b) * (1 + math.cos(math.pi * p)) / 2

    def lr_at(self, s):
        up = max(1, int(self.pct_start * self.total_steps))
        if s <= up:
            return self._cos(self.initial_lr, self.max_lr, s / up)
        p = min(1.0, (s - up) / max(1, self.total_steps - up))
        return self._cos(self.max_lr, self.min_lr, p)

    def params(self):
        return dict(max_lr=self.max_lr, total_steps=self.total_steps, pct_start=self.pct_start,
                    div_factor=self.div_factor, final_div_factor=self.final_div_factor)


class SchedulerFactory:
    @staticmethod
    def create_no_op(opt):
        return NoOpScheduler(opt)

    @staticmethod
    def create_step_lr(opt, step_size, gamma=0.1):
        return StepLR(opt, step_size, gamma)

    @staticmethod
    def create_cosine_annealing(opt, T_max, eta_min=0.0):
        return CosineAnnealingLR(opt, T_max, eta_min)

    @staticmethod
    def create_warmup_cosine(opt, warmup_steps, total_steps, start_lr=0.0, eta_min=0.0):
        return WarmupCosineAnnealing(opt, warmup_steps, total_steps, start_lr, eta_min)

    @staticmethod
    def create(name: str, opt, **p):
        return SchedulerFactory.create_from_config(TConfig(name, name, **p), opt)

    @staticmethod
    def create_from_config(cfg, opt):
        if not isinstance(cfg, TConfig):
            cfg = TConfig.from_json(cfg)
        t, g = cfg.type, cfg.get
        table = {
            "no_op": lambda: NoOpScheduler(opt),
            "step_lr": lambda: StepLR(opt, g("step_size", 10), g("gamma", 0.1)),
            "multi_step_lr": lambda: MultiStepLR(opt, g("milestones", []), g("gamma", 0.1)),
            "exponential_lr": lambda: ExponentialLR(opt, g("gamma", 0.95)),
            "cosine_annealing_lr": lambda: CosineAnnealingLR(opt, g("T_max", 100), g("eta_min", 0.0)),
            "cosine_annealing_warm_restarts": lambda: CosineAnnealingWarmRestarts(opt, g("T_0", 10), g("T_mult", 1), g("eta_min", 0.0)),
            "linear_warmup": lambda: LinearWarmup(opt, g("warmup_steps", 100), g("start_lr", 0.0)),
            "warmup_cosine_annealing": lambda: WarmupCosineAnnealing(opt, g("warmup_steps", 100), g("total_steps", 1000), g("start_lr", 0.0), g("eta_min", 0.0)),
            "reduce_lr_on_plateau": lambda: ReduceLROnPlateau(opt, g("mode", "min"), g("factor", 0.1), g("patience", 10), g("threshold", 1e-4), g("min_lr", 0.0)),
            "polynomial_lr": lambda: PolynomialLR(opt, g("total_steps", 100), g("power", 1.0), g("end_lr", 0.0)),
            "one_cycle_lr": lambda: OneCycleLR(opt, g("max_lr", 0.1), g("total_steps", 100), g("pct_start", 0.3), g("div_factor", 25.0), g("final_div_factor", 1e4)),
        }
        if t not in table:
            raise ValueError(f"unknown scheduler type {t}")
        s = table[t]()
        if cfg.has("base_lr") and t != "one_cycle_lr":
            s.base_lr = float(cfg.get("base_lr"))
        return s
