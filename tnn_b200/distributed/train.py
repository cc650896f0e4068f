"""Distributed (coordinator-driven) training loop.

Reference: include/distributed/train.hpp:19-128 -- train_semi_async_epoch (split each batch into
``num_microbatches``, async_train_batch, update_parameters every ``gradient_accumulation_steps``),
validate_semi_async_epoch (EVAL_MODE, one microbatch = whole batch), train_model(Coordinator&).
"""
from __future__ import annotations

import time

import torch

from ..nn.loss import Loss
from ..nn.train import TrainingConfig
from .coordinator import Coordinator


def split_microbatches(x: torch.Tensor, y: torch.Tensor, m: int):
    m = max(1, min(int(m), x.shape[0]))
    return list(torch.tensor_split(x, m, 0)), list(torch.tensor_split(y, m, 0))


def train_semi_async_epoch(coordinator: Coordinator, loader, criterion: Loss, cfg: TrainingConfig, log=print,
                           epoch: int = 0) -> tuple[float, float]:
    coordinator.set_training(True)
    loader.shuffle()
    loader.reset()
    total_loss = corrects = 0.0
    seen = batches = 0
    t_last = time.time()
    while True:
        b = loader.get_batch(cfg.batch_size)
        if b is None:
            break
        x, y = b
        xs, ys = split_microbatches(x, y, cfg.num_microbatches)
        loss, corr = coordinator.async_train_batch(xs, ys, criterion)
        batches += 1
        if batches % max(1, cfg.gradient_accumulation_steps) == 0:
            coordinator.update_parameters()
        total_loss += loss
        corrects += corr
        seen += y.numel()
        if cfg.progress_print_interval > 0 and batches % cfg.progress_print_interval == 0:
            dt = (time.time() - t_last) / cfg.progress_print_interval * 1e3
            t_last = time.time()
            log(f"epoch {epoch + 1} batch {batches}: loss {total_loss / batches:.4f} acc {100 * corrects / max(1, seen):.2f}% "
                f"{dt:.2f} ms/batch")
        if 0 < cfg.max_steps <= batches:
            break
    return total_loss / max(1, batches), corrects / max(1, seen)


def validate_semi_async_epoch(coordinator: Coordinator, loader, criterion: Loss, cfg: TrainingConfig) -> tuple[float, float]:
    coordinator.set_training(False)
    loader.reset()
    total_loss = corrects = 0.0
    seen = batches = 0
    while True:
        b = loader.get_batch(cfg.batch_size)
        if b is None:
            break
        x, y = b
        l, c = coordinator.async_val_batch(x, y, criterion)
        total_loss += l
        corrects += c
        seen += y.numel()
        batches += 1
    coordinator.set_training(True)
    return total_loss / max(1, batches), corrects / max(1, seen)


def train_model(coordinator: Coordinator, train_loader, val_loader, criterion: Loss, cfg: TrainingConfig | None = None,
                log=print) -> dict:
    cfg = cfg or TrainingConfig()
    hist = {"train_loss": [], "train_acc": [], "val_loss": [], "val_acc": [], "epoch_s": []}
    coordinator.start_profiling()
    for epoch in range(cfg.epochs):
        t0 = time.time()
        tl, ta = train_semi_async_epoch(coordinator, train_loader, criterion, cfg, log, epoch)
        hist["train_loss"].append(tl)
        hist["train_acc"].append(ta)
        hist["epoch_s"].append(time.time() - t0)
        log(f"epoch {epoch + 1}/{cfg.epochs}: train loss {tl:.4f} acc {100 * ta:.2f}% in {hist['epoch_s'][-1]:.2f}s")
        if val_loader is not None:
            vl, va = validate_semi_async_epoch(coordinator, val_loader, criterion, cfg)
            hist["val_loss"].append(vl)
            hist["val_acc"].append(va)
            log(f"validation: loss {vl:.4f} acc {100 * va:.2f}%")
        if 0 < cfg.max_steps:
            break
    coordinator.fetch_profiling()
    return hist
