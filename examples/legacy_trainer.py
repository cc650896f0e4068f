#!/usr/bin/env python
"""Trainer for the legacy (NCHW) model zoo (reference: examples/legacy_trainer.cpp:14-60).

  MODEL_NAME=cifar10_resnet9 DATASET_NAME=cifar10 DATASET_PATH=data python examples/legacy_trainer.py
  python examples/legacy_trainer.py --synthetic          # random NCHW data of the model's input shape

Everything comes from the environment / ``./.env`` like the reference app; hyper-parameters are the
ones it hard-codes: Adam(lr, 0.9, 0.999, eps 1e-5, wd 1e-4), StepLR every 5 epochs x0.1.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.data import SyntheticDataLoader  # noqa: E402
from tnn_b200.models import legacy  # noqa: E402
from tnn_b200.nn import LossFactory, OptimizerFactory, SchedulerFactory  # noqa: E402
from tnn_b200.nn.train import TrainingConfig, train_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--samples", type=int, default=512)
    a = ap.parse_args()
    legacy.ExampleModels.register_defaults()
    cfg = TrainingConfig.load_from_env()
    cfg.print_config()
    model_name, model_path = cfg.model_name, cfg.model_path
    dev = cfg.device()
    if dev.type == "cuda" and not torch.cuda.is_available():
        cfg.device_type = "CPU"
        dev = cfg.device()
    if a.synthetic:
        hw, c, classes = (28, 1, 10) if "mnist" in model_name else ((64, 3, 200) if "tiny" in model_name else (32, 3, 100 if "cifar100" in model_name else 10))
        loaders = []
        for n in (a.samples, max(a.samples // 4, cfg.batch_size)):
            l = SyntheticDataLoader((hw, hw, c), classes, n)
            l.load_data()
            loaders.append(legacy.DataLoaderFactory.wrap(l))
        train, val = loaders
    else:
        if not cfg.dataset_name:
            raise SystemExit("DATASET_NAME environment variable is not set!")
        train, val = legacy.DataLoaderFactory.create(cfg.dataset_name, cfg.dataset_path)
    graph = legacy.load_or_create_model(model_name, model_path or None, device=dev, seed=cfg.seed or None)
    print(f"Training model on device: {'GPU' if dev.type == 'cuda' else 'CPU'}")
    crit = LossFactory.create_logsoftmax_crossentropy()
    opt = OptimizerFactory.create_adam(cfg.lr_initial, 0.9, 0.999, 1e-5, 1e-4, False)
    sched = SchedulerFactory.create_step_lr(opt, max(1, 5 * train.size() // cfg.batch_size), 0.1)
    try:
        hist = train_model(graph, train, val, opt, crit, sched, cfg)
    except Exception as e:  # noqa: BLE001
        print(f"Training failed: {e}", file=sys.stderr)
        return 1
    print(f"[legacy_trainer] done: train loss {hist['train_loss'][-1]:.4f}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
