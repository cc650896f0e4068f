#!/usr/bin/env python
"""Single-process trainer (reference: examples/trainer.cpp:16-80).

  python examples/trainer.py [--config configs/default_config.json] [--synthetic]

Configuration comes from a JSON file (``--config``) or from the environment / ``./.env`` (keys of
``.env.example``).  Hyper-parameters hard-coded by the reference trainer are kept as defaults:
Adam(lr, 0.9, 0.999, eps 1e-3, wd 3e-4), StepLR every 5 epochs x0.1, log-softmax cross-entropy.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.data import DataLoaderFactory, SyntheticDataLoader  # noqa: E402
from tnn_b200.models import ExampleModels, load_or_create_model  # noqa: E402
from tnn_b200.nn import LossFactory, OptimizerFactory, SchedulerFactory  # noqa: E402
from tnn_b200.nn.train import TrainingConfig, train_model  # noqa: E402
from tnn_b200.utils.hardware_info import HardwareInfo  # noqa: E402


def synthetic_for(model_name: str, n=2048):
    if "mnist" in model_name:
        return SyntheticDataLoader((28, 28, 1), 10, n), SyntheticDataLoader((28, 28, 1), 10, n // 4)
    if "gpt2" in model_name:
        return SyntheticDataLoader((1024,), 50257, n // 16, token_vocab=50257), None
    classes = 100 if "cifar100" in model_name else (200 if "tiny" in model_name else 10)
    hw = 64 if "tiny" in model_name else (224 if "imagenet_" in model_name else 32)
    return SyntheticDataLoader((hw, hw, 3), classes, n), SyntheticDataLoader((hw, hw, 3), classes, n // 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None)
    ap.add_argument("--synthetic", action="store_true", help="random data of the model's input shape")
    a = ap.parse_args()
    ExampleModels.register_defaults()
    cfg = TrainingConfig.load_from_json(a.config) if a.config else TrainingConfig.load_from_env()
    dev = cfg.device()
    if dev.type == "cuda" and not torch.cuda.is_available():
        print("[trainer] no CUDA device, falling back to CPU")
        cfg.device_type = "CPU"
        dev = cfg.device()
    print(HardwareInfo().summary())
    print(f"[trainer] model {cfg.model_name} on {dev}, dataset {cfg.dataset_name} @ {cfg.dataset_path}")
    if a.synthetic:
        train, val = synthetic_for(cfg.model_name)
        train.load_data()
        if val is not None:
            val.load_data()
    else:
        train, val = DataLoaderFactory.create(cfg.dataset_name, cfg.dataset_path)
    graph = load_or_create_model(cfg.model_name, cfg.model_path or None, device=dev, seed=cfg.seed or None)
    crit = LossFactory.create_logsoftmax_crossentropy()
    opt = OptimizerFactory.create_adam(cfg.lr_initial, 0.9, 0.999, 1e-3, 3e-4)
    steps_per_epoch = max(1, train.size() // cfg.batch_size)
    sched = SchedulerFactory.create_step_lr(opt, 5 * steps_per_epoch, 0.1)
    hist = train_model(graph, train, val, opt, crit, sched, cfg)
    print(f"[trainer] done: train loss {hist['train_loss'][-1]:.4f}" + (f", val acc {100 * hist['val_acc'][-1]:.2f}%" if hist["val_acc"] else ""))


if __name__ == "__main__":
    main()
