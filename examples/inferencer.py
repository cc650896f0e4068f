#!/usr/bin/env python
"""Load a model snapshot and evaluate it (reference: examples/inferencer.cpp).

  python examples/inferencer.py --model cifar100_wrn16_8 --path model_snapshots/cifar100_wrn16_8 --dataset cifar100 --data data
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.data import DataLoaderFactory, SyntheticDataLoader  # noqa: E402
from tnn_b200.models import load_or_create_model  # noqa: E402
from tnn_b200.nn import LossFactory  # noqa: E402
from tnn_b200.nn.train import TrainingConfig, validate_model  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="cifar10_resnet9")
    ap.add_argument("--path", default="")
    ap.add_argument("--dataset", default="")
    ap.add_argument("--data", default="data")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--device", default="GPU" if torch.cuda.is_available() else "CPU")
    a = ap.parse_args()
    cfg = TrainingConfig(batch_size=a.batch, device_type=a.device, model_name=a.model)
    g = load_or_create_model(a.model, a.path or None, device=cfg.device(), seed=1)
    if a.dataset:
        _, val = DataLoaderFactory.create(a.dataset, a.data)
    else:
        shape = (28, 28, 1) if "mnist" in a.model else (32, 32, 3)
        val = SyntheticDataLoader(shape, 100 if "100" in a.model else 10, 1024)
        val.load_data()
    loss, acc = validate_model(g, val, LossFactory.create_logsoftmax_crossentropy(), cfg)
    print(f"validation: loss {loss:.4f} accuracy {100 * acc:.2f}% over {val.size()} samples")
