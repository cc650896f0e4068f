#!/usr/bin/env python
"""Builds a zoo model node by node with the graph API, collapses the chain with the sequential
reducer, and trains it through the GraphExecutor (reference: examples/graph_test.cpp:20-121).

  python examples/graph_example.py [--config configs/default_config.json] [--model mnist_cnn] [--steps 20]

Without a dataset on disk it trains on synthetic data of the model's input shape.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.data import DataLoaderFactory, SyntheticDataLoader  # noqa: E402
from tnn_b200.models import ExampleModels  # noqa: E402
from tnn_b200.nn import GraphBuilder, GraphExecutor, LayerFactory, LossFactory, OptimizerFactory  # noqa: E402
from tnn_b200.nn.loss import compute_class_corrects  # noqa: E402
from tnn_b200.nn.reducer import ReducerBuilder  # noqa: E402
from tnn_b200.nn.train import TrainingConfig  # noqa: E402


def build_graph(model_name: str, device, seed=123456):
    """One op node per layer of the zoo model, IO nodes between them, then SeqReduction."""
    ExampleModels.register_defaults()
    LayerFactory.register_defaults()
    model = ExampleModels.create(model_name)
    layers = model.get_layers() if hasattr(model, "get_layers") else model.layers
    builder = GraphBuilder(model_name)
    cur = builder.io("input")
    for i, layer in enumerate(layers):
        op = builder.add_layer(LayerFactory.create(layer.get_config()))
        out = builder.io(f"layer_output_{i}")
        builder.add_edge([cur], [out], op)
        cur = out
    n_before = len(builder.edges)
    ReducerBuilder().seq_reduction().build().reduce(builder)
    print(f"[graph] {n_before} op nodes reduced to {len(builder.edges)}")
    return builder.compile(device, seed=seed), "input", cur.uid


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None)
    ap.add_argument("--model", default="mnist_cnn")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    cfg = TrainingConfig.load_from_json(a.config) if a.config else TrainingConfig.load_from_env()
    if a.config is None:
        cfg.model_name = a.model
    dev = cfg.device()
    if dev.type == "cuda" and not torch.cuda.is_available():
        dev = torch.device("cpu")
    graph, in_uid, out_uid = build_graph(cfg.model_name, dev)
    graph.set_training(True)
    executor = GraphExecutor(graph)
    try:
        train, _ = DataLoaderFactory.create(cfg.dataset_name, cfg.dataset_path)
    except Exception as e:  # noqa: BLE001 -- no dataset on this box
        print(f"[graph] dataset unavailable ({type(e).__name__}); using synthetic data")
        shape = (28, 28, 1) if "mnist" in cfg.model_name else (32, 32, 3)
        train = SyntheticDataLoader(shape, 100 if "cifar100" in cfg.model_name else 10, a.batch * a.steps)
        train.load_data()
    train.set_seed(123456)
    crit = LossFactory.create_logsoftmax_crossentropy()
    opt = OptimizerFactory.create_adam(cfg.lr_initial, 0.9, 0.999, 1e-3, 3e-4, False)
    opt.attach(graph.context())
    step = 0
    while step < a.steps:
        batch = train.get_batch(a.batch)
        if batch is None:
            break
        x, y = (t.to(dev) for t in batch)
        out = executor.forward({in_uid: x})[out_uid]
        loss = crit.compute_loss(out, y)
        acc = 100.0 * compute_class_corrects(out, y) / out.shape[0]
        print(f"Loss: {loss:.4f}, Accuracy: {acc:.2f}%")
        executor.backward({out_uid: crit.compute_gradient(out, y)})
        opt.update()
        opt.zero_grads()
        step += 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
