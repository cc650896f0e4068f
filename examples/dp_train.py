#!/usr/bin/env python
"""Data-parallel training with the fused NVLink gradient-reduce + optimizer kernel.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/dp_train.py \
      [--model cifar100_wrn16_8] [--batch 256] [--steps 200] [--dp fused|nccl]

One process per GPU.  Parameters / gradients / bf16 copies of every replica live in a peer-mapped
symmetric slab; each gradient bucket is reduced, updated (Adam) and broadcast by ONE kernel on a side
stream while backward continues (tnn_b200/parallel/dp.py).  ``--dp nccl`` runs the measured baseline
(NCCL all-reduce, then the flat optimizer kernel).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tnn_b200.models import ExampleModels, load_or_create_model  # noqa: E402
from tnn_b200.nn import LossFactory, OptimizerFactory, SchedulerFactory  # noqa: E402
from tnn_b200.nn.train import TrainStep  # noqa: E402
from tnn_b200.parallel.dp import AllReduceDataParallel, FusedDataParallel  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="cifar100_wrn16_8")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--dp", default="fused", choices=["fused", "nccl"])
    a = ap.parse_args()
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    opt = OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4)
    if a.dp == "fused":
        dp = FusedDataParallel(opt)
        graph = dp.build_graph(ExampleModels.create(a.model), dev, seed=1234, name=a.model)
    else:
        graph = load_or_create_model(a.model, device=dev, seed=1234)
        opt.attach(graph.context())
        dp = AllReduceDataParallel(graph, opt)
    step = TrainStep(graph, LossFactory.create_logsoftmax_crossentropy(), opt, SchedulerFactory.create_step_lr(opt, 1000, 0.5), 1, dp)
    classes = 10 if "cifar10_" in a.model else 100
    g = torch.Generator().manual_seed(100 + rank)
    xs = [torch.rand((a.batch, 32, 32, 3), generator=g).pin_memory() for _ in range(8)]
    ys = [torch.randint(0, classes, (a.batch,), generator=g, dtype=torch.int32).pin_memory() for _ in range(8)]
    for i in range(a.steps):
        stats = step(xs[i % 8], ys[i % 8])
        if rank == 0 and i % 20 == 0:
            print(f"step {i}: loss {float(stats[0]):.4f} acc {100 * float(stats[1]) / a.batch:.1f}% lr {opt.get_learning_rate():.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
