#!/usr/bin/env python
"""Pipeline-parallel training driven by a coordinator (reference: examples/tcp_coordinator.cpp:27-108,
roce_coordinator.cpp).  Endpoints come from the environment like the reference:
COORDINATOR_HOST/PORT, LOCAL_WORKER_HOST/PORT/POSITION, WORKER1_HOST/PORT, WORKER2_..., plus the
TrainingConfig keys (MODEL_NAME, DATASET_NAME, BATCH_SIZE, NUM_MICROBATCHES, ...).

  WORKER1_PORT=9001 python examples/tcp_coordinator.py --synthetic --proportions 1,2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.data import DataLoaderFactory  # noqa: E402
from tnn_b200.distributed import (CoordinatorConfig, Endpoint, RoCECoordinator, ParallelMode, RoCEWorker,  # noqa: E402
                                  train_model)
from tnn_b200.models import ExampleModels  # noqa: E402
from tnn_b200.nn import LossFactory, OptimizerFactory, SchedulerFactory  # noqa: E402
from tnn_b200.nn.train import TrainingConfig  # noqa: E402
from tnn_b200.parallel import NaivePartitionerConfig, NaivePipelinePartitioner  # noqa: E402
from tnn_b200.utils.env import Env  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--proportions", default="1,2")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--legacy", action="store_true",
                    help="legacy_tcp_coordinator / legacy_roce_coordinator: NCHW zoo + loaders, Adam eps 1e-5 wd 1e-4")
    a = ap.parse_args()
    env = Env(".env")
    cfg = TrainingConfig.load_from_env()
    props = [int(p) for p in a.proportions.split(",")]
    workers = []
    i = 1
    while env.has(f"WORKER{i}_PORT") or os.environ.get(f"WORKER{i}_PORT"):
        workers.append(Endpoint.roce(env.get(f"WORKER{i}_HOST", "127.0.0.1"), env.get(f"WORKER{i}_PORT", 0, int)))
        i += 1
    local = RoCEWorker(env.get("LOCAL_WORKER_HOST", "127.0.0.1"), env.get("LOCAL_WORKER_PORT", 0, int),
                      use_gpu=a.gpu and torch.cuda.is_available())
    if a.legacy:
        from tnn_b200.models import legacy
        model = legacy.ExampleModels.create(cfg.model_name)
        opt = OptimizerFactory.create_adam(cfg.lr_initial, 0.9, 0.999, 1e-5, 1e-4)
    else:
        model = ExampleModels.create(cfg.model_name)
        opt = OptimizerFactory.create_adam(cfg.lr_initial, 0.9, 0.999, 1e-3, 3e-4)
    ccfg = CoordinatorConfig(ParallelMode.PIPELINE, model, opt, SchedulerFactory.create_step_lr(opt, 10 ** 6, 0.1),
                             NaivePipelinePartitioner(NaivePartitionerConfig(props)), local_worker=local,
                             local_worker_position=env.get("LOCAL_WORKER_POSITION", 0, int), worker_endpoints=workers)
    co = RoCECoordinator(ccfg, env.get("COORDINATOR_HOST", "127.0.0.1"), env.get("COORDINATOR_PORT", 0, int))
    if a.synthetic:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from trainer import synthetic_for
        train, val = synthetic_for(cfg.model_name, 512)
        train.load_data()
        val.load_data()
        if a.legacy:
            train, val = legacy.DataLoaderFactory.wrap(train), legacy.DataLoaderFactory.wrap(val)
    elif a.legacy:
        train, val = legacy.DataLoaderFactory.create(cfg.dataset_name, cfg.dataset_path)
    else:
        train, val = DataLoaderFactory.create(cfg.dataset_name, cfg.dataset_path)
    co.start()
    try:
        train_model(co, train, val, LossFactory.create_logsoftmax_crossentropy(), cfg)
        print("profile log:", co.print_profiling("logs/profiler.log"))
    finally:
        co.stop()
        local.stop()
