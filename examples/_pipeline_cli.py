"""Shared bodies of the coordinator / worker CLIs: ``tcp_*`` and ``roce_*`` differ only in the transport
(reference: examples/tcp_coordinator.cpp:27-108 / roce_coordinator.cpp:20-167, tcp_worker.cpp:39-141 / roce_worker.cpp)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _transport(name):
    from tnn_b200 import distributed as D
    if name == "roce":
        return D.RoCECoordinator, D.RoCEWorker, D.Endpoint.roce
    return D.NetworkCoordinator, D.TCPWorker, D.Endpoint.tcp


def coordinator_main(transport: str) -> None:
    import torch
    from tnn_b200.data import DataLoaderFactory
    from tnn_b200.distributed import CoordinatorConfig, ParallelMode, train_model
    from tnn_b200.models import ExampleModels
    from tnn_b200.nn import LossFactory, OptimizerFactory, SchedulerFactory
    from tnn_b200.nn.train import TrainingConfig
    from tnn_b200.parallel import NaivePartitionerConfig, NaivePipelinePartitioner
    from tnn_b200.utils.env import Env
    Coordinator, Worker, endpoint = _transport(transport)
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
        workers.append(endpoint(env.get(f"WORKER{i}_HOST", "127.0.0.1"), env.get(f"WORKER{i}_PORT", 0, int)))
        i += 1
    local = Worker(env.get("LOCAL_WORKER_HOST", "127.0.0.1"), env.get("LOCAL_WORKER_PORT", 0, int),
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
    co = Coordinator(ccfg, env.get("COORDINATOR_HOST", "127.0.0.1"), env.get("COORDINATOR_PORT", 0, int))
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


def worker_main(transport: str) -> None:
    _, Worker, _ = _transport(transport)
    ap = argparse.ArgumentParser()
    ap.add_argument("port", type=int)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--gpu-index", type=int, default=0)
    ap.add_argument("--io-threads", type=int, default=4)
    ap.add_argument("--num-threads", type=int, default=8)
    ap.add_argument("--device", default="", help="RoCE worker: NIC device name (accepted for CLI compatibility)")
    ap.add_argument("--gid-index", type=int, default=-1, help="RoCE worker: GID index (accepted for CLI compatibility)")
    a = ap.parse_args()
    import torch
    torch.set_num_threads(a.num_threads)
    w = Worker(a.host, a.port, a.gpu, a.gpu_index, a.io_threads)
    print(f"[worker:{transport}] listening on {a.host}:{a.port} ({'GPU ' + str(a.gpu_index) if a.gpu else 'CPU'})", flush=True)
    w.start()
    print("[worker] shutdown", flush=True)
