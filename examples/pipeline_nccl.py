#!/usr/bin/env python
"""1F1B pipeline-parallel training over NCCL p2p, one stage per GPU (the B200-native stage data
path; the coordinator/worker examples show the reference's message protocol).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 examples/pipeline_nccl.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tnn_b200.models import ExampleModels  # noqa: E402
from tnn_b200.nn import LossFactory, OptimizerFactory  # noqa: E402
from tnn_b200.parallel import FTDPartitioner  # noqa: E402
from tnn_b200.parallel.pipeline import PipelineParallel  # noqa: E402

if __name__ == "__main__":
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    cuda = torch.cuda.is_available()
    dev = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local)
    # a collective stuck longer than this aborts the communicator (torch's NCCL watchdog); the engine's own host-side stall
    # watchdog is TNN_PP_WATCHDOG_SEC (parallel/pipeline.py)
    import datetime
    dist.init_process_group("nccl" if cuda else "gloo", timeout=datetime.timedelta(seconds=float(os.environ.get("TNN_PP_TIMEOUT_SEC", "600"))))
    B, M = 256, 8
    model = ExampleModels.create(os.environ.get("MODEL_NAME", "cifar100_wrn16_8"))
    pp = PipelineParallel(model, FTDPartitioner(world, (B // M, 32, 32, 3)), OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4),
                          LossFactory.create_logsoftmax_crossentropy(), M, (B // M, 32, 32, 3), dev)
    g = torch.Generator().manual_seed(0)
    x, y = torch.rand(B, 32, 32, 3, generator=g), torch.randint(0, 100, (B,), generator=g, dtype=torch.int32)
    for step in range(20):
        st = pp.broadcast_stats(pp.train_batch(x, y))
        if rank == 0 and step % 5 == 0:
            print(f"step {step}: loss {float(st[0]):.4f} stages {[(p.start_offset, p.length) for p in pp.partitions]}", flush=True)
    dist.destroy_process_group()
