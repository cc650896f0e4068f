"""Multi-GPU check of the NCCL-p2p 1F1B pipeline (run under torchrun, >= 2 GPUs):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
      --master-port 29544 tests/pp_check.py

1. Loss trajectory of the S-stage pipeline == single-GPU training with the same microbatching
   (bf16 tolerance).
2. Device-timed step (max over ranks) and img/s for WRN-16-8 at global batch 256.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tnn_b200.models import ExampleModels
    from tnn_b200.nn import LossFactory, OptimizerFactory, graph_from_sequential
    from tnn_b200.parallel import FTDPartitioner
    from tnn_b200.parallel.pipeline import PipelineParallel

    model_name = os.environ.get("PP_MODEL", "cifar100_wrn16_8")
    B, M, classes = int(os.environ.get("PP_BATCH", 256)), int(os.environ.get("PP_MICRO", 8)), 100
    seed = 2024
    gen = torch.Generator().manual_seed(11)
    xs = [torch.rand((B, 32, 32, 3), generator=gen).to(dev) for _ in range(3)]
    ys = [torch.randint(0, classes, (B,), generator=gen, dtype=torch.int32).to(dev) for _ in range(3)]
    crit = LossFactory.create_logsoftmax_crossentropy()

    def no_dropout(seq):
        for sub in seq.all_layers():
            if hasattr(sub, "dropout_rate"):
                sub.dropout_rate = 0.0
            if hasattr(sub, "fused_dropout_p"):
                sub.fused_dropout_p = 0.0

    full = ExampleModels.create(model_name)
    part = FTDPartitioner(world, (B // M, 32, 32, 3))
    pp = PipelineParallel(full, part, OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4), crit, M,
                          (B // M, 32, 32, 3), dev, seed=seed)
    no_dropout(pp.stage)
    losses = []
    for i in range(5):
        st = pp.broadcast_stats(pp.train_batch(xs[i % 3], ys[i % 3]))
        losses.append(float(st[0]))
    ok = all(v == v for v in losses)

    if rank == 0:
        ref_model = ExampleModels.create(model_name)
        g = graph_from_sequential(ref_model, dev, seed)
        no_dropout(ref_model)
        opt = OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4)
        opt.attach(g.context())
        ref = []
        for i in range(5):
            tot = 0.0
            for mb, (xm, ym) in enumerate(zip(torch.tensor_split(xs[i % 3], M), torch.tensor_split(ys[i % 3], M))):
                out = ref_model.forward(xm, mb)
                st, grad = crit.loss_and_grad(out, ym, scale=1.0 / M)
                ref_model.backward(grad, mb)
                tot += float(st[0]) / M
            opt.update()
            opt.zero_grads()
            ref.append(tot)
        good = all(abs(a - b) < 5e-2 * max(1.0, abs(a)) for a, b in zip(ref, losses))
        ok &= good
        print(f"pipeline S={world} M={M} partitions={[(p.start_offset, p.length) for p in pp.partitions]}", flush=True)
        print(f"losses pp  {['%.4f' % v for v in losses]}\nlosses ref {['%.4f' % v for v in ref]} -> {'OK' if good else 'FAIL'}", flush=True)
        del g, ref_model
        torch.cuda.empty_cache()

    # timing
    for i in range(3):
        pp.train_batch(xs[i % 3], ys[i % 3])
    dist.barrier()
    torch.cuda.synchronize()
    steps = 15
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        pp.train_batch(xs[i % 3], ys[i % 3])
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"pipeline step {float(t[0]):.3f} ms (max over ranks) -> {B / float(t[0]) * 1e3:.0f} img/s at global batch {B}, "
              f"{world} stages, {M} microbatches; p2p bytes/step/rank0 {pp.p2p_bytes // max(1, pp.steps)}", flush=True)
        print("PP_CHECK", "PASS" if float(flag[0]) > 0 else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if float(flag[0]) > 0 else 1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:  # noqa: BLE001
        import traceback
        print(f"[rank {os.environ.get('RANK')}] " + traceback.format_exc(), flush=True)
        raise
