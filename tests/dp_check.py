"""Multi-GPU check of the fused data-parallel step (run under torchrun, >= 2 GPUs):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 tests/dp_check.py

1. FusedDataParallel (one kernel: peer-memory reduce + Adam + broadcast) must match the
   NCCL all-reduce + flat Adam baseline on the same data (tolerance = summation-order noise),
   with and without the NVLS multicast path, eager and under CUDA-graph replay.
2. Replicas must stay bit-identical across ranks.
3. Reports the device time of the fused step vs the baseline step.
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
    from tnn_b200.core import stepstate
    from tnn_b200.models import ExampleModels, load_or_create_model
    from tnn_b200.nn import LossFactory, OptimizerFactory
    from tnn_b200.nn.train import TrainStep
    from tnn_b200.parallel.dp import AllReduceDataParallel, FusedDataParallel

    model = os.environ.get("DP_MODEL", "cifar10_resnet9")
    B, classes = 64, 10 if "cifar10_" in model else 100
    gen = torch.Generator().manual_seed(1000 + rank)
    xs = [torch.rand((B, 32, 32, 3), generator=gen).to(dev) for _ in range(3)]
    ys = [torch.randint(0, classes, (B,), generator=gen, dtype=torch.int32).to(dev) for _ in range(3)]
    crit = LossFactory.create_logsoftmax_crossentropy()

    def run(kind, steps, use_graph, multicast=None):
        stepstate._STATES.clear()
        opt = OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4)
        if kind == "fused":
            dp = FusedDataParallel(opt, use_multicast=multicast)
            g = dp.build_graph(ExampleModels.create(model), dev, seed=99, name=model)
        else:
            g = load_or_create_model(model, device=dev, seed=99)
            opt.attach(g.context())
            dp = AllReduceDataParallel(g, opt)
        step = TrainStep(g, crit, opt, None, 1, dp, use_graph=use_graph)
        for l in g.get_layers():
            for sub in l.all_layers():
                if hasattr(sub, "dropout_rate"):
                    sub.dropout_rate = 0.0          # compare deterministic math only
                if hasattr(sub, "fused_dropout_p"):
                    sub.fused_dropout_p = 0.0
        losses = [float(step(xs[i % 3], ys[i % 3])[0]) for i in range(steps)]
        torch.cuda.synchronize()
        return g, losses, step, dp

    ok = True
    gb, lb, _, _ = run("nccl", 6, False)
    pb = gb.ctx.param_slab.clone()
    for mc in (False, True):
        for graph in (False, True):
            gf, lf, _, dp = run("fused", 6, graph, multicast=mc)
            pf = gf.ctx.param_slab
            n = min(pf.numel(), pb.numel())
            err = float((pf[:n] - pb[:n]).abs().max())
            rel = float((pf[:n] - pb[:n]).norm() / pb[:n].norm())
            # replicas identical across ranks
            mine = pf[:n].clone()
            other = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(other, mine)
            same = all(torch.equal(o, other[0]) for o in other)
            lp_ok = bool(torch.allclose(gf.ctx.lp_slab[:n].float(), pf[:n], rtol=1e-2, atol=1e-3))
            good = rel < 2e-3 and same and lp_ok and all(abs(a - b) < 5e-2 for a, b in zip(lb, lf))
            ok &= good
            if rank == 0:
                print(f"fused(multicast={'on' if dp.mc else 'off'}, graph={graph}): max|dp|={err:.3e} rel={rel:.3e} "
                      f"replicas_identical={same} bf16_copy_ok={lp_ok} losses_f={['%.4f' % v for v in lf[:3]]} "
                      f"losses_b={['%.4f' % v for v in lb[:3]]} -> {'OK' if good else 'FAIL'}", flush=True)
            del gf, dp
            torch.cuda.empty_cache()

    # timing of the step (device events, max over ranks)
    def time_steps(kind, graph, multicast=None, steps=20):
        g, _, step, dp = run(kind, 4, graph, multicast)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(xs[i % 3], ys[i % 3])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    tb = time_steps("nccl", False)
    tf = time_steps("fused", True, False)
    tm = time_steps("fused", True, True)
    if rank == 0:
        print(f"step ms (max over ranks, {model}, B={B}/gpu, world={world}): nccl-allreduce+adam eager {tb:.3f} | "
              f"fused p2p graph {tf:.3f} | fused nvls graph {tm:.3f}", flush=True)
        print("DP_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
