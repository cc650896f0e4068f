#!/usr/bin/env python
"""Headline benchmark: WRN-16-8 / CIFAR-100 bf16 training throughput (images/s), data parallel
over N B200s of one node (weak scaling, per-GPU batch 256) -- the metric/config named by
BASELINE.json; BASELINE.md's published number is ~2480 img/s (TNN 2-stage pipeline, 2 GPUs).

  python bench.py --gpus N --steps K --warmup W [--impl tnn|reference|torch] [--dp fused|nccl]

For N > 1 launch with ``python -m torch.distributed.run --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...``.

Timing: W warm-up steps, then exactly K steps bracketed by barrier + cuda synchronize, CUDA
events on the compute stream, MAX over ranks.  ``value`` = device-resident inputs;
``e2e.value`` = the same K steps through the public training API with, every step, an H2D copy of
that step's batch from pinned host memory and a D2H read of a step's loss.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  NCCL's banner ("NCCL version ...", a bare printf when the box
# exports NCCL_DEBUG=VERSION) and any other library chatter on fd 1 are sent to stderr: the process keeps
# a private duplicate of the real stdout for the result line and points fd 1 at fd 2.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

BASELINE_IMG_S = 2480.0      # 256 / 0.103 s (TNN C++ pipeline, WRN-16-8 / CIFAR-100)
PER_GPU_BATCH = 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="tnn", choices=["tnn", "reference", "torch"])
    ap.add_argument("--dp", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--model", default="cifar100_wrn16_8")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--host", default="auto", choices=["auto", "cpp", "python"],
                    help="which host drives the training step: the C++ framework (libtnn_b200 through tnn_b200._host; single GPU) "
                         "or the Python host (also the data-parallel path)")
    ap.add_argument("--no-gpt2", action="store_true", help="skip the GPT-2 inference block (BASELINE config 4)")
    ap.add_argument("--gpt2-iters", type=int, default=5)
    return ap.parse_args()


class ClockSampler:
    """SM clock / throttle-reason sampling DURING the timed region: NVML polled every 10 ms from a
    thread (nvidia-smi -lms as a fallback when pynvml is unavailable)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.t = gpu_index, [], None, None
        self.samples, self.stop_flag, self.nvml = [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, mx, pw, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            sm = sorted(s[0] for s in self.samples)
            reasons = sorted(k for k, bit in names.items() if any(s[3] & bit for s in self.samples))
            return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(max((s[1] for s in self.samples), default=0)) or None,
                    "reasons": reasons, "samples": len(sm), "power_w_max": max((s[2] for s in self.samples), default=None),
                    "source": "nvml@10ms"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, power = [], 0, set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None, "source": "nvidia-smi@100ms"}


def emit(d):
    _RESULT_OUT.write(json.dumps(d) + "\n")
    _RESULT_OUT.flush()


def run_reference(args):
    """Unmodified reference code from baseline/_ref through its own training loop (see
    baseline/reference_arm.py for what is run and why it is the reference's torch implementation)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_arm
    reference_arm.run(args, emit, dist_setup, barrier, max_over_ranks, ClockSampler, BASELINE_IMG_S)


def gpt2_block(args, impl: str, rank: int, world: int, dev) -> dict | None:
    """GPT-2 small inference bs=32 seq=1024 (BASELINE config 4) on every rank's GPU; tokens/s are the whole-job
    aggregate (N independent replicas = N x the slowest rank), times are the max over ranks."""
    if args.no_gpt2:
        return None
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import gpt2_lib
    import torch
    try:
        if impl in ("tnn_cpp", "tnn") and args.host != "python":
            # the C++ GPT2Engine (persistent one-launch-per-token decode kernel) on this rank's GPU
            try:
                try:
                    from tnn_b200 import _host  # noqa: F401  (prebuilt in-tree; ranks must not rebuild concurrently)
                except ImportError:
                    from tnn_b200 import build as tb
                    tb.build_host(verbose=False)
                r = gpt2_lib.tnn_gpt2_cpp(dev.index or 0, iters=args.gpt2_iters)
            except Exception as e:  # noqa: BLE001
                if impl == "tnn_cpp" or args.host == "cpp":
                    raise
                print(f"[bench] C++ GPT-2 engine unavailable ({type(e).__name__}: {e}); using the Python engine", file=sys.stderr)
                r = gpt2_lib.tnn_gpt2(dev, iters=args.gpt2_iters)
        elif impl in ("tnn_cpp", "tnn"):
            r = gpt2_lib.tnn_gpt2(dev, iters=args.gpt2_iters)
        else:
            r = gpt2_lib.torch_gpt2(dev, iters=args.gpt2_iters)
    except Exception as e:  # noqa: BLE001
        r = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.synchronize()
    if "error" not in r:
        for k in [k for k in r if k.endswith("_ms") or k.endswith("ms_per_step")]:
            r[k] = max_over_ranks(r[k], world, dev)
        B, L = r["batch"], r["seq"]
        for k in [k for k in r if k.endswith("tok_s")]:
            ms_key = k.replace("tok_s", "ms") if k.startswith("prefill") else "decode_ms_per_step"
            r[k] = world * (B * L if k.startswith("prefill") else B) / r[ms_key] * 1e3
        r["n_gpus"] = world
        r["unit"] = "tokens/s (whole job)"
    return r


def dist_setup(n):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, local, world


def max_over_ranks(ms: float, world: int, dev) -> float:
    import torch
    import torch.distributed as dist
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run_tnn_cpp(args):
    """Single-GPU arm driven by the C++ host: model, executor, workspace allocator, CUDA-graph capture, optimizer and the
    timing loop all live in libtnn_b200 (cpp/src/bench.cpp); Python only samples clocks and prints the line."""
    import torch
    from tnn_b200 import build as tb
    tb.build_host(verbose=False)
    from tnn_b200 import _host
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local)
    sampler.start()
    out = json.loads(_host.bench_train(args.model, args.batch, args.steps, args.warmup, local, True))
    clocks = sampler.stop()
    out["impl"] = "tnn_b200"
    out["clocks"] = clocks
    out["config"]["classes"] = 10 if "cifar10_" in args.model else 100
    out["config"]["image"] = [32, 32, 3]
    out["config"]["seq_len"] = None
    out["config"]["clock_sampling"] = "NVML over the device-timed AND the end-to-end region of the C++ benchmark call"
    gpt2 = gpt2_block(args, "tnn_cpp", 0, 1, dev)
    if gpt2 is not None:
        out["gpt2"] = gpt2
    emit(out)


def run_tnn(args):
    if args.host != "python" and int(os.environ.get("WORLD_SIZE", 1)) == 1 and args.dp == "fused":
        try:
            return run_tnn_cpp(args)
        except Exception as e:  # noqa: BLE001
            if args.host == "cpp":
                raise
            print(f"[bench] C++ host unavailable ({type(e).__name__}: {e}); falling back to the Python host", file=sys.stderr)
    import torch
    from tnn_b200.models import ExampleModels, load_or_create_model
    from tnn_b200.nn import LossFactory, OptimizerFactory
    from tnn_b200.nn.train import TrainStep
    from tnn_b200.ops import native

    rank, local, world = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    Cx = native.require()
    B = args.batch
    # reference trainer hyper-parameters: Adam(lr, 0.9, 0.999, eps 1e-3, wd 3e-4) (examples/trainer.cpp:67-71)
    opt = OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4)
    dp = None
    if world > 1 and args.dp == "fused":
        from tnn_b200.parallel.dp import FusedDataParallel
        dp = FusedDataParallel(opt)
        graph = dp.build_graph(ExampleModels.create(args.model), dev, seed=1234, name=args.model)
        parallelism = f"dp{world}-fused-p2p" + ("-nvls" if dp.mc else "")
    else:
        graph = load_or_create_model(args.model, device=dev, seed=1234)
        opt.attach(graph.context())
        if world > 1:
            from tnn_b200.parallel.dp import AllReduceDataParallel
            dp = AllReduceDataParallel(graph, opt)
        parallelism = f"dp{world}" + ("-nccl-allreduce" if world > 1 else "")
    crit = LossFactory.create_logsoftmax_crossentropy()
    step = TrainStep(graph, crit, opt, None, 1, dp)

    g = torch.Generator().manual_seed(100 + rank)
    NB = 4
    classes = 10 if "cifar10_" in args.model else 100
    host = [(torch.rand((B, 32, 32, 3), generator=g).pin_memory(), torch.randint(0, classes, (B,), generator=g, dtype=torch.int32).pin_memory())
            for _ in range(NB)]
    devb = [(x.to(dev), y.to(dev)) for x, y in host]

    # ---------------- device-resident timing
    for i in range(args.warmup):
        step(*devb[i % NB])
    # TrainStep captures its CUDA graph on the third call: keep going (untimed) until the REPLAYED graph has run at least
    # 8 times, so that the timed region starts on a settled graph (post-capture replays re-upload / re-validate the graph)
    settle = 0
    while step.use_graph and step.replays < 8 and settle < 16:
        step(*devb[settle % NB])
        settle += 1
    barrier(world)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = Cx.launch_count()
    r0 = step.replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        stats = step(*devb[i % NB])
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    launches = (Cx.launch_count() - l0) + (step.replays - r0) * step.captured_launches
    clocks = sampler.stop()
    loss_val = float(stats[0])
    value = world * B * args.steps / (ms / 1e3)

    # ---------------- end to end: pinned host -> device every step, loss read back every step
    e2e = None
    if not args.no_e2e:
        loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        copy_done = torch.cuda.Event()

        def run(n):
            last = 0.0
            for i in range(n):
                hx, hy = host[i % NB]
                st = step(hx, hy)                    # pinned host batch -> H2D inside the step
                if i > 0:
                    copy_done.synchronize()          # D2H read of the previous step's loss
                    last = float(loss_host[0])
                loss_host.copy_(st, non_blocking=True)
                copy_done.record()
            copy_done.synchronize()
            return float(loss_host[0]) + 0.0 * last

        run(max(3, args.warmup))
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        run(args.steps)
        e1.record()
        barrier(world)
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms2 = max_over_ranks(max(e0.elapsed_time(e1), wall_ms), world, dev)
        e2e = {"value": world * B * args.steps / (ms2 / 1e3), "unit": "img/s", "ms_per_step": ms2 / args.steps,
               "h2d_bytes_per_step": B * 32 * 32 * 3 * 4 + B * 4, "d2h_bytes_per_step": 8}

    dp_info = None
    if dp is not None and hasattr(dp, "bucket_report"):
        try:
            dp_info = dp.bucket_report(step._eager, *devb[0])
        except Exception as e:  # noqa: BLE001
            dp_info = {"error": f"{type(e).__name__}: {e}"[:200]}
    launches_wrn = int(launches)
    used_graph = bool(step.use_graph)
    del step, graph, devb
    torch.cuda.empty_cache()
    l0 = Cx.launch_count()
    gpt2 = gpt2_block(args, "tnn", rank, world, dev)
    launches_gpt2 = int(Cx.launch_count() - l0)
    if rank == 0:
        out = {
            "metric": ("WRN-16-8 CIFAR-100" if args.model == "cifar100_wrn16_8" else args.model) + " training throughput", "value": value, "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": value / BASELINE_IMG_S, "dtype": "bf16", "data": "synthetic",
            "impl": "tnn_b200",
            "config": {"model": args.model, "global_batch": world * B, "per_gpu_batch": B, "seq_len": None,
                       "image": [32, 32, 3], "classes": classes, "optimizer": "adam(lr1e-3,eps1e-3,wd3e-4)",
                       "parallelism": parallelism,
                       "l2": "inputs rotate over 4 batches; per-step activation working set (~2 GB) >> 126 MB L2"},
            "settle_replays": int(settle),
            "clocks": clocks, "gpu_launches": int(launches), "cuda_graph": used_graph, "launches_per_step": launches / max(1, args.steps),
            "final_loss": loss_val,
        }
        if e2e:
            out["e2e"] = e2e
        if dp_info is not None:
            out["dp"] = dp_info
        if gpt2 is not None:
            gpt2["gpu_launches_eager"] = launches_gpt2
            out["gpt2"] = gpt2
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_torch(args):
    """torch + cuDNN + NCCL re-creation of the same config (the 'reference-equivalent' measured
    baseline of BASELINE.md): bf16 autocast, channels_last, DDP for N > 1."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from torch_wrn import bench_torch_wrn
    rank, local, world = dist_setup(args.gpus)
    res = bench_torch_wrn(args.batch, args.steps, args.warmup, local, world, args.model)
    import torch
    torch.cuda.empty_cache()
    gpt2 = gpt2_block(args, "torch", rank, world, torch.device("cuda", local))
    if rank == 0:
        res.update({"impl": "torch", "vs_baseline": res["value"] / BASELINE_IMG_S})
        if gpt2 is not None:
            res["gpt2"] = gpt2
        emit(res)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "torch":
        run_torch(a)
    else:
        run_tnn(a)
