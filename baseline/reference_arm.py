"""Reference arm of bench.py: drives the UNMODIFIED reference code from ``baseline/_ref``.

The reference (tungphambasement/TNN) is a C++/CMake tree.  Its C++ build cannot be produced
offline (``include/nn/layer.hpp:16`` includes ``device/engine.hpp`` which the snapshot does not
contain; CMake FetchContent needs the network; libibverbs / TBB are REQUIRED and absent), and
``pip install /root/reference`` fails because there is no setup.py / pyproject.toml.  What the
reference DOES ship and what runs here is its own torch baseline of the headline model:
``torch/torch_wrn16_8_cifar100.py`` (``WRN16_8CIFAR100``, ``WideResidualBlock``) with the training
loop of its ``main()`` (``:176-253``): ``optimizer.zero_grad(); outputs = model(inputs);
loss = criterion(outputs, targets); loss.backward(); optimizer.step(); loss.item()`` with
``nn.CrossEntropyLoss`` and ``optim.Adam(lr, betas=(0.9, 0.999), eps=1e-7, weight_decay=5e-4,
amsgrad=True)``.  This module imports that file verbatim (no edits, nothing of tnn_b200 on the
path) and times that loop on synthetic CIFAR-shaped batches in bf16 autocast (BASELINE.json names
the config "WRN-16-8 CIFAR-100 bf16"); N > 1 wraps the model in torch DDP over NCCL, the data-parallel
mechanism of the reference's ``torch/torch_resnet9_deepspeed.py`` (ZeRO stage 0 == bucketed all-reduce).
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"
WRN_FILE = os.path.join("torch", "torch_wrn16_8_cifar100.py")


def ensure_installed() -> str | None:
    """Return None when baseline/_ref holds the reference tree, else a one-line reason."""
    if os.path.isfile(os.path.join(REF, WRN_FILE)):
        return None
    if os.path.isdir(REF_SRC):
        try:
            shutil.copytree(REF_SRC, REF, dirs_exist_ok=True)
        except Exception as e:  # noqa: BLE001
            return f"copy of {REF_SRC} to baseline/_ref failed: {e}"
        if os.path.isfile(os.path.join(REF, WRN_FILE)):
            return None
    return "baseline/_ref is empty and /root/reference is not mounted on this box"


def load_reference_module():
    """Import baseline/_ref/torch/torch_wrn16_8_cifar100.py unmodified."""
    try:
        import dotenv  # noqa: F401
    except Exception:  # noqa: BLE001  (the file only calls load_dotenv() at import)
        stub = types.ModuleType("dotenv")
        stub.load_dotenv = lambda *a, **k: False
        sys.modules["dotenv"] = stub
    path = os.path.join(REF, WRN_FILE)
    spec = importlib.util.spec_from_file_location("tnn_reference_torch_wrn16_8", path)
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))            # load_dotenv() looks for ./.env
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def run(args, emit, dist_setup, barrier, max_over_ranks, ClockSampler, baseline_img_s):
    why = ensure_installed()
    if why is not None:
        emit({"impl": "reference", "unavailable": why})
        return
    import torch
    import torch.nn as nn
    import torch.optim as optim

    if not torch.cuda.is_available():
        emit({"impl": "reference", "unavailable": "no CUDA device on this box (the reference arm is installed: "
              "baseline/_ref/torch/torch_wrn16_8_cifar100.py)"})
        return
    rank, local, world = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    ref = load_reference_module()
    loaded_native = [m for m in sys.modules if m.startswith("tnn_b200")]
    assert not loaded_native, f"reference arm must not import tnn_b200: {loaded_native}"
    B = args.batch
    torch.manual_seed(1234)
    model = ref.WRN16_8CIFAR100(num_classes=100).to(dev)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    # torch_wrn16_8_cifar100.py:193-203 verbatim
    criterion = nn.CrossEntropyLoss()
    optimizer = optim.Adam(model.parameters(), lr=0.001, betas=(0.9, 0.999), eps=1e-7, weight_decay=5e-4,
                           amsgrad=True)
    model.train()
    g = torch.Generator().manual_seed(100 + rank)
    NB = 4
    host = [(torch.rand((B, 3, 32, 32), generator=g).pin_memory(),
             torch.randint(0, 100, (B,), generator=g, dtype=torch.int64).pin_memory()) for _ in range(NB)]
    devb = [(x.to(dev), y.to(dev)) for x, y in host]

    def train_step(inputs, targets):
        optimizer.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outputs = model(inputs)
            loss = criterion(outputs, targets)
        loss.backward()
        optimizer.step()
        return loss

    W = max(3, args.warmup)
    for i in range(W):
        train_step(*devb[i % NB])
    barrier(world)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = train_step(*devb[i % NB])
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    clocks = sampler.stop()
    value = world * B * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        def run_e2e(n):
            last = 0.0
            for i in range(n):
                hx, hy = host[i % NB]
                inputs, targets = hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True)   # main(): inputs.to(device)
                last = train_step(inputs, targets).item()                                          # main(): loss.item()
            return last
        run_e2e(W)
        barrier(world)
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(args.steps)
        e1.record()
        barrier(world)
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms2 = max_over_ranks(max(e0.elapsed_time(e1), wall_ms), world, dev)
        e2e = {"value": world * B * args.steps / (ms2 / 1e3), "unit": "img/s", "ms_per_step": ms2 / args.steps,
               "h2d_bytes_per_step": B * 3 * 32 * 32 * 4 + B * 8, "d2h_bytes_per_step": 4}

    if rank == 0:
        out = {
            "metric": "WRN-16-8 CIFAR-100 training throughput", "value": value, "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": value / baseline_img_s, "dtype": "bf16", "data": "synthetic",
            "impl": "reference",
            "config": {"model": "cifar100_wrn16_8", "global_batch": world * B, "per_gpu_batch": B, "seq_len": None,
                       "image": [32, 32, 3], "classes": 100,
                       "optimizer": "adam(lr1e-3,eps1e-7,wd5e-4,amsgrad) [torch_wrn16_8_cifar100.py:196-203]",
                       "parallelism": f"ddp{world}" if world > 1 else "dp1",
                       "reference_class": "baseline/_ref/torch/torch_wrn16_8_cifar100.py::WRN16_8CIFAR100 (the reference's own "
                                          "shipped torch implementation of the headline model, imported unmodified; its C++ tree cannot "
                                          "compile offline: include/nn/layer.hpp:16 includes a header missing from the snapshot, CMake "
                                          "FetchContent needs the network, libibverbs/TBB absent)",
                       "stack": "torch eager + cuDNN + NCCL DDP, bf16 autocast, NCHW as in the reference script",
                       "l2": "inputs rotate over 4 batches; per-step activation working set >> 126 MB L2"},
            "clocks": clocks, "gpu_launches": 0, "final_loss": float(loss),
            "gpt2": {"unavailable": "the reference ships GPT-2 only in its C++ tree (examples/gpt2_inference.cpp), which cannot be built offline"},
        }
        if e2e:
            out["e2e"] = e2e
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
