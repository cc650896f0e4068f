"""Single-process training: TrainingConfig, TrainStep, train_model / validate_model.

Reference: include/nn/train.hpp:45-86, src/nn/train.cpp:50-428.  Differences by design: the step
issues no host synchronisation (the reference syncs the stream after every Sequential pass and
every step, train.cpp:200); loss / accuracy are accumulated on the device by the fused CE kernel
and read back every ``progress_print_interval`` steps; inputs are staged through pinned host
buffers on a copy stream, double buffered.
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import asdict, dataclass, field

import torch

from ..core import stepstate
from ..core.dtype import dtype_to_string, string_to_dtype, to_torch
from ..utils.env import Env
from .graph import Graph, GraphExecutor
from .loss import Loss
from .optimizers import Optimizer


@dataclass
class TrainingConfig:
    epochs: int = 10
    batch_size: int = 128
    max_steps: int = -1
    lr_initial: float = 1e-3
    gradient_accumulation_steps: int = 1
    progress_print_interval: int = 50
    num_threads: int = 8
    profiler_type: str = "NONE"            # NONE | NORMAL | CUMULATIVE
    print_layer_profiling: bool = False
    print_layer_memory_usage: bool = False
    model_name: str = "cifar10_resnet9"
    model_path: str = ""
    dataset_name: str = "cifar10"
    dataset_path: str = "data"
    device_type: str = "GPU"
    io_dtype: str = "FP32"
    param_dtype: str = "FP32"
    compute_dtype: str = "FP32"
    num_microbatches: int = 2
    snapshot_dir: str = "model_snapshots"
    seed: int = 0

    # env keys follow .env.example of the reference (src/nn/train.cpp:50-82)
    @staticmethod
    def load_from_env(path: str = ".env") -> "TrainingConfig":
        env = Env(path)
        c = TrainingConfig()
        c.epochs = env.get("EPOCHS", c.epochs, int)
        c.batch_size = env.get("BATCH_SIZE", c.batch_size, int)
        c.max_steps = env.get("MAX_STEPS", c.max_steps, int)
        c.lr_initial = env.get("LR_INITIAL", c.lr_initial, float)
        c.gradient_accumulation_steps = env.get("GRADIENT_ACCUMULATION_STEPS", c.gradient_accumulation_steps, int)
        c.progress_print_interval = env.get("PROGRESS_PRINT_INTERVAL", c.progress_print_interval, int)
        c.num_threads = env.get("NUM_THREADS", c.num_threads, int)
        c.profiler_type = env.get("PROFILER_TYPE", c.profiler_type, str)
        c.print_layer_profiling = env.get("PRINT_LAYER_PROFILING", c.print_layer_profiling, bool)
        c.print_layer_memory_usage = env.get("PRINT_LAYER_MEMORY_USAGE", c.print_layer_memory_usage, bool)
        c.model_name = env.get("MODEL_NAME", c.model_name, str)
        c.model_path = env.get("MODEL_PATH", c.model_path, str)
        c.dataset_name = env.get("DATASET_NAME", c.dataset_name, str)
        c.dataset_path = env.get("DATASET_PATH", c.dataset_path, str)
        c.device_type = env.get("DEVICE_TYPE", c.device_type, str)
        c.io_dtype = env.get("IO_DTYPE", c.io_dtype, str)
        c.param_dtype = env.get("PARAM_DTYPE", c.param_dtype, str)
        c.compute_dtype = env.get("COMPUTE_DTYPE", c.compute_dtype, str)
        c.num_microbatches = env.get("NUM_MICROBATCHES", c.num_microbatches, int)
        return c

    @staticmethod
    def load_from_json(path: str) -> "TrainingConfig":
        with open(path) as f:
            j = json.load(f)
        c = TrainingConfig()
        for k, v in j.items():
            if hasattr(c, k):
                setattr(c, k, type(getattr(c, k))(v))
        return c

    def to_json(self) -> dict:
        return asdict(self)

    def print_config(self, out=None):
        """One ``key: value`` line per field (train.hpp print_config)."""
        print("Training configuration:", file=out)
        for k, v in asdict(self).items():
            print(f"  {k}: {v}", file=out)

    def device(self) -> torch.device:
        t = self.device_type.upper()
        if t.startswith("GPU") or t.startswith("CUDA"):
            idx = int(t.split(":")[1]) if ":" in t else 0
            return torch.device("cuda", idx)
        return torch.device("cpu")


class TrainStep:
    """One optimisation step through the public API: forward, fused loss+grad, backward,
    (every ``accum`` calls) optimizer update + scheduler step.  Returns the device-resident
    [loss, n_correct] tensor; nothing here synchronises the host.

    On a GPU (and ``accum == 1``) the whole step is captured into ONE CUDA graph after two eager
    warm-up calls and replayed afterwards: per-step scalars (lr, Adam bias corrections, dropout
    seed counter) live in the device step state that the host refreshes with a single 64-byte
    H2D copy before each replay, so schedules and dropout masks keep evolving under replay.
    ``x`` / ``y`` may be device tensors or PINNED host tensors (copied straight into the graph's
    static input buffers)."""

    def __init__(self, graph: Graph, criterion: Loss, optimizer: Optimizer, scheduler=None, accum: int = 1,
                 dp=None, use_graph: bool | None = None):
        self.graph, self.criterion, self.optimizer, self.scheduler = graph, criterion, optimizer, scheduler
        self.executor = GraphExecutor(graph)
        self.accum = max(1, int(accum))
        self.dp = dp                      # parallel.dp.FusedDataParallel / AllReduceDataParallel or None
        self._micro = 0
        if optimizer.ctx is None:
            optimizer.attach(graph.context())
        optimizer.grad_scale = 1.0 / self.accum
        self.device = graph.device()
        if use_graph is None:
            use_graph = os.environ.get("TNN_CUDA_GRAPH", "1") == "1"
        # NCCL collectives are not captured here: the all-reduce baseline stays eager
        self.use_graph = bool(use_graph) and self.device.type == "cuda" and self.accum == 1 and \
            not type(dp).__name__.startswith("AllReduce")
        self._cg = None
        self._warm = 0
        self.captured_launches = 0
        self.replays = 0
        self._sx = self._sy = self._sstats = None
        self._state = stepstate.get(self.device)
        # fused DP with bucket overlap: the top-level Sequential reports per-layer backward completion
        if dp is not None and getattr(dp, "side", None) is not None and self.accum == 1:
            top = graph.get_layers()[0]
            if hasattr(top, "backward_hook"):
                top.backward_hook = dp.backward_hook

    # ---- the work itself (capturable: no host sync, no Python-visible per-step scalars)
    def _core(self, x, y, mb_id=0):
        out = self.executor.forward({"input": x}, mb_id=mb_id)["output"]
        stats, grad = self.criterion.loss_and_grad(out, y)
        self.executor.backward({"output": grad}, mb_id=mb_id)
        return stats

    def _update(self):
        self.optimizer.external_state = True
        try:
            if self.dp is not None:
                self.dp.step()
            else:
                self.optimizer.update()
        finally:
            self.optimizer.external_state = False

    def _eager(self, x, y, mb_id=0):
        if x.device != self.device:
            x, y = x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True)
        self._micro += 1
        do_update = self._micro % self.accum == 0
        self._state.advance(self.optimizer if do_update else None)
        stats = self._core(x, y, mb_id)
        if do_update:
            self._update()
            if self.scheduler is not None:
                self.scheduler.step()
        return stats

    def __call__(self, x: torch.Tensor, y: torch.Tensor, mb_id: int = 0) -> torch.Tensor:
        if not self.use_graph:
            return self._eager(x, y, mb_id)
        if self._cg is None:
            if self._warm < 2:
                self._warm += 1
                return self._eager(x, y, mb_id)
            self._capture(x, y)
        elif tuple(x.shape) != tuple(self._sx.shape):
            return self._eager(x, y, mb_id)      # ragged last batch: run it eagerly
        self._sx.copy_(x, non_blocking=True)
        self._sy.copy_(y, non_blocking=True)
        self._micro += 1
        self._state.advance(self.optimizer)
        self._cg.replay()
        self.replays += 1
        self.optimizer.step_count += 1
        if self.scheduler is not None:
            self.scheduler.step()
        return self._sstats

    def _capture(self, x, y):
        self._sx = torch.empty(x.shape, dtype=x.dtype, device=self.device)
        self._sy = torch.empty(y.shape, dtype=y.dtype, device=self.device)
        self._sx.copy_(x, non_blocking=True)
        self._sy.copy_(y, non_blocking=True)
        torch.cuda.synchronize(self.device)
        count = self.optimizer.step_count
        from ..ops import native
        l0 = native.require().launch_count()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            self._sstats = self._core(self._sx, self._sy, 0)
            self._update()
        self.captured_launches = int(native.require().launch_count() - l0)   # kernels per replay
        self.optimizer.step_count = count     # capture does not execute; replays advance it
        self._cg = cg


class DevicePrefetcher:
    """Stages (x, y) host batches through pinned buffers on a dedicated copy stream, two deep."""

    def __init__(self, device: torch.device):
        self.device = device
        self.cuda = device.type == "cuda"
        self.stream = torch.cuda.Stream(device) if self.cuda else None
        self.slots = [dict(), dict()]
        self.i = 0

    def stage(self, x: torch.Tensor, y: torch.Tensor):
        if not self.cuda:
            return x, y, None
        s = self.slots[self.i]
        self.i ^= 1
        if "ev_free" in s:
            s["ev_free"].synchronize()
        for k, t in (("x", x), ("y", y)):
            hk = "h" + k
            if hk not in s or s[hk].shape != t.shape or s[hk].dtype != t.dtype:
                s[hk] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                s["d" + k] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            if t.data_ptr() != s[hk].data_ptr():
                s[hk].copy_(t)
        with torch.cuda.stream(self.stream):
            s["dx"].copy_(s["hx"], non_blocking=True)
            s["dy"].copy_(s["hy"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        s["ev_ready"] = ev
        return s["dx"], s["dy"], s

    def acquire(self, slot):
        if slot is not None:
            torch.cuda.current_stream(self.device).wait_event(slot["ev_ready"])

    def release(self, slot):
        if slot is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            slot["ev_free"] = ev


def validate_model(graph: Graph, loader, criterion: Loss, cfg: TrainingConfig) -> tuple[float, float]:
    """Returns (mean loss, accuracy) over the loader (src/nn/train.cpp:388-428)."""
    dev = graph.device()
    graph.set_training(False)
    ex = GraphExecutor(graph)
    loader.reset()
    tot = torch.zeros(2, dtype=torch.float32, device=dev)
    n = batches = 0
    while True:
        b = loader.get_batch(cfg.batch_size)
        if b is None:
            break
        x, y = b
        out = ex.forward({"input": x.to(dev, non_blocking=True)})["output"]
        stats, _ = criterion.loss_and_grad(out, y.to(dev, non_blocking=True), want_grad=False)
        tot += stats.to(dev)
        n += y.numel()
        batches += 1
    graph.set_training(True)
    tot = tot.cpu()
    return (float(tot[0]) / max(batches, 1), float(tot[1]) / max(n, 1))


def train_model(graph: Graph, train_loader, val_loader, optimizer: Optimizer, criterion: Loss, scheduler=None,
                cfg: TrainingConfig | None = None, dp=None, log=print) -> dict:
    """Epoch + validation loop with best-validation checkpointing, or a fixed-step loop when
    ``cfg.max_steps > 0`` (src/nn/train.cpp:219-363)."""
    cfg = cfg or TrainingConfig()
    dev = graph.device()
    graph.set_training(True)
    lprof = None
    if cfg.print_layer_profiling or cfg.print_layer_memory_usage or str(cfg.profiler_type).upper() not in ("", "NONE"):
        from ..utils.profiler import LayerProfiler
        mode = str(cfg.profiler_type).upper()
        lprof = LayerProfiler("NORMAL" if mode in ("", "NONE") else mode)
        for layer in graph.get_layers():
            lprof.attach(layer)
    # per-layer events cannot be recorded inside a replayed CUDA graph: profile the eager step
    step_fn = TrainStep(graph, criterion, optimizer, scheduler, cfg.gradient_accumulation_steps, dp,
                        use_graph=False if lprof is not None else None)
    pre = DevicePrefetcher(dev)
    history = {"train_loss": [], "val_acc": [], "step_ms": []}
    best_acc = -1.0
    global_step = 0
    os.makedirs(cfg.snapshot_dir, exist_ok=True)
    epochs = cfg.epochs if cfg.max_steps <= 0 else 10 ** 9
    for epoch in range(epochs):
        train_loader.shuffle()
        train_loader.reset()
        acc = torch.zeros(2, dtype=torch.float32, device=dev)
        seen = 0
        epoch_batches = 0                 # acc is per epoch: divide by THIS epoch's batches
        t_epoch = time.time()
        t_last = time.time()
        nxt = train_loader.get_batch(cfg.batch_size)
        staged = pre.stage(*nxt) if nxt is not None else None
        while staged is not None:
            dx, dy, slot = staged
            nxt = train_loader.get_batch(cfg.batch_size)
            pre.acquire(slot)
            stats = step_fn(dx, dy)
            pre.release(slot)
            staged = pre.stage(*nxt) if nxt is not None else None
            acc += stats
            seen += dy.numel()
            global_step += 1
            epoch_batches += 1
            if cfg.progress_print_interval > 0 and global_step % cfg.progress_print_interval == 0:
                a = acc.cpu()
                dt = (time.time() - t_last) / cfg.progress_print_interval * 1e3
                t_last = time.time()
                history["step_ms"].append(dt)
                log(f"epoch {epoch + 1} step {global_step}: loss {float(a[0]) / max(1, epoch_batches):.4f} "
                    f"acc {100 * float(a[1]) / max(1, seen):.2f}% {dt:.2f} ms/step lr {optimizer.get_learning_rate():.2e}")
            if 0 < cfg.max_steps <= global_step:
                break
        history["train_loss"].append(float(acc[0].cpu()) / max(1, epoch_batches))
        log(f"epoch {epoch + 1} done in {time.time() - t_epoch:.2f}s")
        if lprof is not None:
            if cfg.print_layer_profiling or str(cfg.profiler_type).upper() not in ("", "NONE"):
                log(lprof.report(f"layer profile, epoch {epoch + 1}"))
            if cfg.print_layer_memory_usage:
                log(lprof.memory_report(f"layer memory, epoch {epoch + 1}"))
        if val_loader is not None:
            vloss, vacc = validate_model(graph, val_loader, criterion, cfg)
            history["val_acc"].append(vacc)
            log(f"validation: loss {vloss:.4f} acc {100 * vacc:.2f}%")
            if vacc > best_acc:
                best_acc = vacc
                save_checkpoint(graph, os.path.join(cfg.snapshot_dir, graph.name), optimizer, scheduler, global_step, dp)
        if 0 < cfg.max_steps <= global_step:
            save_checkpoint(graph, os.path.join(cfg.snapshot_dir, graph.name), optimizer, scheduler, global_step, dp)
            break
    if lprof is not None:
        history["layer_profile"] = lprof.rows()
        lprof.detach()
    return history


def save_checkpoint(graph: Graph, path: str, optimizer: Optimizer | None = None, scheduler=None, step: int = 0, dp=None):
    """Model file in the reference format (JSON + tensor records) plus an optional trailer with
    optimizer / scheduler state so that resume is exact.  With a data-parallel wrapper the call is
    COLLECTIVE (sharded optimizer state is gathered) and only rank 0 writes the file."""
    extra = {"step": int(step)}
    state_src = dp if (dp is not None and hasattr(dp, "state_dict")) else optimizer
    if state_src is not None:
        for k, v in state_src.state_dict().items():
            if v is not None:
                extra["opt." + k] = v
    if dp is not None and getattr(dp, "world", 1) > 1:
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_rank(getattr(dp, "group", None)) != 0:
            return
    if optimizer is not None:
        extra["opt.config"] = optimizer.get_config().to_json()
    if scheduler is not None:
        extra["sched.state"] = scheduler.state_dict()
        extra["sched.config"] = scheduler.get_config().to_json()
    graph.save_state(path, extra)


def optimizer_state_from_extra(extra: dict, device="cpu") -> dict:
    return {k[4:]: (v.to(device) if isinstance(v, torch.Tensor) else v)
            for k, v in extra.items() if k.startswith("opt.") and k != "opt.config" and v is not None}


def load_checkpoint(path: str, device="cpu", optimizer: Optimizer | None = None, scheduler=None, alloc_fn=None):
    graph, extra = Graph.load_state(path, device, alloc_fn, return_extra=True)
    if optimizer is not None:
        optimizer.attach(graph.context())
        optimizer.load_state_dict(optimizer_state_from_extra(extra, device))
    if scheduler is not None and "sched.state" in extra:
        scheduler.load_state_dict(extra["sched.state"])
    return graph, extra
