"""Pipeline parallelism over NCCL p2p: one process per GPU (torchrun), one stage per rank, 1F1B.

This is the B200-native replacement for the reference's data path between pipeline stages
(serialise -> TCP/RoCE -> deserialise on every hop): activations and their
gradients move device-to-device with ``ncclSend/ncclRecv`` (``torch.distributed`` p2p, backend
``nccl``; ``gloo`` in the CPU tests), shapes and dtypes are static per schedule so no metadata
travels, and the receive for the next microbatch is posted before the current one is computed so
the NVLink transfer overlaps compute.  The schedule is non-interleaved 1F1B (the reference
injects all forwards first -- GPipe with eager backward, coordinator.hpp:165-223 -- which holds
M microbatches of activations on stage 0; 1F1B holds at most S - rank).

The model is split by a ``Partitioner`` exactly as the coordinator does (fresh layers from
config, per-stage seed => identical initial weights to the single-device model), every stage
owns its optimizer (worker.hpp:194-207) and the last stage computes loss and gradient scaled
by 1/M (coordinator.hpp:198,206).

Launch overhead: a stage runs ~150 small kernels per microbatch; issued one by one from Python the
stage is launch-bound long before the GPU is busy.  On a GPU every (forward, microbatch) and
(backward, microbatch) unit is therefore captured ONCE into a CUDA graph (after two eager
iterations) and replayed from then on: activations arrive by ``ncclRecv`` straight into the
graph's static input buffer, the graph's output buffer is what ``ncclSend`` reads, all graphs share
one memory pool and are captured -- and always replayed -- in schedule order, so a unit's scratch
memory can be reused by later units while the per-microbatch layer caches stay alive between a
forward and its backward.  ``TNN_PP_GRAPH=0`` keeps the eager path.

Failure detection (the reference's un-timed ``cv.wait`` deadlocks on a lost microbatch,
coordinator.hpp:182-184): a neighbour that died or fell out of the schedule leaves this rank waiting in a
send / receive forever.  Two layers catch that.  (1) Device side: torch's NCCL watchdog aborts the communicator
when a collective exceeds the process group's ``timeout`` -- pass one to ``init_process_group`` (the examples
use ``TNN_PP_TIMEOUT_SEC``, default 600).  (2) Host side: ``StallWatchdog`` below notices when THIS rank has sat in
one schedule step for ``stall_timeout`` seconds (``TNN_PP_WATCHDOG_SEC``, default 600, 0 = off), prints which
unit it is stuck in (phase, micro-batch, peer rank) and ends the process with exit code 86, so that a launcher
(torchrun) tears the job down instead of leaving N - 1 GPUs spinning.  The watchdog is a plain Python thread that
only reads a timestamp: it makes no CUDA or NCCL call of its own (safe next to stream capture).
"""
from __future__ import annotations

import os
import sys
import threading
import time

import torch
import torch.distributed as dist

from ..core import stepstate
from ..nn.blocks import Sequential
from ..nn.graph import Graph, graph_from_sequential
from ..nn.loss import Loss
from ..nn.optimizers import Optimizer, OptimizerFactory
from .partitioner import Partitioner, flatten_pipeline_layers


class StallWatchdog:
    """Ends the process when the owner reports no progress for ``timeout`` seconds while it is inside a schedule.

    ``enter(what)`` / ``progress(what)`` stamp the current position, ``leave()`` disarms; ``on_stall(message)`` defaults to
    printing the message and ``os._exit(86)`` (a hung rank cannot be recovered in-process: its peers are blocked in
    their own sends / receives and the launcher must take the job down)."""

    EXIT_CODE = 86

    def __init__(self, timeout: float, name: str, on_stall=None):
        self.timeout, self.name = float(timeout), name
        self.on_stall = on_stall or self._die
        self._lock = threading.Lock()
        self._since: float | None = None
        self._what = ""
        self._stop = threading.Event()
        self.fired = False
        self._thread = None
        if self.timeout > 0:
            self._thread = threading.Thread(target=self._run, name=f"{name}-stall-watchdog", daemon=True)
            self._thread.start()

    def enter(self, what):
        with self._lock:
            self._since, self._what = time.monotonic(), what

    progress = enter

    def leave(self):
        with self._lock:
            self._since = None

    def close(self):
        self._stop.set()

    def _die(self, message: str):
        print(message, file=sys.stderr, flush=True)
        os._exit(self.EXIT_CODE)

    def _run(self):
        while not self._stop.wait(min(1.0, max(0.05, self.timeout / 4))):
            with self._lock:
                since, what = self._since, self._what
            if since is not None and time.monotonic() - since > self.timeout:
                self.fired = True
                self.on_stall(f"[{self.name}] pipeline stall watchdog: no progress for {self.timeout:.0f} s in {' '.join(str(w) for w in what if w != '') if isinstance(what, tuple) else what}; a neighbouring stage "
                              f"died or left the schedule (TNN_PP_WATCHDOG_SEC changes the budget, 0 disables)")
                return


class PipelineParallel:
    def __init__(self, model: Sequential, partitioner: Partitioner, optimizer: Optimizer, criterion: Loss,
                 num_microbatches: int, microbatch_shape: tuple, device, seed: int = 123456, group=None,
                 scheduler_factory=None, stall_timeout: float | None = None, on_stall=None):
        assert dist.is_initialized(), "init torch.distributed first (nccl on GPUs, gloo on CPU)"
        self.group = group or dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.device = torch.device(device)
        self.M = int(num_microbatches)
        self.criterion = criterion
        layers = flatten_pipeline_layers(model)
        self.partitions = partitioner.partition_model(layers)
        if len(self.partitions) != self.world:
            raise ValueError(f"{len(self.partitions)} partitions for world size {self.world}")
        part = self.partitions[self.rank]
        mine = Partitioner.split(layers, [part], fresh=True)[0]
        self.stage = Sequential(mine, f"stage_{self.rank}")
        self.graph: Graph = graph_from_sequential(self.stage, self.device, seed, name=f"stage{self.rank}")
        self.optimizer = OptimizerFactory.create_from_config(optimizer.get_config())
        self.optimizer.attach(self.graph.context())
        self.scheduler = scheduler_factory(self.optimizer) if scheduler_factory else None
        # static boundary shapes for this microbatch size
        s = tuple(microbatch_shape)
        self.in_shape = self.out_shape = None
        for i, p in enumerate(self.partitions):
            if i == self.rank:
                self.in_shape = s
            for l in layers[p.start_offset:p.end]:
                s = tuple(l.output_shape(s))
            if i == self.rank:
                self.out_shape = s
        self.act_dtype = torch.bfloat16 if self.device.type == "cuda" else torch.float32
        self.is_first, self.is_last = self.rank == 0, self.rank == self.world - 1
        self.prev, self.next = self.rank - 1, self.rank + 1
        self.steps = 0
        self.p2p_bytes = 0
        self._fwd_i = self._bwd_i = 0
        self._rbufs = {"prev": {}, "next": {}}      # persistent receive buffers per microbatch slot
        self.use_graphs = self.device.type == "cuda" and os.environ.get("TNN_PP_GRAPH", "1") == "1"
        self._graph_mode = False
        self._graphs: dict = {}
        self._statics: dict = {}
        self._pool = None
        self._gstats = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._stat_scale = torch.tensor([1.0 / self.M, 1.0], dtype=torch.float32).to(self.device)
        self.graph_replays = 0
        if stall_timeout is None:
            stall_timeout = float(os.environ.get("TNN_PP_WATCHDOG_SEC", "600"))
        self.watchdog = StallWatchdog(stall_timeout, f"stage {self.rank}/{self.world}", on_stall)

    # ------------------------------------------------------------------ p2p helpers
    def _ops(self, send_next=None, send_prev=None, recv_prev=False, recv_next=False):
        ops, rp, rn = [], None, None
        if send_prev is not None:
            ops.append(dist.P2POp(dist.isend, send_prev, self.prev, self.group))
        if recv_prev:
            rp = self._recv_buf("prev", self._fwd_i, self.in_shape)
            ops.append(dist.P2POp(dist.irecv, rp, self.prev, self.group))
        if send_next is not None:
            ops.append(dist.P2POp(dist.isend, send_next, self.next, self.group))
        if recv_next:
            rn = self._recv_buf("next", self._bwd_i, self.out_shape)
            ops.append(dist.P2POp(dist.irecv, rn, self.next, self.group))
        if ops:
            # a tuple, formatted only if the watchdog fires (this runs once per p2p group on the hot path)
            self.watchdog.progress(("p2p", "forward mb", self._fwd_i, "backward mb", self._bwd_i, "send->next" if send_next is not None else "",
                                    "send->prev" if send_prev is not None else "", "recv<-prev" if recv_prev else "", "recv<-next" if recv_next else "",
                                    "prev/next rank", self.prev, self.next))
            for t in (send_next, send_prev):
                if t is not None:
                    self.p2p_bytes += t.numel() * t.element_size()
            for r in dist.batch_isend_irecv(ops):
                r.wait()          # NCCL: orders the CURRENT stream after the transfer; the host does not block
        return rp, rn

    def _recv_buf(self, side: str, mb: int, shape) -> torch.Tensor:
        """The receive buffer of microbatch slot ``mb`` (the same storage every iteration: it is the
        static input of that slot's CUDA graph and it is what the layer caches of the slot refer to)."""
        d = self._rbufs[side]
        b = d.get(mb)
        if b is None:
            b = d[mb] = torch.empty(shape, dtype=self.act_dtype, device=self.device)
        return b

    def _boundary(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(self.act_dtype).contiguous()

    # ------------------------------------------------------------------ CUDA-graph units
    def _static(self, kind: str, mb: int, src: torch.Tensor) -> torch.Tensor:
        """A fixed-address copy of a per-iteration input (first stage's images, last stage's labels)."""
        key = (kind, mb)
        b = self._statics.get(key)
        if b is None or b.shape != src.shape or b.dtype != src.dtype:
            if b is not None:
                raise ValueError(f"pipeline graphs were captured for {tuple(b.shape)} {b.dtype}, got {tuple(src.shape)} {src.dtype}")
            b = self._statics[key] = torch.empty_like(src, memory_format=torch.contiguous_format)
        b.copy_(src, non_blocking=True)
        return b

    def _run_graph(self, key, body):
        ent = self._graphs.get(key)
        if ent is None:
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, pool=self._pool, capture_error_mode="thread_local"):
                out = body()
            ent = self._graphs[key] = (cg, out)
        ent[0].replay()
        self.graph_replays += 1
        return ent[1]

    # ------------------------------------------------------------------ one microbatch of work
    def _forward(self, x, mb, labels, stats):
        if self._graph_mode:
            sx = self._static("x", mb, x) if self.is_first else x         # non-first: x IS the slot's receive buffer
            sl = self._static("y", mb, labels[mb]) if self.is_last else None

            def body():
                y = self.stage.forward(sx, mb)
                if self.is_last:
                    st, grad = self.criterion.loss_and_grad(y, sl, scale=1.0 / self.M)
                    self._gstats.add_(st.to(torch.float32) * self._stat_scale)
                    return grad
                return self._boundary(y)
            return self._run_graph(("f", mb), body)
        y = self.stage.forward(x, mb)
        if self.is_last:
            st, grad = self.criterion.loss_and_grad(y, labels[mb], scale=1.0 / self.M)
            stats += st.to(stats.dtype) * self._stat_scale.to(stats.dtype)
            return grad
        return self._boundary(y)

    def _backward(self, g, mb):
        if self._graph_mode:
            # g is the slot's receive buffer (or, on the last stage, the forward graph's own output)
            def body():
                dx = self.stage.backward(g, mb)
                return None if (self.is_first or dx is None) else self._boundary(dx)
            return self._run_graph(("b", mb), body)
        dx = self.stage.backward(g, mb)
        return None if (self.is_first or dx is None) else self._boundary(dx)

    # ------------------------------------------------------------------ schedule
    def train_batch(self, x: torch.Tensor | None, y: torch.Tensor | None, update: bool = True) -> torch.Tensor:
        """1F1B over ``M`` microbatches.  ``x`` is consumed on the first stage, ``y`` on the last.
        Returns the device tensor [mean loss, #correct] (meaningful on the last stage)."""
        M, S, r = self.M, self.world, self.rank
        self.watchdog.enter(f"train_batch {self.steps} (start)")
        xs = list(torch.tensor_split(x.to(self.device, non_blocking=True), M, 0)) if self.is_first else None
        ys = list(torch.tensor_split(y.to(self.device, non_blocking=True), M, 0)) if self.is_last else None
        if self.device.type == "cuda":
            stepstate.get(self.device).advance(None)         # new dropout masks every iteration
        # two eager iterations first: lazy workspaces, bf16 weight copies and the conv->BN statistics
        # fusion (which switches on from the second forward) must be settled before capture
        self._graph_mode = self.use_graphs and self.steps >= 2 and self.graph.training
        if self._graph_mode:
            stats = self._gstats
            stats.zero_()
        else:
            stats = torch.zeros(2, dtype=torch.float32, device=self.device)
        warm = min(S - r - 1, M)
        remaining = M - warm
        self._fwd_i = self._bwd_i = 0

        def next_input(recv):
            return xs[self._fwd_i] if self.is_first else recv

        # warm-up forwards
        for _ in range(warm):
            inp, _ = self._ops(recv_prev=not self.is_first)
            out = self._forward(next_input(inp), self._fwd_i, ys, stats)
            self._fwd_i += 1
            self._ops(send_next=out)
        inp = None
        if remaining > 0 and not self.is_first:
            inp, _ = self._ops(recv_prev=True)
        # steady state: one forward, one backward
        for k in range(remaining):
            out = self._forward(next_input(inp), self._fwd_i, ys, stats)
            self._fwd_i += 1
            if self.is_last:
                gout = out
            else:
                _, gout = self._ops(send_next=out, recv_next=True)
            gin = self._backward(gout, self._bwd_i)
            self._bwd_i += 1
            last = k == remaining - 1
            if self.is_first:
                inp = None
            elif last:
                self._ops(send_prev=gin)
                inp = None
            else:
                inp, _ = self._ops(send_prev=gin, recv_prev=True)
        # cool-down backwards
        for _ in range(warm):
            _, gout = self._ops(recv_next=True)
            gin = self._backward(gout, self._bwd_i)
            self._bwd_i += 1
            if not self.is_first:
                self._ops(send_prev=gin)
        if update:
            self.optimizer.update()
            self.optimizer.zero_grads()
            if self.scheduler is not None:
                self.scheduler.step()
        self.steps += 1
        self.watchdog.leave()
        return stats.clone() if self._graph_mode else stats

    def eval_batch(self, x, y):
        self.graph.set_training(False)
        self.watchdog.enter("eval_batch")
        inp, _ = self._ops(recv_prev=not self.is_first)
        out = self.stage.forward(x.to(self.device) if self.is_first else inp, 0)
        stats = torch.zeros(2, dtype=torch.float32, device=self.device)
        if self.is_last:
            st, _ = self.criterion.loss_and_grad(out, y.to(self.device), want_grad=False)
            stats += st
        else:
            self._ops(send_next=self._boundary(out))
        self.graph.set_training(True)
        self.watchdog.leave()
        return stats

    def broadcast_stats(self, stats: torch.Tensor) -> torch.Tensor:
        """Make the last stage's [loss, corrects] visible on every rank."""
        s = stats.clone()
        dist.broadcast(s, self.world - 1, group=self.group)
        return s
