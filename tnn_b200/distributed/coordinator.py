"""Coordinator: owns the full model and the partitioner, deploys stages to workers and drives
micro-batched training / validation.

Reference: include/distributed/coordinator.hpp:50-483 -- initialize_topology :398-445,
deploy_stages :368-395, async_train_batch :165-223 (inject all microbatch forwards, compute
loss + grad (x 1/M) on the coordinator as each output returns, wait for M BACKWARD_COMPLETE),
update_parameters :129-138, join :146-157, profiling fan-in :306-352; tcp_coordinator.hpp,
roce_coordinator.hpp; train loop include/distributed/train.hpp:19-128.

Additions over the reference: every wait has a deadline and raises with the list of
missing acknowledgements; ERROR_REPORT from a worker aborts the batch instead of being ignored;
``health_check`` / ``checkpoint`` / ``gather_parameters`` use the commands the reference only
declares.  In DATA mode the coordinator averages replica gradients (the reference has no
gradient reduction at all); on one NVSwitch node use parallel.dp.FusedDataParallel.
"""
from __future__ import annotations

import json
import threading
import time
from dataclasses import dataclass, field

import torch

from ..nn.blocks import Sequential
from ..nn.graph import Graph
from ..nn.loss import Loss
from ..parallel.partitioner import Partitioner, flatten_pipeline_layers
from ..utils.profiler import GlobalProfiler, Profiler
from .communicator import Communicator, InProcessCommunicator, TCPCommunicator
from .message import CommandType, Endpoint, Job, Message, ParallelMode, StageConfig
from .worker import Worker


class DistributedError(RuntimeError):
    pass


@dataclass
class CoordinatorConfig:
    parallel_mode: ParallelMode = ParallelMode.PIPELINE
    model: object = None                       # Graph or Sequential (full model)
    optimizer: object = None
    scheduler: object = None
    partitioner: Partitioner | None = None
    local_worker: Worker | None = None         # optional in-process stage (position = local_worker_position)
    local_worker_position: int = 0
    coordinator_endpoint: Endpoint | None = None
    worker_endpoints: list = field(default_factory=list)
    stage_devices: list = field(default_factory=list)   # optional per-stage device strings ("cuda:1")
    seed: int = 123456


class Coordinator:
    def __init__(self, config: CoordinatorConfig, communicator: Communicator | None = None, name: str = "coordinator"):
        self.cfg = config
        self.name = name
        self.comm = communicator or InProcessCommunicator(name)
        self.stage_endpoints: list[Endpoint] = []
        self.stage_configs: list[StageConfig] = []
        self.partitions = []
        self.num_stages = 0
        self.training = True
        self._local_thread: threading.Thread | None = None
        self.merged_profiler: Profiler | None = None
        self.timeout_deploy, self.timeout_update, self.timeout_profile, self.timeout_job = 60.0, 60.0, 30.0, 120.0

    # ------------------------------------------------------------------ topology
    def _model_layers(self):
        m = self.cfg.model
        if isinstance(m, Graph):
            layers = []
            for l in m.get_layers():
                layers += flatten_pipeline_layers(l)
            return layers
        return flatten_pipeline_layers(m)

    def initialize(self):
        """Partition the model and build one StageConfig per stage (initialize_topology)."""
        self.comm.start()
        cfg = self.cfg
        layers = self._model_layers()
        eps = list(cfg.worker_endpoints)       # how the coordinator reaches each stage
        public = list(cfg.worker_endpoints)    # how stages reach each other
        if cfg.local_worker is not None:
            lw = cfg.local_worker.comm
            lw.start()
            pos = min(cfg.local_worker_position, len(eps))
            eps.insert(pos, Endpoint.in_process(lw))      # the local stage is driven by reference
            public.insert(pos, lw.endpoint() if isinstance(lw, TCPCommunicator) else Endpoint.in_process(lw))
        if not eps:
            raise DistributedError("no worker endpoints")
        if cfg.parallel_mode == ParallelMode.PIPELINE:
            parts = cfg.partitioner.partition_model(layers)
            if len(parts) != len(eps):
                raise DistributedError(f"{len(parts)} partitions for {len(eps)} workers")
            stage_layers = Partitioner.split(layers, parts, fresh=True)
        else:
            parts = [cfg.partitioner.partition_model(layers)[0]] * len(eps)
            stage_layers = [Partitioner.split(layers, [parts[0]], fresh=True)[0] for _ in eps]
        self.partitions, self.stage_endpoints, self.num_stages = parts, eps, len(eps)
        me = self.coordinator_endpoint()
        self.stage_configs = []
        for i, ls in enumerate(stage_layers):
            seq = Sequential(ls, f"stage_{i}")
            pipeline = cfg.parallel_mode == ParallelMode.PIPELINE
            sc = StageConfig(
                model_config=seq.get_config().to_json(),
                optimizer_config=cfg.optimizer.get_config().to_json() if cfg.optimizer is not None else {},
                scheduler_config=_sched_json(cfg.scheduler),
                next_stage_endpoint=public[i + 1] if pipeline and i + 1 < len(eps) else Endpoint("NONE"),
                prev_stage_endpoint=public[i - 1] if pipeline and i > 0 else Endpoint("NONE"),
                coordinator_endpoint=me, stage_id=i, seed=cfg.seed,
                device=cfg.stage_devices[i] if i < len(cfg.stage_devices) else "")
            self.stage_configs.append(sc)
        return self

    def coordinator_endpoint(self) -> Endpoint:
        if self.cfg.coordinator_endpoint is not None and not isinstance(self.comm, InProcessCommunicator):
            return self.cfg.coordinator_endpoint
        return self.comm.endpoint()

    def deploy_stages(self):
        if self.cfg.local_worker is not None and self._local_thread is None:
            self.cfg.local_worker.start(background=True)
            self._local_thread = self.cfg.local_worker._thread
        for ep, sc in zip(self.stage_endpoints, self.stage_configs):
            self.comm.connect(ep, self.timeout_deploy)
            # a TCP stage must be able to reach an in-process coordinator: rewrite to our TCP endpoint
            self.comm.send_message(Message(CommandType.CONFIG_TRANSFER, sc), ep)
        self.join(CommandType.CONFIG_RECEIVED, self.num_stages, self.timeout_deploy, what="deploy_stages")
        return self

    def start(self):
        return self.initialize().deploy_stages()

    def stop(self):
        for ep in self.stage_endpoints:
            try:
                self.comm.send_message(Message(CommandType.SHUTDOWN), ep)
            except Exception:  # noqa: BLE001
                pass
        if self.cfg.local_worker is not None:
            self.cfg.local_worker.stop()
        self.comm.stop()

    # ------------------------------------------------------------------ waiting
    def _next(self, timeout: float, types=None) -> Message | None:
        msg = self.comm.dequeue_input_message(timeout=timeout, types=types)
        return msg

    def join(self, cmd: CommandType, expected: int, timeout: float, what: str = "") -> list[Message]:
        """Collect `expected` messages of type `cmd`; a worker ERROR_REPORT or the deadline raise."""
        got: list[Message] = []
        deadline = time.monotonic() + timeout
        while len(got) < expected:
            left = deadline - time.monotonic()
            if left <= 0:
                raise DistributedError(f"{what or cmd.name}: timed out after {timeout:.0f}s with {len(got)}/{expected} "
                                       f"{cmd.name} (received from {[m.sender for m in got]})")
            m = self._next(min(left, 0.5), types=[cmd, CommandType.ERROR_REPORT])
            if m is None:
                continue
            if m.command_type == CommandType.ERROR_REPORT:
                raise DistributedError(f"worker error during {what or cmd.name}: {m.payload}")
            got.append(m)
        return got

    def broadcast(self, cmd: CommandType, payload=None):
        for ep in self.stage_endpoints:
            self.comm.send_message(Message(cmd, payload), ep)

    # ------------------------------------------------------------------ data path
    def set_training(self, flag: bool):
        self.training = bool(flag)
        self.broadcast(CommandType.TRAIN_MODE if flag else CommandType.EVAL_MODE)

    def forward(self, x: torch.Tensor, mb_id: int = 0, stage: int = 0):
        self.comm.send_message(Message(CommandType.FORWARD_JOB, Job(x, mb_id)), self.stage_endpoints[stage])

    def backward(self, grad: torch.Tensor, mb_id: int = 0, stage: int | None = None):
        s = self.num_stages - 1 if stage is None else stage
        self.comm.send_message(Message(CommandType.BACKWARD_JOB, Job(grad, mb_id)), self.stage_endpoints[s])

    def update_parameters(self):
        if self.cfg.parallel_mode == ParallelMode.DATA and self.num_stages > 1:
            self._average_replica_gradients()
        self.broadcast(CommandType.UPDATE_PARAMETERS)
        self.join(CommandType.PARAMETERS_UPDATED, self.num_stages, self.timeout_update, what="update_parameters")

    def async_train_batch(self, mb_inputs: list, mb_labels: list, criterion: Loss) -> tuple[float, float]:
        """Semi-asynchronous micro-batched step; returns (sum of microbatch mean losses / M, corrects)."""
        M = len(mb_inputs)
        data_mode = self.cfg.parallel_mode == ParallelMode.DATA
        for i, x in enumerate(mb_inputs):
            self.forward(x, i, stage=(i % self.num_stages) if data_mode else 0)
        loss_total, corrects, outputs, completes = 0.0, 0.0, 0, 0
        deadline = time.monotonic() + self.timeout_job
        while outputs < M or completes < M:
            left = deadline - time.monotonic()
            if left <= 0:
                raise DistributedError(f"async_train_batch: timed out with {outputs}/{M} outputs and {completes}/{M} "
                                       f"BACKWARD_COMPLETE (a lost microbatch deadlocks the reference; here it raises)")
            m = self._next(min(left, 0.5), types=[CommandType.FORWARD_JOB, CommandType.BACKWARD_COMPLETE, CommandType.ERROR_REPORT])
            if m is None:
                continue
            if m.command_type == CommandType.ERROR_REPORT:
                raise DistributedError(f"worker error during async_train_batch: {m.payload}")
            if m.command_type == CommandType.BACKWARD_COMPLETE:
                completes += 1
                continue
            job: Job = m.payload
            y = mb_labels[job.mb_id].to(job.data.device, non_blocking=True)
            stats, grad = criterion.loss_and_grad(job.data, y, scale=1.0 / M)
            loss_total += float(stats[0]) / M
            corrects += float(stats[1])
            self.backward(grad, job.mb_id, stage=(job.mb_id % self.num_stages) if data_mode else None)
            outputs += 1
        return loss_total, corrects

    def async_val_batch(self, x: torch.Tensor, y: torch.Tensor, criterion: Loss) -> tuple[float, float]:
        self.forward(x, 0)
        m = self.join(CommandType.FORWARD_JOB, 1, self.timeout_job, what="async_val_batch")[0]
        stats, _ = criterion.loss_and_grad(m.payload.data, y.to(m.payload.data.device), want_grad=False)
        return float(stats[0]), float(stats[1])

    # ------------------------------------------------------------------ DATA mode gradient averaging
    def gather_parameters(self) -> list[torch.Tensor]:
        self.broadcast(CommandType.SEND_PARAMS)
        got = self.join(CommandType.PARAMS_TRANSFER, self.num_stages, self.timeout_update, what="gather_parameters")
        got.sort(key=lambda m: m.payload.mb_id)
        return [m.payload.data for m in got]

    def _average_replica_gradients(self):
        """All-reduce (sum of 1/M-scaled microbatch gradients = batch mean) across replicas.

        In-process replicas expose their gradient slab by reference (summed in place).  Remote (TCP / RoCE)
        replicas go through the message bus: SEND_PARAMS("grads") makes every worker reply PARAMS_TRANSFER with
        its gradient slab, the coordinator sums them and returns the total as PARAMS_TRANSFER tagged
        ``GRAD_SLAB_TAG``; workers overwrite their gradient slab and acknowledge with PARAMS_LOADED, so
        UPDATE_PARAMETERS applies the same batch gradient on every replica (the reference has no gradient
        reduction at all)."""
        workers = [ep.communicator() for ep in self.stage_endpoints if ep.type == Endpoint.IN_PROCESS]
        owners = [getattr(c, "_owner", None) for c in workers]
        if len(owners) == self.num_stages and all(o is not None and o.graph is not None for o in owners):
            slabs = [o.graph.ctx.grad_slab for o in owners]
            total = slabs[0].detach().clone().to("cpu", torch.float32)
            for s in slabs[1:]:
                total += s.detach().to("cpu", torch.float32)
            # every microbatch gradient already carries the 1/M factor (async_train_batch), so the
            # batch gradient is the SUM over replicas
            for s in slabs:
                s.copy_(total.to(s.device))
            return
        from .message import GRAD_SLAB_TAG
        self.broadcast(CommandType.SEND_PARAMS, "grads")
        got = self.join(CommandType.PARAMS_TRANSFER, self.num_stages, self.timeout_update, what="gather replica gradients")
        if len({m.payload.data.numel() for m in got}) != 1:
            raise DistributedError("DATA mode: replicas report gradient slabs of different sizes")
        total = got[0].payload.data.detach().to("cpu", torch.float32).clone()
        for m in got[1:]:
            total += m.payload.data.detach().to("cpu", torch.float32)
        self.broadcast(CommandType.PARAMS_TRANSFER, Job(total, GRAD_SLAB_TAG))
        self.join(CommandType.PARAMS_LOADED, self.num_stages, self.timeout_update, what="scatter summed gradients")

    # ------------------------------------------------------------------ monitoring
    def health_check(self, timeout: float = 5.0) -> list[dict]:
        self.broadcast(CommandType.HEALTH_CHECK)
        return [json.loads(m.payload) for m in self.join(CommandType.STATUS_RESPONSE, self.num_stages, timeout, "health_check")]

    def barrier(self, timeout: float = 30.0):
        self.broadcast(CommandType.BARRIER_SYNC)
        self.join(CommandType.BARRIER_SYNC, self.num_stages, timeout, "barrier")

    def checkpoint(self, path: str, timeout: float = 60.0) -> list[str]:
        self.broadcast(CommandType.CHECKPOINT_REQUEST, path)
        return sorted(m.payload for m in self.join(CommandType.CHECKPOINT_COMPLETE, self.num_stages, timeout, "checkpoint"))

    def save_to_file(self, path: str):
        self.broadcast(CommandType.SAVE_TO_FILE, path)

    def report_load(self, timeout: float = 10.0) -> list[dict]:
        self.broadcast(CommandType.REPORT_LOAD)
        return [json.loads(m.payload) for m in self.join(CommandType.LOAD_REPORT, self.num_stages, timeout, "report_load")]

    # ------------------------------------------------------------------ profiling
    def start_profiling(self):
        GlobalProfiler.init_start_time()
        self.broadcast(CommandType.START_PROFILING)
        self.join(CommandType.PROFILING_STARTED, self.num_stages, self.timeout_profile, "start_profiling")

    def fetch_profiling(self) -> Profiler:
        self.broadcast(CommandType.REPORT_PROFILING)
        got = self.join(CommandType.PROFILING_REPORTED, self.num_stages, self.timeout_profile, "fetch_profiling")
        merged = Profiler()
        merged.init_start_time()
        seen = set()
        for m in got:
            p = m.payload
            if id(p) in seen or (hasattr(p, "_c") and id(p._c) in seen):
                continue          # in-process workers share the process-global profiler
            seen.add(id(p))
            seen.add(id(getattr(p, "_c", p)))
            merged.merge(p)
        merged.sort()
        self.merged_profiler = merged
        return merged

    def print_profiling(self, path: str = "logs/profiler.log"):
        p = self.merged_profiler or self.fetch_profiling()
        p.log(path)
        return path

    def clear_profiling(self):
        self.broadcast(CommandType.CLEAR_PROFILING)
        self.join(CommandType.PROFILING_CLEARED, self.num_stages, self.timeout_profile, "clear_profiling")


def _sched_json(s):
    if s is None:
        return {}
    c = s.get_config()
    return c.to_json() if hasattr(c, "to_json") else c


class NetworkCoordinator(Coordinator):
    """Coordinator that listens on TCP (examples/tcp_coordinator.cpp:27-108)."""

    def __init__(self, config: CoordinatorConfig, host="127.0.0.1", port=0, device=None):
        comm = TCPCommunicator(host, port, device=device, name="coordinator")
        comm.start()
        config.coordinator_endpoint = comm.endpoint()
        super().__init__(config, comm)


class RoCECoordinator(Coordinator):
    """Coordinator on the rendezvous transport (examples/roce_coordinator.cpp); every worker of the job must be a
    ``RoCEWorker`` (the two transports frame their packets differently, as TCP and RoCE binaries do upstream)."""

    def __init__(self, config: CoordinatorConfig, host="127.0.0.1", port=0, device=None, slab_bytes=256 << 20):
        from .rendezvous import RendezvousCommunicator
        comm = RendezvousCommunicator(host, port, device=device, name="coordinator", slab_bytes=slab_bytes)
        comm.start()
        config.coordinator_endpoint = comm.endpoint()
        super().__init__(config, comm)
