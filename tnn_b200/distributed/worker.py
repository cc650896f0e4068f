"""Pipeline-stage worker: a blocking event loop that executes coordinator commands.

Reference: include/distributed/worker.hpp:41-346 (event loop :64-89, set_config :98-130,
FORWARD_JOB :145-167, BACKWARD_JOB :168-193, UPDATE_PARAMETERS :194-207, profiling :226-236,
SAVE_TO_FILE :287-304, ERROR_REPORT :317-328).  Commands the reference declares but leaves as
``throw "Not implemented"`` (HEALTH_CHECK, STATUS_REQUEST, LOAD_PARAMS, SEND_PARAMS,
CHECKPOINT_REQUEST, BARRIER_SYNC, REPORT_LOAD) are implemented here.
"""
from __future__ import annotations

import json
import os
import threading
import time
import traceback

import torch

from ..core.config import TConfig
from ..nn.graph import Graph, graph_from_sequential
from ..nn.layer import LayerFactory
from ..nn.optimizers import OptimizerFactory
from ..nn.schedulers import SchedulerFactory
from ..utils.profiler import EventType, GlobalProfiler, Profiler
from .communicator import Communicator, InProcessCommunicator, TCPCommunicator
from .message import GRAD_SLAB_TAG, CommandType, Endpoint, Job, Message, StageConfig


def _parse_fault_env(spec: str):
    if not spec:
        return None
    p = spec.split(":")
    if len(p) < 3:
        return None
    return {"stage": p[0], "cmd": p[1].upper(), "action": p[2], "nth": int(p[3]) if len(p) > 3 else 0, "seen": 0}


class Worker:
    def __init__(self, communicator: Communicator, device="cpu", name: str = "worker"):
        self.comm = communicator
        self.comm._owner = self
        self.device = torch.device(device)
        self.name = name
        self.graph: Graph | None = None
        self.model = None
        self.optimizer = None
        self.scheduler = None
        self.config: StageConfig | None = None
        self.next_ep = self.prev_ep = self.coord_ep = Endpoint("NONE")
        self.should_stop = False
        self.training = True
        self._thread: threading.Thread | None = None
        self.stage_id = -1
        self._busy_ns = 0
        self._jobs = 0
        self._fwd_ns = self._bwd_ns = 0
        self._fwd_jobs = self._bwd_jobs = 0
        self._cpu_t0 = (time.process_time(), time.perf_counter())
        self.fault_inject = None          # callable(Message) -> None; tests use it to drop / delay / raise
        self._env_fault = _parse_fault_env(os.environ.get("TNN_FAULT_INJECT", ""))
        # directories a peer may name in SAVE_TO_FILE / LOAD_PARAMS / CHECKPOINT_REQUEST (the bus is unauthenticated):
        # the working directory, the system temp directory and TNN_WORKER_IO_ROOT (os.pathsep separated)
        import tempfile
        roots = [os.getcwd(), tempfile.gettempdir()] + [r for r in os.environ.get("TNN_WORKER_IO_ROOT", "").split(os.pathsep) if r]
        self.io_roots = [os.path.realpath(r) for r in roots]

    def _checked_path(self, path: str) -> str:
        """Reject paths outside the configured I/O roots (after resolving symlinks and '..')."""
        real = os.path.realpath(os.path.abspath(path))
        for root in self.io_roots:
            if real == root or real.startswith(root.rstrip(os.sep) + os.sep):
                return real
        raise PermissionError(f"worker {self.name}: path {path!r} is outside the allowed I/O roots {self.io_roots} "
                              f"(extend with TNN_WORKER_IO_ROOT)")

    # ------------------------------------------------------------------ lifecycle
    def start(self, background: bool = False):
        self.comm.start()
        if background:
            self._thread = threading.Thread(target=self.run, name=f"{self.name}-loop", daemon=True)
            self._thread.start()
            return self
        try:
            self.run()
        finally:
            self.comm.stop()
        return self

    def run(self):
        # the kernels of this repo launch on the CURRENT device's stream: a worker that owns cuda:1 must make it current in
        # the thread that runs its loop (the device is per thread in the CUDA runtime)
        if torch.device(self.device).type == "cuda":
            torch.cuda.set_device(self.device)
        while not self.should_stop:
            msg = self.comm.dequeue_input_message(timeout=0.25)
            if msg is None:
                continue
            try:
                if self.fault_inject is not None:
                    self.fault_inject(msg)
                if self._env_fault and self._apply_env_fault(msg):
                    continue
                self.process_message(msg)
            except Exception as e:  # noqa: BLE001
                err = f"[{self.name} stage {self.stage_id}] {type(e).__name__}: {e}\n{traceback.format_exc(limit=4)}"
                if self.coord_ep:
                    try:
                        self.comm.send_message(Message(CommandType.ERROR_REPORT, err), self.coord_ep)
                    except Exception:  # noqa: BLE001
                        pass

    def _apply_env_fault(self, msg: Message) -> bool:
        """TNN_FAULT_INJECT="<stage|*>:<COMMAND>:<drop|delay=SEC|raise|kill>[:<nth>]" (SURVEY 5.3).
        Returns True when the message must be dropped."""
        f = self._env_fault
        if f["stage"] not in ("*", str(self.stage_id)) or msg.command_type.name != f["cmd"]:
            return False
        f["seen"] += 1
        if f["nth"] and f["seen"] != f["nth"]:
            return False
        act = f["action"]
        if act == "drop":
            return True
        if act.startswith("delay="):
            time.sleep(float(act.split("=", 1)[1]))
            return False
        if act == "kill":
            os._exit(17)
        raise RuntimeError(f"TNN_FAULT_INJECT raised on {msg.command_type.name}")

    def stop(self):
        self.should_stop = True
        if self._thread is not None:
            self._thread.join(timeout=3)
        self.comm.stop()

    # ------------------------------------------------------------------ configuration
    def set_config(self, cfg: StageConfig):
        self.config = cfg
        self.stage_id = cfg.stage_id
        if cfg.device:
            self.device = torch.device(cfg.device)
        seq = LayerFactory.create(TConfig.from_json(cfg.model_config))
        self.model = seq
        self.graph = graph_from_sequential(seq, self.device, cfg.seed, name=f"stage{cfg.stage_id}")
        self.graph.set_training(True)
        if cfg.optimizer_config:
            self.optimizer = OptimizerFactory.create_from_config(cfg.optimizer_config)
            self.optimizer.attach(self.graph.context())
        if cfg.scheduler_config and self.optimizer is not None:
            self.scheduler = SchedulerFactory.create_from_config(cfg.scheduler_config, self.optimizer)
        self.coord_ep, self.next_ep, self.prev_ep = cfg.coordinator_endpoint, cfg.next_stage_endpoint, cfg.prev_stage_endpoint
        for ep in (self.coord_ep, self.next_ep, self.prev_ep):
            if ep and ep.type != Endpoint.IN_PROCESS:
                self.comm.connect(ep)

    def load_tracker(self) -> dict:
        """The reference's LoadTracker record (include/distributed/load_tracker.hpp:10-22): average forward /
        backward time per job in ms, CPU utilisation of this worker process (0..1 per core since start; -1 when
        unavailable) and peak memory in MB (host RSS, plus the CUDA allocator peak on a GPU stage)."""
        cpu0, wall0 = self._cpu_t0
        wall = time.perf_counter() - wall0
        cpu = (time.process_time() - cpu0) / wall if wall > 0 else -1.0
        try:
            import resource
            rss_mb = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0
        except Exception:  # noqa: BLE001
            rss_mb = -1.0
        gpu_mb = torch.cuda.max_memory_allocated(self.device) / 2 ** 20 if self.device.type == "cuda" else 0.0
        return {"avg_forward_time": self._fwd_ns / 1e6 / max(1, self._fwd_jobs),
                "avg_backward_time": self._bwd_ns / 1e6 / max(1, self._bwd_jobs),
                "avg_cpu_utilization": cpu, "max_memory_usage": rss_mb, "max_device_memory_usage": gpu_mb}

    # ------------------------------------------------------------------ dispatch
    def _reply(self, cmd: CommandType, payload=None, to: Endpoint | None = None):
        self.comm.send_message(Message(cmd, payload), to or self.coord_ep)

    def process_message(self, msg: Message):
        c = msg.command_type
        if c == CommandType.FORWARD_JOB:
            job: Job = msg.get(Job)
            t0 = Profiler.now_ns()
            x = job.data.to(self.device, non_blocking=True) if job.data.device != self.device else job.data
            y = self.model.forward(x, job.mb_id)
            GlobalProfiler.add_event(EventType.COMPUTE, t0, Profiler.now_ns(), f"forward mb {job.mb_id}", self.name)
            self._busy_ns += Profiler.now_ns() - t0
            self._fwd_ns += Profiler.now_ns() - t0
            self._fwd_jobs += 1
            self._jobs += 1
            self._reply(CommandType.FORWARD_JOB, Job(y, job.mb_id), self.next_ep if self.next_ep else self.coord_ep)
        elif c == CommandType.BACKWARD_JOB:
            job = msg.get(Job)
            t0 = Profiler.now_ns()
            g = job.data.to(self.device, non_blocking=True) if job.data.device != self.device else job.data
            dx = self.model.backward(g, job.mb_id)
            GlobalProfiler.add_event(EventType.COMPUTE, t0, Profiler.now_ns(), f"backward mb {job.mb_id}", self.name)
            self._busy_ns += Profiler.now_ns() - t0
            self._bwd_ns += Profiler.now_ns() - t0
            self._bwd_jobs += 1
            if self.prev_ep and dx is not None:
                self._reply(CommandType.BACKWARD_JOB, Job(dx, job.mb_id), self.prev_ep)
            else:
                self._reply(CommandType.BACKWARD_COMPLETE, None)
        elif c == CommandType.UPDATE_PARAMETERS:
            t0 = Profiler.now_ns()
            if self.optimizer is not None:
                self.optimizer.update()
                self.optimizer.zero_grads()
                if self.scheduler is not None:
                    self.scheduler.step()
            if self.graph is not None and self.graph.device().type == "cuda":
                from ..core import stepstate
                stepstate.get(self.graph.device()).advance(None)     # fresh dropout masks for the next batch
            GlobalProfiler.add_event(EventType.COMPUTE, t0, Profiler.now_ns(), "update", self.name)
            self._reply(CommandType.PARAMETERS_UPDATED, True)
        elif c == CommandType.TRAIN_MODE:
            self.training = True
            if self.graph is not None:
                self.graph.set_training(True)
        elif c == CommandType.EVAL_MODE:
            self.training = False
            if self.graph is not None:
                self.graph.set_training(False)
        elif c == CommandType.CONFIG_TRANSFER:
            self.set_config(msg.get(StageConfig))
            self._reply(CommandType.CONFIG_RECEIVED, True)
        elif c == CommandType.SHUTDOWN:
            self.should_stop = True
        elif c == CommandType.START_PROFILING:
            GlobalProfiler.get_profiler().clear()
            GlobalProfiler.init_start_time()
            self._reply(CommandType.PROFILING_STARTED, True)
        elif c == CommandType.REPORT_PROFILING:
            self._reply(CommandType.PROFILING_REPORTED, GlobalProfiler.get_profiler())
        elif c == CommandType.CLEAR_PROFILING:
            GlobalProfiler.get_profiler().clear()
            self._reply(CommandType.PROFILING_CLEARED, True)
        elif c == CommandType.PRINT_PROFILING:
            for e in GlobalProfiler.get_profiler().events():
                print(f"[{self.name}] {e['type']} {e['name']} {(e['end_ns'] - e['start_ns']) / 1e6:.3f} ms")
            self._reply(CommandType.PROFILING_PRINTED, True)
        elif c == CommandType.SAVE_TO_FILE:
            path = self._checked_path(msg.get(str))
            self.graph.save_state(path + f".stage{self.stage_id}" if not path.endswith(f".stage{self.stage_id}") else path)
        elif c == CommandType.LOAD_PARAMS:
            path = self._checked_path(msg.get(str))
            with open(path, "rb") as f:
                data = f.read()
            _, end = json.JSONDecoder().raw_decode(data.decode("latin-1"))
            import io
            self.graph.load_state_tensors(io.BytesIO(data[end:]))
            self._reply(CommandType.PARAMS_LOADED, True)
        elif c == CommandType.SEND_PARAMS:
            # payload "grads" selects the gradient slab (DATA-mode reduction over the bus), default = parameters
            what = msg.get(str) if msg.has_type(str) else "params"
            slab = self.graph.ctx.grad_slab if what == "grads" else self.graph.ctx.param_slab
            self._reply(CommandType.PARAMS_TRANSFER, Job(slab.detach().clone(), self.stage_id))
        elif c == CommandType.PARAMS_TRANSFER:
            job = msg.get(Job)
            if job.mb_id == GRAD_SLAB_TAG:
                self.graph.ctx.grad_slab.copy_(job.data.to(self.device))
            else:
                self.graph.ctx.param_slab.copy_(job.data.to(self.device))
                self.graph.ctx.refresh_lp()
            self._reply(CommandType.PARAMS_LOADED, True)
        elif c in (CommandType.HEALTH_CHECK, CommandType.STATUS_REQUEST):
            status = {"name": self.name, "stage": self.stage_id, "configured": self.graph is not None, "training": self.training,
                      "jobs": self._jobs, "device": str(self.device), "time": time.time()}
            self._reply(CommandType.STATUS_RESPONSE, json.dumps(status))
        elif c == CommandType.CHECKPOINT_REQUEST:
            path = self._checked_path(msg.get(str))
            from ..nn.train import save_checkpoint
            save_checkpoint(self.graph, f"{path}.stage{self.stage_id}", self.optimizer, self.scheduler,
                            self.optimizer.step_count if self.optimizer else 0)
            self._reply(CommandType.CHECKPOINT_COMPLETE, f"{path}.stage{self.stage_id}")
        elif c == CommandType.BARRIER_SYNC:
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            self._reply(CommandType.BARRIER_SYNC, True)
        elif c == CommandType.REPORT_LOAD:
            avg_ms = self._busy_ns / 1e6 / max(1, self._jobs)
            self._reply(CommandType.LOAD_REPORT, json.dumps({"stage": self.stage_id, "avg_forward_backward_ms": avg_ms,
                                                             "jobs": self._jobs, **self.load_tracker()}))
        elif c == CommandType.HANDSHAKE:
            self._reply(CommandType.HANDSHAKE_ACK, self.name, Endpoint.from_json(json.loads(msg.get(str))) if msg.has_type(str) else None)
        else:
            raise NotImplementedError(f"worker does not handle {c.name}")


class InProcessWorker(Worker):
    def __init__(self, device="cpu", name="local-worker"):
        super().__init__(InProcessCommunicator(name), device, name)


class TCPWorker(Worker):
    """``tcp_worker <port> [--gpu]`` of the reference (examples/tcp_worker.cpp:121-141)."""

    def __init__(self, host="127.0.0.1", port=0, use_gpu=False, gpu_index=0, io_threads=4, name=None):
        dev = torch.device("cuda", gpu_index) if use_gpu and torch.cuda.is_available() else torch.device("cpu")
        comm = TCPCommunicator(host, port, io_threads, device=dev, name=name or f"tcp-worker-{port}")
        super().__init__(comm, dev, name or f"tcp-worker-{port}")


class RoCEWorker(Worker):
    """``roce_worker`` of the reference (examples/roce_worker.cpp): same worker over the one-sided rendezvous
    transport (``distributed/rendezvous.py`` -- registered slab, PREPARE / READY_TO_WRITE / write + immediate)."""

    def __init__(self, host="127.0.0.1", port=0, use_gpu=False, gpu_index=0, io_threads=4, name=None, slab_bytes=256 << 20):
        from .rendezvous import RendezvousCommunicator
        dev = torch.device("cuda", gpu_index) if use_gpu and torch.cuda.is_available() else torch.device("cpu")
        comm = RendezvousCommunicator(host, port, io_threads, device=dev, name=name or f"roce-worker-{port}", slab_bytes=slab_bytes)
        super().__init__(comm, dev, name or f"roce-worker-{port}")
