"""Wire format of the coordinator / worker runtime.

Reference: include/distributed/command_type.hpp:20-79 (CommandType, ordered by priority),
message.hpp:21-141 (Message = header{CommandType u16} + Variant payload), job.hpp:12-59
(Job{mb_id u64, Tensor}), stage_config.hpp:11-50, endpoint.hpp:19-236, packet.hpp:10-70.

Encoding (produced by the C++ ``Writer`` / parsed by the C++ ``Reader`` of
csrc/runtime/archiver.hpp, i.e. the same archiver that writes checkpoints):
  Message   := u16 command_type, u32 variant_index, payload
  payload   := 0 monostate: nothing | 1 Job: u64 mb_id, Tensor | 2 string: u64 len + bytes
             | 3 bool: u8 | 4 Profiler: string(serialised events) | 5 StageConfig: string(JSON)
  Tensor    := u32 dtype, u64 ndims + u64 dims[], raw bytes          (tensor.hpp:612-621)
A GPU tensor inside a Job is staged to the host by the writer (what the reference does for every
hop); on-node stage traffic does not use this path (parallel/pipeline.py moves it with NCCL).
"""
from __future__ import annotations

import enum
import json
import math
from dataclasses import dataclass, field

import torch

from ..core.config import TConfig
from ..core.dtype import DType, from_torch, to_torch
from ..ops import native


class CommandType(enum.IntEnum):
    _START = 0
    # core commands
    FORWARD_JOB = 1
    BACKWARD_JOB = 2
    BACKWARD_COMPLETE = 3
    UPDATE_PARAMETERS = 4
    # mode switching
    TRAIN_MODE = 5
    EVAL_MODE = 6
    SHUTDOWN = 7
    # configuration
    CONFIG_TRANSFER = 8
    CONFIG_RECEIVED = 9
    LOAD_PARAMS = 10
    PARAMS_LOADED = 11
    SEND_PARAMS = 12
    PARAMS_TRANSFER = 13
    SAVE_TO_FILE = 14
    # status and monitoring
    STATUS_REQUEST = 15
    STATUS_RESPONSE = 16
    PARAMETERS_UPDATED = 17
    HEALTH_CHECK = 18
    # error handling
    ERROR_REPORT = 19
    JOB_FAILURE = 20
    # advanced features
    BARRIER_SYNC = 21
    CHECKPOINT_REQUEST = 22
    CHECKPOINT_COMPLETE = 23
    # load balancing
    UPDATE_LOAD = 24
    REPORT_LOAD = 25
    LOAD_REPORT = 26
    # profiling
    START_PROFILING = 27
    PROFILING_STARTED = 28
    PRINT_PROFILING = 29
    PROFILING_PRINTED = 30
    CLEAR_PROFILING = 31
    PROFILING_CLEARED = 32
    REPORT_PROFILING = 33
    PROFILING_REPORTED = 34
    # handshake
    HANDSHAKE = 35
    HANDSHAKE_ACK = 36
    _COUNT = 37


class ParallelMode(enum.IntEnum):
    DATA = 0
    PIPELINE = 1


# ------------------------------------------------------------------------------- endpoints
class Endpoint:
    """Where a communicator can be reached: IN_PROCESS (object reference), TCP (host, port) or
    ROCE (kept for API parity; on a B200 node it resolves to the TCP control channel because the
    data path is NVLink)."""
    IN_PROCESS, TCP, ROCE = "IN_PROCESS", "TCP", "ROCE"
    _registry: dict[str, object] = {}

    def __init__(self, type: str = "TCP", **params):
        self.type = type
        self.params = dict(params)

    @staticmethod
    def tcp(host: str, port: int) -> "Endpoint":
        return Endpoint(Endpoint.TCP, host=str(host), port=int(port))

    @staticmethod
    def roce(host: str, port: int, device_name: str = "", gid_index: int = -1) -> "Endpoint":
        return Endpoint(Endpoint.ROCE, host=str(host), port=int(port), device_name=device_name, gid_index=int(gid_index))

    @staticmethod
    def in_process(comm) -> "Endpoint":
        key = f"inproc-{id(comm)}"
        Endpoint._registry[key] = comm
        return Endpoint(Endpoint.IN_PROCESS, key=key)

    def communicator(self):
        return Endpoint._registry.get(self.params.get("key", ""))

    def id(self) -> str:
        if self.type == Endpoint.IN_PROCESS:
            return self.params["key"]
        return f"{self.params.get('host')}:{self.params.get('port')}"

    def to_json(self) -> dict:
        return {"type": self.type, "parameters": dict(self.params)}

    @staticmethod
    def from_json(j) -> "Endpoint":
        if isinstance(j, str):
            j = json.loads(j)
        if not j:
            return Endpoint("NONE")
        return Endpoint(j.get("type", "TCP"), **j.get("parameters", {}))

    def __eq__(self, o):
        return isinstance(o, Endpoint) and self.type == o.type and self.params == o.params

    def __hash__(self):
        return hash((self.type, self.id()))

    def __bool__(self):
        return self.type != "NONE"

    def __repr__(self):
        return f"Endpoint({self.type}, {self.params})"


# ------------------------------------------------------------------------------- payloads
@dataclass
class Job:
    data: torch.Tensor
    mb_id: int = 0


# mb_id of a PARAMS_TRANSFER job that carries a GRADIENT slab (DATA-mode reduction over the bus)
GRAD_SLAB_TAG = (1 << 62) | 0x6772


@dataclass
class StageConfig:
    model_config: dict = field(default_factory=dict)        # Sequential TConfig JSON
    optimizer_config: dict = field(default_factory=dict)
    scheduler_config: dict = field(default_factory=dict)
    next_stage_endpoint: Endpoint = field(default_factory=lambda: Endpoint("NONE"))
    prev_stage_endpoint: Endpoint = field(default_factory=lambda: Endpoint("NONE"))
    coordinator_endpoint: Endpoint = field(default_factory=lambda: Endpoint("NONE"))
    stage_id: int = 0
    seed: int = 123456                                       # worker.hpp:98-130 seeds every stage with 123456
    device: str = ""                                         # "" = worker default

    def to_json(self) -> dict:
        return {"model_config": self.model_config, "optimizer_config": self.optimizer_config,
                "scheduler_config": self.scheduler_config, "next_stage_endpoint": self.next_stage_endpoint.to_json(),
                "prev_stage_endpoint": self.prev_stage_endpoint.to_json(),
                "coordinator_endpoint": self.coordinator_endpoint.to_json(), "stage_id": self.stage_id,
                "seed": self.seed, "device": self.device}

    @staticmethod
    def from_json(j) -> "StageConfig":
        if isinstance(j, (str, bytes)):
            j = json.loads(j)
        return StageConfig(j.get("model_config", {}), j.get("optimizer_config", {}), j.get("scheduler_config", {}),
                           Endpoint.from_json(j.get("next_stage_endpoint")), Endpoint.from_json(j.get("prev_stage_endpoint")),
                           Endpoint.from_json(j.get("coordinator_endpoint")), int(j.get("stage_id", 0)),
                           int(j.get("seed", 123456)), j.get("device", ""))


_V_NONE, _V_JOB, _V_STR, _V_BOOL, _V_PROF, _V_STAGE = range(6)


def _put_bytes(a, b: bytes):
    """std::string encoding of the archiver: u64 length + raw bytes (archiver.hpp:123-134)."""
    a.u64(len(b))
    a.raw(b)


class Message:
    """header{command_type} + one payload (None | Job | str | bool | Profiler | StageConfig)."""

    def __init__(self, command_type: CommandType = CommandType._START, payload=None, sender: str = ""):
        self.command_type = CommandType(command_type)
        self.payload = payload
        self.sender = sender            # filled in by the receiving communicator

    def has_type(self, t) -> bool:
        return isinstance(self.payload, t)

    def get(self, t=None):
        if t is not None and not isinstance(self.payload, t):
            raise TypeError(f"message {self.command_type.name} holds {type(self.payload).__name__}, not {t.__name__}")
        return self.payload

    def __repr__(self):
        return f"Message({self.command_type.name}, {type(self.payload).__name__})"

    # ---- archiver
    def _variant(self) -> int:
        from ..utils.profiler import Profiler
        p = self.payload
        if p is None:
            return _V_NONE
        if isinstance(p, Job):
            return _V_JOB
        if isinstance(p, bool):
            return _V_BOOL
        if isinstance(p, str):
            return _V_STR
        if isinstance(p, Profiler):
            return _V_PROF
        if isinstance(p, StageConfig):
            return _V_STAGE
        raise TypeError(f"unsupported payload {type(p)}")

    def _emit(self, a, sizing: bool):
        v = self._variant()
        a.u16(int(self.command_type))
        a.u32(v)
        p = self.payload
        if v == _V_JOB:
            a.u64(int(p.mb_id))
            t = p.data
            if t.is_cuda:
                t = t.cpu()                       # host staging (reference: Writer does cd_copy, io.hpp:46-55)
            t = t.contiguous()
            raw = t.view(torch.uint8) if t.dtype != torch.bool else t.to(torch.uint8)
            dt = int(from_torch(p.data.dtype))
            if sizing:
                a.tensor(dt, list(t.shape), raw.numel())
            else:
                a.tensor(dt, list(t.shape), raw.data_ptr(), raw.numel())
                self._keep = raw                  # keep alive until getvalue()
        elif v == _V_STR:
            _put_bytes(a, p.encode("utf-8"))
        elif v == _V_BOOL:
            a.boolean(bool(p))
        elif v == _V_PROF:
            _put_bytes(a, p.serialize())
        elif v == _V_STAGE:
            _put_bytes(a, json.dumps(p.to_json()).encode("utf-8"))

    def byte_size(self) -> int:
        s = native.require().Sizer()
        self._emit(s, True)
        return s.size()

    def serialize(self) -> bytes:
        w = native.require().Writer(self.byte_size())
        self._emit(w, False)
        return w.getvalue()

    @staticmethod
    def deserialize(data: bytes, swap: bool = False, device=None, sender: str = "") -> "Message":
        from ..utils.profiler import Profiler
        r = native.require().Reader(data, bool(swap))
        cmd = CommandType(r.u16())
        v = r.u32()
        payload = None
        if v == _V_JOB:
            mb = r.u64()
            dt = to_torch(DType(r.u32()))
            shape = r.sizes()
            t = torch.empty(shape, dtype=dt)
            n = t.numel() * t.element_size()
            if n:
                r.raw_into(t.data_ptr(), n, t.element_size())
            if device is not None and torch.device(device).type == "cuda":
                t = t.pin_memory().to(device, non_blocking=True)
            payload = Job(t, int(mb))
        elif v == _V_STR:
            payload = r.string().decode("utf-8")
        elif v == _V_BOOL:
            payload = r.boolean()
        elif v == _V_PROF:
            payload = Profiler.deserialize(r.string())
        elif v == _V_STAGE:
            payload = StageConfig.from_json(r.string().decode("utf-8"))
        return Message(cmd, payload, sender)
