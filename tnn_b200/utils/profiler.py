"""Event profiler (reference: include/profiling/event.hpp:11-62, profiler.hpp:13-157).

Host spans live in the native thread-safe store (csrc/runtime ProfilerCore): ``Event{type in
COMPUTE|COMMUNICATION|OTHER, start, end, name, source}`` on the system clock, ``merge`` re-bases
another rank's events by the start-time delta, the whole profiler is serialisable so workers can
ship it to the coordinator.  GPU work is timed with ``cudaEvent`` pairs on the launching stream
(``device_span``) -- the reference has no GPU timing at all."""
from __future__ import annotations

import contextlib
import enum
import os

import torch

from ..ops import native


class EventType(enum.IntEnum):
    COMPUTE = 0
    COMMUNICATION = 1
    OTHER = 2


class Profiler:
    def __init__(self, core=None):
        self._c = core if core is not None else native.require().ProfilerCore()
        self._pending = []       # (type, name, source, ev0, ev1, host_start_ns)

    def init_start_time(self):
        self._c.init_start_time()

    @staticmethod
    def now_ns() -> int:
        return native.require().ProfilerCore.now_ns()

    def add_event(self, type: EventType, start_ns: int, end_ns: int, name: str, source: str = ""):
        self._c.add_event(int(type), int(start_ns), int(end_ns), name, source)

    @contextlib.contextmanager
    def span(self, name: str, type: EventType = EventType.COMPUTE, source: str = ""):
        t0 = self.now_ns()
        try:
            yield
        finally:
            self.add_event(type, t0, self.now_ns(), name, source)

    @contextlib.contextmanager
    def device_span(self, name: str, type: EventType = EventType.COMPUTE, source: str = "", stream=None):
        """Times the enclosed GPU work with CUDA events; resolved lazily by ``flush_device``."""
        if not torch.cuda.is_available():
            with self.span(name, type, source):
                yield
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = self.now_ns()
        e0.record(stream)
        try:
            yield
        finally:
            e1.record(stream)
            self._pending.append((type, name, source, e0, e1, t0))

    def flush_device(self):
        for type, name, source, e0, e1, t0 in self._pending:
            e1.synchronize()
            self.add_event(type, t0, t0 + int(e0.elapsed_time(e1) * 1e6), name, source)
        self._pending.clear()

    def merge(self, other: "Profiler"):
        self._c.merge(other._c)

    def sort(self):
        self._c.sort()

    def clear(self):
        self._c.clear()
        self._pending.clear()

    def events(self):
        return [dict(type=EventType(t).name, start_ns=s, end_ns=e, name=n, source=src) for t, s, e, n, src in self._c.events()]

    def serialize(self) -> bytes:
        self.flush_device()
        return self._c.serialize()

    @staticmethod
    def deserialize(b: bytes) -> "Profiler":
        return Profiler(native.require().ProfilerCore.deserialize(b))

    def log(self, path: str = "logs/profiler.log"):
        """One line per event, the format ``visualizers/visualize_profiler.py`` parses."""
        self.flush_device()
        self.sort()
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        base = self._c.start_ns()
        with open(path, "w") as f:
            for ev in self.events():
                f.write(f"[{ev['source']}] {ev['type']} {ev['name']} start={(ev['start_ns'] - base) / 1e6:.3f}ms "
                        f"end={(ev['end_ns'] - base) / 1e6:.3f}ms dur={(ev['end_ns'] - ev['start_ns']) / 1e6:.3f}ms\n")


class GlobalProfiler:
    _p: Profiler | None = None

    @classmethod
    def get_profiler(cls) -> Profiler:
        if cls._p is None:
            cls._p = Profiler()
            cls._p.init_start_time()
        return cls._p

    @classmethod
    def init_start_time(cls):
        cls.get_profiler().init_start_time()

    @classmethod
    def add_event(cls, *a, **k):
        cls.get_profiler().add_event(*a, **k)
