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


class LayerProfiler:
    """Per-layer forward / backward timing and memory table for the single-process trainer -- the
    ``PROFILER_TYPE = NONE | NORMAL | CUMULATIVE``, ``PRINT_LAYER_PROFILING`` and
    ``PRINT_LAYER_MEMORY_USAGE`` switches of TrainingConfig (include/nn/train.hpp:31,54-56), which
    the reference declares without implementing the per-layer part.

    ``attach(model)`` wraps ``forward_impl`` / ``backward_impl`` of every leaf layer INSTANCE (the
    classes and the un-profiled hot path are untouched); ``detach()`` restores them.  On a GPU each
    call is bracketed by CUDA events on the launching stream and resolved when the table is
    built, so profiling adds no synchronisation inside the step (it does need the eager step: the
    trainer turns CUDA-graph replay off while a profiler is attached).  NORMAL reports the interval
    since the last report, CUMULATIVE the totals since attach."""

    def __init__(self, mode: str = "NORMAL"):
        self.mode = (mode or "NORMAL").upper()
        self._layers = []            # (path, layer)
        self._saved = {}
        self._host = {}              # path -> [fwd_ms, bwd_ms, fwd_calls, bwd_calls]
        self._pending = []           # (path, slot, ev0, ev1)

    @staticmethod
    def _leaves(layer, prefix=""):
        kids = layer.children() if hasattr(layer, "children") else []
        path = f"{prefix}/{layer.name}" if prefix else layer.name
        if not kids:
            yield path, layer
            return
        for k in kids:
            yield from LayerProfiler._leaves(k, path)

    def attach(self, model) -> "LayerProfiler":
        import time
        for path, layer in self._leaves(model):
            if id(layer) in self._saved:
                continue
            self._layers.append((path, layer))
            self._host[path] = [0.0, 0.0, 0, 0]
            self._saved[id(layer)] = (layer, layer.__dict__.get("forward_impl"), layer.__dict__.get("backward_impl"))

            def wrap(fn, slot, path=path):
                def timed(*a, **k):
                    t = a[0] if a and isinstance(a[0], torch.Tensor) else (a[0][0] if a and isinstance(a[0], (list, tuple)) and a[0] else None)
                    if t is not None and t.is_cuda:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        out = fn(*a, **k)
                        e1.record()
                        self._pending.append((path, slot, e0, e1))
                    else:
                        t0 = time.perf_counter()
                        out = fn(*a, **k)
                        self._host[path][slot] += (time.perf_counter() - t0) * 1e3
                    self._host[path][2 + slot] += 1
                    return out
                return timed

            layer.forward_impl = wrap(layer.forward_impl, 0)
            layer.backward_impl = wrap(layer.backward_impl, 1)
        return self

    def detach(self):
        for layer, f, b in self._saved.values():
            for name, orig in (("forward_impl", f), ("backward_impl", b)):
                if orig is None:
                    layer.__dict__.pop(name, None)           # fall back to the class method
                else:
                    layer.__dict__[name] = orig
        self._saved.clear()
        self._layers.clear()

    def _resolve(self):
        if self._pending:
            torch.cuda.synchronize()
            for path, slot, e0, e1 in self._pending:
                self._host[path][slot] += e0.elapsed_time(e1)
            self._pending.clear()

    def rows(self):
        self._resolve()
        return [(p, *self._host[p]) for p, _ in self._layers]

    def reset(self):
        self._resolve()
        for v in self._host.values():
            v[:] = [0.0, 0.0, 0, 0]

    def report(self, title: str = "layer profile") -> str:
        rows = self.rows()
        tot = sum(r[1] + r[2] for r in rows) or 1.0
        w = max([len(r[0]) for r in rows] + [5])
        lines = [f"-- {title} ({self.mode.lower()}) --", f"{'layer':<{w}} {'fwd ms':>10} {'bwd ms':>10} {'calls':>7} {'share':>7}"]
        for p, f, b, nf, nb in rows:
            lines.append(f"{p:<{w}} {f:>10.3f} {b:>10.3f} {nf:>7d} {100 * (f + b) / tot:>6.1f}%")
        lines.append(f"{'total':<{w}} {sum(r[1] for r in rows):>10.3f} {sum(r[2] for r in rows):>10.3f}")
        if self.mode != "CUMULATIVE":
            self.reset()
        return "\n".join(lines)

    def memory_report(self, title: str = "layer memory") -> str:
        """Parameter bytes and the bytes currently held in each layer's per-microbatch caches."""
        def nbytes(o):
            if isinstance(o, torch.Tensor):
                return o.numel() * o.element_size()
            if isinstance(o, dict):
                return sum(nbytes(v) for v in o.values())
            if isinstance(o, (list, tuple)):
                return sum(nbytes(v) for v in o)
            return 0
        w = max([len(p) for p, _ in self._layers] + [5])
        lines = [f"-- {title} --", f"{'layer':<{w}} {'params KiB':>12} {'cached KiB':>12}"]
        tp = tc = 0
        for p, layer in self._layers:
            pb = sum(nbytes(t) for t in getattr(layer, "params", {}).values())
            cb = nbytes(getattr(layer, "_cache", {}))
            tp, tc = tp + pb, tc + cb
            lines.append(f"{p:<{w}} {pb / 1024:>12.1f} {cb / 1024:>12.1f}")
        lines.append(f"{'total':<{w}} {tp / 1024:>12.1f} {tc / 1024:>12.1f}")
        return "\n".join(lines)
