"""Small helpers of the reference's include/utils/misc.hpp and parser.hpp: ``benchmark`` (time a callable, on
the device when the work is CUDA work), ``hash_combine`` / ``get_shape_hash`` (boost-style hash mixing used for
plan caches), ``from_str`` (typed parsing of configuration strings), ``get_enum_vector``."""
from __future__ import annotations

import enum
import time

import torch

from .hardware_info import get_shape_hash  # noqa: F401  (same definition, one implementation)


def benchmark(fn, iters: int = 10, warmup: int = 2, device=None) -> float:
    """Mean milliseconds per call of ``fn()``.  With a CUDA ``device`` the interval is measured with CUDA events on
    the current stream after ``warmup`` untimed calls (synchronised on both sides); otherwise wall clock."""
    for _ in range(warmup):
        fn()
    dev = torch.device(device) if device is not None else None
    if dev is not None and dev.type == "cuda":
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / iters
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) * 1e3 / iters


def hash_combine(seed: int, value: int) -> int:
    """64-bit boost::hash_combine."""
    seed ^= (value + 0x9E3779B97F4A7C15 + ((seed << 6) & 0xFFFFFFFFFFFFFFFF) + (seed >> 2)) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFFFFFFFFFF


def from_str(text: str, typ):
    """Typed parse of a configuration string (parser.hpp ``from_str<T>``): bool accepts 1/0/true/false/yes/no/on/off,
    enums accept member names (case-insensitive) or values."""
    t = text.strip()
    if typ is bool:
        low = t.lower()
        if low in ("1", "true", "yes", "on"):
            return True
        if low in ("0", "false", "no", "off"):
            return False
        raise ValueError(f"not a boolean: {text!r}")
    if isinstance(typ, type) and issubclass(typ, enum.Enum):
        for m in typ:
            if m.name.lower() == t.lower():
                return m
        return typ(type(next(iter(typ)).value)(t))
    return typ(t)


def get_enum_vector(typ) -> list:
    return list(typ)
