"""Thread-safe map (reference: include/common/concurrent_map.hpp:10-60 -- a mutex around an unordered_map with
emplace / try_get / erase / contains / operator[] / size), used for the message-reassembly and peer tables
of the communicators."""
from __future__ import annotations

import threading
from typing import Callable, Generic, Hashable, TypeVar

K = TypeVar("K", bound=Hashable)
V = TypeVar("V")


class ConcurrentMap(Generic[K, V]):
    def __init__(self, default_factory: Callable[[], V] | None = None):
        self._m: dict = {}
        self._lock = threading.RLock()
        self._default = default_factory

    def emplace(self, key: K, value: V) -> bool:
        """Insert if absent (C++ emplace semantics); returns whether the insertion happened."""
        with self._lock:
            if key in self._m:
                return False
            self._m[key] = value
            return True

    def try_get(self, key: K, default: V | None = None):
        """(found, value)"""
        with self._lock:
            if key in self._m:
                return True, self._m[key]
            return False, default

    def erase(self, key: K) -> bool:
        with self._lock:
            return self._m.pop(key, _MISSING) is not _MISSING

    def contains(self, key: K) -> bool:
        with self._lock:
            return key in self._m

    __contains__ = contains

    def __getitem__(self, key: K) -> V:
        """operator[]: default-constructs the value when a ``default_factory`` was given."""
        with self._lock:
            if key not in self._m:
                if self._default is None:
                    raise KeyError(key)
                self._m[key] = self._default()
            return self._m[key]

    def __setitem__(self, key: K, value: V) -> None:
        with self._lock:
            self._m[key] = value

    def update_with(self, key: K, fn: Callable[[V | None], V]) -> V:
        """Atomic read-modify-write of one entry (``fn`` receives the current value or None)."""
        with self._lock:
            v = fn(self._m.get(key))
            self._m[key] = v
            return v

    def size(self) -> int:
        with self._lock:
            return len(self._m)

    __len__ = size

    def snapshot(self) -> dict:
        with self._lock:
            return dict(self._m)

    def clear(self) -> None:
        with self._lock:
            self._m.clear()


_MISSING = object()
