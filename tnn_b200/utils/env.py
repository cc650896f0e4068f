"""``.env`` loader + typed lookups (reference: include/utils/env.hpp:13-129).  Process environment
variables take precedence over the file."""
from __future__ import annotations

import os


class _InstanceOrGlobal:
    """``Env.get(...)`` on the class reads the process-wide environment (./.env loaded once), like the reference's static
    ``Env::get<T>`` (env.hpp:121); on an instance it reads that instance's file."""

    def __init__(self, fn):
        self.fn = fn

    def __get__(self, obj, cls):
        target = obj if obj is not None else cls.instance()
        return lambda *a, **k: self.fn(target, *a, **k)


class Env:
    _global = None

    @classmethod
    def instance(cls) -> "Env":
        if cls._global is None:
            cls._global = Env(".env")
        return cls._global

    @classmethod
    def reload(cls, path: str = ".env") -> "Env":
        cls._global = Env(path)
        return cls._global

    def __init__(self, path: str = ".env"):
        self.values: dict[str, str] = {}
        if path and os.path.exists(path):
            with open(path) as f:
                for line in f:
                    line = line.strip()
                    if not line or line.startswith("#") or "=" not in line:
                        continue
                    k, v = line.split("=", 1)
                    v = v.strip()
                    if len(v) >= 2 and v[0] == v[-1] and v[0] in "\"'":
                        v = v[1:-1]
                    self.values[k.strip()] = v

    @_InstanceOrGlobal
    def has(self, key: str) -> bool:
        return key in os.environ or key in self.values

    @_InstanceOrGlobal
    def get(self, key: str, default=None, typ=None):
        if typ is None:                                   # Env::get<T>: the type of the default decides
            typ = type(default) if default is not None else str
        raw = os.environ.get(key, self.values.get(key))
        if raw is None:
            return default
        try:
            if typ is bool:
                return raw.strip().lower() in ("1", "true", "yes", "on")
            return typ(raw)
        except (TypeError, ValueError):
            return default
