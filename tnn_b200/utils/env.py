"""``.env`` loader + typed lookups (reference: include/utils/env.hpp:13-129).  Process environment
variables take precedence over the file."""
from __future__ import annotations

import os


class Env:
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

    def has(self, key: str) -> bool:
        return key in os.environ or key in self.values

    def get(self, key: str, default=None, typ=str):
        raw = os.environ.get(key, self.values.get(key))
        if raw is None:
            return default
        try:
            if typ is bool:
                return raw.strip().lower() in ("1", "true", "yes", "on")
            return typ(raw)
        except (TypeError, ValueError):
            return default
