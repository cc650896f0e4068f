"""Console + file logger (reference: include/logging/logger.hpp:16-145) and a structured JSON-lines
step log (device-timed step ms, throughput, lr)."""
from __future__ import annotations

import json
import logging
import os
import sys
import time


class Logger:
    def __init__(self, name: str = "tnn", log_file: str | None = None, level: str | None = None):
        self._l = logging.getLogger(name)
        self._l.setLevel(getattr(logging, (level or os.environ.get("LOG_LEVEL", "INFO")).upper(), logging.INFO))
        self._l.propagate = False
        if not self._l.handlers:
            fmt = logging.Formatter("[%(asctime)s] [%(name)s] [%(levelname)s] %(message)s")
            h = logging.StreamHandler(sys.stdout)
            h.setFormatter(fmt)
            self._l.addHandler(h)
            if log_file:
                os.makedirs(os.path.dirname(log_file) or ".", exist_ok=True)
                fh = logging.FileHandler(log_file)
                fh.setFormatter(fmt)
                self._l.addHandler(fh)

    def info(self, *a):
        self._l.info(" ".join(str(x) for x in a))

    def warn(self, *a):
        self._l.warning(" ".join(str(x) for x in a))

    def error(self, *a):
        self._l.error(" ".join(str(x) for x in a))

    def debug(self, *a):
        self._l.debug(" ".join(str(x) for x in a))


class GlobalLogger:
    _inst: Logger | None = None

    @classmethod
    def get(cls) -> Logger:
        if cls._inst is None:
            cls._inst = Logger("tnn")
        return cls._inst


class StepLog:
    """JSON-lines metrics sink."""

    def __init__(self, path: str):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        self.f = open(path, "a")

    def write(self, **kv):
        kv.setdefault("ts", time.time())
        self.f.write(json.dumps(kv) + "\n")
        self.f.flush()

    def close(self):
        self.f.close()
