"""Loader for the in-tree native extension ``tnn_b200/_C.so``.

On a GPU box the extension is mandatory: ops raise instead of silently falling back to
eager PyTorch (the driver records which .so files were loaded).  On a CPU-only box the
pure-PyTorch fp32 reference implementations are used (they are also the test oracles).
"""
from __future__ import annotations

import importlib
import os

import torch

_C = None
_ERR: Exception | None = None


def load(build_if_missing: bool = True):
    global _C, _ERR
    if _C is not None:
        return _C
    try:
        _C = importlib.import_module("tnn_b200._C")
        return _C
    except Exception as e:  # noqa: BLE001
        _ERR = e
    if build_if_missing and os.environ.get("TNN_NO_AUTOBUILD", "0") != "1":
        try:
            from .. import build as _b

            _b.build(verbose=False)
            _C = importlib.import_module("tnn_b200._C")
            return _C
        except Exception as e:  # noqa: BLE001
            _ERR = e
    return None


def require():
    c = load()
    if c is None:
        raise RuntimeError(
            "tnn_b200 native extension (_C.so) is not available; run `python -m tnn_b200.build`. "
            f"Last error: {_ERR!r}")
    return c


def available() -> bool:
    return load(build_if_missing=False) is not None


def use_native(t: torch.Tensor) -> bool:
    """True when `t` lives on a GPU -> the hand-written kernels must run."""
    if t.is_cuda:
        require()
        return True
    return False
