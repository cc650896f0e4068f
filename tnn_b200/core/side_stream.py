"""Side stream for work that hangs off the backward critical path.

A convolution's weight gradient needs only (dy, x); nothing later in the backward pass reads it
until the optimizer.  Issued on the compute stream it sits between the layer's dgrad and the next
BatchNorm backward and everything waits for it; issued on a second stream, the tensor-core-bound
wgrad kernels (one 160 KB CTA per SM, 192 threads) share the SMs with the memory-bound BatchNorm
backward kernels and fill the last-wave gaps of the persistent dgrad GEMMs.  Inside a captured
training step the two streams become parallel branches of the CUDA graph.  (Reference: one ``Flow`` = one
CUDA stream per device context, include/device/flow.hpp:28-65, and every layer's dgrad / wgrad are issued
back to back on it, e.g. src/nn/layers_impl/conv2d_layer.cpp backward.)

Usage (layers):   side_stream.run(device, fn, keep=(tensors...))
Joined automatically when the OUTERMOST ``Layer.backward`` returns (``enter``/``exit`` depth
counting), so callers of ``backward`` always observe complete gradients on the current stream.
``TNN_WGRAD_STREAM=0`` disables it.
"""
from __future__ import annotations

import os

import torch

_depth = 0
_streams: dict = {}
_pending: dict = {}


def enabled() -> bool:
    return os.environ.get("TNN_WGRAD_STREAM", "1") == "1"


def _key(device) -> int:
    device = torch.device(device)
    return device.index if device.index is not None else torch.cuda.current_device()


def stream(device) -> "torch.cuda.Stream":
    k = _key(device)
    s = _streams.get(k)
    if s is None:
        s = _streams[k] = torch.cuda.Stream(torch.device("cuda", k))
    return s


def run(device, fn, keep=()):
    """Run ``fn()`` on the side stream after everything already queued on the current stream;
    ``keep`` holds references to the tensors it reads until the join (their storage must not be
    recycled by the allocator while the side stream is still reading them)."""
    k = _key(device)
    side = stream(device)
    main = torch.cuda.current_stream(torch.device("cuda", k))
    side.wait_stream(main)
    with torch.cuda.stream(side):
        fn()
    _pending.setdefault(k, []).append(keep)


def has_pending(device) -> bool:
    return bool(_pending.get(_key(device)))


def join(device=None):
    """Make the current stream wait for the side stream(s) and drop the kept references."""
    keys = list(_pending.keys()) if device is None else [_key(device)]
    for k in keys:
        if _pending.get(k):
            torch.cuda.current_stream(torch.device("cuda", k)).wait_stream(_streams[k])
            _pending[k] = []


def enter():
    global _depth
    _depth += 1


def exit_():
    global _depth
    _depth -= 1
    if _depth <= 0:
        _depth = 0
        if _pending:
            join()
