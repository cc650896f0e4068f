"""Per-device "step state": 16 floats in device memory that every graph-replayable kernel reads
its per-step scalars from (learning rate, Adam bias corrections, dropout seed counter).  The host
refreshes it with ONE small pinned->device copy per step, outside any captured CUDA graph, so a
captured training step can be replayed while lr schedules, bias corrections and dropout masks
keep changing.  (The reference passes these scalars BY VALUE to a kernel per parameter tensor and step --
src/nn/optimizers_impl/cuda/adam_kernels.cu:22,59 take ``bias_correction1/2`` as arguments -- which is exactly what
cannot be baked into a CUDA graph.)

layout: [0] lr [1] beta1 [2] beta2 [3] eps [4] weight_decay [5] decoupled [6] grad_scale
        [7] momentum [8] 1-beta1^t [9] 1-beta2^t [12..13] uint64 seed counter
"""
from __future__ import annotations

import torch

_STATES: dict = {}


class StepState:
    RING = 2048

    def __init__(self, device: torch.device):
        self.device = device
        # The H2D copy is asynchronous and the host may run many replays ahead of the device, so every
        # step gets its OWN pinned source row (a ring far deeper than the driver's launch queue lets
        # the host get ahead): a single reused row could be rewritten before an earlier copy reads it.
        self.ring = torch.zeros((self.RING, 16), dtype=torch.float32, pin_memory=(device.type == "cuda"))
        self.host = self.ring[0]
        self.dev = torch.zeros(16, dtype=torch.float32, device=device)
        self.counter = 0

    def advance(self, optimizer=None):
        """Bump the seed counter, refresh optimizer scalars, push to the device (async)."""
        self.counter += 1
        h = self.ring[self.counter % self.RING]
        h.copy_(self.host)                      # scalars persist from step to step unless refreshed
        self.host = h
        h[12:14].view(torch.int64)[0] = self.counter
        if optimizer is not None:
            t = optimizer.step_count + 1
            h[0] = optimizer.lr
            h[1] = getattr(optimizer, "beta1", 0.0)
            h[2] = getattr(optimizer, "beta2", 0.0)
            h[3] = getattr(optimizer, "epsilon", 0.0)
            h[4] = getattr(optimizer, "weight_decay", 0.0)
            h[5] = 1.0 if getattr(optimizer, "decoupled", False) else 0.0
            h[6] = optimizer.grad_scale
            h[7] = getattr(optimizer, "momentum", 0.0)
            h[8] = 1.0 - getattr(optimizer, "beta1", 0.0) ** t
            h[9] = 1.0 - getattr(optimizer, "beta2", 0.0) ** t
        self.dev.copy_(h, non_blocking=True)


def get(device) -> StepState:
    device = torch.device(device)
    key = (device.type, device.index if device.index is not None else (torch.cuda.current_device() if device.type == "cuda" else 0))
    st = _STATES.get(key)
    if st is None:
        st = StepState(device)
        _STATES[key] = st
    return st
