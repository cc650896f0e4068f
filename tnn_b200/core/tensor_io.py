"""Tensor record (de)serialisation of the checkpoint format.

Record layout (reference: include/tensor/tensor.hpp:585-606, tensor_factory.hpp:95-128):
``u32 dtype tag | u64 ndims | u64 dims[ndims] | raw little-endian element bytes``.
A model file is the pretty-printed JSON graph config immediately followed by one record per
parameter in GraphContext registration order.  Optional trailer (new): magic ``TNNX`` + u64
length + JSON index + records, carrying optimizer state / step so a resume is exact.
"""
from __future__ import annotations

import io
import json
import struct

import numpy as np
import torch

from .dtype import DType, dtype_size, from_torch, to_torch

TRAILER_MAGIC = b"TNNX"


def _to_bytes(t: torch.Tensor) -> bytes:
    t = t.detach().contiguous().cpu()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().tobytes()
    if t.dtype == torch.bool:
        return t.to(torch.uint8).numpy().tobytes()
    return t.numpy().tobytes()


def save_tensor(f, t: torch.Tensor):
    f.write(struct.pack("<I", int(from_torch(t.dtype))))
    f.write(struct.pack("<Q", t.dim()))
    for s in t.shape:
        f.write(struct.pack("<Q", int(s)))
    f.write(_to_bytes(t))


def _read_exact(f, n: int) -> bytes:
    b = f.read(n)
    if len(b) != n:
        raise EOFError(f"expected {n} bytes, got {len(b)}")
    return b


def load_tensor(f, device="cpu") -> torch.Tensor:
    (tag,) = struct.unpack("<I", _read_exact(f, 4))
    (nd,) = struct.unpack("<Q", _read_exact(f, 8))
    shape = [struct.unpack("<Q", _read_exact(f, 8))[0] for _ in range(nd)]
    dt = DType(tag)
    n = 1
    for s in shape:
        n *= s
    raw = _read_exact(f, n * dtype_size(dt))
    tdt = to_torch(dt)
    if tdt == torch.bfloat16:
        t = torch.frombuffer(bytearray(raw), dtype=torch.int16).view(torch.bfloat16)
    elif tdt == torch.bool:
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(torch.bool)
    else:
        t = torch.frombuffer(bytearray(raw), dtype=tdt) if n > 0 else torch.empty(0, dtype=tdt)
    return t.reshape(shape).to(device)


def load_into(f, dst: torch.Tensor):
    """Reads one record and copies it into `dst`, checking dtype and shape (tensor_factory.hpp:95)."""
    t = load_tensor(f)
    if tuple(t.shape) != tuple(dst.shape):
        raise ValueError(f"checkpoint shape {tuple(t.shape)} != parameter shape {tuple(dst.shape)}")
    if t.dtype != dst.dtype:
        raise ValueError(f"checkpoint dtype {t.dtype} != parameter dtype {dst.dtype}")
    dst.copy_(t)


def write_trailer(f, sections: dict):
    """sections: name -> tensor | json-able value."""
    body = io.BytesIO()
    index = []
    for k, v in sections.items():
        if isinstance(v, torch.Tensor):
            index.append({"name": k, "kind": "tensor"})
            save_tensor(body, v)
        else:
            index.append({"name": k, "kind": "json", "value": v})
    idx = json.dumps(index).encode()
    f.write(TRAILER_MAGIC)
    f.write(struct.pack("<Q", len(idx)))
    f.write(idx)
    f.write(body.getvalue())


def read_trailer(f) -> dict:
    magic = f.read(4)
    if magic != TRAILER_MAGIC:
        return {}
    (n,) = struct.unpack("<Q", _read_exact(f, 8))
    index = json.loads(_read_exact(f, n).decode())
    out = {}
    for e in index:
        out[e["name"]] = load_tensor(f) if e["kind"] == "tensor" else e["value"]
    return out
