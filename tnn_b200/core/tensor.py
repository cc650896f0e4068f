"""Tensor facade (reference: include/tensor/tensor.hpp:41-610, tensor_factory.hpp:12-159).

Storage and stream plumbing are ``torch.Tensor``; this module provides the reference's factory
and member surface on top of it: ``make_tensor`` overloads (dtype / shape / device / allocator),
``create_like``, ``dtype_cast``, ``load`` / ``load_into`` / ``save`` in the reference record format
``{u32 dtype, u64 ndims, u64 dims[], raw}``, device moves, in-place arithmetic through the native
op family, and the debugging helpers (head / print_data / min / max / mean / variance).
"""
from __future__ import annotations

import math

import torch

from ..ops import tensor_ops as ops
from . import tensor_io
from .dtype import DType, string_to_dtype, to_torch

Tensor = torch.Tensor
DEFAULT_ALIGNMENT = 256   # dptr default alignment (include/device/dptr.hpp)


def _torch_dtype(dt) -> torch.dtype:
    if isinstance(dt, torch.dtype):
        return dt
    if isinstance(dt, str):
        dt = string_to_dtype(dt)
    return to_torch(DType(dt))


def _torch_device(device) -> torch.device:
    if device is None:
        return torch.device("cpu")
    if hasattr(device, "torch_device"):
        return device.torch_device()
    if hasattr(device, "type") and hasattr(device.type, "name") and not isinstance(device, torch.device):
        return torch.device("cuda", device.index) if device.type.name == "GPU" else torch.device("cpu")
    return torch.device(device)


def make_tensor(*args, device=None, allocator=None, dtype=None) -> torch.Tensor:
    """make_tensor(shape) | make_tensor(dtype, shape) | make_tensor(allocator, dtype, shape): a zero-filled
    row-major tensor.  With an allocator (DELAllocator / PoolAllocator) the storage comes from it."""
    a = list(args)
    if a and hasattr(a[0], "allocate"):
        allocator = a.pop(0)
    if len(a) == 2:
        dtype, shape = a
    elif len(a) == 1:
        shape = a[0]
    else:
        raise TypeError("make_tensor([allocator,] [dtype,] shape)")
    dt = _torch_dtype(dtype if dtype is not None else torch.float32)
    shape = tuple(int(s) for s in shape)
    if allocator is not None:
        t = allocator.allocate(shape, dt)
        t.zero_()
        return t
    return torch.zeros(shape, dtype=dt, device=_torch_device(device))


def create_like(t: torch.Tensor, dtype=None) -> torch.Tensor:
    return torch.zeros(t.shape, dtype=_torch_dtype(dtype) if dtype is not None else t.dtype, device=t.device)


def dtype_cast(t: torch.Tensor, dtype) -> torch.Tensor:
    return ops.cast(t.contiguous(), _torch_dtype(dtype))


def to_device(t: torch.Tensor, device, non_blocking: bool = True) -> torch.Tensor:
    return t.to(_torch_device(device), non_blocking=non_blocking)


def to_host(t: torch.Tensor) -> torch.Tensor:
    return t.cpu()


def span(t: torch.Tensor, start: int, length: int) -> torch.Tensor:
    """Contiguous view over ``length`` leading-dimension rows (the only view the reference has)."""
    return t[start:start + length]


def copy_batch(dst: torch.Tensor, dst_index: int, src: torch.Tensor, src_index: int) -> torch.Tensor:
    dst[dst_index].copy_(src[src_index])
    return dst


def save(t: torch.Tensor, f) -> None:
    if isinstance(f, (str, bytes)) or hasattr(f, "__fspath__"):
        with open(f, "wb") as fh:
            tensor_io.save_tensor(fh, t)
    else:
        tensor_io.save_tensor(f, t)


def load(f, device="cpu") -> torch.Tensor:
    if isinstance(f, (str, bytes)) or hasattr(f, "__fspath__"):
        with open(f, "rb") as fh:
            return tensor_io.load_tensor(fh, _torch_device(device))
    return tensor_io.load_tensor(f, _torch_device(device))


def load_into(f, dst: torch.Tensor) -> torch.Tensor:
    tensor_io.load_into(f, dst)
    return dst


def fill(t: torch.Tensor, v: float) -> torch.Tensor:
    return ops.set_scalar(t, v)


def fill_random_uniform(t, lo=0.0, hi=1.0, seed=0):
    return ops.fill_random_uniform(t, lo, hi, seed)


def fill_random_normal(t, mean=0.0, std=1.0, seed=0):
    return ops.fill_random_normal(t, mean, std, seed)


def stats(t: torch.Tensor) -> dict:
    """min / max / mean / variance (population), computed on the device and read once."""
    n = max(1, t.numel())
    mean = float(ops.sum(t.contiguous())) / n
    var = float(ops.sum_squared_diff(t.contiguous(), mean)) / n
    return {"min": float(t.min()), "max": float(t.max()), "mean": mean, "variance": var}


def head(t: torch.Tensor, n: int = 8) -> list:
    return t.reshape(-1)[:n].float().cpu().tolist()


def print_data(t: torch.Tensor, n: int = 16) -> str:
    s = f"Tensor(shape={tuple(t.shape)}, dtype={t.dtype}, device={t.device}) head={head(t, n)}"
    print(s)
    return s


def nbytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


def aligned(t: torch.Tensor, alignment: int = DEFAULT_ALIGNMENT) -> bool:
    return t.data_ptr() % alignment == 0


def shape_size(shape) -> int:
    return int(math.prod(shape)) if len(shape) else 1


# ---- the remaining members of the reference Tensor (include/tensor/tensor.hpp: clone, resize, ensure,
# ---- share_from, same_shape, shape_str, the scalar arithmetic and the min / max / mean / variance queries)
def clone(t: torch.Tensor) -> torch.Tensor:
    return t.detach().clone()


def same_shape(a: torch.Tensor, b: torch.Tensor) -> bool:
    return tuple(a.shape) == tuple(b.shape)


def shape_str(t: torch.Tensor) -> str:
    return "[" + ", ".join(str(int(d)) for d in t.shape) + "]"


def capacity(t: torch.Tensor) -> int:
    """Elements the underlying storage can hold from this tensor's offset (reference: capacity())."""
    return (t.untyped_storage().nbytes() // t.element_size()) - t.storage_offset()


def ensure(t: torch.Tensor, shape) -> torch.Tensor:
    """A tensor of ``shape`` that reuses ``t``'s storage when it is large enough (grow-only buffer, the
    reference's ensure()); otherwise a fresh allocation on the same device.  Contents are unspecified."""
    n = shape_size(shape)
    if t.is_contiguous() and capacity(t) >= n:
        return t.as_strided(tuple(shape), torch.empty(tuple(shape), device="meta").stride())
    return torch.empty(tuple(shape), dtype=t.dtype, device=t.device)


def resize(t: torch.Tensor, shape) -> torch.Tensor:
    """Reference resize(): new shape, old contents dropped unless the element count is unchanged."""
    if shape_size(shape) == t.numel() and t.is_contiguous():
        return t.reshape(tuple(shape))
    return ensure(t, shape)


def share_from(src: torch.Tensor, shape=None) -> torch.Tensor:
    """A second handle on ``src``'s storage (reference share_from()): writes through either are visible in both."""
    return src.view(src.shape if shape is None else tuple(shape))


def _scalar_like(t: torch.Tensor, v: float) -> torch.Tensor:
    return torch.full((1,), float(v), dtype=t.dtype, device=t.device).expand_as(t).contiguous()


def add_scalar(t, v, out=None): return ops.add(t, _scalar_like(t, v), out=out)          # noqa: E704
def sub_scalar(t, v, out=None): return ops.sub(t, _scalar_like(t, v), out=out)          # noqa: E704
def mul_scalar(t, v, out=None): return ops.mul(t, _scalar_like(t, v), out=out)          # noqa: E704


def div_scalar(t, v, out=None):
    if v == 0:
        raise ZeroDivisionError("Division by zero.")
    return ops.mul(t, _scalar_like(t, 1.0 / v), out=out)


def tmin(t: torch.Tensor) -> float: return stats(t)["min"]        # reference Tensor::min()  # noqa: E704
def tmax(t: torch.Tensor) -> float: return stats(t)["max"]        # reference Tensor::max()  # noqa: E704
def mean(t: torch.Tensor) -> float: return stats(t)["mean"]                              # noqa: E704
def variance(t: torch.Tensor) -> float: return stats(t)["variance"]                      # noqa: E704
