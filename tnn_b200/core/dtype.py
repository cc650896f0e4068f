"""Runtime dtype enum (reference: include/type/type.hpp:76-224, src/type/type.cpp:8-11)."""
from __future__ import annotations

import enum

import numpy as np
import torch


class DType(enum.IntEnum):
    # numeric values are the on-disk / on-wire u32 tags of the archiver format
    BYTE = 0
    UINT8_T = 1
    BOOL = 2
    FP16 = 3
    BF16 = 4
    FP32 = 5
    FP64 = 6
    INT32_T = 7
    INT64_T = 8
    SIZE_T = 9
    UNKNOWN = 10


_TORCH = {
    DType.BYTE: torch.uint8, DType.UINT8_T: torch.uint8, DType.BOOL: torch.bool, DType.FP16: torch.float16,
    DType.BF16: torch.bfloat16, DType.FP32: torch.float32, DType.FP64: torch.float64,
    DType.INT32_T: torch.int32, DType.INT64_T: torch.int64, DType.SIZE_T: torch.int64,
}
_FROM_TORCH = {torch.uint8: DType.UINT8_T, torch.bool: DType.BOOL, torch.float16: DType.FP16,
               torch.bfloat16: DType.BF16, torch.float32: DType.FP32, torch.float64: DType.FP64,
               torch.int32: DType.INT32_T, torch.int64: DType.INT64_T}
_SIZE = {DType.BYTE: 1, DType.UINT8_T: 1, DType.BOOL: 1, DType.FP16: 2, DType.BF16: 2, DType.FP32: 4,
         DType.FP64: 8, DType.INT32_T: 4, DType.INT64_T: 8, DType.SIZE_T: 8}
_EPS = {DType.FP16: 1e-3, DType.BF16: 1e-2, DType.FP32: 1e-5, DType.FP64: 1e-8}
_STR = {"FP32": DType.FP32, "FLOAT": DType.FP32, "BF16": DType.BF16, "FP16": DType.FP16, "FP64": DType.FP64,
        "DOUBLE": DType.FP64, "INT32": DType.INT32_T, "INT64": DType.INT64_T, "UINT8": DType.UINT8_T,
        "BYTE": DType.BYTE, "BOOL": DType.BOOL, "SIZE_T": DType.SIZE_T}
_NAME = {DType.FP32: "FP32", DType.BF16: "BF16", DType.FP16: "FP16", DType.FP64: "FP64", DType.INT32_T: "INT32",
         DType.INT64_T: "INT64", DType.UINT8_T: "UINT8", DType.BYTE: "BYTE", DType.BOOL: "BOOL",
         DType.SIZE_T: "SIZE_T", DType.UNKNOWN: "UNKNOWN"}


def to_torch(d: DType) -> torch.dtype:
    return _TORCH[DType(d)]


def from_torch(t: torch.dtype) -> DType:
    return _FROM_TORCH[t]


def dtype_size(d: DType) -> int:
    return _SIZE[DType(d)]


def dtype_eps(d) -> float:
    """Per-dtype epsilon used by the model zoo for BatchNorm (type.cpp:8-11)."""
    if isinstance(d, torch.dtype):
        d = from_torch(d)
    return _EPS.get(DType(d), 1e-5)


def string_to_dtype(s: str) -> DType:
    try:
        return _STR[s.upper()]
    except KeyError as e:
        raise ValueError(f"unknown dtype string {s!r}") from e


def dtype_to_string(d) -> str:
    if isinstance(d, torch.dtype):
        d = from_torch(d)
    return _NAME[DType(d)]


def to_numpy(d: DType):
    return {DType.FP32: np.float32, DType.FP64: np.float64, DType.FP16: np.float16, DType.INT32_T: np.int32,
            DType.INT64_T: np.int64, DType.UINT8_T: np.uint8, DType.BYTE: np.uint8, DType.BOOL: np.bool_,
            DType.SIZE_T: np.uint64, DType.BF16: np.uint16}[DType(d)]
