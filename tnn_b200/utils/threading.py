"""Intra-op CPU threading (reference: include/threading/thread_handler.hpp:37-140 parallel_for /
parallel_for_2d over TBB|OpenMP, thread_wrapper.hpp:18-49 ThreadWrapper = task_arena(n)).
Backed by the native std::thread pool (csrc/runtime/cpu_math.hpp) that the CPU SGEMM uses."""
from __future__ import annotations

import torch

from ..ops import native


def parallel_for(n: int, fn, grain: int = 1, threads: int = 0):
    """fn(begin, end) over [0, n) in grain-sized chunks on the native pool."""
    C = native.load()
    if C is None:
        fn(0, n)
        return
    C.parallel_for(int(n), int(grain), fn, int(threads))


def parallel_for_2d(rows: int, cols: int, fn, block: int = 32, threads: int = 0):
    """fn(r0, r1, c0, c1) over a rows x cols grid in block x block tiles."""
    cb = (cols + block - 1) // block
    rb = (rows + block - 1) // block

    def body(b, e):
        for t in range(b, e):
            r, c = divmod(t, cb)
            fn(r * block, min(rows, (r + 1) * block), c * block, min(cols, (c + 1) * block))
    parallel_for(rb * cb, body, 1, threads)


class ThreadWrapper:
    """Runs a callable with a bounded number of CPU compute threads (torch intra-op + native pool)."""

    def __init__(self, num_threads: int):
        self.num_threads = max(1, int(num_threads))

    def execute(self, fn, *a, **k):
        prev = torch.get_num_threads()
        torch.set_num_threads(self.num_threads)
        C = native.load()
        if C is not None:
            C.pool_threads(self.num_threads)
        try:
            return fn(*a, **k)
        finally:
            torch.set_num_threads(prev)


def cpu_gemm(a: torch.Tensor, b: torch.Tensor, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, out=None, threads=0):
    """fp32 CPU GEMM through the native kernel: out = alpha * op(a) @ op(b) + beta * out
    (reference: tnn::gemm, include/math/gemm.hpp:19-44)."""
    C = native.require()
    a, b = a.contiguous().float(), b.contiguous().float()
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    if out is None:
        out = torch.zeros((M, N), dtype=torch.float32)
    C.cpu_sgemm(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, trans_a, trans_b, float(alpha), float(beta),
                a.stride(0), b.stride(0), out.stride(0), threads)
    return out


def cpu_dot(a: torch.Tensor, b: torch.Tensor) -> float:
    """fp32 dot product through the native AVX2/FMA kernel (reference: the hand-written assembly
    ``simd_dot_product_asm``, include/utils/simd_asm.hpp:13-15)."""
    a, b = a.contiguous().float().reshape(-1), b.contiguous().float().reshape(-1)
    assert a.numel() == b.numel()
    return float(native.require().cpu_dot(a.data_ptr(), b.data_ptr(), a.numel()))
