"""Legacy 2-D ``Matrix`` value type (reference: include/matrix/matrix.hpp:24-229): an owning
rows x cols buffer with element-wise arithmetic, scalar scaling, a matrix product, clone / reshape /
resize and random fills.  The reference keeps it for the pre-Tensor CPU layers; here it is a thin
value wrapper over a 2-D tensor whose arithmetic goes through the same element-wise kernels as the
``ops`` family (``csrc/kernels/tensor_ops.cu`` on a GPU, the fp32 oracle on the host) and whose
``@`` / ``matmul`` is the tcgen05 GEMM for bf16 operands on a GPU.

Difference on purpose: the reference's ``operator*(Matrix)`` checks ``cols == other.rows`` but then
calls the ELEMENT-WISE ``ops::mul`` over ``size()`` elements (matrix.hpp:185-194), which reads out of
bounds for non-square shapes; ``Matrix * Matrix`` here is the real product.
"""
from __future__ import annotations

import torch

from ..ops import gemm as _gemm
from ..ops import tensor_ops as _ops


class Matrix:
    __slots__ = ("_t",)

    def __init__(self, rows: int = 0, cols: int = 0, data: torch.Tensor | None = None, device="cpu",
                 dtype: torch.dtype = torch.float32):
        if data is not None:
            if data.numel() != rows * cols:
                raise ValueError("Matrix: data does not hold rows*cols elements")
            self._t = data.detach().reshape(rows, cols).clone()
        else:
            self._t = torch.zeros((int(rows), int(cols)), device=device, dtype=dtype)

    # ------------------------------------------------------------------ shape / storage
    @classmethod
    def wrap(cls, t: torch.Tensor) -> "Matrix":
        """Adopt a 2-D tensor without copying."""
        if t.dim() != 2:
            raise ValueError("Matrix.wrap expects a 2-D tensor")
        m = cls.__new__(cls)
        m._t = t
        return m

    def rows(self) -> int: return self._t.shape[0]                      # noqa: E704
    def cols(self) -> int: return self._t.shape[1]                      # noqa: E704
    def size(self) -> int: return self._t.numel()                       # noqa: E704
    def data(self) -> torch.Tensor: return self._t                      # noqa: E704
    @property
    def device(self): return self._t.device                             # noqa: E704
    @property
    def dtype(self): return self._t.dtype                               # noqa: E704

    def __getitem__(self, rc):
        return self._t[rc]

    def __setitem__(self, rc, v):
        self._t[rc] = v

    def clone(self) -> "Matrix":
        return Matrix(self.rows(), self.cols(), self._t)

    def reshape(self, rows: int, cols: int) -> "Matrix":
        if rows * cols != self.size():
            raise ValueError("Total number of elements must remain the same for reshape.")
        return Matrix(rows, cols, self._t)

    def resize(self, rows: int, cols: int) -> None:
        """Reallocate (contents are NOT preserved, as in the reference) unless the shape is unchanged."""
        if (rows, cols) != tuple(self._t.shape):
            self._t = torch.zeros((rows, cols), device=self._t.device, dtype=self._t.dtype)

    def to(self, device) -> "Matrix":
        return Matrix.wrap(self._t.to(device))

    # ------------------------------------------------------------------ fills
    def fill(self, value: float) -> "Matrix":
        _ops.set_scalar(self._t, float(value))
        return self

    def fill_random_uniform(self, rng: float, seed: int | None = None) -> "Matrix":
        seed = int(torch.seed() & 0x7FFFFFFF) if seed is None else int(seed)
        _ops.fill_random_uniform(self._t, -float(rng), float(rng), seed)
        return self

    def fill_random_normal(self, mean: float, stddev: float, seed: int | None = None) -> "Matrix":
        seed = int(torch.seed() & 0x7FFFFFFF) if seed is None else int(seed)
        _ops.fill_random_normal(self._t, float(mean), float(stddev), seed)
        return self

    # ------------------------------------------------------------------ arithmetic
    def _same(self, o: "Matrix", what: str):
        if self._t.shape != o._t.shape:
            raise ValueError(f"Matrix dimensions must match for {what}.")

    def __add__(self, o: "Matrix") -> "Matrix":
        self._same(o, "addition")
        return Matrix.wrap(_ops.add(self._t, o._t))

    def __iadd__(self, o: "Matrix") -> "Matrix":
        self._same(o, "addition")
        _ops.add(self._t, o._t, out=self._t)
        return self

    def __sub__(self, o: "Matrix") -> "Matrix":
        self._same(o, "subtraction")
        return Matrix.wrap(_ops.sub(self._t, o._t))

    def __isub__(self, o: "Matrix") -> "Matrix":
        self._same(o, "subtraction")
        _ops.sub(self._t, o._t, out=self._t)
        return self

    def _scaled(self, s: float, out=None) -> torch.Tensor:
        k = torch.full((1,), float(s), device=self._t.device, dtype=self._t.dtype).expand_as(self._t)
        return _ops.mul(self._t, k.contiguous(), out=out)

    def __mul__(self, o):
        if isinstance(o, Matrix):
            return self.matmul(o)
        return Matrix.wrap(self._scaled(o))

    __rmul__ = __mul__

    def __imul__(self, s: float) -> "Matrix":
        self._scaled(s, out=self._t)
        return self

    def __truediv__(self, s: float) -> "Matrix":
        if s == 0:
            raise ZeroDivisionError("Division by zero.")
        return Matrix.wrap(self._scaled(1.0 / s))

    def __itruediv__(self, s: float) -> "Matrix":
        if s == 0:
            raise ZeroDivisionError("Division by zero.")
        self._scaled(1.0 / s, out=self._t)
        return self

    def hadamard(self, o: "Matrix") -> "Matrix":
        self._same(o, "element-wise multiplication")
        return Matrix.wrap(_ops.mul(self._t, o._t))

    def matmul(self, o: "Matrix") -> "Matrix":
        if self.cols() != o.rows():
            raise ValueError("Matrix dimensions must match for multiplication.")
        a, b = self._t, o._t
        if a.is_cuda and a.dtype == torch.bfloat16 and a.shape[1] % 8 == 0:
            # C = A @ B = linear_fwd(A, W) with W = B^T ([N, K], K contiguous)
            return Matrix.wrap(_gemm.linear_fwd(a.contiguous(), _ops.transpose_2d(b.contiguous())))
        return Matrix.wrap((a.float() @ b.float()).to(a.dtype))

    __matmul__ = matmul

    def transpose(self) -> "Matrix":
        return Matrix.wrap(_ops.transpose_2d(self._t.contiguous()))

    def __repr__(self) -> str:
        return f"Matrix({self.rows()}x{self.cols()}, {self._t.dtype}, {self._t.device})"
