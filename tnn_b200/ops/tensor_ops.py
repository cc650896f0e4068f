"""``ops::`` free-function family (reference: include/ops/ops.hpp:17-835) and the tensor layout ops
(include/tensor/tensor_ops.hpp:13-354).  GPU fp32 / bf16 tensors run the hand-written kernels of
csrc/kernels/tensor_ops.cu (128-bit vectorised, reductions finish on the device); contiguous CPU
fp32 / fp64 tensors run the native multi-versioned (AVX2+FMA | baseline) loops of
csrc/runtime/cpu_ops.hpp on the shared thread pool (the reference's ops/cpu/skernels.cpp,
dkernels.cpp); anything else (integer dtypes, strided views) uses the torch reference math that
the tests treat as the oracle.
"""
from __future__ import annotations

import torch

from . import native

_BIN = {"add": 0, "sub": 1, "mul": 2, "div": 3, "min": 4, "max": 5, "eq": 6, "gt": 7}
_UN = {"sqrt": 0, "rsqrt": 1, "rcp": 2, "abs": 3, "clamp": 4, "neg": 5, "exp": 6, "log": 7, "square": 8}
_TORCH_BIN = {"add": torch.add, "sub": torch.sub, "mul": torch.mul, "div": torch.div, "min": torch.minimum,
              "max": torch.maximum, "eq": lambda a, b: (a == b).to(a.dtype), "gt": lambda a, b: (a > b).to(a.dtype)}


# element types of the templated GPU kernels (csrc/kernels/tensor_ops.cu: bf16 / fp32 / fp16 / fp64)
_GPU_DTYPES = (torch.float32, torch.bfloat16, torch.float16, torch.float64)


def _native_ok(*ts) -> bool:
    t0 = ts[0]
    return t0.is_cuda and t0.dtype in _GPU_DTYPES and all(
        t is None or (t.is_cuda and t.dtype == t0.dtype and t.is_contiguous() and t.shape == t0.shape) for t in ts)


def _cpu_ok(*ts) -> bool:
    """Contiguous same-shape CPU fp32 / fp64 operands -> the native loops."""
    t0 = ts[0]
    return (not t0.is_cuda) and t0.dtype in (torch.float32, torch.float64) and t0.numel() > 0 and all(
        t is None or ((not t.is_cuda) and t.dtype == t0.dtype and t.is_contiguous() and t.shape == t0.shape) for t in ts)


def _cpu():
    """The native module for the CPU family (None when it could not be built: torch math is used)."""
    return native.load()


def _binary(name, a, b, out=None):
    out = torch.empty_like(a) if out is None else out
    if isinstance(b, torch.Tensor):
        if _native_ok(a, b, out):
            native.require().tensor_binary(a, b, 0.0, out, _BIN[name])
            return out
        if _cpu_ok(a, b, out) and _cpu() is not None:
            _cpu().cpu_binary(a.data_ptr(), b.data_ptr(), 0.0, out.data_ptr(), a.numel(), _BIN[name], a.dtype == torch.float64)
            return out
        out.copy_(_TORCH_BIN[name](a, b))
        return out
    if _native_ok(a, out):
        native.require().tensor_binary(a, None, float(b), out, _BIN[name])
        return out
    if _cpu_ok(a, out) and _cpu() is not None:
        _cpu().cpu_binary(a.data_ptr(), 0, float(b), out.data_ptr(), a.numel(), _BIN[name], a.dtype == torch.float64)
        return out
    bb = torch.full_like(a, b)
    out.copy_(_TORCH_BIN[name](a, bb))
    return out


def add(a, b, out=None): return _binary("add", a, b, out)            # noqa: E704
def sub(a, b, out=None): return _binary("sub", a, b, out)            # noqa: E704
def mul(a, b, out=None): return _binary("mul", a, b, out)            # noqa: E704
def div(a, b, out=None): return _binary("div", a, b, out)            # noqa: E704
def minimum(a, b, out=None): return _binary("min", a, b, out)        # noqa: E704
def maximum(a, b, out=None): return _binary("max", a, b, out)        # noqa: E704
def equal(a, b, out=None): return _binary("eq", a, b, out)           # noqa: E704
def greater(a, b, out=None): return _binary("gt", a, b, out)         # noqa: E704
add_scalar, sub_scalar, mul_scalar, div_scalar = add, sub, mul, div


def _unary(name, a, out=None, lo=0.0, hi=0.0):
    out = torch.empty_like(a) if out is None else out
    if _native_ok(a, out):
        native.require().tensor_unary(a, out, _UN[name], float(lo), float(hi))
        return out
    if _cpu_ok(a, out) and _cpu() is not None:
        _cpu().cpu_unary(a.data_ptr(), out.data_ptr(), a.numel(), _UN[name], float(lo), float(hi), a.dtype == torch.float64)
        return out
    f = {"sqrt": torch.sqrt, "rsqrt": torch.rsqrt, "rcp": torch.reciprocal, "abs": torch.abs,
         "clamp": lambda t: torch.clamp(t, lo, hi), "neg": torch.neg, "exp": torch.exp, "log": torch.log,
         "square": torch.square}[name]
    out.copy_(f(a))
    return out


def sqrt(a, out=None): return _unary("sqrt", a, out)                 # noqa: E704
def rsqrt(a, out=None): return _unary("rsqrt", a, out)               # noqa: E704
def rcp(a, out=None): return _unary("rcp", a, out)                   # noqa: E704
def abs(a, out=None): return _unary("abs", a, out)                   # noqa: E704,A001
def neg(a, out=None): return _unary("neg", a, out)                   # noqa: E704
def exp(a, out=None): return _unary("exp", a, out)                   # noqa: E704
def log(a, out=None): return _unary("log", a, out)                   # noqa: E704
def square(a, out=None): return _unary("square", a, out)             # noqa: E704
def clamp(a, lo, hi, out=None): return _unary("clamp", a, out, lo, hi)  # noqa: E704


def _ternary(op, a, b, c, out=None):
    out = torch.empty_like(a) if out is None else out
    if _native_ok(a, b, c, out):
        native.require().tensor_ternary(a, b, c, out, op)
        return out
    if _cpu_ok(a, b, c, out) and _cpu() is not None:
        _cpu().cpu_ternary(a.data_ptr(), b.data_ptr(), c.data_ptr(), out.data_ptr(), a.numel(), op, a.dtype == torch.float64)
        return out
    out.copy_(a * b + c if op == 0 else (a * b - c if op == 1 else c - a * b))
    return out


def fmadd(a, b, c, out=None): return _ternary(0, a, b, c, out)       # noqa: E704
def fmsub(a, b, c, out=None): return _ternary(1, a, b, c, out)       # noqa: E704
def fnmadd(a, b, c, out=None): return _ternary(2, a, b, c, out)      # noqa: E704


def set_scalar(a: torch.Tensor, v: float) -> torch.Tensor:
    if a.is_cuda and a.dtype == torch.float32 and a.is_contiguous():
        native.require().fill_f32(a, float(v))
        return a
    if _cpu_ok(a) and _cpu() is not None:
        _cpu().cpu_fill(a.data_ptr(), float(v), a.numel(), a.dtype == torch.float64)
        return a
    return a.fill_(v)


def axpy(y: torch.Tensor, x: torch.Tensor, alpha: float) -> torch.Tensor:
    """y += alpha * x."""
    if y.is_cuda and y.dtype == x.dtype == torch.float32 and y.is_contiguous() and x.is_contiguous():
        native.require().axpy_f32(y, x, float(alpha))
        return y
    if _cpu_ok(y, x) and _cpu() is not None:
        _cpu().cpu_axpy(y.data_ptr(), x.data_ptr(), float(alpha), y.numel(), y.dtype == torch.float64)
        return y
    return y.add_(x, alpha=alpha)


def _reduce(mode, a, b=None, mu=0.0) -> torch.Tensor:
    """Device scalar tensor (fp32); nothing is copied to the host."""
    if _native_ok(a, b) if b is not None else _native_ok(a):
        out = torch.empty(1, dtype=torch.float32, device=a.device)
        native.require().tensor_reduce(a, b, float(mu), mode, out)
        return out[0]
    if (_cpu_ok(a, b) if b is not None else _cpu_ok(a)) and _cpu() is not None:
        r = _cpu().cpu_reduce(a.data_ptr(), b.data_ptr() if b is not None else 0, float(mu), a.numel(), mode,
                              a.dtype == torch.float64)
        return torch.tensor(r, dtype=torch.float32)
    af = a.double()
    if mode == 0:
        return af.sum().float()
    if mode == 1:
        return (af * b.double()).sum().float()
    return ((af - mu) ** 2).sum().float()


def sum(a): return _reduce(0, a)                                      # noqa: E704,A001
def dot_product(a, b): return _reduce(1, a, b)                        # noqa: E704
def sum_squared_diff(a, mean): return _reduce(2, a, None, mean)       # noqa: E704


def fill_random_uniform(a: torch.Tensor, lo: float, hi: float, seed: int) -> torch.Tensor:
    if a.is_cuda and a.dtype == torch.float32 and a.is_contiguous():
        native.require().philox_uniform(a, float(lo), float(hi), int(seed))
        return a
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    a.copy_(torch.empty(a.shape, dtype=torch.float32).uniform_(lo, hi, generator=g))
    return a


def fill_random_normal(a: torch.Tensor, mean: float, std: float, seed: int) -> torch.Tensor:
    if a.is_cuda and a.dtype == torch.float32 and a.is_contiguous():
        native.require().philox_normal(a, float(mean), float(std), int(seed))
        return a
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    a.copy_(torch.empty(a.shape, dtype=torch.float32).normal_(mean, std, generator=g))
    return a


def cast(a: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if a.is_cuda and a.is_contiguous():
        if a.dtype == torch.float32 and dtype == torch.bfloat16:
            out = torch.empty(a.shape, dtype=dtype, device=a.device)
            native.require().cast_f32_to_bf16(a, out)
            return out
        if a.dtype == torch.bfloat16 and dtype == torch.float32:
            out = torch.empty(a.shape, dtype=dtype, device=a.device)
            native.require().cast_bf16_to_f32(a, out)
            return out
        if a.dtype in _GPU_DTYPES and dtype in _GPU_DTYPES:
            if a.dtype == dtype:
                return a
            out = torch.empty(a.shape, dtype=dtype, device=a.device)
            native.require().cast_any(a, out)          # any pair of bf16 / fp32 / fp16 / fp64
            return out
    return a.to(dtype)


def cd_copy(src: torch.Tensor, dst: torch.Tensor, non_blocking: bool = True) -> torch.Tensor:
    """Cross-device copy (host<->device, device<->device over NVLink), async on the current stream."""
    dst.copy_(src, non_blocking=non_blocking)
    return dst


def bswap(a: torch.Tensor) -> torch.Tensor:
    """In-place byte swap of every element (endianness conversion)."""
    if a.is_cuda and a.is_contiguous():
        native.require().tensor_bswap(a)
        return a
    es = a.element_size()
    if es > 1:
        raw = a.view(torch.uint8).reshape(-1, es)
        raw.copy_(raw.flip(1).clone())
    return a


# ------------------------------------------------------------------------------- layout ops (NHWC)
def _window(x: torch.Tensor, DH: int, DW: int, off_h: int, off_w: int, fill: float = 0.0) -> torch.Tensor:
    N, SH, SW, C = x.shape
    out = torch.empty((N, DH, DW, C), dtype=x.dtype, device=x.device)
    if _native_ok(x):
        native.require().tensor_window_copy(x, out, N, SH, SW, DH, DW, C, off_h, off_w, float(fill))
        return out
    out.fill_(fill)
    h0, w0 = max(0, -off_h), max(0, -off_w)
    h1, w1 = min(DH, SH - off_h), min(DW, SW - off_w)
    if h1 > h0 and w1 > w0:
        out[:, h0:h1, w0:w1] = x[:, h0 + off_h:h1 + off_h, w0 + off_w:w1 + off_w]
    return out


def pad(x: torch.Tensor, pad_h: int, pad_w: int, value: float = 0.0) -> torch.Tensor:
    return _window(x, x.shape[1] + 2 * pad_h, x.shape[2] + 2 * pad_w, -pad_h, -pad_w, value)


def unpad(x: torch.Tensor, pad_h: int, pad_w: int) -> torch.Tensor:
    return _window(x, x.shape[1] - 2 * pad_h, x.shape[2] - 2 * pad_w, pad_h, pad_w)


def crop(x: torch.Tensor, top: int, left: int, height: int, width: int) -> torch.Tensor:
    return _window(x, height, width, top, left)


def slice_batch(x: torch.Tensor, start: int, length: int) -> torch.Tensor:
    return x[start:start + length]


def split(x: torch.Tensor, n: int) -> list[torch.Tensor]:
    return list(torch.tensor_split(x, n, 0))


def transpose_2d(x: torch.Tensor) -> torch.Tensor:
    if x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous():
        out = torch.empty((x.shape[1], x.shape[0]), dtype=x.dtype, device=x.device)
        native.require().transpose_bf16(x, out, x.shape[0], x.shape[1])
        return out
    if x.is_cuda and x.dtype in _GPU_DTYPES and x.dim() == 2 and x.is_contiguous():
        out = torch.empty((x.shape[1], x.shape[0]), dtype=x.dtype, device=x.device)
        native.require().transpose_any(x, out)
        return out
    return x.t().contiguous()


def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    N, C, H, W = x.shape
    if x.is_cuda and x.dtype in (torch.float32, torch.bfloat16):
        out = torch.empty((N, H, W, C), dtype=x.dtype, device=x.device)
        native.require().nchw_to_nhwc(x.contiguous(), out, N, C, H, W)
        return out
    return x.permute(0, 2, 3, 1).contiguous()


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    N, H, W, C = x.shape
    if x.is_cuda and x.dtype in (torch.float32, torch.bfloat16):
        out = torch.empty((N, C, H, W), dtype=x.dtype, device=x.device)
        native.require().nhwc_to_nchw(x.contiguous(), out, N, C, H, W)
        return out
    return x.permute(0, 3, 1, 2).contiguous()


def im2col(x: torch.Tensor, R: int, S: int, stride: int, pad_: int) -> torch.Tensor:
    """[N,H,W,C] -> [N*OH*OW, R*S*C] (legacy conv path; the tcgen05 conv needs no col buffer)."""
    N, H, W, C = x.shape
    OH, OW = (H + 2 * pad_ - R) // stride + 1, (W + 2 * pad_ - S) // stride + 1
    if x.is_cuda and x.dtype == torch.bfloat16 and (R * S * C) % 8 == 0:
        col = torch.empty((N * OH * OW, R * S * C), dtype=x.dtype, device=x.device)
        native.require().im2col_nhwc(x.contiguous(), col, R * S * C, [N, H, W, C, R, S, stride, pad_, OH, OW])
        return col
    cols = torch.nn.functional.unfold(x.permute(0, 3, 1, 2).float(), (R, S), padding=pad_, stride=stride)   # [N, C*R*S, L]
    cols = cols.reshape(N, C, R * S, OH * OW).permute(0, 3, 2, 1).reshape(N * OH * OW, R * S * C)
    return cols.to(x.dtype)


def softmax(x: torch.Tensor) -> torch.Tensor:
    """Softmax over the last (channel) dimension."""
    C = x.shape[-1]
    if _native_ok(x):
        y = torch.empty_like(x)
        native.require().softmax_rows(x, y, x.numel() // C, C)
        return y
    return torch.softmax(x.float(), -1).to(x.dtype)


def softmax_backward(y: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    C = y.shape[-1]
    if _native_ok(y, dy):
        dx = torch.empty_like(y)
        native.require().softmax_rows_bwd(y, dy, dx, y.numel() // C, C)
        return dx
    yf, g = y.float(), dy.float()
    return (yf * (g - (g * yf).sum(-1, keepdim=True))).to(y.dtype)


# ---- remaining utilities of include/tensor/tensor_utils.hpp / tensor_ops.hpp --------------------------------
def col2im(col: torch.Tensor, x_shape, R: int, S: int, stride: int, pad_: int) -> torch.Tensor:
    """[N*OH*OW, R*S*C] -> [N,H,W,C]: the adjoint of ``im2col`` (every input element sums the col entries that
    map onto it; gather formulation on the GPU)."""
    N, H, W, C = x_shape
    OH, OW = (H + 2 * pad_ - R) // stride + 1, (W + 2 * pad_ - S) // stride + 1
    if col.is_cuda and col.dtype == torch.bfloat16 and (R * S * C) % 8 == 0:
        dx = torch.empty((N, H, W, C), dtype=col.dtype, device=col.device)
        native.require().col2im_nhwc(col.contiguous(), R * S * C, dx, [N, H, W, C, R, S, stride, pad_, OH, OW])
        return dx
    c = col.float().reshape(N, OH * OW, R * S, C).permute(0, 3, 2, 1).reshape(N, C * R * S, OH * OW)
    x = torch.nn.functional.fold(c, (H, W), (R, S), padding=pad_, stride=stride)      # [N, C, H, W]
    return x.permute(0, 2, 3, 1).contiguous().to(col.dtype)


def nchw_to_cnhw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(1, 0, 2, 3).contiguous()


def cnhw_to_nchw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(1, 0, 2, 3).contiguous()


def check_nan_and_inf(x: torch.Tensor, name: str = "tensor", raise_on_error: bool = False) -> bool:
    """True when every element is finite; otherwise reports the counts (reference: check_nan_and_inf)."""
    finite = torch.isfinite(x)
    if bool(finite.all()):
        return True
    n_nan, n_inf = int(torch.isnan(x).sum()), int(torch.isinf(x).sum())
    msg = f"{name}: {n_nan} NaN, {n_inf} Inf of {x.numel()} elements"
    if raise_on_error:
        raise FloatingPointError(msg)
    print("[tnn_b200] " + msg)
    return False


def buckets(x: torch.Tensor, n: int = 10, lo: float | None = None, hi: float | None = None):
    """Histogram of the finite values: (edges [n+1], counts [n])."""
    v = x.detach().float().reshape(-1)
    v = v[torch.isfinite(v)]
    if v.numel() == 0:
        return torch.zeros(n + 1), torch.zeros(n, dtype=torch.int64)
    lo = float(v.min()) if lo is None else lo
    hi = float(v.max()) if hi is None else hi
    if hi <= lo:
        hi = lo + 1.0
    counts = torch.histc(v.cpu(), bins=n, min=lo, max=hi).to(torch.int64)
    return torch.linspace(lo, hi, n + 1), counts


def print_data_distribution(x: torch.Tensor, n: int = 10, name: str = "tensor") -> str:
    edges, counts = buckets(x, n)
    total = max(1, int(counts.sum()))
    lines = [f"{name}: {x.numel()} elements, distribution over {n} buckets"]
    for i in range(n):
        bar = "#" * int(40 * int(counts[i]) / total)
        lines.append(f"  [{float(edges[i]):+.4g}, {float(edges[i + 1]):+.4g}) {int(counts[i]):>9d} {bar}")
    s = "\n".join(lines)
    print(s)
    return s
