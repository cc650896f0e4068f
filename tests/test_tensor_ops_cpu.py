"""Tensor facade + ops family on CPU (the oracle path; the GPU kernels are compared against the
same functions in tests/test_kernels_gpu.py).  Mirrors unit_tests/tensor_test.cpp and the op
coverage of gpu_tensor_ops_test.cpp."""
import io

import pytest
import torch

from tnn_b200.core import tensor as T
from tnn_b200.core.dtype import DType
from tnn_b200.ops import tensor_ops as ops


def test_factories_and_record_round_trip(tmp_path):
    t = T.make_tensor(DType.FP32, [2, 3, 4])
    assert t.shape == (2, 3, 4) and t.dtype == torch.float32 and float(t.abs().sum()) == 0
    assert T.make_tensor([5]).shape == (5,)
    assert T.make_tensor("BF16", [2, 2]).dtype == torch.bfloat16
    T.fill_random_uniform(t, -1, 1, seed=3)
    assert -1 <= float(t.min()) and float(t.max()) <= 1 and float(t.abs().sum()) > 0
    u = T.create_like(t)
    assert u.shape == t.shape and float(u.abs().sum()) == 0
    T.save(t, tmp_path / "t.bin")
    assert torch.equal(T.load(tmp_path / "t.bin"), t)
    buf = io.BytesIO()
    T.save(t, buf)
    buf.seek(0)
    assert torch.equal(T.load_into(buf, torch.empty(2, 3, 4)), t)
    assert T.dtype_cast(t, "BF16").dtype == torch.bfloat16
    s = T.stats(t)
    assert abs(s["mean"] - float(t.mean())) < 1e-6 and abs(s["variance"] - float(t.var(unbiased=False))) < 1e-6
    assert len(T.head(t, 5)) == 5 and T.span(t, 1, 1).shape == (1, 3, 4) and T.nbytes(t) == 96
    d = torch.zeros(3, 4)
    T.copy_batch(d, 2, torch.ones(2, 4), 1)
    assert float(d[2].sum()) == 4


def test_allocator_backed_tensor():
    from tnn_b200.core.device import DELAllocator
    from tnn_b200.ops import native
    if native.load() is None:
        return
    a = DELAllocator("cpu", slab_bytes=1 << 20)
    t = T.make_tensor(a, DType.FP32, [16, 16])
    assert T.aligned(t) and float(t.abs().sum()) == 0


def test_elementwise_family():
    a, b, c = torch.randn(37), torch.randn(37).abs() + 0.5, torch.randn(37)
    torch.testing.assert_close(ops.add(a, b), a + b)
    torch.testing.assert_close(ops.sub(a, b), a - b)
    torch.testing.assert_close(ops.mul(a, b), a * b)
    torch.testing.assert_close(ops.div(a, b), a / b)
    torch.testing.assert_close(ops.minimum(a, b), torch.minimum(a, b))
    torch.testing.assert_close(ops.maximum(a, b), torch.maximum(a, b))
    torch.testing.assert_close(ops.greater(a, b), (a > b).float())
    torch.testing.assert_close(ops.equal(a, a), torch.ones(37))
    torch.testing.assert_close(ops.mul_scalar(a, 2.5), a * 2.5)
    torch.testing.assert_close(ops.add_scalar(a, -1.0), a - 1.0)
    torch.testing.assert_close(ops.sqrt(b), b.sqrt())
    torch.testing.assert_close(ops.rsqrt(b), b.rsqrt())
    torch.testing.assert_close(ops.rcp(b), 1 / b)
    torch.testing.assert_close(ops.abs(a), a.abs())
    torch.testing.assert_close(ops.clamp(a, -0.5, 0.5), a.clamp(-0.5, 0.5))
    torch.testing.assert_close(ops.fmadd(a, b, c), a * b + c)
    torch.testing.assert_close(ops.fmsub(a, b, c), a * b - c)
    torch.testing.assert_close(ops.fnmadd(a, b, c), c - a * b)
    y = c.clone()
    torch.testing.assert_close(ops.axpy(y, a, 0.5), c + 0.5 * a)
    assert abs(float(ops.sum(a)) - float(a.sum())) < 1e-4
    assert abs(float(ops.dot_product(a, b)) - float((a * b).sum())) < 1e-4
    assert abs(float(ops.sum_squared_diff(a, 0.25)) - float(((a - 0.25) ** 2).sum())) < 1e-4
    z = torch.zeros(8)
    assert float(ops.set_scalar(z, 3.0).sum()) == 24
    assert ops.cast(a, torch.bfloat16).dtype == torch.bfloat16
    w = torch.tensor([1, 256], dtype=torch.int32)
    ops.bswap(w)
    assert w.tolist() == [1 << 24, 1 << 16]
    n1 = ops.fill_random_normal(torch.empty(1000), 1.0, 2.0, 7)
    n2 = ops.fill_random_normal(torch.empty(1000), 1.0, 2.0, 7)
    assert torch.equal(n1, n2) and abs(float(n1.mean()) - 1.0) < 0.3


def test_layout_ops():
    x = torch.arange(2 * 4 * 5 * 3, dtype=torch.float32).reshape(2, 4, 5, 3)
    p = ops.pad(x, 1, 2, -1.0)
    assert p.shape == (2, 6, 9, 3) and float(p[0, 0, 0, 0]) == -1.0
    torch.testing.assert_close(p[:, 1:5, 2:7], x)
    torch.testing.assert_close(ops.unpad(p, 1, 2), x)
    torch.testing.assert_close(ops.crop(x, 1, 2, 2, 3), x[:, 1:3, 2:5])
    torch.testing.assert_close(ops.nhwc_to_nchw(x), x.permute(0, 3, 1, 2))
    torch.testing.assert_close(ops.nchw_to_nhwc(ops.nhwc_to_nchw(x)), x)
    torch.testing.assert_close(ops.transpose_2d(x.reshape(8, 15)), x.reshape(8, 15).t())
    assert [t.shape[0] for t in ops.split(x, 2)] == [1, 1] and ops.slice_batch(x, 1, 1).shape[0] == 1
    # im2col agrees with a conv expressed as a GEMM
    w = torch.randn(6, 3, 3, 3)                                   # [K, R, S, C]
    col = ops.im2col(x, 3, 3, 1, 1)
    y = (col @ w.reshape(6, -1).t()).reshape(2, 4, 5, 6)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    torch.testing.assert_close(y, ref, atol=1e-3, rtol=1e-4)
    s = ops.softmax(x)
    torch.testing.assert_close(s, torch.softmax(x, -1))
    g = torch.randn_like(x)
    xx = x.clone().requires_grad_(True)
    torch.softmax(xx, -1).backward(g)
    torch.testing.assert_close(ops.softmax_backward(s, g), xx.grad, atol=1e-6, rtol=1e-4)


def test_tensor_facade_remaining_members():
    """clone / resize / ensure / share_from / same_shape / shape_str / scalar arithmetic / statistics
    (include/tensor/tensor.hpp)."""
    import pytest
    import torch
    from tnn_b200.core import tensor as T
    a = T.make_tensor([2, 3, 4])
    T.fill(a, 1.5)
    b = T.clone(a)
    b[0, 0, 0] = 9.0
    assert float(a[0, 0, 0]) == 1.5 and T.same_shape(a, b) and T.shape_str(a) == "[2, 3, 4]"
    s = T.share_from(a, [6, 4])
    s[0, 0] = -2.0
    assert float(a[0, 0, 0]) == -2.0                                # shared storage
    assert T.capacity(a) >= 24
    small = T.ensure(a, [3, 4])
    assert small.data_ptr() == a.data_ptr() and tuple(small.shape) == (3, 4)     # reuses the buffer
    big = T.ensure(a, [10, 10])
    assert big.data_ptr() != a.data_ptr() and tuple(big.shape) == (10, 10)
    assert tuple(T.resize(a, [4, 6]).shape) == (4, 6) and T.resize(a, [4, 6]).data_ptr() == a.data_ptr()
    x = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    assert torch.allclose(T.add_scalar(x, 2.0), x + 2) and torch.allclose(T.sub_scalar(x, 1.0), x - 1)
    assert torch.allclose(T.mul_scalar(x, 0.5), x * 0.5) and torch.allclose(T.div_scalar(x, 4.0), x / 4)
    with pytest.raises(ZeroDivisionError):
        T.div_scalar(x, 0.0)
    assert T.tmin(x) == 0.0 and T.tmax(x) == 11.0 and abs(T.mean(x) - 5.5) < 1e-6
    assert abs(T.variance(x) - float(x.var(unbiased=False))) < 1e-4


def test_tensor_utils_col2im_layout_and_diagnostics(capsys):
    """col2im is the adjoint of im2col; CNHW transposes; NaN/Inf check; value distribution
    (include/tensor/tensor_utils.hpp, tensor_ops.hpp)."""
    import pytest
    import torch
    from tnn_b200.ops import tensor_ops as ops
    x = torch.randn(2, 6, 5, 3)
    for (R, S, st, pd) in [(3, 3, 1, 1), (2, 2, 2, 0), (3, 3, 2, 1)]:
        col = ops.im2col(x, R, S, st, pd)
        g = torch.randn_like(col)
        back = ops.col2im(g, x.shape, R, S, st, pd)
        # <im2col(x), g> == <x, col2im(g)>
        assert abs(float((col * g).sum()) - float((x * back).sum())) < 1e-3 * max(1.0, float((col * g).abs().sum()))
    y = torch.randn(2, 3, 4, 5)
    assert torch.equal(ops.cnhw_to_nchw(ops.nchw_to_cnhw(y)), y) and ops.nchw_to_cnhw(y).shape == (3, 2, 4, 5)
    assert ops.check_nan_and_inf(y, "y")
    bad = y.clone()
    bad[0, 0, 0, 0] = float("nan")
    bad[1, 1, 1, 1] = float("inf")
    assert not ops.check_nan_and_inf(bad, "bad")
    assert "1 NaN, 1 Inf" in capsys.readouterr().out
    with pytest.raises(FloatingPointError):
        ops.check_nan_and_inf(bad, "bad", raise_on_error=True)
    edges, counts = ops.buckets(torch.arange(100, dtype=torch.float32), 10)
    assert counts.tolist() == [10] * 10 and float(edges[0]) == 0.0 and float(edges[-1]) == 99.0
    assert "distribution over 4 buckets" in ops.print_data_distribution(torch.randn(50), 4, "w")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_native_cpu_elementwise_family_matches_torch(dtype):
    """ops/cpu/{s,d}kernels.cpp twin: the native multi-versioned loops behind every ops.* call on contiguous CPU
    fp32 / fp64 tensors (sizes straddle the per-task grain so the thread-pool path and the tails are covered)."""
    from tnn_b200.ops import native, tensor_ops as ops
    C = native.require()
    assert hasattr(C, "cpu_binary") and hasattr(C, "cpu_reduce")
    g = torch.Generator().manual_seed(5)
    tol = dict(rtol=1e-6, atol=1e-6) if dtype == torch.float32 else dict(rtol=1e-12, atol=1e-12)
    for n in (1, 7, 1000, (1 << 15) + 3, 200_001):
        a = torch.randn(n, generator=g, dtype=dtype)
        b = torch.randn(n, generator=g, dtype=dtype) + 0.1
        c = torch.randn(n, generator=g, dtype=dtype)
        for name, ref in (("add", a + b), ("sub", a - b), ("mul", a * b), ("div", a / b), ("minimum", torch.minimum(a, b)),
                          ("maximum", torch.maximum(a, b)), ("equal", (a == b).to(dtype)), ("greater", (a > b).to(dtype))):
            torch.testing.assert_close(getattr(ops, name)(a, b), ref, **tol)
        torch.testing.assert_close(ops.add(a, 2.5), a + 2.5, **tol)
        torch.testing.assert_close(ops.div(a, 4.0), a / 4.0, **tol)
        torch.testing.assert_close(ops.maximum(a, 0.0), a.clamp_min(0), **tol)
        p = a.abs() + 0.5
        for name, ref in (("sqrt", p.sqrt()), ("rsqrt", p.rsqrt()), ("rcp", 1 / p), ("abs", a.abs()), ("neg", -a),
                          ("exp", a.exp()), ("log", p.log()), ("square", a * a)):
            src = p if name in ("sqrt", "rsqrt", "rcp", "log") else a
            torch.testing.assert_close(getattr(ops, name)(src), ref, rtol=tol["rtol"] * 10, atol=tol["atol"] * 10)
        torch.testing.assert_close(ops.clamp(a, -0.5, 0.25), a.clamp(-0.5, 0.25), **tol)
        torch.testing.assert_close(ops.fmadd(a, b, c), a * b + c, **tol)
        torch.testing.assert_close(ops.fmsub(a, b, c), a * b - c, **tol)
        torch.testing.assert_close(ops.fnmadd(a, b, c), c - a * b, **tol)
        y = c.clone()
        ops.axpy(y, a, -1.5)
        torch.testing.assert_close(y, c - 1.5 * a, **tol)
        assert float(ops.set_scalar(y, 3.25).min()) == 3.25 == float(y.max())
        inplace = a.clone()
        ops.add(inplace, b, out=inplace)                                     # aliasing output
        torch.testing.assert_close(inplace, a + b, **tol)
        ad = a.double()
        assert float(ops.sum(a)) == pytest.approx(float(ad.sum()), rel=1e-6, abs=1e-6)
        assert float(ops.dot_product(a, b)) == pytest.approx(float((ad * b.double()).sum()), rel=1e-6, abs=1e-5)
        assert float(ops.sum_squared_diff(a, 0.3)) == pytest.approx(float(((ad - 0.3) ** 2).sum()), rel=1e-6)
    # thread count does not change a reduction (per-grain partials combined in order)
    big = torch.randn(500_000, generator=g, dtype=dtype)
    r1 = C.cpu_reduce(big.data_ptr(), 0, 0.0, big.numel(), 0, dtype == torch.float64, 1)
    r8 = C.cpu_reduce(big.data_ptr(), 0, 0.0, big.numel(), 0, dtype == torch.float64, 8)
    assert r1 == r8
    # strided views and integer tensors keep working through the torch path
    v = torch.arange(12, dtype=dtype).reshape(3, 4).t()
    torch.testing.assert_close(ops.add(v, v), v + v)
    assert ops.add(torch.tensor([1, 2]), torch.tensor([3, 4])).tolist() == [4, 6]
