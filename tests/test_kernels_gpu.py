"""GPU numerics: every hand-written kernel vs a plain PyTorch fp32 reference of the same op.

Mirrors the reference's CPU-oracle strategy (unit_tests/cuda_dense_ops_test.cpp,
cuda_conv2d_nchw_ops_test.cpp, cuda_batchnorm_ops_test.cpp, cuda_loss_funcs_test.cpp).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


# ----------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 256), (200, 100, 72), (1000, 512, 768),
                                   (4096, 768, 3072), (37, 24, 40)])
def test_linear_fwd(M, N, K):
    from tnn_b200.ops import gemm
    torch.manual_seed(0)
    x = torch.randn(M, K, device=_dev()).bfloat16()
    w = (torch.randn(N, K, device=_dev()) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device=_dev())
    y = gemm.linear_fwd(x, w, b)
    ref = x.float() @ w.float().t() + b
    assert _rel_err(y, ref) < 1e-2
    y2 = gemm.linear_fwd(x, w, b, act="relu")
    assert _rel_err(y2, torch.relu(ref)) < 1e-2
    y3 = gemm.linear_fwd(x, w, None, act="gelu")
    assert _rel_err(y3, torch.nn.functional.gelu(x.float() @ w.float().t(), approximate="tanh")) < 1e-2
    res = torch.randn(M, N, device=_dev()).bfloat16()
    y4 = gemm.linear_fwd(x, w, b, residual=res)
    assert _rel_err(y4, ref + res.float()) < 1e-2


@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (512, 768, 3072), (200, 104, 72)])
def test_linear_dgrad_wgrad(M, N, K):
    from tnn_b200.ops import gemm
    torch.manual_seed(1)
    x = torch.randn(M, K, device=_dev()).bfloat16()
    w = (torch.randn(N, K, device=_dev()) / math.sqrt(K)).bfloat16()
    dy = torch.randn(M, N, device=_dev()).bfloat16()
    dx = gemm.linear_dgrad(dy, w)
    assert _rel_err(dx, dy.float() @ w.float()) < 1e-2
    dw = torch.zeros(N, K, device=_dev())
    gemm.linear_wgrad(dy, x, dw)
    ref = dy.float().t() @ x.float()
    assert _rel_err(dw, ref) < 1e-2
    gemm.linear_wgrad(dy, x, dw)  # accumulate (beta = 1)
    assert _rel_err(dw, 2 * ref) < 1e-2
    db = torch.zeros(N, device=_dev())
    gemm.bias_grad(dy, db)
    assert _rel_err(db, dy.float().sum(0)) < 1e-2


# ----------------------------------------------------------------------------- Conv
CONV_CASES = [
    # N, H, W, C, K, R, stride, pad
    (8, 32, 32, 16, 128, 3, 1, 1),
    (8, 32, 32, 128, 128, 3, 1, 1),
    (8, 16, 16, 256, 256, 3, 1, 1),
    (8, 8, 8, 512, 512, 3, 1, 1),
    (8, 32, 32, 128, 256, 3, 2, 1),
    (8, 16, 16, 256, 512, 3, 2, 1),
    (8, 32, 32, 16, 128, 1, 1, 0),
    (8, 32, 32, 128, 256, 1, 2, 0),
    (6, 4, 4, 512, 512, 3, 1, 1),
    (4, 32, 32, 3, 16, 3, 1, 1),      # direct path
    (4, 28, 28, 8, 16, 5, 1, 0),      # im2col path
]


@pytest.mark.parametrize("N,H,W,C,K,R,stride,pad", CONV_CASES)
def test_conv2d(N, H, W, C, K, R, stride, pad):
    from tnn_b200.ops import conv
    torch.manual_seed(2)
    dev = _dev()
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = torch.randn(K, R, R, C, device=dev) / math.sqrt(R * R * C)
    w = w.bfloat16().float()
    b = torch.randn(K, device=dev)
    plan = conv.make_plan(N, H, W, C, K, R, R, stride, pad, dev)
    st = conv.ConvState()
    cache = {}
    y = conv.conv2d_fwd(plan, x, w, w.bfloat16(), b, cache=cache)
    ref = conv.conv2d_fwd_ref(x, w, b, stride, pad)
    assert y.shape == ref.shape
    assert _rel_err(y, ref) < 1e-2, plan.path
    dy = torch.randn_like(ref).bfloat16()
    if plan.path != "direct":
        dx = conv.conv2d_dgrad(plan, dy, w, st)
        dref = conv.conv2d_dgrad_ref(dy, w, (N, H, W, C), stride, pad)
        assert _rel_err(dx, dref) < 1e-2, plan.path
        res = torch.randn(N, H, W, C, device=dev).bfloat16()
        dx2 = conv.conv2d_dgrad(plan, dy, w, st, residual=res)
        assert _rel_err(dx2, dref + res.float()) < 1e-2, plan.path
    dw = torch.zeros(K, R, R, C, device=dev)
    db = torch.zeros(K, device=dev)
    conv.conv2d_wgrad(plan, dy, x, dw, db, cache=cache)
    wref = conv.conv2d_wgrad_ref(dy, x, (K, R, R, C), stride, pad)
    assert _rel_err(dw, wref) < 1e-2, plan.path
    assert _rel_err(db, dy.float().sum((0, 1, 2))) < 1e-2


def test_conv_residual_relu_epilogue():
    from tnn_b200.ops import conv
    dev = _dev()
    N, H, W, C, K = 8, 16, 16, 128, 128
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = (torch.randn(K, 3, 3, C, device=dev) / math.sqrt(9 * C)).bfloat16().float()
    res = torch.randn(N, H, W, K, device=dev).bfloat16()
    plan = conv.make_plan(N, H, W, C, K, 3, 3, 1, 1, dev)
    assert plan.path == "tma"
    y = conv.conv2d_fwd(plan, x, w, w.bfloat16(), None, residual=res, act="relu")
    ref = torch.relu(conv.conv2d_fwd_ref(x, w, None, 1, 1) + res.float())
    assert _rel_err(y, ref) < 1e-2


# ----------------------------------------------------------------------------- attention (tcgen05 flash fwd + decode)
@pytest.mark.parametrize("B,L,H,causal", [(2, 128, 2, False), (2, 128, 2, True), (1, 384, 3, True), (2, 200, 4, True),
                                          (2, 200, 4, False), (1, 1024, 12, True), (3, 72, 1, True)])
def test_attention_fwd_matches_reference(B, L, H, causal):
    from tnn_b200.ops import attention as A
    from tnn_b200.ops import native
    assert hasattr(native.require(), "attention_fwd")
    torch.manual_seed(1)
    qkv = (torch.randn(B, L, 3, H, 64, device=_dev()) * 0.8).bfloat16()
    o, lse = A.attention_fwd(qkv, causal)
    oref, lref = A.attention_ref(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal)
    assert o.shape == (B, L, H * 64) and o.dtype == torch.bfloat16
    assert _rel_err(o.reshape(B, L, H, 64), oref) < 2e-2
    assert (lse - lref).abs().max().item() < 2e-2


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("B,L,H", [(2, 128, 2), (1, 200, 3), (2, 384, 2), (1, 1024, 2)])
def test_attention_bwd_matches_reference(B, L, H, causal, monkeypatch):
    """tcgen05 flash backward vs the fp32 recompute-everything formula (autograd of attention_ref)."""
    from tnn_b200.ops import attention as A
    from tnn_b200.ops import native
    assert hasattr(native.require(), "attention_bwd")
    torch.manual_seed(3)
    qkv = (torch.randn(B, L, 3, H, 64, device=_dev()) * 0.8).bfloat16()
    do = (torch.randn(B, L, H * 64, device=_dev()) * 0.5).bfloat16()
    o, lse = A.attention_fwd(qkv, causal)
    got = A.attention_bwd(qkv, do, o, lse, causal)
    ref_in = qkv.float().requires_grad_(True)
    oref, _ = A.attention_ref(ref_in[:, :, 0], ref_in[:, :, 1], ref_in[:, :, 2], causal)
    (oref.reshape(B, L, H * 64) * do.float()).sum().backward()
    ref = ref_in.grad
    assert got.shape == qkv.shape and got.dtype == torch.bfloat16
    for i, name in enumerate(("dq", "dk", "dv")):
        err = _rel_err(got[:, :, i], ref[:, :, i])
        assert err < 3e-2, f"{name} rel err {err}"
    # and it agrees with the op-by-op fallback it replaces
    monkeypatch.setenv("TNN_ATTN_BWD_KERNEL", "0")
    slow = A.attention_bwd(qkv, do, o, lse, causal)
    assert _rel_err(got, slow.float()) < 3e-2


def test_attention_kv_cache_prefill_and_decode():
    from tnn_b200.ops import attention as A
    torch.manual_seed(2)
    B, H, Lp, Lmax = 2, 3, 150, 256
    cache = A.KVCache(B, Lmax, H, 64, _dev(), torch.bfloat16)
    full = (torch.randn(B, Lp + 5, 3, H, 64, device=_dev()) * 0.7).bfloat16()
    o, _ = A.attention_fwd(full[:, :Lp], True, kv_cache=cache, pos=0)
    oref, _ = A.attention_ref(full[:, :Lp, 0], full[:, :Lp, 1], full[:, :Lp, 2], True)
    assert _rel_err(o.reshape(B, Lp, H, 64), oref) < 2e-2
    for t in range(5):
        pos = Lp + t
        step = full[:, pos:pos + 1]
        o1, _ = A.attention_fwd(step, True, kv_cache=cache, pos=pos)
        ref, _ = A.attention_ref(step[:, :, 0], full[:, :pos + 1, 1], full[:, :pos + 1, 2], True, q_offset=pos)
        assert o1.shape == (B, 1, H * 64)
        assert _rel_err(o1.reshape(B, 1, H, 64), ref) < 2e-2
    # chunked prefill continuation (L_q > 1 with an offset) goes through the tensor-core kernel
    cache2 = A.KVCache(B, Lmax, H, 64, _dev(), torch.bfloat16)
    A.attention_fwd(full[:, :100], True, kv_cache=cache2, pos=0)
    o2, _ = A.attention_fwd(full[:, 100:Lp], True, kv_cache=cache2, pos=100)
    assert _rel_err(o2.reshape(B, Lp - 100, H, 64), oref[:, 100:]) < 2e-2


# ----------------------------------------------------------------------------- generic tensor op family
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [8, 1000, 4099, 1 << 20])
def test_tensor_op_family_matches_cpu_oracle(dtype, n):
    from tnn_b200.ops import tensor_ops as ops
    torch.manual_seed(n)
    a = torch.randn(n).to(dtype)
    b = (torch.randn(n).abs() + 0.5).to(dtype)
    c = torch.randn(n).to(dtype)
    A, B, Cc = a.to(_dev()), b.to(_dev()), c.to(_dev())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for f in (ops.add, ops.sub, ops.mul, ops.div, ops.minimum, ops.maximum, ops.greater):
        torch.testing.assert_close(f(A, B).cpu().float(), f(a, b).float(), atol=tol, rtol=tol)
    for f in (ops.sqrt, ops.rsqrt, ops.rcp):
        torch.testing.assert_close(f(B).cpu().float(), f(b).float(), atol=tol, rtol=tol)
    torch.testing.assert_close(ops.abs(A).cpu(), a.abs())
    torch.testing.assert_close(ops.clamp(A, -0.5, 0.5).cpu(), a.clamp(-0.5, 0.5))
    torch.testing.assert_close(ops.mul_scalar(A, 1.5).cpu().float(), (a.float() * 1.5).to(dtype).float(), atol=tol, rtol=tol)
    torch.testing.assert_close(ops.fmadd(A, B, Cc).cpu().float(), (a.float() * b.float() + c.float()), atol=4 * tol, rtol=4 * tol)
    ref = a.double().sum().item()
    assert abs(float(ops.sum(A)) - ref) < 1e-3 * max(1.0, abs(ref)) + 1e-2 * (n ** 0.5) * (dtype == torch.bfloat16)
    refd = (a.double() * b.double()).sum().item()
    assert abs(float(ops.dot_product(A, B)) - refd) < 2e-3 * max(1.0, abs(refd)) + 1e-3 * n ** 0.5
    refs = ((a.double() - 0.25) ** 2).sum().item()
    assert abs(float(ops.sum_squared_diff(A, 0.25)) - refs) < 2e-3 * refs + 1e-3


def test_layout_and_softmax_kernels():
    from tnn_b200.ops import tensor_ops as ops
    x = torch.randn(3, 6, 7, 16)
    for dt in (torch.float32, torch.bfloat16):
        xc, xg = x.to(dt), x.to(dt).to(_dev())
        torch.testing.assert_close(ops.pad(xg, 2, 1, 0.0).cpu(), ops.pad(xc, 2, 1, 0.0))
        torch.testing.assert_close(ops.unpad(xg, 1, 2).cpu(), ops.unpad(xc, 1, 2))
        torch.testing.assert_close(ops.crop(xg, 1, 2, 3, 4).cpu(), ops.crop(xc, 1, 2, 3, 4))
        s = ops.softmax(xg)
        assert _rel_err(s.cpu(), torch.softmax(x, -1)) < (1e-5 if dt == torch.float32 else 1e-2)
        g = torch.randn_like(x).to(dt)
        assert _rel_err(ops.softmax_backward(s, g.to(_dev())).cpu(), ops.softmax_backward(s.cpu(), g)) < 2e-2
    w = torch.tensor([1, 256, 65536], dtype=torch.int32).to(_dev())
    assert ops.bswap(w).cpu().tolist() == [1 << 24, 1 << 16, 1 << 8]


@pytest.mark.parametrize("M,N,K", [(32, 2304, 768), (32, 768, 3072), (1, 768, 768), (7, 100, 72), (32, 50257 // 8 * 8, 768), (20, 136, 1032)])
def test_skinny_decode_gemm(M, N, K):
    from tnn_b200.ops import gemm
    x = torch.randn(M, K, device=_dev()).bfloat16()
    w = (torch.randn(N, K, device=_dev()) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device=_dev())
    res = torch.randn(M, N, device=_dev()).bfloat16()
    ref = x.float() @ w.float().t() + b
    assert _rel_err(gemm.linear_fwd(x, w, b), ref) < 1e-2
    assert _rel_err(gemm.linear_fwd(x, w, b, act="gelu"), torch.nn.functional.gelu(ref, approximate="tanh")) < 1e-2
    assert _rel_err(gemm.linear_fwd(x, w, None, act="relu", residual=res), torch.relu(ref - b + res.float())) < 1e-2


@pytest.mark.parametrize("rows,V", [(64, 5003), (300, 2048), (16, 50257)])
def test_softmax_ce_vectorised_rows_with_padded_stride(rows, V):
    """bf16 logits living in rows padded to a multiple of 8 elements (what the Dense layer emits for an
    odd vocabulary): the 16-byte-vector online-softmax kernel vs the fp32 formula, no copy of the logits."""
    from tnn_b200.nn import LossFactory
    torch.manual_seed(4)
    ld = (V + 7) // 8 * 8
    buf = torch.zeros(rows, ld, device=_dev(), dtype=torch.bfloat16)
    logits = buf[:, :V]
    logits.copy_(torch.randn(rows, V, device=_dev()) * 3)
    y = torch.randint(0, V, (rows,), device=_dev(), dtype=torch.int32)
    crit = LossFactory.create_logsoftmax_crossentropy()
    pred3 = logits.reshape(rows // 4, 4, V) if rows % 4 == 0 else logits
    stats, grad = crit.loss_and_grad(pred3, y.reshape(pred3.shape[:-1]))
    assert grad.shape == pred3.shape and grad.stride(-2) == ld          # gradient keeps the padded stride
    lf = logits.float()
    ref_loss = torch.nn.functional.cross_entropy(lf, y.long())
    ref_grad = (torch.softmax(lf, -1) - torch.nn.functional.one_hot(y.long(), V)) / rows
    assert abs(float(stats[0]) - float(ref_loss)) < 2e-2
    assert float(stats[1]) == float((lf.argmax(-1) == y.long()).sum())
    assert _rel_err(grad.reshape(rows, V), ref_grad) < 2e-2


@pytest.mark.parametrize("M,N", [(8192, 768), (4096, 1001), (2048, 72)])
def test_tall_column_sums(M, N):
    from tnn_b200.ops import gemm
    torch.manual_seed(6)
    ld = (N + 7) // 8 * 8
    buf = torch.zeros(M, ld, device=_dev(), dtype=torch.bfloat16)
    dy = buf[:, :N]
    dy.copy_(torch.randn(M, N, device=_dev()))
    db = torch.full((N,), 0.5, device=_dev())
    gemm.bias_grad(dy, db, 0.25)
    ref = 0.5 + 0.25 * dy.float().sum(0)
    assert _rel_err(db, ref) < 2e-3
