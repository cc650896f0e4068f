"""fp16 / fp32 (tf32 tensor cores) / fp64 on the GPU: GEMMs (csrc/kernels/gemm_xt.cu), the templated tensor ops,
casts, activations and the Dense layer's explicit IO dtype contract.

Reference behaviour: include/type/type.hpp (fp16 / bf16 / fp32 / fp64 element types), src/math/cuda/cublas_gemm.cu:20-164
(one cuBLAS call per dtype), src/nn/layers_impl/dense_layer.cpp:336-340 (DenseLayer instantiated per dtype).
Every check is against a plain PyTorch fp32 / fp64 computation of the same op on the SAME (already rounded) inputs.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _operand(rows, k, mn, dtype, dev, g):
    """[rows, k] logical operand stored K-major ([rows, k]) or MN-major ([k, rows])."""
    t = torch.randn((k, rows) if mn else (rows, k), generator=g, device=dev, dtype=torch.float32)
    t = t.to(dtype)
    logical = (t.t() if mn else t).double()
    return t, logical


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 72, 136), (384, 264, 520), (64, 40, 24)])
def test_gemm_xt_matches_fp64_reference(dtype, a_mn, b_mn, M, N, K):
    from tnn_b200.ops import gemm as G
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K)
    a, la = _operand(M, K, a_mn, dtype, dev, g)
    b, lb = _operand(N, K, b_mn, dtype, dev, g)
    out = G.gemm_xt(a, b, a_mn, b_mn)
    torch.cuda.synchronize()
    ref = la @ lb.t()
    assert out.dtype == dtype and out.shape == (M, N)
    err = (out.double() - ref).abs().max().item()
    bound = (la.abs() @ lb.abs().t()).max().item()          # sum |a||b| of the worst element
    tol = {torch.float16: 2e-3, torch.float32: 1.5e-3, torch.float64: 1e-13}[dtype] * bound
    assert err <= tol, f"{dtype} M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn}: max err {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_gemm_xt_epilogue_bias_residual_act_accumulate(dtype):
    from tnn_b200.ops import gemm as G
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(5)
    M, N, K = 160, 200, 192
    a = torch.randn(M, K, generator=g, device=dev).to(dtype)
    b = (torch.randn(N, K, generator=g, device=dev) * 0.1).to(dtype)
    bias = torch.randn(N, generator=g, device=dev)
    res = torch.randn(M, N, generator=g, device=dev).to(dtype)
    for act, fn in (("relu", torch.relu), ("gelu", lambda t: torch.nn.functional.gelu(t, approximate="tanh")), (None, lambda t: t)):
        y = G.gemm_xt(a, b, bias=bias, residual=res, act=act, alpha=0.5)
        ref = fn(0.5 * (a.double() @ b.double().t()) + bias.double() + res.double())
        tol = 2e-2 if dtype == torch.float16 else 1e-2
        assert (y.double() - ref).abs().max().item() < tol, act
    # accumulate into an fp32 buffer (weight-gradient form): out += alpha * A^T-ish product
    acc = torch.randn(M, N, generator=g, device=dev)
    want = acc.double() + 0.25 * (a.double() @ b.double().t())
    G.gemm_xt(a, b, out=acc, alpha=0.25, accumulate=True)
    assert (acc.double() - want).abs().max().item() < (1e-2 if dtype == torch.float16 else 6e-3)
    if dtype == torch.float16:
        y32 = G.gemm_xt(a, b, out_f32=True)
        assert y32.dtype == torch.float32
        assert (y32.double() - a.double() @ b.double().t()).abs().max().item() < 5e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.float64, torch.float32, torch.bfloat16])
def test_tensor_ops_every_dtype(dtype):
    from tnn_b200.ops import tensor_ops as T
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(11)
    n = 4099                                                # not a multiple of any vector width
    a = (torch.rand(n, generator=g, device=dev) + 0.5).to(dtype)
    b = (torch.rand(n, generator=g, device=dev) + 0.5).to(dtype)
    c = torch.randn(n, generator=g, device=dev).to(dtype)
    ad, bd, cd = a.double(), b.double(), c.double()
    eps = {torch.float16: 2e-3, torch.bfloat16: 1.6e-2, torch.float32: 1e-6, torch.float64: 1e-14}[dtype]

    def close(got, want):
        assert got.dtype == dtype
        w = want.to(dtype).double()
        assert ((got.double() - w).abs() <= eps * (1 + w.abs())).all()

    close(T.add(a, b), ad + bd)
    close(T.sub(a, b), ad - bd)
    close(T.mul(a, b), ad * bd)
    close(T.div(a, b), ad / bd)
    close(T.maximum(a, b), torch.maximum(ad, bd))
    close(T.mul(a, 3.0), ad * 3.0)
    close(T.sqrt(a), ad.sqrt())
    close(T.rsqrt(a), ad.rsqrt())
    close(T.exp(c.clamp(-3, 3)), c.clamp(-3, 3).double().exp())
    close(T.log(a), ad.log())
    close(T.clamp(c, -0.5, 0.5), cd.clamp(-0.5, 0.5))
    close(T.fmadd(a, b, c), ad * bd + cd)
    s = T.sum(a)
    assert abs(s.item() - ad.sum().item()) < 2e-3 * n * (1e-3 if dtype in (torch.float32, torch.float64) else 1)
    d = T.dot_product(a, b)
    assert abs(d.item() - (ad * bd).sum().item()) < 1e-3 * n
    x = torch.randn(37, 129, generator=g, device=dev).to(dtype)
    y = T.softmax(x)
    ref = torch.softmax(x.double(), -1)
    assert y.dtype == dtype and (y.double() - ref).abs().max().item() < 10 * eps
    dy = torch.randn(37, 129, generator=g, device=dev).to(dtype)
    dx = T.softmax_backward(y, dy)
    yd, gd = y.double(), dy.double()
    refdx = yd * (gd - (gd * yd).sum(-1, keepdim=True))
    assert (dx.double() - refdx).abs().max().item() < 20 * eps
    p = T.pad(x.reshape(1, 37, 129, 1).contiguous(), 2, 1)
    assert p.dtype == dtype and p.shape == (1, 41, 131, 1) and torch.equal(p[0, 2:-2, 1:-1, 0], x)


def test_cast_and_transpose_every_pair():
    from tnn_b200.ops import tensor_ops as T
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(3)
    src = torch.randn(1000, generator=g, device=dev, dtype=torch.float64)
    kinds = (torch.float16, torch.bfloat16, torch.float32, torch.float64)
    for s in kinds:
        a = src.to(s)
        for d in kinds:
            got = T.cast(a, d)
            assert got.dtype == d and torch.equal(got, a.to(d)), (s, d)
        m = a[:999].reshape(27, 37).contiguous()
        assert torch.equal(T.transpose_2d(m), m.t().contiguous())


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64])
def test_dense_layer_explicit_io_dtype(dtype):
    """DenseLayer with io_dtype fp16 / fp32 / fp64: forward, input gradient, weight and bias gradients."""
    from tnn_b200.nn.layers import DenseLayer
    dev = _dev()
    torch.manual_seed(0)
    lay = DenseLayer(96, 72, True, "d").set_io_dtype(dtype).init(dev, 4)
    lay.set_training(True)
    x = torch.randn(50, 96, device=dev).to(dtype)
    y = lay.forward(x, 0)
    assert y.dtype == dtype
    w, b = lay.params["weight"].double(), lay.params["bias"].double()
    wd = w.to(dtype).double() if dtype == torch.float16 else w
    ref = x.double() @ wd.t() + b
    tol = {torch.float16: 2e-2, torch.float32: 5e-3, torch.float64: 1e-6}[dtype]
    assert (y.double() - ref).abs().max().item() < tol
    dy = torch.randn(50, 72, device=dev).to(dtype)
    for gname in lay.grads:
        lay.grads[gname].zero_()
    dx = lay.backward(dy, 0)
    assert dx.dtype == dtype
    assert (dx.double() - dy.double() @ wd).abs().max().item() < tol * 2
    assert (lay.grads["weight"].double() - dy.double().t() @ x.double()).abs().max().item() < tol * 8
    assert (lay.grads["bias"].double() - dy.double().sum(0)).abs().max().item() < tol * 4


@pytest.mark.parametrize("dtype", [torch.float16, torch.float64])
def test_activation_layer_fp16_fp64(dtype):
    from tnn_b200.nn.layers import ActivationLayer
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(2)
    x = torch.randn(33, 65, generator=g, device=dev).to(dtype)
    dy = torch.randn(33, 65, generator=g, device=dev).to(dtype)
    for name in ("relu", "gelu", "tanh", "sigmoid", "softmax"):
        lay = ActivationLayer(name).init(dev, 0)
        lay.set_training(True)
        y = lay.forward(x, 0)
        dx = lay.backward(dy, 0)
        xr = x.detach().double().clone().requires_grad_(True)       # a fresh leaf (x.double() aliases x for fp64)
        ref = {"relu": torch.relu, "gelu": lambda t: torch.nn.functional.gelu(t, approximate="tanh"), "tanh": torch.tanh,
               "sigmoid": torch.sigmoid, "softmax": lambda t: torch.softmax(t, -1)}[name](xr)
        ref.backward(dy.double())
        tol = 4e-3 if dtype == torch.float16 else 1e-12           # fp64 activations evaluate in double precision
        assert y.dtype == dtype and (y.double() - ref.detach()).abs().max().item() < tol, name
        assert dx.dtype == dtype and (dx.double() - xr.grad).abs().max().item() < 3 * tol, name


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("B,L,H", [(2, 128, 2), (1, 200, 3), (1, 1024, 2)])
def test_fp16_attention_kernels_match_fp32_reference(B, L, H, causal):
    """The fused tcgen05 attention forward / backward with fp16 operands (the dtype the reference's flash attention
    requires, cudnn_flash_attention_ops.cu:123) against the fp32 formula."""
    from tnn_b200.ops import attention as A
    dev = _dev()
    torch.manual_seed(4)
    qkv = (torch.randn(B, L, 3, H, 64, device=dev) * 0.8).half()
    do = (torch.randn(B, L, H * 64, device=dev) * 0.5).half()
    o, lse = A.attention_fwd(qkv, causal)
    assert o.dtype == torch.float16 and o.shape == (B, L, H * 64)
    ref_in = qkv.float().requires_grad_(True)
    oref, lref = A.attention_ref(ref_in[:, :, 0], ref_in[:, :, 1], ref_in[:, :, 2], causal)

    def rel(a, b):
        return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()

    assert rel(o.reshape(B, L, H, 64), oref.detach()) < 5e-3                    # fp16: 11-bit mantissa (bf16 bound is 2e-2)
    assert (lse - lref.detach()).abs().max().item() < 5e-3
    got = A.attention_bwd(qkv, do, o, lse, causal)
    (oref.reshape(B, L, H * 64) * do.float()).sum().backward()
    assert got.dtype == torch.float16 and got.shape == qkv.shape
    for i, name in enumerate(("dq", "dk", "dv")):
        assert rel(got[:, :, i], ref_in.grad[:, :, i]) < 8e-3, name


def test_attention_block_fp16_io_dtype_trains():
    """AttentionBlock / FlashAttentionBlock with io_dtype fp16: fp16 projections (gemm_xt) + fp16 attention, parity with CPU."""
    from tnn_b200.nn import FlashAttentionBlock
    dev = _dev()
    torch.manual_seed(5)
    cpu = FlashAttentionBlock(128, 2, True, "fa").init("cpu", 7)
    gpu = FlashAttentionBlock(128, 2, True, "fa").set_io_dtype(torch.float16).init("cpu", 7).to(dev)
    cpu.set_training(True), gpu.set_training(True)
    x = torch.randn(2, 96, 128).half().float()
    dy = torch.randn(2, 96, 128).half().float()
    yc = cpu.forward(x, 0)
    yg = gpu.forward(x.to(dev).half(), 0)
    assert yg.dtype == torch.float16
    assert ((yg.float().cpu() - yc).norm() / yc.norm()).item() < 5e-3
    dc = cpu.backward(dy, 0)
    dg = gpu.backward(dy.to(dev).half(), 0)
    assert dg.dtype == torch.float16 and ((dg.float().cpu() - dc).norm() / dc.norm()).item() < 1e-2
    for k in cpu.grads:
        assert ((gpu.grads[k].cpu() - cpu.grads[k]).norm() / (cpu.grads[k].norm() + 1e-12)).item() < 2e-2, k


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64])
def test_layernorm_explicit_io_dtype(dtype):
    """LayerNormLayer with io_dtype fp16 / fp32 / fp64 (any-dtype warp-per-row kernels): forward with the fused residual add,
    backward with the fused extra gradient, dgamma / dbeta -- against autograd in fp64."""
    from tnn_b200.nn.layers import LayerNormLayer
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(8)
    lay = LayerNormLayer(200, 1e-5, True, "ln").set_io_dtype(dtype).init(dev, 2)
    lay.params["gamma"].copy_(torch.rand(200, generator=g, device=dev) + 0.5)
    lay.params["beta"].copy_(torch.randn(200, generator=g, device=dev))
    lay.set_training(True)
    x = torch.randn(37, 200, generator=g, device=dev).to(dtype)
    res = torch.randn(37, 200, generator=g, device=dev).to(dtype)
    dy = torch.randn(37, 200, generator=g, device=dev).to(dtype)
    extra = torch.randn(37, 200, generator=g, device=dev).to(dtype)
    y = lay.forward_impl(x, 0, residual=res)
    assert y.dtype == dtype
    xr = (x.double() + res.double()).to(dtype).double().requires_grad_(True)     # the layer stores the sum in the io dtype
    gam = lay.params["gamma"].double().clone().requires_grad_(True)
    bet = lay.params["beta"].double().clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (200,), gam, bet, 1e-5)
    tol = {torch.float16: 4e-3, torch.float32: 2e-5, torch.float64: 2e-6}[dtype]      # statistics are kept in fp32
    assert (y.double() - ref.detach()).abs().max().item() < tol * 4
    for k in lay.grads:
        lay.grads[k].zero_()
    dx = lay.backward_impl(dy, 0, extra=extra)
    ref.backward(dy.double())
    assert dx.dtype == dtype and (dx.double() - (xr.grad + extra.double())).abs().max().item() < tol * 8
    assert (lay.grads["gamma"].double() - gam.grad).abs().max().item() < tol * 40
    assert (lay.grads["beta"].double() - bet.grad).abs().max().item() < tol * 40


def test_tiny_gpt2_runs_in_fp16_end_to_end():
    """A 2-layer GPT-2 (head_dim 64) with io_dtype fp16 on every layer: embeddings -> LayerNorm (any-dtype kernel) -> fp16 QKV /
    attention / projection -> fp16 MLP -> LM head; forward logits and parameter gradients against the CPU fp32 run."""
    from tnn_b200.models.zoo import _gpt2
    dev = _dev()
    cpu = _gpt2(128, 2, 2, True, "tiny", None, seq_len=64, vocab=97, dropout=0.0).init("cpu", 11)
    gpu = _gpt2(128, 2, 2, True, "tiny", torch.float16, seq_len=64, vocab=97, dropout=0.0).init("cpu", 11).to(dev)
    cpu.set_training(True), gpu.set_training(True)
    ids = (torch.arange(3 * 64).reshape(3, 64) * 7 % 97).to(torch.int32)
    yc = cpu.forward(ids, 0)
    yg = gpu.forward(ids.to(dev), 0)
    assert yg.dtype == torch.float16 and yg.shape == yc.shape
    assert ((yg.float().cpu() - yc).norm() / yc.norm()).item() < 2e-2
    dy = torch.randn_like(yc) * 0.1
    cpu.backward(dy, 0)
    gpu.backward(dy.to(dev).half(), 0)
    worst = 0.0
    for lc, lg in zip(cpu.all_layers(), gpu.all_layers()):
        for k, gc in lc.grads.items():
            if gc.norm() > 1e-6:
                worst = max(worst, ((lg.grads[k].cpu() - gc).norm() / gc.norm()).item())
    assert worst < 6e-2, worst


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("cfg", [(8, 16, 3, 3, 1, 1, 1, 1, (2, 10, 10, 8)), (3, 8, 5, 5, 2, 2, 2, 2, (2, 11, 11, 3)), (16, 24, 3, 2, 2, 1, 1, 0, (3, 9, 8, 16))])
def test_conv2d_explicit_io_dtype(dtype, cfg):
    """Conv2DLayer with io_dtype fp16 / fp32: im2col + the fp16 / tf32 tcgen05 GEMM (forward, dgrad through col2im, wgrad with fp32
    accumulation, bias gradient) against the CPU fp32 layer; padded (R*S*C % 8 != 0) and anisotropic geometries included."""
    from tnn_b200.nn.layers import Conv2DLayer
    dev = _dev()
    cin, cout, kh, kw, sh, sw, ph, pw, shape = cfg
    cpu = Conv2DLayer(cin, cout, kh, kw, sh, sw, ph, pw, True, "c").init("cpu", 5)
    gpu = Conv2DLayer(cin, cout, kh, kw, sh, sw, ph, pw, True, "c").set_io_dtype(dtype).init("cpu", 5).to(dev)
    cpu.set_training(True), gpu.set_training(True)
    torch.manual_seed(9)
    x = torch.randn(*shape).half().float()
    yc = cpu.forward(x, 0)
    yg = gpu.forward(x.to(dev).to(dtype), 0)
    assert yg.dtype == dtype and yg.shape == yc.shape
    tol = 4e-3 if dtype == torch.float16 else 2e-3
    assert ((yg.float().cpu() - yc).norm() / yc.norm()).item() < tol
    dy = torch.randn_like(yc).half().float()
    dc = cpu.backward(dy, 0)
    dg = gpu.backward(dy.to(dev).to(dtype), 0)
    assert dg.dtype == dtype and ((dg.float().cpu() - dc).norm() / dc.norm()).item() < 2 * tol
    for k in cpu.grads:
        assert ((gpu.grads[k].cpu() - cpu.grads[k]).norm() / cpu.grads[k].norm()).item() < 2 * tol, k
