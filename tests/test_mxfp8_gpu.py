"""Block-scaled fp8 kernels (csrc/kernels/mxfp8.cu) against plain fp32 PyTorch oracles."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


def _ref_quant(x):
    """OCP MX rule in torch: per 32-block scale 2^(floor(log2 amax) - 8), e4m3 round-to-nearest, saturate."""
    rows, K = x.shape
    xb = x.float().view(rows, K // 32, 32)
    amax = xb.abs().amax(-1, keepdim=True)
    e = torch.where(amax > 0, torch.floor(torch.log2(amax.clamp_min(1e-38))) - 8, torch.full_like(amax, -8.0))
    scale = torch.exp2(e)
    q = (xb / scale).clamp(-448, 448).to(torch.float8_e4m3fn).float()
    return (q * scale).view(rows, K), (e.view(rows, K // 32) + 127).to(torch.uint8)


@pytest.mark.parametrize("rows,K,dtype", [(128, 128, torch.bfloat16), (200, 384, torch.bfloat16), (777, 768, torch.float32),
                                          (33, 3072, torch.bfloat16)])
def test_quantize_matches_mx_rule_and_layout(rows, K, dtype):
    from tnn_b200.ops import mxfp8
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(rows * 7 + K)
    x = (torch.randn(rows, K, generator=g) * torch.exp(torch.randn(rows, 1, generator=g) * 2)).to(dtype).to(dev)
    x[0, :32] = 0                                                    # an all-zero block
    q, sf = mxfp8.quantize(x)
    deq_ref, sf_ref = _ref_quant(x)
    r = torch.arange(rows, device=dev).view(-1, 1).expand(rows, K // 32)
    gi = torch.arange(K // 32, device=dev).view(1, -1).expand(rows, K // 32)
    got_sf = sf[mxfp8.sf_index(r, gi, K)]
    assert torch.equal(got_sf[1:], sf_ref[1:]) and torch.equal(got_sf[0, 1:], sf_ref[0, 1:])
    torch.testing.assert_close(mxfp8.dequantize_torch(q, sf), deq_ref, rtol=0, atol=0)
    torch.testing.assert_close(mxfp8.dequantize(q, sf), deq_ref, rtol=0, atol=0)
    # padding rows of the spare scale block decode to 1.0 (never NaN = 0xFF)
    assert int(sf.max()) < 255
    # quantisation error bound: e4m3 has 3 mantissa bits -> relative step 2^-3 of the block maximum's binade
    err = (deq_ref - x.float()).abs().view(rows, K // 32, 32).amax(-1)
    amax = x.float().abs().view(rows, K // 32, 32).amax(-1)
    assert bool((err <= amax / 8 + 1e-30).all())


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (256, 192, 256), (300, 768, 768), (1000, 2304, 768), (513, 3072, 768),
                                   (384, 768, 3072), (64, 50257, 768), (129, 100, 384)])
@pytest.mark.parametrize("block_n,pair", [(None, 0), (128, 0), (None, 1), (128, 1)])
def test_gemm_mxfp8_equals_fp32_matmul_of_dequantised_operands(M, N, K, block_n, pair, monkeypatch):
    """The tensor-core result must equal (A_dq @ B_dq^T) computed in fp32 from the SAME quantised operands:
    any slip in the scale layout, the scale-id bytes or the tile / scale-block offsets shows up as O(1) errors."""
    from tnn_b200.ops import mxfp8
    dev = _dev()
    if block_n is not None:
        monkeypatch.setenv("TNN_FP8_BLOCK_N", str(block_n))
    monkeypatch.setenv("TNN_FP8_PAIR", str(pair))      # 1: two-CTA clusters multicasting the shared B tile (odd tile counts included)
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + 5 * K)
    a = (torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, K // 32, generator=g).repeat_interleave(32, 1))).bfloat16().to(dev)
    b = (torch.randn(N, K, generator=g) * torch.exp(torch.randn(N, K // 32, generator=g).repeat_interleave(32, 1) * 0.5)).to(dev)
    aq, asf = mxfp8.quantize(a)
    bq, bsf = mxfp8.quantize(b)
    ref = mxfp8.dequantize_torch(aq, asf).double() @ mxfp8.dequantize_torch(bq, bsf).double().t()
    out, _, _ = mxfp8.gemm(aq, asf, bq, bsf)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs()
    # bf16 output rounding (2^-8 relative) + fp32 accumulation noise
    assert float(err.max()) <= 6e-3 * scale, (float(err.max()), scale)
    rel = float((err / (ref.abs() + 1e-3 * scale)).mean())
    assert rel < 3e-3, rel


def test_gemm_mxfp8_epilogue_bias_residual_gelu_and_requantised_output():
    from tnn_b200.ops import mxfp8
    dev = _dev()
    torch.manual_seed(5)
    M, N, K = 640, 3072, 768
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05)
    bias = torch.randn(N, device=dev) * 0.1
    res = torch.randn(M, N, device=dev).bfloat16()
    aq, asf = mxfp8.quantize(a)
    wq, wsf = mxfp8.quantize(w)
    lin = mxfp8.dequantize_torch(aq, asf) @ mxfp8.dequantize_torch(wq, wsf).t()
    # bias + residual, no activation
    out, _, _ = mxfp8.gemm(aq, asf, wq, wsf, bias=bias, residual=res)
    torch.testing.assert_close(out.float(), lin + bias + res.float(), rtol=2e-2, atol=2e-2)
    # GELU (tanh form) + second output re-quantised for the next GEMM
    out, oq, osf = mxfp8.gemm(aq, asf, wq, wsf, bias=bias, act=mxfp8.ACT_GELU, want_fp8=True)
    ref = torch.nn.functional.gelu(lin + bias, approximate="tanh")
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
    deq = mxfp8.dequantize_torch(oq, osf)
    # the fp8 copy is the MX quantisation of the fp32 epilogue value: within one e4m3 step of the block maximum
    amax = ref.abs().view(M, N // 32, 32).amax(-1, keepdim=True).expand(M, N // 32, 32).reshape(M, N)
    assert bool(((deq - ref).abs() <= amax / 8 + 2e-2).all())
    # and it feeds a second GEMM
    w2 = torch.randn(K, N, device=dev) * 0.05
    w2q, w2sf = mxfp8.quantize(w2)
    out2, _, _ = mxfp8.gemm(oq, osf, w2q, w2sf)
    ref2 = deq @ mxfp8.dequantize_torch(w2q, w2sf).t()
    torch.testing.assert_close(out2.float(), ref2, rtol=2e-2, atol=2e-2 * float(ref2.abs().max()))


@pytest.mark.parametrize("C", [256, 768, 1024])
def test_layernorm_mxfp8_matches_layernorm_then_quantise(C):
    from tnn_b200.ops import mxfp8
    dev = _dev()
    torch.manual_seed(C)
    rows = 515
    x = torch.randn(rows, C, device=dev).bfloat16()
    resid = torch.randn(rows, C, device=dev).bfloat16()
    gamma = torch.rand(C, device=dev) + 0.5
    beta = torch.randn(C, device=dev) * 0.1
    s = torch.empty_like(x)
    y = torch.empty_like(x)
    q, sf = mxfp8.layernorm_quant(x, gamma, beta, 1e-5, residual=resid, sum_out=s, y_bf16=y)
    sref = (x.float() + resid.float()).bfloat16()
    assert torch.equal(s, sref)
    yref = torch.nn.functional.layer_norm(sref.float(), (C,), gamma, beta, 1e-5)
    torch.testing.assert_close(y.float(), yref, rtol=2e-2, atol=2e-2)
    deq = mxfp8.dequantize_torch(q, sf)
    amax = yref.abs().view(rows, C // 32, 32).amax(-1, keepdim=True).expand(rows, C // 32, 32).reshape(rows, C)
    assert bool(((deq - yref).abs() <= amax / 8 + 1e-2).all())
    assert int(sf.max()) < 255
