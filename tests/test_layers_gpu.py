"""Layer CPU<->GPU parity ("device agnosticity", reference: unit_tests/layer_device_agnosticity_test.cpp):
the same layer with the same parameters must produce the same forward/backward on the fp32
CPU reference path and on the sm_100a kernel path (bf16 tolerance)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from tnn_b200.nn import layers as L  # noqa: E402
from tnn_b200.nn import blocks as B  # noqa: E402
from tnn_b200.nn.builder import LayerBuilder  # noqa: E402


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _pair(make, seed=3):
    cpu = make().init("cpu", seed)
    gpu = make().init("cpu", seed).to("cuda")
    return cpu, gpu


def _check(make, x, tol=2e-2, train=True, int_input=False, grad_tol=6e-2, worst_element=True):
    """``worst_element``: also bound the largest single-element error (single layers).  Off for multi-layer blocks with
    ReLU / BatchNorm inside: a pre-activation that rounds across zero in bf16 legitimately flips one mask element."""
    torch.manual_seed(1234)          # dy below must not depend on the process-wide RNG state
    cpu, gpu = _pair(make)
    cpu.set_training(train), gpu.set_training(train)
    xq = x if int_input else x.bfloat16().float()
    yc = cpu.forward(xq)
    yg = gpu.forward(xq.cuda() if int_input else xq.cuda().bfloat16())
    assert yc.shape == yg.shape
    assert _rel(yg, yc) < tol, f"forward mismatch {_rel(yg, yc)}"
    # norm-relative bounds would hide a wrong tile on a small fraction of elements: bound the WORST element too
    # (bf16 inputs / outputs: 2^-8 relative per element plus accumulation noise)
    worst = float((yg.float().cpu() - yc.float()).abs().max()) if worst_element else 0.0
    assert worst <= 4 * tol * max(1.0, float(yc.float().abs().max())), f"forward max-abs mismatch {worst}"
    if not train:
        return
    dy = torch.randn_like(yc).bfloat16().float()
    dc = cpu.backward(dy)
    dg = gpu.backward(dy.cuda().bfloat16())
    if dc is not None and dg is not None:
        assert _rel(dg, dc) < tol, f"dgrad mismatch {_rel(dg, dc)}"
        worst = float((dg.float().cpu() - dc.float()).abs().max()) if worst_element else 0.0
        assert worst <= 4 * tol * max(1.0, float(dc.float().abs().max())), f"dgrad max-abs mismatch {worst}"
    for lc, lg in zip(cpu.all_layers(), gpu.all_layers()):
        for k in lc.grads:
            gc, gg = lc.grads[k], lg.grads[k]
            if gc.norm() > 1e-3 * max(1.0, gc.numel() ** 0.5) * 1e-1:
                # BN affine grads are sums over few (here 64..256) bf16 samples per channel: looser bound
                lim = 0.15 if k in ("gamma", "beta") else grad_tol
                assert _rel(gg, gc) < lim, f"{lc.name}.{k} grad mismatch {_rel(gg, gc)}"


def test_dense():
    _check(lambda: L.DenseLayer(96, 100, True, "fc"), torch.randn(64, 96))


def test_dense_with_fused_activation_trains():
    """activation='gelu' / 'relu' in the Dense epilogue: inference is one kernel, training keeps the pre-activation
    (round 1 raised NotImplementedError in backward)."""
    _check(lambda: L.DenseLayer(96, 100, True, "fc", activation="gelu"), torch.randn(64, 96))
    # relu: a pre-activation that rounds across zero in bf16 flips one mask element and with it a whole dx row
    _check(lambda: L.DenseLayer(96, 100, True, "fc", activation="relu"), torch.randn(64, 96), tol=5e-2, worst_element=False)
    for act in ("gelu", "relu"):
        _check(lambda: L.DenseLayer(96, 100, True, "fc", activation=act), torch.randn(64, 96), train=False)


def test_legacy_batchnorm_runs_natively_on_nchw():
    """legacy_batchnorm normalises the NCHW tensor in its own layout (norm_nchw.cu): 3 launches forward, 3 backward, no
    NHWC transposes (round 1: transpose -> NHWC kernels -> transpose)."""
    from tnn_b200.ops import native
    _check(lambda: L.LegacyBatchNormLayer(32, 1e-5, 0.1, True, False, name="lbn"), torch.randn(8, 32, 12, 12))       # vector path
    _check(lambda: L.LegacyBatchNormLayer(24, 1e-5, 0.1, True, False, name="lbn"), torch.randn(5, 24, 7, 7))         # H*W = 49: scalar path
    _check(lambda: L.LegacyBatchNormLayer(16, 1e-3, 0.2, True, True, name="lbn"), torch.randn(8, 16, 8, 8), worst_element=False)
    _check(lambda: L.LegacyBatchNormLayer(16, 1e-3, 0.2, False, False, name="lbn"), torch.randn(4, 16, 8, 8), train=False)
    lay = L.LegacyBatchNormLayer(32, name="lbn").init("cpu", 1).to("cuda")
    lay.set_training(True)
    x = torch.randn(8, 32, 12, 12, device="cuda").bfloat16()
    n0 = native.require().launch_count()
    y = lay.forward(x, 0)
    n1 = native.require().launch_count()
    lay.backward(torch.randn_like(y), 0)
    n2 = native.require().launch_count()
    assert (n1 - n0, n2 - n1) == (3, 3), (n1 - n0, n2 - n1)
    # running statistics follow the unbiased-variance rule of the reference
    xf = x.float()
    want_var = 0.9 * 1.0 + 0.1 * xf.permute(1, 0, 2, 3).reshape(32, -1).var(1, unbiased=True)
    assert torch.allclose(lay.inner.params["running_var"], want_var, atol=2e-3)
    # fp32 activations stay fp32
    y32 = lay.forward(x.float(), 1)
    assert y32.dtype == torch.float32 and _rel(y32, y) < 1e-2


def test_conv_3x3():
    _check(lambda: L.Conv2DLayer(32, 64, 3, 3, 1, 1, 1, 1, True, "c"), torch.randn(4, 16, 16, 32))


def test_conv_anisotropic_stride_and_padding_runs_on_the_own_kernels():
    """stride (2, 1), padding (1, 0), a 3 x 2 window: im2col + tcgen05 GEMM of this repo (round 1 fell back to F.conv2d)."""
    import inspect
    assert "F.conv2d" not in inspect.getsource(L.Conv2DLayer)
    _check(lambda: L.Conv2DLayer(8, 16, 3, 2, 2, 1, 1, 0, True, "c"), torch.randn(2, 13, 10, 8))


def test_dense_with_fused_residual_when_rows_are_not_multiples_of_8():
    cpu = L.DenseLayer(96, 100, True, "fc").init("cpu", 3)
    gpu = L.DenseLayer(96, 100, True, "fc").init("cpu", 3).to("cuda")
    cpu.set_training(False), gpu.set_training(False)
    x = torch.randn(64, 96).bfloat16().float()
    r = torch.randn(64, 100).bfloat16().float()
    yc = cpu.forward_impl(x, 0, residual=r)
    yg = gpu.forward_impl(x.cuda().bfloat16(), 0, residual=r.cuda().bfloat16())
    assert _rel(yg, yc) < 2e-2
    assert float((yg.float().cpu() - yc).abs().max()) < 0.08 * float(yc.abs().max())


def test_conv_first_layer_direct():
    _check(lambda: L.Conv2DLayer(3, 16, 3, 3, 1, 1, 1, 1, True, "c"), torch.randn(4, 32, 32, 3))


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("C", [16, 128, 24])
def test_batchnorm(relu, C):
    _check(lambda: L.BatchNormLayer(C, 1e-5, 0.1, True, relu, "bn"), torch.randn(8, 8, 8, C) * 2 + 0.5)


def test_batchnorm_eval():
    _check(lambda: L.BatchNormLayer(64, 1e-5, 0.1, True, True, "bn"), torch.randn(4, 4, 4, 64), train=False)


def test_batchnorm_running_stats():
    cpu, gpu = _pair(lambda: L.BatchNormLayer(32, 1e-5, 0.1, True, False, "bn"))
    x = (torch.randn(16, 4, 4, 32) * 3 + 1).bfloat16().float()
    cpu.forward(x), gpu.forward(x.cuda().bfloat16())
    assert _rel(gpu.params["running_mean"], cpu.params["running_mean"]) < 1e-2
    assert _rel(gpu.params["running_var"], cpu.params["running_var"]) < 1e-2


def test_maxpool_avgpool():
    _check(lambda: L.MaxPool2DLayer(2, 2, 2, 2, 0, 0, "mp"), torch.randn(4, 8, 8, 16))
    _check(lambda: L.MaxPool2DLayer(3, 3, 2, 2, 1, 1, "mp"), torch.randn(2, 9, 9, 8))
    _check(lambda: L.AvgPool2DLayer(8, 8, 1, 1, 0, 0, "ap"), torch.randn(4, 8, 8, 32))
    _check(lambda: L.AvgPool2DLayer(2, 2, 2, 2, 0, 0, "ap"), torch.randn(4, 8, 8, 16))


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "sigmoid", "tanh", "elu", "gelu"])
def test_activations(act):
    _check(lambda: L.ActivationLayer(act, act), torch.randn(4, 33, 7))


def test_layernorm_groupnorm():
    _check(lambda: L.LayerNormLayer(96, 1e-5, True, "ln"), torch.randn(6, 10, 96))
    _check(lambda: L.LayerNormLayer(768, 1e-5, True, "ln"), torch.randn(4, 40, 768))      # 3 vectors per lane
    _check(lambda: L.LayerNormLayer(1280, 1e-5, True, "ln"), torch.randn(2, 33, 1280))    # 5 -> 8-vector variant
    _check(lambda: L.LayerNormLayer(100, 1e-5, True, "ln"), torch.randn(7, 9, 100))       # scalar fallback (C % 8 != 0)
    _check(lambda: L.GroupNormLayer(4, 32, 1e-5, True, "gn"), torch.randn(3, 5, 5, 32))


def test_embedding():
    ids = torch.randint(0, 50, (4, 12))
    _check(lambda: L.EmbeddingLayer(50, 32, -1, "emb"), ids, int_input=True)


def test_residual_block_fused():
    def make():
        b = LayerBuilder([[8, 8, 32]])
        b.wide_residual_block(32, 64, 2, 0.0, "wrb")
        return b.build()[0]
    # a chain of conv -> BN -> ReLU -> conv -> BN with bf16 activations between the kernels vs an fp32
    # chain: per-layer rounding compounds through the BN normalisation, hence the wider bounds and
    # the larger batch (more samples per channel statistic)
    _check(make, torch.randn(16, 8, 8, 32), tol=8e-2, grad_tol=0.1, worst_element=False)

    def make2():
        b = LayerBuilder([[8, 8, 32]])
        b.basic_residual_block(32, 32, 1, "brb")
        return b.build()[0]
    _check(make2, torch.randn(16, 8, 8, 32), tol=8e-2, grad_tol=0.1, worst_element=False)


def test_softmax_ce_and_adam():
    from tnn_b200.nn import LossFactory
    crit = LossFactory.create_logsoftmax_crossentropy()
    logits = torch.randn(64, 100).bfloat16().float()
    y = torch.randint(0, 100, (64,))
    sc, gc = crit.loss_and_grad(logits, y)
    sg, gg = crit.loss_and_grad(logits.cuda().bfloat16(), y.cuda().int())
    assert abs(float(sc[0]) - float(sg[0])) < 2e-2
    assert float(sc[1]) == float(sg[1])
    assert _rel(gg, gc) < 2e-2
    # flat Adam kernel vs reference math
    from tnn_b200.ops import native
    Cx = native.require()
    n = 10007
    p = torch.randn(n); g = torch.randn(n); m = torch.zeros(n); v = torch.zeros(n)
    pg, gg2, mg, vg = p.cuda(), g.cuda(), m.cuda(), v.cuda()
    lp = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    for t in range(1, 4):
        Cx.adam_step_flat(pg, gg2, mg, vg, lp, 1e-2, 0.9, 0.999, 1e-3, 3e-4, False, t, 1.0, False)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        upd = 1e-2 * (m / (1 - 0.9 ** t)) / ((v / (1 - 0.999 ** t)).sqrt() + 1e-3) + 3e-4 * 1e-2 * p
        p = p - upd
    assert _rel(pg, p) < 1e-5
    assert _rel(lp, p) < 1e-2


def test_batchnorm_statistics_fused_into_conv_epilogue():
    """conv -> BN: after the first step the conv epilogue accumulates the BN statistics; results must
    match the stand-alone statistics kernel (and the running averages must keep evolving)."""
    import os
    from tnn_b200.nn import BatchNormLayer, Conv2DLayer, ResidualBlock, Sequential
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    for (N, H, C, K, st) in [(8, 16, 64, 128, 1), (4, 32, 128, 256, 2), (16, 8, 256, 512, 1)]:
        def build():
            conv = Conv2DLayer(C, K, 3, 3, st, st, 1, 1, True, "c").init(dev, 3)
            bn = BatchNormLayer(K, 1e-2, 0.1, True, True, "b").init(dev, 3)
            return conv, bn
        x = torch.randn(N, H, H, C, device=dev).bfloat16()
        outs = {}
        for mode in ("0", "1"):
            os.environ["TNN_BN_FUSE_STATS"] = mode
            conv, bn = build()
            for it in range(3):
                y = bn.forward(conv.forward(x, 0), 0)
            outs[mode] = (y.float(), bn.params["running_mean"].clone(), bn.params["running_var"].clone(), conv._stats_ws)
        os.environ["TNN_BN_FUSE_STATS"] = "1"
        assert outs["1"][3] is not None and outs["0"][3] is None          # the conv only accumulates when asked to
        assert (outs["1"][0] - outs["0"][0]).abs().max().item() < 3e-2
        torch.testing.assert_close(outs["1"][1], outs["0"][1], atol=2e-3, rtol=2e-3)
        torch.testing.assert_close(outs["1"][2], outs["0"][2], atol=2e-3, rtol=2e-3)
    # residual block output (conv epilogue with residual add) feeding the next block's BatchNorm
    main = Sequential([BatchNormLayer(128, 1e-2, use_relu=True, name="b1"), Conv2DLayer(128, 128, 3, 3, 1, 1, 1, 1, True, "c1")], "m")
    blk = ResidualBlock(main, None, "linear", "r").init(dev, 5)
    nxt = BatchNormLayer(128, 1e-2, use_relu=True, name="b2").init(dev, 5)
    x = torch.randn(8, 16, 16, 128, device=dev).bfloat16()
    ref = None
    for it in range(3):
        out = blk.forward(x, 0)
        used = getattr(out, "_tnn_bn_stats", None) is not None
        y = nxt.forward(out, 0)
        if it == 0:
            assert not used
            ref = y.float()
        else:
            assert used
            assert (y.float() - ref).abs().max().item() < 3e-2
    os.environ.pop("TNN_BN_FUSE_STATS", None)


@pytest.mark.parametrize("relu,p", [(True, 0.0), (True, 0.3), (False, 0.3)])
def test_batchnorm_backward_mask_recompute_equals_stored_mask(relu, p):
    """Backward rebuilds the ReLU/dropout mask from x + statistics + the dropout counter; it must give
    the same gradients as reading the mask from the stored forward output."""
    import os
    from tnn_b200.nn import BatchNormLayer
    dev = torch.device("cuda", 0)
    x = (torch.randn(16, 12, 12, 64, device=dev) * 1.5 + 0.3).bfloat16()
    dy = torch.randn(16, 12, 12, 64, device=dev).bfloat16()
    res = {}
    os.environ["TNN_BN_BITMASK"] = "0"          # compare the two y-based / x-based mask sources here
    for mode in ("0", "1"):
        os.environ["TNN_BN_RECOMPUTE_MASK"] = mode
        bn = BatchNormLayer(64, 1e-2, 0.1, True, relu, "bn").init(dev, 7)
        torch.manual_seed(99)
        bn.params["gamma"].uniform_(0.5, 1.5)
        bn.params["beta"].uniform_(-0.5, 0.5)
        bn.fused_dropout_p = p
        y = bn.forward(x, 0)
        assert (bn.cache(0)["y"] is None) == (mode == "1")
        dx = bn.backward(dy, 0)
        res[mode] = (y.float(), dx.float(), bn.grads["gamma"].clone(), bn.grads["beta"].clone())
    os.environ.pop("TNN_BN_RECOMPUTE_MASK", None)
    os.environ.pop("TNN_BN_BITMASK", None)
    # (the batch statistics are accumulated with atomics: the two forwards agree to rounding, not bitwise)
    assert _rel(res["1"][0], res["0"][0]) < 5e-3
    assert (res["0"][1] - res["1"][1]).abs().max().item() <= 2e-2 * res["0"][1].abs().max().item()
    # identical masks except for a handful of elements exactly on the ReLU boundary; the statistics of the two
    # forwards differ in the last bits (atomics), so compare in norm rather than element by element
    assert _rel(res["1"][1], res["0"][1]) < 1e-2
    torch.testing.assert_close(res["0"][2], res["1"][2], atol=5e-2, rtol=1e-2)
    torch.testing.assert_close(res["0"][3], res["1"][3], atol=5e-2, rtol=1e-2)


@pytest.mark.parametrize("relu,p,C,K", [(True, 0.0, 64, 128), (True, 0.3, 128, 128), (False, 0.3, 32, 64), (True, 0.3, 256, 256)])
def test_batchnorm_backward_fused_into_dgrad_epilogue(relu, p, C, K):
    """BN(+ReLU+dropout) -> conv3x3: the conv's dgrad epilogue does the BatchNorm backward reduction
    (mask, scale, per-channel sums); gradients must match the separate reduce + apply kernels."""
    import os
    from tnn_b200.nn.builder import LayerBuilder
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    x = (torch.randn(8, 16, 16, C, device=dev) * 1.5 + 0.3).bfloat16()
    dy = torch.randn(8, 16, 16, K, device=dev).bfloat16()
    res = {}
    for mode in ("0", "1"):
        os.environ["TNN_BN_FUSE_BWD"] = mode
        b = LayerBuilder([[16, 16, C]])
        b.batchnorm(1e-2, 0.1, True, relu, "bn")
        if p > 0:
            b.dropout(p, "drop")
        b.conv2d(K, 3, 3, 1, 1, 1, 1, True, "conv")
        seq = B.Sequential(b.build(), "s").init(dev, 11)
        bn, conv = seq.layers[0], seq.layers[-1]
        assert conv._bn_prev is bn and bn._bwd_consumer is conv
        torch.manual_seed(99)
        bn.params["gamma"].uniform_(0.5, 1.5)
        bn.params["beta"].uniform_(-0.5, 0.5)
        y = seq.forward(x, 0)
        dx = seq.backward(dy, 0)
        res[mode] = (y.float(), dx.float(), bn.grads["gamma"].clone(), bn.grads["beta"].clone(),
                     conv.grads["weight"].clone(), conv.grads["bias"].clone())
    os.environ.pop("TNN_BN_FUSE_BWD", None)
    a, f = res["0"], res["1"]
    assert _rel(f[0], a[0]) < 1e-2          # (statistics are atomically accumulated: not bit-reproducible)
    assert _rel(f[1], a[1]) < 2e-2, f"dx {_rel(f[1], a[1])}"
    assert _rel(f[2], a[2]) < 2e-2 and _rel(f[3], a[3]) < 2e-2, f"dgamma {_rel(f[2], a[2])} dbeta {_rel(f[3], a[3])}"
    assert _rel(f[4], a[4]) < 1e-2 and _rel(f[5], a[5]) < 1e-2      # the conv's own grads do not depend on the fusion


@pytest.mark.parametrize("relu,p,C", [(True, 0.0, 64), (True, 0.3, 128), (False, 0.3, 40)])
def test_batchnorm_backward_keep_bits_equal_stored_output_mask(relu, p, C):
    """The forward writes one byte of keep bits per 8-channel vector; backward from those bits must
    equal backward from the mask read off the stored output (y != 0)."""
    import os
    from tnn_b200.nn import BatchNormLayer
    dev = torch.device("cuda", 0)
    torch.manual_seed(8)
    x = (torch.randn(16, 12, 12, C, device=dev) * 1.5 + 0.3).bfloat16()
    dy = torch.randn(16, 12, 12, C, device=dev).bfloat16()
    extra = torch.randn(16, 12, 12, C, device=dev).bfloat16()
    res = {}
    for mode in ("0", "1"):
        os.environ["TNN_BN_BITMASK"] = mode
        bn = BatchNormLayer(C, 1e-2, 0.1, True, relu, "bn").init(dev, 7)
        torch.manual_seed(99)
        bn.params["gamma"].uniform_(0.5, 1.5)
        bn.params["beta"].uniform_(-0.5, 0.5)
        bn.fused_dropout_p = p
        y = bn.forward(x, 0)
        assert (bn.cache(0)["bits"] is not None) == (mode == "1") and (bn.cache(0)["y"] is None) == (mode == "1")
        dx = bn.backward_impl(dy, 0, extra=extra)
        res[mode] = (y.float(), dx.float(), bn.grads["gamma"].clone(), bn.grads["beta"].clone())
    os.environ.pop("TNN_BN_BITMASK", None)
    assert _rel(res["1"][0], res["0"][0]) < 5e-3            # (batch statistics are atomically accumulated)
    assert _rel(res["1"][1], res["0"][1]) < 1e-2            # identical masks; only the atomic summation order differs
    torch.testing.assert_close(res["0"][2], res["1"][2], atol=2e-2, rtol=1e-3)   # (atomic accumulation order)
    torch.testing.assert_close(res["0"][3], res["1"][3], atol=2e-2, rtol=1e-3)


def test_transformer_ffn_residual_block_fuses_shortcut_gradient_into_layernorm_backward():
    """LN -> Dense -> GELU -> Dense inside a residual block: the shortcut gradient is added by the LayerNorm
    dx kernel (no separate add); forward / backward must match the fp32 CPU path."""
    def make():
        b = LayerBuilder([[24, 256]])
        fb = b._sub().layernorm(1e-5, True, "ln").dense(512, True, "fc1").activation("gelu", "act").dense(256, True, "fc2")
        b.residual_block(fb.build(), None, "linear", "res")
        return b.build()[0]
    _check(make, torch.randn(4, 24, 256), tol=4e-2, grad_tol=8e-2, worst_element=False)
