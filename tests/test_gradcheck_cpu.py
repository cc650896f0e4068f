"""Finite-difference gradient checks of the layer backward passes (CPU fp32 oracle path).

SURVEY 4 asks for what the reference lacks: every analytic backward (input gradient AND parameter
gradients) is compared with central differences of a scalar loss L = sum(forward(x) * r) for a fixed
random r.  The CPU path is the oracle that every GPU kernel is compared against elsewhere, so its own
derivatives are pinned here."""
import pytest
import torch

from tnn_b200.nn import blocks as B
from tnn_b200.nn import layers as L
from tnn_b200.nn.builder import LayerBuilder


def _loss(layer, x, r):
    return float((layer.forward(x, 0).double() * r).sum())


def _check_layer(make, x, eps=1e-2, tol=2e-2, n_probe=12, int_input=False):
    torch.manual_seed(0)
    layer = make().init("cpu", 5)
    layer.set_training(True)
    # dropout-free, deterministic forward is required for differences to make sense
    y = layer.forward(x, 0)
    r = torch.randn(y.shape, dtype=torch.float64)
    dx = layer.backward(r.float(), 0)
    g = torch.Generator().manual_seed(1)
    # ---- input gradient
    if not int_input and dx is not None:
        flat = x.reshape(-1)
        for idx in torch.randperm(flat.numel(), generator=g)[:n_probe].tolist():
            old = float(flat[idx])
            flat[idx] = old + eps
            lp = _loss(layer, x, r)
            flat[idx] = old - eps
            lm = _loss(layer, x, r)
            flat[idx] = old
            fd = (lp - lm) / (2 * eps)
            an = float(dx.reshape(-1)[idx])
            assert abs(fd - an) <= tol * max(1.0, abs(fd), abs(an)), f"dx[{idx}]: fd {fd} vs analytic {an}"
    # ---- parameter gradients
    for sub in layer.all_layers():
        for name, grad in sub.grads.items():
            p = sub.params[name].reshape(-1)
            ga = grad.reshape(-1).clone()
            for idx in torch.randperm(p.numel(), generator=g)[:max(2, n_probe // 3)].tolist():
                old = float(p[idx])
                p[idx] = old + eps
                if hasattr(sub, "mark_params_dirty"):
                    sub.mark_params_dirty()
                lp = _loss(layer, x, r)
                p[idx] = old - eps
                lm = _loss(layer, x, r)
                p[idx] = old
                fd = (lp - lm) / (2 * eps)
                an = float(ga[idx])
                assert abs(fd - an) <= tol * max(1.0, abs(fd), abs(an)), f"{sub.name}.{name}[{idx}]: fd {fd} vs analytic {an}"


def test_dense_and_activations():
    _check_layer(lambda: L.DenseLayer(12, 7, True, "fc"), torch.randn(5, 12))
    for act in ("tanh", "sigmoid", "gelu", "elu", "leaky_relu"):
        _check_layer(lambda a=act: L.ActivationLayer(a, "act"), torch.randn(4, 9) * 0.7 + 0.2)


@pytest.mark.parametrize("stride", [1, 2])
def test_conv2d(stride):
    _check_layer(lambda: L.Conv2DLayer(3, 4, 3, 3, stride, stride, 1, 1, True, "c"), torch.randn(2, 6, 6, 3))


def test_normalisation_layers():
    _check_layer(lambda: L.BatchNormLayer(5, 1e-3, 0.1, True, False, "bn"), torch.randn(6, 3, 3, 5) * 1.5 + 0.3)
    _check_layer(lambda: L.LayerNormLayer(10, 1e-3, True, "ln"), torch.randn(3, 4, 10))
    _check_layer(lambda: L.GroupNormLayer(2, 6, 1e-3, True, "gn"), torch.randn(2, 3, 3, 6))


def test_pooling_and_embedding():
    _check_layer(lambda: L.AvgPool2DLayer(2, 2, 2, 2, 0, 0, "ap"), torch.randn(2, 4, 4, 3))
    # max pooling: keep the maxima well separated so +-eps never changes the arg-max
    x = torch.randn(2, 4, 4, 3)
    x = x + torch.arange(x.numel(), dtype=torch.float32).reshape(x.shape) * 0.05
    _check_layer(lambda: L.MaxPool2DLayer(2, 2, 2, 2, 0, 0, "mp"), x, eps=1e-3)
    _check_layer(lambda: L.EmbeddingLayer(11, 6, -1, "emb"), torch.randint(0, 11, (3, 5)), int_input=True)


def test_attention_and_residual_blocks():
    _check_layer(lambda: B.AttentionBlock(16, 2, True, "attn"), torch.randn(2, 5, 16) * 0.5, tol=3e-2)

    def res():
        b = LayerBuilder([[6, 8]])
        fb = b._sub().layernorm(1e-3, True, "ln").dense(12, True, "fc1").activation("tanh", "act").dense(8, True, "fc2")
        b.residual_block(fb.build(), None, "linear", "res")
        return b.build()[0]
    _check_layer(res, torch.randn(3, 6, 8) * 0.5)
