"""Hand-computed expectations on tiny tensors -- the style of the reference's layer tests
(unit_tests/legacy_dense_layer_test.cpp, legacy_conv2d_layer_test.cpp, legacy_maxpool2d_layer_test.cpp,
legacy_avgpool_layer_test.cpp, legacy_batchnorm_layer_test.cpp, groupnorm_layer_test.cpp,
residual_block_test.cpp, slice_ops_test.cpp): fixed weights, inputs whose outputs can be written down
by hand, edge cases (stride, padding, non-square windows, single element), gradient shapes and values.

The autograd-oracle versions of the same layers are in test_layers_cpu.py / test_gradcheck_cpu.py;
nothing here asks torch what the answer is.
"""
import pytest
import torch

from tnn_b200.core.config import TConfig
from tnn_b200.nn import (ActivationLayer, AvgPool2DLayer, BatchNormLayer, Conv2DLayer, DenseLayer, FlattenLayer,
                         GroupNormLayer, LayerFactory, LayerNormLayer, MaxPool2DLayer, ResidualBlock, Sequential,
                         SliceLayer)


def T(*v, shape=None):
    t = torch.tensor(v, dtype=torch.float32)
    return t.reshape(shape) if shape is not None else t


def _set(layer, **params):
    layer.init("cpu", seed=0)
    for k, v in params.items():
        layer.params[k].copy_(v.reshape(layer.params[k].shape))
    layer.zero_grads()
    return layer


# ------------------------------------------------------------------------------------- dense
def test_dense_hand_computed_forward_and_grads():
    # y = x W^T + b with W = [[1,2,3],[0,-1,1]], b = [0.5,-0.5]
    l = _set(DenseLayer(3, 2, True, "fc"), weight=T(1, 2, 3, 0, -1, 1), bias=T(0.5, -0.5)).set_training(True)
    x = T(1, 0, -1, 2, 2, 2, shape=(2, 3))
    y = l.forward(x, 0)
    assert y.tolist() == [[-1.5, -1.5], [12.5, -0.5]]
    dy = T(1, 0, 0, 2, shape=(2, 2))
    dx = l.backward(dy, 0)
    assert dx.tolist() == [[1, 2, 3], [0, -2, 2]]                       # dy W
    assert l.grads["weight"].tolist() == [[1, 0, -1], [4, 4, 4]]         # dy^T x
    assert l.grads["bias"].tolist() == [1, 2]


def test_dense_without_bias_has_no_bias_parameter_and_single_row_batch():
    l = _set(DenseLayer(2, 1, False, "fc"), weight=T(3, -2)).set_training(True)
    assert "bias" not in l.params and "bias" not in l.grads
    assert l.forward(T(1, 1, shape=(1, 2)), 0).tolist() == [[1.0]]
    assert l.backward(T(2, shape=(1, 1)), 0).tolist() == [[6, -4]]
    assert l.output_shape((7, 2)) == (7, 1)


# ------------------------------------------------------------------------------------- conv2d
def _arange_img(h, w):
    return torch.arange(1, h * w + 1, dtype=torch.float32).reshape(1, h, w, 1)


def test_conv2d_ones_kernel_is_a_window_sum():
    l = _set(Conv2DLayer(1, 1, 2, 2, 1, 1, 0, 0, False, "c"), weight=torch.ones(4)).set_training(True)
    y = l.forward(_arange_img(3, 3), 0)                                  # 1 2 3 / 4 5 6 / 7 8 9
    assert y.reshape(2, 2).tolist() == [[12, 16], [24, 28]]
    dx = l.backward(torch.ones(1, 2, 2, 1), 0)                           # how many windows cover each pixel
    assert dx.reshape(3, 3).tolist() == [[1, 2, 1], [2, 4, 2], [1, 2, 1]]
    assert l.grads["weight"].reshape(2, 2).tolist() == [[12, 16], [24, 28]]


def test_conv2d_padding_stride_and_bias():
    # 3x3 ones kernel, pad 1, stride 2 on a 4x4 ramp: output 2x2, each = sum of the in-bounds 3x3 window + bias
    l = _set(Conv2DLayer(1, 1, 3, 3, 2, 2, 1, 1, True, "c"), weight=torch.ones(9), bias=T(10.0)).set_training(True)
    x = _arange_img(4, 4)
    assert l.output_shape((1, 4, 4, 1)) == (1, 2, 2, 1)
    y = l.forward(x, 0).reshape(2, 2)
    # windows centred at (0,0), (0,2), (2,0), (2,2)
    assert y.tolist() == [[1 + 2 + 5 + 6 + 10, 2 + 3 + 4 + 6 + 7 + 8 + 10],
                          [5 + 6 + 9 + 10 + 13 + 14 + 10, 6 + 7 + 8 + 10 + 11 + 12 + 14 + 15 + 16 + 10]]
    l.backward(torch.ones(1, 2, 2, 1), 0)
    assert float(l.grads["bias"]) == 4.0


def test_conv2d_non_square_kernel_and_multi_channel_mixing():
    # 1x2 kernel over 2 input channels -> 2 output channels; weight layout [C_out, KH, KW, C_in]
    w = torch.zeros(2, 1, 2, 2)
    w[0, 0, 0, 0] = 1.0       # out0 = left pixel, channel 0
    w[0, 0, 1, 1] = 2.0       #       + 2 * right pixel, channel 1
    w[1, 0, 1, 0] = -1.0      # out1 = -(right pixel, channel 0)
    l = _set(Conv2DLayer(2, 2, 1, 2, 1, 1, 0, 0, False, "c"), weight=w).set_training(True)
    x = T(1, 10, 2, 20, 3, 30, shape=(1, 1, 3, 2))                       # pixels (1,10) (2,20) (3,30)
    y = l.forward(x, 0)
    assert y.shape == (1, 1, 2, 2)
    assert y.reshape(2, 2).tolist() == [[1 + 2 * 20, -2], [2 + 2 * 30, -3]]


def test_conv2d_1x1_stride2_subsamples():
    l = _set(Conv2DLayer(1, 1, 1, 1, 2, 2, 0, 0, False, "c"), weight=T(2.0)).set_training(True)
    y = l.forward(_arange_img(4, 4), 0)
    assert y.reshape(2, 2).tolist() == [[2, 6], [18, 22]]
    dx = l.backward(torch.ones(1, 2, 2, 1), 0).reshape(4, 4)
    assert dx.tolist() == [[2, 0, 2, 0], [0, 0, 0, 0], [2, 0, 2, 0], [0, 0, 0, 0]]


# ------------------------------------------------------------------------------------- pooling
def test_maxpool_2x2_values_and_gradient_routing():
    l = MaxPool2DLayer(2, 2, name="mp").init("cpu").set_training(True)
    x = T(1, 3, 2, 4,
          5, 6, 8, 7,
          9, 2, 1, 0,
          3, 4, 5, 6, shape=(1, 4, 4, 1))
    y = l.forward(x, 0)
    assert y.reshape(2, 2).tolist() == [[6, 8], [9, 6]]
    dx = l.backward(T(1, 2, 3, 4, shape=(1, 2, 2, 1)), 0).reshape(4, 4)
    assert dx.tolist() == [[0, 0, 0, 0], [0, 1, 2, 0], [3, 0, 0, 0], [0, 0, 0, 4]]


def test_maxpool_overlapping_windows_accumulate_gradient():
    l = MaxPool2DLayer(2, 2, 1, 1, name="mp").init("cpu").set_training(True)
    x = T(0, 0, 0,
          0, 9, 0,
          0, 0, 0, shape=(1, 3, 3, 1))
    assert l.forward(x, 0).reshape(-1).tolist() == [9, 9, 9, 9]
    dx = l.backward(torch.ones(1, 2, 2, 1), 0).reshape(3, 3)
    assert dx.tolist() == [[0, 0, 0], [0, 4, 0], [0, 0, 0]]


def test_maxpool_non_square_window_padding_and_negative_inputs():
    l = MaxPool2DLayer(1, 3, 1, 1, 0, 1, name="mp").init("cpu").set_training(True)
    x = T(-5, -1, -3, -2, shape=(1, 1, 4, 1))
    assert l.output_shape((1, 1, 4, 1)) == (1, 1, 4, 1)
    # padding never wins: it behaves as -inf, not as 0
    assert l.forward(x, 0).reshape(-1).tolist() == [-1, -1, -1, -2]


def test_avgpool_values_and_gradient():
    l = AvgPool2DLayer(2, 2, name="ap").init("cpu").set_training(True)
    y = l.forward(_arange_img(4, 4), 0)
    assert y.reshape(2, 2).tolist() == [[3.5, 5.5], [11.5, 13.5]]
    dx = l.backward(T(4, 8, 12, 16, shape=(1, 2, 2, 1)), 0).reshape(4, 4)
    assert dx.tolist() == [[1, 1, 2, 2], [1, 1, 2, 2], [3, 3, 4, 4], [3, 3, 4, 4]]


def test_avgpool_padding_divisor_nhwc_counts_in_bounds_pixels_legacy_counts_the_window():
    # 3x3 window, pad 1, on a 2x2 image of ones: every window sees the 4 real pixels
    lp = AvgPool2DLayer(3, 3, 1, 1, 1, 1, "ap").init("cpu").set_training(True)
    yp = lp.forward(torch.ones(1, 2, 2, 1), 0)
    assert yp.shape == (1, 2, 2, 1)
    torch.testing.assert_close(yp, torch.ones(1, 2, 2, 1))                # sum 4 / count 4
    dx = lp.backward(torch.ones(1, 2, 2, 1), 0)                          # each pixel is in 4 windows, 1/4 each
    torch.testing.assert_close(dx, torch.ones(1, 2, 2, 1))
    leg = LayerFactory.create(TConfig("lp", "legacy_avgpool2d", pool_h=3, pool_w=3, stride_h=1, stride_w=1, pad_h=1,
                                      pad_w=1)).init("cpu").set_training(True)
    yl = leg.forward(torch.ones(1, 1, 2, 2), 0)                           # NCHW, sum 4 / window 9
    torch.testing.assert_close(yl, torch.full((1, 1, 2, 2), 4.0 / 9.0))
    torch.testing.assert_close(leg.backward(torch.ones(1, 1, 2, 2), 0), torch.full((1, 1, 2, 2), 4.0 / 9.0))
    # the per-position window/in-bounds table that wraps the GPU kernel: corner windows of a 3x3 image
    d = AvgPool2DLayer(3, 3, 1, 1, 1, 1, "ap")._dims((1, 3, 3, 8))
    tab = AvgPool2DLayer(3, 3, 1, 1, 1, 1, "ap")._border_scale(d, "cpu", torch.float32).reshape(3, 3)
    torch.testing.assert_close(tab, T(9 / 4, 9 / 6, 9 / 4, 9 / 6, 1, 9 / 6, 9 / 4, 9 / 6, 9 / 4, shape=(3, 3)))


def test_global_average_pool_window_equals_input():
    l = AvgPool2DLayer(3, 2, name="gap").init("cpu").set_training(True)
    x = torch.arange(12, dtype=torch.float32).reshape(1, 3, 2, 2)        # 2 channels
    y = l.forward(x, 0)
    assert y.shape == (1, 1, 1, 2)
    assert y.reshape(-1).tolist() == [5.0, 6.0]
    assert l.backward(T(6, 12, shape=(1, 1, 1, 2)), 0).reshape(6, 2).tolist() == [[1, 2]] * 6


# ------------------------------------------------------------------------------------- batchnorm
def test_batchnorm_training_normalises_and_updates_running_stats():
    l = BatchNormLayer(2, 0.0, 0.5, True, False, "bn").init("cpu").set_training(True)
    # channel 0: {1,3} mean 2 var 1 ; channel 1: {10,30} mean 20 var 100
    x = T(1, 10, 3, 30, shape=(2, 1, 1, 2))
    y = l.forward(x, 0)
    torch.testing.assert_close(y.reshape(2, 2), T(-1, -1, 1, 1, shape=(2, 2)))
    # running = (1-m)*old + m*batch, with the unbiased variance (n/(n-1) = 2)
    torch.testing.assert_close(l.params["running_mean"], T(1.0, 10.0))
    torch.testing.assert_close(l.params["running_var"], T(0.5 + 0.5 * 2.0, 0.5 + 0.5 * 200.0))


def test_batchnorm_affine_relu_and_eval_mode_use_running_stats():
    l = BatchNormLayer(1, 0.0, 0.1, True, True, "bn").init("cpu")
    l.params["gamma"].fill_(2.0)
    l.params["beta"].fill_(1.0)
    l.set_training(True)
    x = T(0, 2, 4, 6, shape=(4, 1, 1, 1))                                 # mean 3, var 5
    y = l.forward(x, 0).reshape(-1)
    s = 5.0 ** 0.5
    torch.testing.assert_close(y, T(0.0, 1 - 2 / s, 2 / s + 1, 6 / s + 1))  # relu clips 2*(-3/s)+1 = -1.68
    l.set_training(False)
    l.params["running_mean"].fill_(1.0)
    l.params["running_var"].fill_(4.0)
    torch.testing.assert_close(l.forward(T(5.0, shape=(1, 1, 1, 1)), 0).reshape(-1), T(2 * (5 - 1) / 2 + 1.0))


def test_batchnorm_gradients_sum_to_zero_and_param_grads():
    l = BatchNormLayer(1, 1e-5, 0.1, True, False, "bn").init("cpu").set_training(True)
    x = T(1, 2, 4, 9, shape=(4, 1, 1, 1))
    y = l.forward(x, 0)
    dy = T(1, -2, 0.5, 3, shape=(4, 1, 1, 1))
    dx = l.backward(dy, 0)
    assert abs(float(dx.sum())) < 1e-5                                    # dx is orthogonal to constants
    assert abs(float((dx * x).sum())) < 1e-4                              # ... and to x itself (scale invariance)
    torch.testing.assert_close(l.grads["beta"], dy.sum().reshape(1))
    torch.testing.assert_close(l.grads["gamma"], (dy * y).sum().reshape(1), atol=1e-5, rtol=1e-5)


def test_batchnorm_constant_channel_is_finite():
    l = BatchNormLayer(1, 1e-5, 0.1, True, False, "bn").init("cpu").set_training(True)
    y = l.forward(torch.full((3, 2, 2, 1), 7.0), 0)
    assert torch.isfinite(y).all() and float(y.abs().max()) < 1e-3


# ------------------------------------------------------------------------------------- group / layer norm
def test_groupnorm_groups_are_normalised_independently():
    l = GroupNormLayer(2, 4, 0.0, True, "gn").init("cpu").set_training(True)
    # one pixel, 4 channels: groups {1,3} and {10,30}
    y = l.forward(T(1, 3, 10, 30, shape=(1, 1, 1, 4)), 0).reshape(-1)
    torch.testing.assert_close(y, T(-1, 1, -1, 1))
    l.params["gamma"].copy_(T(1, 2, 3, 4))
    l.params["beta"].copy_(T(0, 1, 0, -1))
    y = l.forward(T(1, 3, 10, 30, shape=(1, 1, 1, 4)), 0).reshape(-1)
    torch.testing.assert_close(y, T(-1, 3, -3, 3))


def test_groupnorm_single_group_spans_space_and_channels_per_sample():
    l = GroupNormLayer(1, 2, 0.0, False, "gn").init("cpu").set_training(True)
    x = T(0, 2, 4, 6, 100, 100, 100, 104, shape=(2, 1, 2, 2))             # sample 0: mean 3 var 5 ; sample 1: mean 101 var 3
    y = l.forward(x, 0).reshape(2, 4)
    torch.testing.assert_close(y[0], T(-3, -1, 1, 3) / 5.0 ** 0.5)
    torch.testing.assert_close(y[1], T(-1, -1, -1, 3) / 3.0 ** 0.5)
    assert "gamma" not in l.params


def test_layernorm_rows():
    l = LayerNormLayer(3, 0.0, True, "ln").init("cpu").set_training(True)
    y = l.forward(T(1, 2, 3, -2, 0, 2, shape=(2, 3)), 0)
    r = 1.5 ** 0.5
    torch.testing.assert_close(y, T(-r, 0, r, -r, 0, r, shape=(2, 3)))
    dx = l.backward(torch.ones(2, 3), 0)                                  # constant upstream grad -> zero input grad
    assert float(dx.abs().max()) < 1e-6
    torch.testing.assert_close(l.grads["beta"], T(2.0, 2.0, 2.0))


# ------------------------------------------------------------------------------------- slice / flatten / activation
def test_slice_forward_and_zero_padded_backward():
    l = SliceLayer(1, 1, 2, "s").init("cpu").set_training(True)
    x = torch.arange(8, dtype=torch.float32).reshape(2, 4)
    assert l.forward(x, 0).tolist() == [[1, 2], [5, 6]]
    assert l.backward(T(1, 1, 2, 2, shape=(2, 2)), 0).tolist() == [[0, 1, 1, 0], [0, 2, 2, 0]]
    assert l.output_shape((2, 4)) == (2, 2)
    l0 = SliceLayer(0, 1, 1, "s0").init("cpu").set_training(True)
    assert l0.forward(x, 0).tolist() == [[4, 5, 6, 7]]


def test_flatten_and_activation_values():
    f = FlattenLayer(1, -1, "f").init("cpu").set_training(True)
    assert f.forward(torch.zeros(2, 3, 4, 5), 0).shape == (2, 60) == f.output_shape((2, 3, 4, 5))
    assert f.backward(torch.zeros(2, 60), 0).shape == (2, 3, 4, 5)
    x = T(-2, -0.5, 0, 3)
    relu = ActivationLayer("relu").init("cpu").set_training(True)
    assert relu.forward(x, 0).tolist() == [0, 0, 0, 3]
    assert relu.backward(torch.ones(4), 0).tolist() == [0, 0, 0, 1]
    leaky = ActivationLayer("leaky_relu").init("cpu").set_training(True)
    torch.testing.assert_close(leaky.forward(x, 0), T(-0.02, -0.005, 0, 3))
    sig = ActivationLayer("sigmoid").init("cpu").set_training(True)
    assert float(sig.forward(T(0.0), 0)) == 0.5
    assert float(sig.backward(T(1.0), 0)) == 0.25
    sm = ActivationLayer("softmax").init("cpu").set_training(True)
    torch.testing.assert_close(sm.forward(T(0, 0, 0, 0, shape=(1, 4)), 0), torch.full((1, 4), 0.25))


# ------------------------------------------------------------------------------------- residual block
def test_residual_identity_shortcut_adds_input_and_its_gradient():
    main = Sequential([DenseLayer(2, 2, False, "fc")], "main")
    blk = ResidualBlock(main, None, "linear", "res").init("cpu", seed=0)
    main.layers[0].params["weight"].copy_(T(2, 0, 0, 3, shape=(2, 2)))
    blk.set_training(True)
    y = blk.forward(T(1, 1, shape=(1, 2)), 0)
    assert y.tolist() == [[3, 4]]                                        # F(x) + x
    assert blk.backward(T(1, 1, shape=(1, 2)), 0).tolist() == [[3, 4]]    # W^T dy + dy


def test_residual_relu_gates_both_paths():
    main = Sequential([DenseLayer(2, 2, False, "fc")], "main")
    blk = ResidualBlock(main, None, "relu", "res").init("cpu", seed=0)
    main.layers[0].params["weight"].copy_(T(-3, 0, 0, 1, shape=(2, 2)))
    blk.set_training(True)
    y = blk.forward(T(1, 1, shape=(1, 2)), 0)                            # pre-activation (-2, 2)
    assert y.tolist() == [[0, 2]]
    assert blk.backward(T(5, 7, shape=(1, 2)), 0).tolist() == [[0, 14]]


def test_residual_projection_shortcut_and_config_round_trip():
    main = Sequential([DenseLayer(2, 3, False, "fc")], "main")
    short = Sequential([DenseLayer(2, 3, False, "proj")], "short")
    blk = ResidualBlock(main, short, "linear", "res").init("cpu", seed=0)
    main.layers[0].params["weight"].fill_(1.0)
    short.layers[0].params["weight"].fill_(2.0)
    blk.set_training(True)
    assert blk.forward(T(1, 2, shape=(1, 2)), 0).tolist() == [[9, 9, 9]]
    assert blk.output_shape((4, 2)) == (4, 3)
    cfg = blk.get_config()
    again = LayerFactory.create(TConfig.from_json(cfg.to_json()))
    assert type(again) is ResidualBlock and again.get_config() == cfg


# ------------------------------------------------------------------------------------- config contents
@pytest.mark.parametrize("layer,expect", [
    (Conv2DLayer(3, 8, 3, 5, 2, 1, 1, 2, False, "c"), dict(in_channels=3, out_channels=8, kernel_h=3, kernel_w=5, stride_h=2,
                                                           stride_w=1, pad_h=1, pad_w=2, use_bias=False)),
    (DenseLayer(4, 6, True, "d"), dict(input_features=4, output_features=6, use_bias=True)),
    (MaxPool2DLayer(3, 2, 2, 1, 1, 0, "m"), dict(pool_h=3, pool_w=2, stride_h=2, stride_w=1, pad_h=1, pad_w=0)),
    (BatchNormLayer(5, 1e-3, 0.2, False, True, "b"), dict(num_features=5, epsilon=1e-3, momentum=0.2, affine=False, use_relu=True)),
])
def test_config_carries_every_constructor_argument(layer, expect):
    cfg = layer.get_config()
    assert cfg.name == layer.name and cfg.type == layer.TYPE_NAME
    for k, v in expect.items():
        assert cfg[k] == v, k


# ------------------------------------------------------------------------------------- losses (hand-computed)
def test_loss_values_gradients_and_config_round_trip():
    """loss.hpp / loss_ops.cpp: means over every element, 2/n MSE gradient, clamped probabilities."""
    import math
    from tnn_b200.nn.loss import LossFactory
    pred, tgt = T(1, 2, 3, 5, shape=(2, 2)), T(0, 2, 5, 1, shape=(2, 2))       # diffs 1 0 -2 4
    mse = LossFactory.create_mse()
    assert mse.compute_loss(pred, tgt) == pytest.approx((1 + 0 + 4 + 16) / 4)
    assert mse.compute_gradient(pred, tgt).tolist() == [[0.5, 0.0], [-1.0, 2.0]]
    mae = LossFactory.create_mae()
    assert mae.compute_loss(pred, tgt) == pytest.approx(7 / 4)
    assert mae.compute_gradient(pred, tgt).tolist() == [[0.25, 0.0], [-0.25, 0.25]]
    hub = LossFactory.create_huber(1.5)
    # |d| <= delta: d^2/2 ; else delta*(|d| - delta/2)
    assert hub.compute_loss(pred, tgt) == pytest.approx((0.5 + 0 + 1.5 * (2 - 0.75) + 1.5 * (4 - 0.75)) / 4)
    assert hub.compute_gradient(pred, tgt).tolist() == [[0.25, 0.0], [-0.375, 0.375]]
    # cross entropy on probabilities: -log p[label] averaged over rows, gradient only at the label
    ce = LossFactory.create_crossentropy(use_logits=False, eps=1e-6)
    probs = T(0.5, 0.5, 0.0, 1.0, shape=(2, 2))
    labels = torch.tensor([0, 0], dtype=torch.int32)
    assert ce.compute_loss(probs, labels) == pytest.approx((math.log(2) - math.log(1e-6)) / 2, rel=1e-4)
    g = ce.compute_gradient(probs, labels)
    assert g[0].tolist() == [-1.0, 0.0] and g[1, 1] == 0 and float(g[1, 0]) == pytest.approx(-0.5e6, rel=1e-3)
    # logits: uniform logits -> log(C); gradient (softmax - onehot)/rows
    cel = LossFactory.create("ce")
    assert cel.compute_loss(torch.zeros(3, 4), torch.tensor([0, 1, 2])) == pytest.approx(math.log(4), rel=1e-6)
    gl = cel.compute_gradient(torch.zeros(1, 4), torch.tensor([2]))
    torch.testing.assert_close(gl, T(0.25, 0.25, -0.75, 0.25, shape=(1, 4)))
    for loss in (mse, mae, hub, ce, cel, LossFactory.create_logsoftmax_crossentropy()):
        cfg = loss.get_config()
        again = LossFactory.create_from_config(TConfig.from_json(cfg.to_json()))
        assert type(again) is type(loss) and again.get_config() == cfg and type(loss.clone()) is type(loss)
    assert hub.get_config()["delta"] == 1.5 and ce.get_config()["epsilon"] == 1e-6 and ce.get_config()["use_logits"] is False
    with pytest.raises(ValueError):
        LossFactory.create("hinge")
