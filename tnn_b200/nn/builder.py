"""LayerBuilder: fluent model DSL with shape inference (reference: include/nn/layer_builder.hpp).
Shapes are batch-less ({H, W, C} for images, {L} for token ids, {L, E} for sequences); a batch
dimension of 1 is prepended for inference like the reference does."""
from __future__ import annotations

import torch

from ..core.dtype import dtype_eps
from . import blocks as B
from . import layers as L
from .layer import Layer


class LayerBuilder:
    def __init__(self, input_shape=None):
        if input_shape is None:
            input_shape = [[]]
        if input_shape and not isinstance(input_shape[0], (list, tuple)):
            input_shape = [list(input_shape)]
        self.input_shape = [list(s) for s in input_shape]
        self.shape = (1, *self.input_shape[0])
        self.layers: list[Layer] = []
        self.io_dtype = None

    # ------------------------------------------------------------------ helpers
    def dtype(self, dt):
        self.io_dtype = dt
        return self

    def _eps(self):
        return dtype_eps(self.io_dtype or torch.float32)

    def _push(self, layer: Layer):
        if self.io_dtype is not None:
            layer.set_io_dtype(self.io_dtype)
        self.shape = tuple(layer.output_shape(self.shape))
        self.layers.append(layer)
        return self

    def _name(self, name, default):
        return name if name else f"{default}_{len(self.layers)}"

    def get_current_shape(self):
        return [list(self.shape)]

    def get_batchless_current_shape(self):
        return [list(self.shape[1:])]

    def add_layer(self, layer: Layer):
        return self._push(layer)

    def build(self) -> list[Layer]:
        out, self.layers = self.layers, []
        return out

    # ------------------------------------------------------------------ primitive layers
    def dense(self, output_features, use_bias=True, name=""):
        return self._push(L.DenseLayer(self.shape[-1], output_features, use_bias, self._name(name, "dense")))

    def conv2d(self, out_channels, kernel_h, kernel_w, stride_h=1, stride_w=1, pad_h=0, pad_w=0, use_bias=True, name=""):
        return self._push(L.Conv2DLayer(self.shape[-1], out_channels, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                        pad_w, use_bias, self._name(name, "conv2d")))

    def batchnorm(self, epsilon=None, momentum=0.1, affine=True, use_relu=False, name=""):
        eps = self._eps() if epsilon is None else epsilon
        return self._push(L.BatchNormLayer(self.shape[-1], eps, momentum, affine, bool(use_relu), self._name(name, "batchnorm")))

    def groupnorm(self, num_groups, epsilon=1e-5, affine=True, name=""):
        return self._push(L.GroupNormLayer(num_groups, self.shape[-1], epsilon, affine, self._name(name, "groupnorm")))

    def layernorm(self, epsilon=1e-5, affine=True, name=""):
        return self._push(L.LayerNormLayer(self.shape[-1], epsilon, affine, self._name(name, "layer_norm")))

    def maxpool2d(self, pool_h, pool_w, stride_h=None, stride_w=None, pad_h=0, pad_w=0, name=""):
        return self._push(L.MaxPool2DLayer(pool_h, pool_w, stride_h, stride_w, pad_h, pad_w, self._name(name, "maxpool2d")))

    def avgpool2d(self, pool_h, pool_w, stride_h=None, stride_w=None, pad_h=0, pad_w=0, name=""):
        return self._push(L.AvgPool2DLayer(pool_h, pool_w, stride_h, stride_w, pad_h, pad_w, self._name(name, "avgpool2d")))

    def activation(self, activation, name=""):
        return self._push(L.ActivationLayer(activation, self._name(name, activation)))

    def dropout(self, dropout_rate, name=""):
        return self._push(L.DropoutLayer(dropout_rate, self._name(name, "dropout")))

    def flatten(self, start_dim=1, end_dim=-1, name=""):
        return self._push(L.FlattenLayer(start_dim, end_dim, self._name(name, "flatten")))

    def class_token(self, embed_dim=None, name=""):
        return self._push(L.ClassTokenLayer(embed_dim or self.shape[-1], self._name(name, "class_token")))

    def positional_embedding(self, embed_dim=None, seq_len=None, name=""):
        return self._push(L.PositionalEmbeddingLayer(embed_dim or self.shape[-1], seq_len or self.shape[-2],
                                                     self._name(name, "pos_embedding")))

    def embedding(self, vocab_size, embed_dim, padding_idx=-1, name=""):
        return self._push(L.EmbeddingLayer(vocab_size, embed_dim, padding_idx, self._name(name, "embedding")))

    def slice(self, axis, start, length, name=""):
        return self._push(L.SliceLayer(axis, start, length, self._name(name, "slice")))

    def transpose(self, name=""):
        return self._push(L.TransposeLayer(self._name(name, "transpose")))

    def identity(self, name=""):
        return self._push(L.IdentityLayer(self._name(name, "identity")))

    def attention(self, embed_dim, num_heads, is_causal=True, name=""):
        return self._push(B.AttentionBlock(embed_dim, num_heads, is_causal, self._name(name, "attention")))

    def flash_attention(self, embed_dim, num_heads, is_causal=True, name=""):
        return self._push(B.FlashAttentionBlock(embed_dim, num_heads, is_causal, self._name(name, "flash_attention")))

    def broadcast_m_seq(self, sequences, join_layer, name=""):
        return self._push(B.MSequential(sequences, join_layer, self._name(name, "msequential")))

    # legacy NCHW twins
    def legacy_dense(self, output_features, use_bias=True, name=""):
        return self._push(L.LegacyDenseLayer(self.shape[-1], output_features, use_bias, self._name(name, "legacy_dense")))

    def legacy_conv2d(self, out_channels, kernel_h, kernel_w, stride_h=1, stride_w=1, pad_h=0, pad_w=0, use_bias=True, name=""):
        return self._push(L.LegacyConv2DLayer(self.shape[1], out_channels, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                              pad_w, use_bias, name=self._name(name, "legacy_conv2d")))

    def legacy_batchnorm(self, epsilon=1e-5, momentum=0.1, affine=True, use_relu=False, name=""):
        return self._push(L.LegacyBatchNormLayer(self.shape[1], epsilon, momentum, affine, use_relu,
                                                 name=self._name(name, "legacy_batchnorm")))

    def legacy_maxpool2d(self, pool_h, pool_w, stride_h=None, stride_w=None, pad_h=0, pad_w=0, name=""):
        return self._push(L.LegacyMaxPool2DLayer(pool_h, pool_w, stride_h, stride_w, pad_h, pad_w,
                                                 name=self._name(name, "legacy_maxpool2d")))

    def legacy_avgpool2d(self, pool_h, pool_w, stride_h=None, stride_w=None, pad_h=0, pad_w=0, name=""):
        return self._push(L.LegacyAvgPool2DLayer(pool_h, pool_w, stride_h, stride_w, pad_h, pad_w,
                                                 name=self._name(name, "legacy_avgpool2d")))

    # ------------------------------------------------------------------ composite blocks
    def _sub(self):
        b = LayerBuilder(self.get_batchless_current_shape())
        b.io_dtype = self.io_dtype
        return b

    def residual_block(self, main_layers, shortcut_layers=None, activation="relu", name=""):
        nm = self._name(name, "residual_block")
        main = B.Sequential(main_layers, nm + "_main_path")
        sc = B.Sequential(shortcut_layers, nm + "_shortcut_path") if shortcut_layers else None
        return self._push(B.ResidualBlock(main, sc, activation, nm))

    def basic_residual_block(self, in_channels, out_channels, stride=1, name="basic_residual_block"):
        main = (self._sub().conv2d(out_channels, 3, 3, stride, stride, 1, 1, False, name + "_conv1")
                .batchnorm(None, 0.1, True, True, name + "_bn0")
                .conv2d(out_channels, 3, 3, 1, 1, 1, 1, False, name + "_conv2")
                .batchnorm(None, 0.1, True, False, name + "_bn1").build())
        sc = None
        if stride != 1 or in_channels != out_channels:
            sc = (self._sub().conv2d(out_channels, 1, 1, stride, stride, 0, 0, False, name + "_conv0")
                  .batchnorm(None, 0.1, True, False, name + "_bn0s").build())
        return self.residual_block(main, sc, "relu", name)

    def wide_residual_block(self, in_channels, out_channels, stride=1, dropout_rate=0.0, name="wide_residual_block"):
        """Pre-activation WRN block: BN-ReLU-conv3x3(stride)-BN-ReLU-dropout-conv3x3, 1x1 shortcut
        when the shape changes, linear after the add (layer_builder.hpp:367-403)."""
        mb = self._sub().batchnorm(None, 0.1, True, True, name + "_bn1")
        mb.conv2d(out_channels, 3, 3, stride, stride, 1, 1, True, name + "_conv1")
        mb.batchnorm(None, 0.1, True, True, name + "_bn2")
        if dropout_rate > 0:
            mb.dropout(dropout_rate, name + "_dropout")
        mb.conv2d(out_channels, 3, 3, 1, 1, 1, 1, True, name + "_conv2")
        sc = None
        if stride != 1 or in_channels != out_channels:
            sc = self._sub().conv2d(out_channels, 1, 1, stride, stride, 0, 0, False, name + "_shortcut").build()
        return self.residual_block(mb.build(), sc, "linear", name)

    def bottleneck_residual_block(self, in_channels, mid_channels, out_channels, stride=1, name="bottleneck_residual_block"):
        main = (self._sub().conv2d(mid_channels, 1, 1, 1, 1, 0, 0, False, name + "_conv0").batchnorm(None, 0.1, True, True, name + "_bn0")
                .conv2d(mid_channels, 3, 3, stride, stride, 1, 1, False, name + "_conv1").batchnorm(None, 0.1, True, True, name + "_bn1")
                .conv2d(out_channels, 1, 1, 1, 1, 0, 0, False, name + "_conv2").batchnorm(None, 0.1, True, True, name + "_bn2").build())
        sc = None
        if stride != 1 or in_channels != out_channels:
            sc = (self._sub().conv2d(out_channels, 1, 1, stride, stride, 0, 0, False, name + "_conv3")
                  .batchnorm(None, 0.1, True, False, name + "_bn3").build())
        return self.residual_block(main, sc, "none", name)

    def gpt_block(self, embed_dim, num_heads, ffn_dim, dropout_rate=0.1, is_causal=False, activation_fn="gelu", name="",
                  flash=False):
        nm = name or f"gpt_block_{len(self.layers)}"
        ab = self._sub().layernorm(1e-5, True, nm + "_ln_1")
        (ab.flash_attention if flash else ab.attention)(embed_dim, num_heads, is_causal, nm + "_attn")
        if dropout_rate > 0:
            ab.dropout(dropout_rate, nm + "_attn_dropout")
        self.residual_block(ab.build(), None, "linear", nm + "_attn_res")
        fb = (self._sub().layernorm(1e-5, True, nm + "_ln_2").dense(ffn_dim, True, nm + "_mlp_fc1")
              .activation(activation_fn, nm + "_mlp_act").dense(embed_dim, True, nm + "_mlp_fc2"))
        if dropout_rate > 0:
            fb.dropout(dropout_rate, nm + "_mlp_dropout")
        return self.residual_block(fb.build(), None, "linear", nm + "_ffn_res")

    def flash_gpt_block(self, embed_dim, num_heads, ffn_dim, dropout_rate=0.1, is_causal=False, activation_fn="gelu", name=""):
        return self.gpt_block(embed_dim, num_heads, ffn_dim, dropout_rate, is_causal, activation_fn, name, flash=True)
