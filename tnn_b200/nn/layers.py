"""Concrete layers.  GPU path = hand-written sm_100a kernels (ops/ + csrc/kernels), CPU path =
fp32 PyTorch reference math (the oracle).  Layout is NHWC / feature-last (reference current
generation); ``legacy_*`` types are NCHW facades over the same kernels.

Reference files: src/nn/layers_impl/{dense,conv2d,batchnorm,maxpool2d,avgpool2d,dropout,flatten,
groupnorm,layer_norm,embedding,positional_embedding,class_token,slice,transpose,identity,n_ary,
sdpa}_layer.cpp and include/nn/activations.hpp.
"""
from __future__ import annotations

import math
import os
import zlib

import torch
import torch.nn.functional as F

from ..core import side_stream, stepstate
from ..core.config import TConfig
from ..ops import conv as conv_ops
from ..ops import gemm as gemm_ops
from ..ops import native
from ..ops import tensor_ops
from .layer import Layer, ParamDescriptor, register

ACT_KIND = {"none": 0, "linear": 0, "relu": 1, "leaky_relu": 2, "sigmoid": 3, "tanh": 4, "elu": 5, "gelu": 6}


def _uniform_init(bound: float):
    def f(t: torch.Tensor, gen: torch.Generator):
        t.uniform_(-bound, bound, generator=gen)
    return f


def _const_init(v: float):
    def f(t: torch.Tensor, gen: torch.Generator):
        t.fill_(v)
    return f


def _normal_init(std: float):
    def f(t: torch.Tensor, gen: torch.Generator):
        t.normal_(0.0, std, generator=gen)
    return f


def _round8(n: int) -> int:
    return (n + 7) // 8 * 8


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 / fp16 / fp64 -> bf16 on the GPU through the hand-written cast kernels (no-op when already bf16)."""
    return to_dtype(x, torch.bfloat16)


def to_dtype(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Activation cast between bf16 / fp32 / fp16 / fp64 (csrc/kernels: cast_f32_to_bf16, cast_any)."""
    if x.dtype == dtype:
        return x
    if x.is_cuda and x.dtype in (torch.float32, torch.float16, torch.float64, torch.bfloat16):
        x = x.contiguous()
        y = torch.empty(x.shape, dtype=dtype, device=x.device)
        if x.dtype == torch.float32 and dtype == torch.bfloat16:
            native.require().cast_f32_to_bf16(x, y)
        else:
            native.require().cast_any(x, y)
        return y
    return x.to(dtype)


def _aligned_rows(t2d: torch.Tensor) -> torch.Tensor:
    """Return a [M, N] bf16 matrix whose row stride is a multiple of 8 elements (TMA needs
    16-byte-aligned rows); pads through a copy only when the incoming view is not."""
    if t2d.stride(1) == 1 and t2d.stride(0) % 8 == 0 and t2d.data_ptr() % 16 == 0:
        return t2d
    M, N = t2d.shape
    buf = torch.empty((M, _round8(N)), dtype=t2d.dtype, device=t2d.device)
    out = buf[:, :N]
    out.copy_(t2d)
    return out


# =============================================================================== Dense
@register()
class DenseLayer(Layer):
    """y = x W^T + b, W [out, in], init U(+-1/sqrt(in)) (dense_layer.cpp:46-61)."""
    TYPE_NAME = "dense"

    def __init__(self, input_features: int, output_features: int, use_bias: bool = True, name: str = "dense",
                 activation: str | None = None):
        super().__init__(name)
        self.input_features, self.output_features, self.use_bias = int(input_features), int(output_features), bool(use_bias)
        self.activation = activation  # optional epilogue fusion (relu / gelu)

    def param_descriptors(self):
        b = 1.0 / math.sqrt(self.input_features)
        d = [ParamDescriptor("weight", (self.output_features, self.input_features), _uniform_init(b))]
        if self.use_bias:
            d.append(ParamDescriptor("bias", (self.output_features,), _uniform_init(b)))
        return d

    def forward_impl(self, x, mb_id, residual=None):
        fused = self.activation not in (None, "none", "linear")
        if fused and self.training:
            # training needs the PRE-activation for the backward pass: GEMM (+bias, +residual) without the activation
            # epilogue, then the activation kernel; inference keeps the one-kernel fused epilogue
            act, self.activation = self.activation, None
            try:
                z = self.forward_impl(x, mb_id, residual)
            finally:
                self.activation = act
            self.cache(mb_id)["z"] = z
            if z.is_cuda:
                zc = z.contiguous()
                y = torch.empty_like(zc)
                native.require().act_fwd(zc, y, ACT_KIND[act], 1.0)
                return y
            return _act_ref(z.float(), act, 1.0).to(z.dtype)
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        N = self.output_features
        bias = self.params.get("bias")
        if x2.is_cuda and self.io_dtype in gemm_ops.XT_DTYPES:
            # explicit fp16 / fp32 / fp64 IO contract (dense_layer.cpp:336-340 instantiates the layer per dtype): fp16 and
            # fp32 (tf32) run on the tensor cores through csrc/kernels/gemm_xt.cu, fp64 on the DFMA kernel
            dt = self.io_dtype
            x2 = to_dtype(x2, dt)
            res2 = to_dtype(residual.reshape(-1, N), dt) if residual is not None else None
            y = gemm_ops.linear_fwd(x2, self.lp_as("weight", dt), bias, self.activation, res2)
        elif x2.is_cuda:
            x2 = _aligned_rows(to_bf16(x2))
            M = x2.shape[0]
            res2 = residual.reshape(-1, N) if residual is not None else None
            # the epilogue reads the residual with the OUTPUT's row stride: rows are padded to 8 elements unless a
            # residual with a different stride has to be fused (N % 8 != 0: element-wise epilogue of the same kernel)
            ld = _round8(N) if (res2 is None or res2.stride(0) == _round8(N)) else res2.stride(0)
            buf = torch.empty((M, ld), dtype=torch.bfloat16, device=x.device)
            y = buf[:, :N]
            gemm_ops.linear_fwd(x2, self.lp("weight"), bias, self.activation, res2, out=y)
        else:
            y = gemm_ops.linear_fwd(x2.float(), self.params["weight"], bias, self.activation, residual.reshape(-1, N) if residual is not None else None)
        if self.training:
            c = self.cache(mb_id)
            c["x"] = x2
        return y.reshape(*shp[:-1], N)

    def backward_impl(self, dy, mb_id):
        c = self.cache(mb_id)
        x2 = c["x"]
        N = self.output_features
        dy2 = dy.reshape(-1, N)
        if self.activation not in (None, "none", "linear"):
            z = c["z"]
            if dy.is_cuda:
                zc = z.contiguous()
                g = to_dtype(dy.reshape(zc.shape), zc.dtype).contiguous()
                dz = torch.empty_like(zc)
                native.require().act_bwd(zc, g, dz, ACT_KIND[self.activation], 1.0)
                dy2 = dz.reshape(-1, N)
            else:
                zf = z.float().detach().requires_grad_(True)
                with torch.enable_grad():
                    _act_ref(zf, self.activation, 1.0).backward(dy.reshape(zf.shape).float())
                dy2 = zf.grad.reshape(-1, N)
        if dy2.is_cuda and self.io_dtype in gemm_ops.XT_DTYPES:
            dt = self.io_dtype
            dy2 = to_dtype(dy2, dt).contiguous()
            gemm_ops.linear_wgrad(dy2, x2, self.grads["weight"])
            if self.use_bias:
                gemm_ops.bias_grad(dy2, self.grads["bias"])
            dx = gemm_ops.linear_dgrad(dy2, self.lp_as("weight", dt))
        elif dy2.is_cuda:
            dy2 = _aligned_rows(to_bf16(dy2))
            gemm_ops.linear_wgrad(dy2, x2, self.grads["weight"])
            if self.use_bias:
                gemm_ops.bias_grad(dy2, self.grads["bias"])
            dx = gemm_ops.linear_dgrad(dy2, self.lp("weight"))
        else:
            dy2 = dy2.float()
            gemm_ops.linear_wgrad(dy2, x2, self.grads["weight"])
            if self.use_bias:
                gemm_ops.bias_grad(dy2, self.grads["bias"])
            dx = gemm_ops.linear_dgrad(dy2, self.params["weight"])
        return dx.reshape(*dy.shape[:-1], self.input_features)

    def output_shape(self, s):
        return (*s[:-1], self.output_features)

    def forward_flops(self, s):
        return 2 * int(math.prod(s[:-1])) * self.input_features * self.output_features

    def config_params(self):
        return dict(input_features=self.input_features, output_features=self.output_features, use_bias=self.use_bias)

    @classmethod
    def create_from_config(cls, c: TConfig):
        return cls(c["input_features"], c["output_features"], c.get("use_bias", True), c.name)


# =============================================================================== Conv2D (NHWC)
@register()
class Conv2DLayer(Layer):
    """NHWC convolution, weights [K,R,S,C], init U(+-1/sqrt(C*R*S)) (conv2d_layer.cpp:70-86)."""
    TYPE_NAME = "conv2d"

    def __init__(self, in_channels, out_channels, kernel_h, kernel_w, stride_h=1, stride_w=1, pad_h=0, pad_w=0,
                 use_bias=True, name="conv2d"):
        super().__init__(name)
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_h, self.kernel_w = int(kernel_h), int(kernel_w)
        self.stride_h, self.stride_w, self.pad_h, self.pad_w = int(stride_h), int(stride_w), int(pad_h), int(pad_w)
        self.use_bias = bool(use_bias)
        # anisotropic stride / padding: im2col path only (no TMA-window plan, no BatchNorm-backward fusion into the dgrad)
        self._generic = self.stride_h != self.stride_w or self.pad_h != self.pad_w
        self._plans = {}
        self._state = conv_ops.ConvState()
        # BatchNorm statistics fusion: a consumer BatchNorm that finds this layer as the producer of
        # its input asks for the statistics to be accumulated in the conv epilogue from then on
        self._stats_ws = None
        self._stats_requested = False
        self._bn_prev = None         # set by Sequential: BatchNorm(+ReLU+dropout) whose output is this layer's only input

    def param_descriptors(self):
        fan_in = self.in_channels * self.kernel_h * self.kernel_w
        b = 1.0 / math.sqrt(fan_in)
        d = [ParamDescriptor("weight", (self.out_channels, self.kernel_h, self.kernel_w, self.in_channels), _uniform_init(b))]
        if self.use_bias:
            d.append(ParamDescriptor("bias", (self.out_channels,), _uniform_init(b)))
        return d

    def _plan(self, x):
        key = (tuple(x.shape), x.device.type)
        p = self._plans.get(key)
        if p is None:
            N, H, W, C = x.shape
            # anisotropic stride / padding run through the im2col + tcgen05 GEMM path of the same kernels (no library call)
            p = conv_ops.make_plan(N, H, W, C, self.out_channels, self.kernel_h, self.kernel_w, self.stride_h,
                                   self.pad_h, x.device, stride_w=self.stride_w, pad_w=self.pad_w)
            self._plans[key] = p
        return p

    # ---- explicit fp16 / fp32 io dtype: im2col + the fp16 / tf32 tcgen05 GEMMs of gemm_xt.cu (the reference instantiates
    # its conv layer per dtype over cuDNN, conv2d_layer.cpp; bf16 -- the default -- runs the TMA-window implicit GEMM)
    def _xt_geo(self, x):
        N, H, W, C = x.shape
        OH = conv_ops.out_size(H, self.kernel_h, self.stride_h, self.pad_h)
        OW = conv_ops.out_size(W, self.kernel_w, self.stride_w, self.pad_w)
        geo = [N, H, W, C, self.kernel_h, self.kernel_w, self.stride_h, self.stride_w, self.pad_h, self.pad_w, OH, OW]
        rsc = self.kernel_h * self.kernel_w * C
        per16 = 16 // x.element_size()
        return geo, N * OH * OW, rsc, (rsc + per16 - 1) // per16 * per16, OH, OW

    def _xt_weight(self, dt, rsc, ld):
        """[K, ld] weight matrix in the io dtype, rows padded with zeros to the 16-byte pitch the TMA descriptors need."""
        w = self.lp_as("weight", dt).reshape(self.out_channels, rsc)
        if ld == rsc:
            return w
        key = f"weight@{dt}@pad"
        wp = self.params_lp.get(key)
        if wp is None or self._lp_version.get(key, -1) != self.param_version:
            wp = self.params_lp[key] = torch.zeros((self.out_channels, ld), dtype=dt, device=w.device)
            wp[:, :rsc].copy_(w)
            self._lp_version[key] = self.param_version
        return wp

    def _forward_xt(self, x, mb_id, residual, act):
        Cx, dt = native.require(), self.io_dtype
        x = to_dtype(x, dt).contiguous()
        geo, M, rsc, ld, OH, OW = self._xt_geo(x)
        col = torch.zeros((M, ld), dtype=dt, device=x.device) if ld != rsc else torch.empty((M, ld), dtype=dt, device=x.device)
        Cx.im2col_nhwc(x, col, ld, geo)
        res2 = to_dtype(residual, dt).reshape(M, self.out_channels).contiguous() if residual is not None else None
        y = gemm_ops.gemm_xt(col, self._xt_weight(dt, rsc, ld), bias=self.params.get("bias"), residual=res2, act=act)
        if self.training:
            self.cache(mb_id).update(x=x, col=col, xt=(geo, M, rsc, ld))
        return y.reshape(x.shape[0], OH, OW, self.out_channels)

    def _backward_xt(self, dy, mb_id, need_dx):
        Cx = native.require()
        c = self.cache(mb_id)
        x, col = c["x"], c["col"]
        geo, M, rsc, ld = c["xt"]
        dt = x.dtype
        dy2 = to_dtype(dy, dt).reshape(M, self.out_channels).contiguous()
        gw = self.grads["weight"].reshape(self.out_channels, rsc)
        if ld == rsc:
            gemm_ops.gemm_xt(dy2, col, True, True, out=gw, accumulate=True)              # dW[K, RSC] += dy^T col, fp32 accumulation
        else:
            t = gemm_ops.gemm_xt(dy2, col, True, True, out_f32=True)
            tensor_ops.add(gw, t[:, :rsc].contiguous(), out=gw)
        if self.grads.get("bias") is not None:
            Cx.colsum_any(dy2, self.grads["bias"], 1.0, True)
        if not need_dx:
            return None
        dcol = gemm_ops.gemm_xt(dy2, self._xt_weight(dt, rsc, ld), False, True)         # [M, ld] = dy W
        dx = torch.empty_like(x)
        Cx.col2im_nhwc(dcol, ld, dx, geo)
        return dx

    def forward_impl(self, x, mb_id, residual=None, act=None):
        if x.is_cuda and self.io_dtype in (torch.float16, torch.float32):
            return self._forward_xt(x, mb_id, residual, act)
        plan = self._plan(x)
        c = self.cache(mb_id) if self.training else {}
        if x.is_cuda:
            x = to_bf16(x).contiguous()
            ws = None
            if self.training and self._stats_requested and plan.path == "tma":
                if self._stats_ws is None or self._stats_ws.device != x.device:
                    self._stats_ws = torch.zeros(native.require().bn_ws_floats(self.out_channels), dtype=torch.float32,
                                                 device=x.device)
                ws = self._stats_ws
            y = conv_ops.conv2d_fwd(plan, x, self.params["weight"], self.lp("weight"), self.params.get("bias"),
                                    residual=residual, act=act, cache=c, stats_ws=ws)
            if self.training:
                y._tnn_producer = self
                y._tnn_bn_stats = ws          # None until a BatchNorm consumer asked for them
        else:
            y = conv_ops.conv2d_fwd(plan, x.float(), self.params["weight"], None, self.params.get("bias"),
                                    residual=residual, act=act)
        if self.training:
            c["x"] = x
            c["plan"] = plan
        return y

    def backward_impl(self, dy, mb_id, residual=None, need_dx=True):
        c = self.cache(mb_id)
        if "xt" in c:
            dx = self._backward_xt(dy, mb_id, need_dx)
            if dx is not None and residual is not None:
                dx = tensor_ops.add(dx, to_dtype(residual, dx.dtype).contiguous())
            return dx
        x, plan = c["x"], c["plan"]
        w = self.params["weight"]
        colsum = getattr(dy, "_tnn_colsum", None)
        if dy.is_cuda:
            dy = to_bf16(dy).contiguous()
        else:
            dy = dy.float()
        db = self.grads.get("bias")
        if db is not None and colsum is not None and dy.is_cuda and plan.path != "direct":
            native.require().axpy_f32(db, colsum, 1.0)       # bias grad came fused out of the BN backward
            db = None
        if dy.is_cuda and plan.path == "tma" and side_stream.enabled():
            # weight gradient on the side stream: overlaps the dgrad / BatchNorm backward that follow
            gw = self.grads["weight"]
            side_stream.run(dy.device, lambda: conv_ops.conv2d_wgrad(plan, dy, x, gw, db, cache=c), keep=(dy, x, c, gw, db))
        else:
            conv_ops.conv2d_wgrad(plan, dy, x, self.grads["weight"], db, cache=c)
        if not need_dx or plan.path == "direct":
            return None
        bn = self._bn_prev
        if bn is not None and dy.is_cuda and residual is None and plan.path == "tma" and plan.stride == 1 and \
                not plan.d_needs_zero and len(plan.d_launches) == 1 and bn.training and \
                os.environ.get("TNN_BN_FUSE_BWD", "0") == "1":
            bc = bn._cache.get(mb_id)
            if bc is not None and bc.get("pack") is not None and bc["x"].shape == x.shape and self.in_channels % 32 == 0:
                # dgrad + BatchNorm-backward reduction in one kernel: the epilogue masks / scales the
                # gradient, stores g and accumulates (sum g, sum g*xhat) into the BatchNorm's workspace.
                # OFF by default (TNN_BN_FUSE_BWD=1 enables): measured in the same run on WRN-16-8 it is
                # 4.5 % SLOWER per step (5.63 vs 5.38 ms) -- the 4 epilogue warps become the GEMM's
                # bottleneck (dgrad 107 -> 165 us) and that costs more than the reduction kernel it removes
                bnb = dict(x=bc["x"], pack=bc["pack"], relu=bn.use_relu, p=bc["p"], seed=bc.get("seed", 0),
                           state=stepstate.get(dy.device).dev if bc["p"] > 0 else None, ws=bn._bufs(dy.device)["ws"])
                g = conv_ops.conv2d_dgrad(plan, dy, w, self._state, self.param_version, bn_bwd=bnb)
                g._tnn_bn_g = bn
                return g
        return conv_ops.conv2d_dgrad(plan, dy, w, self._state, self.param_version, residual=residual)

    def output_shape(self, s):
        N, H, W, C = s
        return (N, conv_ops.out_size(H, self.kernel_h, self.stride_h, self.pad_h),
                conv_ops.out_size(W, self.kernel_w, self.stride_w, self.pad_w), self.out_channels)

    def forward_flops(self, s):
        o = self.output_shape(s)
        return 2 * o[0] * o[1] * o[2] * o[3] * self.kernel_h * self.kernel_w * self.in_channels

    def config_params(self):
        return dict(in_channels=self.in_channels, out_channels=self.out_channels, kernel_h=self.kernel_h,
                    kernel_w=self.kernel_w, stride_h=self.stride_h, stride_w=self.stride_w, pad_h=self.pad_h,
                    pad_w=self.pad_w, use_bias=self.use_bias)

    @classmethod
    def create_from_config(cls, c: TConfig):
        return cls(c["in_channels"], c["out_channels"], c["kernel_h"], c["kernel_w"], c.get("stride_h", 1),
                   c.get("stride_w", 1), c.get("pad_h", 0), c.get("pad_w", 0), c.get("use_bias", True), c.name)


# =============================================================================== BatchNorm
@register()
class BatchNormLayer(Layer):
    """BatchNorm over the last (channel) dim with optional fused ReLU (batchnorm_layer.cpp:59-375).
    Biased batch variance for normalisation, unbiased for running_var; momentum 0.1."""
    TYPE_NAME = "batchnorm"

    def __init__(self, num_features, epsilon=1e-5, momentum=0.1, affine=True, use_relu=False, name="batchnorm"):
        super().__init__(name)
        self.num_features, self.epsilon, self.momentum = int(num_features), float(epsilon), float(momentum)
        self.affine, self.use_relu = bool(affine), bool(use_relu)
        self._gpu = None
        self.fused_dropout_p = 0.0   # set by Sequential when a Dropout layer directly follows
        self._bwd_consumer = None    # set by Sequential: the Conv2DLayer whose dgrad epilogue does this layer's backward reduction

    def param_descriptors(self):
        C = self.num_features
        d = []
        if self.affine:
            d += [ParamDescriptor("gamma", (C,), _const_init(1.0)), ParamDescriptor("beta", (C,), _const_init(0.0))]
        d += [ParamDescriptor("running_mean", (C,), _const_init(0.0), trainable=False),
              ParamDescriptor("running_var", (C,), _const_init(1.0), trainable=False)]
        return d

    def _bufs(self, dev):
        if self._gpu is None or self._gpu["ws"].device != dev:
            C = self.num_features
            z = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)  # noqa: E731
            nws = native.require().bn_ws_floats(C)
            self._gpu = dict(ws=z(nws), cws=z(nws), mean=z(C), invstd=z(C), ss=z(2 * C), coef=z(3 * C))
        return self._gpu

    def forward_impl(self, x, mb_id):
        C = self.num_features
        shp = x.shape
        assert shp[-1] == C, f"{self.name}: expected {C} channels, got {shp}"
        gamma, beta = self.params.get("gamma"), self.params.get("beta")
        p_drop = self.fused_dropout_p if self.training else 0.0
        if x.is_cuda:
            Cx = native.require()
            fused_ws = getattr(x, "_tnn_bn_stats", None)
            producer = getattr(x, "_tnn_producer", None)
            x = to_bf16(x).contiguous()
            M = x.numel() // C
            b = self._bufs(x.device)
            y = torch.empty_like(x)
            seed = 0
            if self.training:
                seed = (self.seed * 1000003 + (zlib.crc32(self.name.encode()) & 0xFFFFFF) * 7919 + mb_id) & 0x7FFFFFFFFFFFFFFF
            rec = None
            if self.training:
                # this microbatch's statistics record, written by the statistics kernel itself (no clones):
                # [0:4C] the (invstd, -mean*invstd, scale, shift) pack per channel (16-byte aligned), [4C:5C] mean, [5C:6C] invstd
                rec = torch.empty(6 * C, dtype=torch.float32, device=x.device)
            if self.training and fused_ws is not None and os.environ.get("TNN_BN_FUSE_STATS", "0") == "1":
                # the producing convolution already accumulated sum / sum-of-squares in its epilogue
                # (TNN_BN_FUSE_STATS=1; off by default since the persistent GEMMs became epilogue-bound: the
                # same conv costs 93-127 us with the statistics epilogue and 67-72 us without, which is more
                # than the stand-alone statistics kernel -- same-run A/B 4.269 vs 4.284 ms/step)
                Cx.bn_finalize_stats(fused_ws, M, C, gamma, beta, self.epsilon, self.momentum, self.params["running_mean"],
                                     self.params["running_var"], rec[4 * C:5 * C], rec[5 * C:], b["ss"], rec[:4 * C])
                x._tnn_bn_stats = None                    # the workspace is reset by the fold: single consumer
            elif self.training:
                if producer is not None and getattr(producer, "out_channels", -1) == C:
                    # takes effect from the next forward on
                    producer._stats_requested = os.environ.get("TNN_BN_FUSE_STATS", "0") == "1"
                Cx.bn_train_stats(x, M, C, gamma, beta, self.epsilon, self.momentum, self.params["running_mean"],
                                  self.params["running_var"], rec[4 * C:5 * C], rec[5 * C:], b["ss"], b["ws"], rec[:4 * C])
            else:
                Cx.bn_eval_scale_shift(C, gamma, beta, self.params["running_mean"], self.params["running_var"],
                                       self.epsilon, b["ss"])
            # keep bits (one byte per 8-channel vector): what backward needs of y, at 1/16 of its bytes
            bits = None
            if self.training and (self.use_relu or p_drop > 0) and C % 8 == 0 and os.environ.get("TNN_BN_BITMASK", "1") == "1":
                bits = torch.empty(M * (C // 8), dtype=torch.uint8, device=x.device)
            Cx.bn_apply(x, y, M, C, b["ss"], self.use_relu, p_drop, seed, stepstate.get(x.device).dev if p_drop > 0 else None,
                        bits)
            if self.training:
                c = self.cache(mb_id)
                c["x"] = x
                # TNN_BN_RECOMPUTE_MASK=1: backward rebuilds the ReLU / dropout mask from x + statistics +
                # the dropout counter instead of reading y (saves the y activation; measured 2.8% SLOWER
                # per WRN step on B200 -- the extra ALU work costs more than the saved read -- so off by default)
                keep_y = os.environ.get("TNN_BN_RECOMPUTE_MASK", "0") != "1"
                # BN-backward fusion (the consumer conv's dgrad epilogue rebuilds the mask from x): y is not
                # needed by backward at all then -- but keep it unless the consumer is known to fuse
                fuse_bwd = self._bwd_consumer is not None and os.environ.get("TNN_BN_FUSE_BWD", "0") == "1"
                c["y"] = y if ((self.use_relu or p_drop > 0) and keep_y and not fuse_bwd and bits is None) else None
                c["bits"] = bits
                c["mean"], c["invstd"], c["pack"] = rec[4 * C:5 * C], rec[5 * C:], rec[:4 * C]
                c["p"] = p_drop
                c["seed"] = seed
            return y
        xf = x.float().reshape(-1, C)
        if self.training:
            mean = xf.mean(0)
            var = xf.var(0, unbiased=False)
            M = xf.shape[0]
            with torch.no_grad():
                self.params["running_mean"].mul_(1 - self.momentum).add_(self.momentum * mean)
                self.params["running_var"].mul_(1 - self.momentum).add_(self.momentum * var * (M / max(M - 1, 1)))
        else:
            mean, var = self.params["running_mean"], self.params["running_var"]
        invstd = torch.rsqrt(var + self.epsilon)
        xh = (xf - mean) * invstd
        y = xh * gamma + beta if self.affine else xh
        if self.use_relu:
            y = torch.relu(y)
        if p_drop > 0:
            g = torch.Generator().manual_seed((self.seed * 1000003 + self._step * 7919 + mb_id) & 0x7FFFFFFF)
            self._step += 1
            keep = (torch.rand(y.shape, generator=g) >= p_drop).to(y.dtype) / (1 - p_drop)
            y = y * keep
        if self.training:
            c = self.cache(mb_id)
            c.update(x=xf, y=y, mean=mean, invstd=invstd, p=p_drop)
        return y.reshape(shp).to(x.dtype)

    def backward_impl(self, dy, mb_id, extra=None):
        c = self.cache(mb_id)
        C = self.num_features
        gamma = self.params.get("gamma")
        masked = self.use_relu or c["p"] > 0
        if dy.is_cuda:
            Cx = native.require()
            dy = to_bf16(dy).contiguous()
            x, y = c["x"], c["y"]
            M = x.numel() // C
            b = self._bufs(dy.device)
            state = stepstate.get(dy.device).dev if c["p"] > 0 else None
            beta = self.params.get("beta")
            if getattr(dy, "_tnn_bn_g", None) is self:
                # dy is already g = mask * d(out) / (1-p) and b["ws"] holds (sum g, sum g*xhat): the
                # consumer convolution's dgrad epilogue did the reduction pass (umma_gemm.cuh, Epilogue::bnb_*)
                Cx.bn_bwd_coef(b["ws"], M, C, gamma, c["invstd"], self.grads.get("gamma"), self.grads.get("beta"), b["coef"])
                dx = torch.empty_like(x)
                colsum = torch.empty(C, dtype=torch.float32, device=dy.device)
                Cx.bn_bwd_apply(dy, x, None, extra, dx, M, C, c["mean"], c["invstd"], b["coef"], False, 0.0, colsum,
                                b["cws"], gamma, beta, False, 0, None)
                dx = dx.reshape(dy.shape)
                dx._tnn_colsum = colsum
                return dx
            # (y is None with a mask = the forward dropped y expecting the fused backward or
            # TNN_BN_RECOMPUTE_MASK=1: the kernels below then rebuild the mask from x, the statistics and
            # the dropout counter)
            bits = c.get("bits")
            Cx.bn_bwd_reduce(dy, x, y, M, C, gamma, c["mean"], c["invstd"], masked, c["p"], self.grads.get("gamma"),
                             self.grads.get("beta"), b["coef"], b["ws"], beta, self.use_relu, c.get("seed", 0), state, bits)
            dx = torch.empty_like(x)
            colsum = torch.empty(C, dtype=torch.float32, device=dy.device)
            Cx.bn_bwd_apply(dy, x, y, extra, dx, M, C, c["mean"], c["invstd"], b["coef"], masked, c["p"], colsum,
                            b["cws"], gamma, beta, self.use_relu, c.get("seed", 0), state, bits)
            dx = dx.reshape(dy.shape)
            dx._tnn_colsum = colsum      # column sums of dx = bias gradient of the producing conv/dense
            return dx
        g = dy.float().reshape(-1, C)
        xf, y = c["x"], c["y"]
        if masked:
            scale = 1.0 / (1.0 - c["p"]) if c["p"] > 0 else 1.0
            g = g * (y != 0).to(g.dtype) * scale
        xh = (xf - c["mean"]) * c["invstd"]
        db, dg = g.sum(0), (g * xh).sum(0)
        if self.affine:
            self.grads["gamma"].add_(dg)
            self.grads["beta"].add_(db)
        M = g.shape[0]
        coef = (gamma if self.affine else 1.0) * c["invstd"]
        dx = coef * (g - db / M - xh * dg / M)
        if extra is not None:
            dx = dx + extra.float().reshape(-1, C)
        return dx.reshape(dy.shape).to(dy.dtype)

    def config_params(self):
        return dict(num_features=self.num_features, epsilon=self.epsilon, momentum=self.momentum, affine=self.affine,
                    use_relu=self.use_relu)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["num_features"], c.get("epsilon", 1e-5), c.get("momentum", 0.1), c.get("affine", True),
                   c.get("use_relu", False), c.name)


# =============================================================================== Pooling
class _Pool2D(Layer):
    def __init__(self, pool_h, pool_w, stride_h=None, stride_w=None, pad_h=0, pad_w=0, name="pool"):
        super().__init__(name)
        self.pool_h, self.pool_w = int(pool_h), int(pool_w)
        self.stride_h = int(stride_h) if stride_h else self.pool_h
        self.stride_w = int(stride_w) if stride_w else self.pool_w
        self.pad_h, self.pad_w = int(pad_h), int(pad_w)

    def _dims(self, s):
        N, H, W, C = s
        OH = (H + 2 * self.pad_h - self.pool_h) // self.stride_h + 1
        OW = (W + 2 * self.pad_w - self.pool_w) // self.stride_w + 1
        return [N, H, W, C, self.pool_h, self.pool_w, self.stride_h, self.stride_w, self.pad_h, self.pad_w, OH, OW]

    def output_shape(self, s):
        d = self._dims(s)
        return (s[0], d[10], d[11], s[3])

    def config_params(self):
        return dict(pool_h=self.pool_h, pool_w=self.pool_w, stride_h=self.stride_h, stride_w=self.stride_w,
                    pad_h=self.pad_h, pad_w=self.pad_w)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["pool_h"], c["pool_w"], c.get("stride_h"), c.get("stride_w"), c.get("pad_h", 0), c.get("pad_w", 0), c.name)


@register()
class MaxPool2DLayer(_Pool2D):
    TYPE_NAME = "maxpool2d"

    def forward_impl(self, x, mb_id):
        d = self._dims(x.shape)
        if x.is_cuda and x.shape[3] % 8 == 0:
            Cx = native.require()
            x = to_bf16(x).contiguous()
            y = torch.empty((d[0], d[10], d[11], d[3]), dtype=torch.bfloat16, device=x.device)
            mask = torch.empty(y.shape, dtype=torch.uint8, device=x.device) if self.training else None
            Cx.maxpool_fwd(x, y, mask, d)
            if self.training:
                self.cache(mb_id).update(mask=mask, d=d)
            return y
        xf = x.permute(0, 3, 1, 2).float()
        y = F.max_pool2d(xf, (self.pool_h, self.pool_w), (self.stride_h, self.stride_w), (self.pad_h, self.pad_w))
        if self.training:
            self.cache(mb_id).update(xf=xf)
        return y.permute(0, 2, 3, 1).contiguous().to(x.dtype)

    def backward_impl(self, dy, mb_id):
        c = self.cache(mb_id)
        if "mask" in c:
            Cx = native.require()
            d = c["d"]
            dy = to_bf16(dy).contiguous()
            dx = torch.empty((d[0], d[1], d[2], d[3]), dtype=torch.bfloat16, device=dy.device)
            Cx.maxpool_bwd(dy, c["mask"], dx, d)
            return dx
        xx = c["xf"].detach().clone().requires_grad_(True)
        y = F.max_pool2d(xx, (self.pool_h, self.pool_w), (self.stride_h, self.stride_w), (self.pad_h, self.pad_w))
        (gx,) = torch.autograd.grad(y, xx, dy.permute(0, 3, 1, 2).float())
        return gx.permute(0, 2, 3, 1).contiguous().to(dy.dtype)


@register()
class AvgPool2DLayer(_Pool2D):
    """Average pooling.  Like the reference's NHWC layer (src/nn/layers_impl/cpu/avgpool_ops.cpp:27-48,75-78)
    the divisor is the number of IN-BOUNDS pixels of the window, so zero padding does not dilute border
    outputs; the legacy NCHW layer divides by pool_h*pool_w instead (avgpool_nchw_ops.cpp:19,56) and is
    served by ``count_include_pad = True``.  The kernel divides by the window size; with padding the
    per-position ratio window/in-bounds is applied around it (a [OH,OW] table, cached per shape)."""
    TYPE_NAME = "avgpool2d"
    count_include_pad = False

    def _border_scale(self, d, device, dtype):
        """[1,OH,OW,1] factor (pool_h*pool_w / in-bounds count), or None when every window is full."""
        if self.count_include_pad or (self.pad_h == 0 and self.pad_w == 0):
            return None
        key = (tuple(d), str(device), dtype)
        tab = getattr(self, "_border_tab", None)
        if tab is None or tab[0] != key:
            H, W, OH, OW = d[1], d[2], d[10], d[11]
            h0 = torch.arange(OH) * self.stride_h - self.pad_h
            w0 = torch.arange(OW) * self.stride_w - self.pad_w
            nh = (torch.clamp(h0 + self.pool_h, max=H) - torch.clamp(h0, min=0)).clamp(min=0)
            nw = (torch.clamp(w0 + self.pool_w, max=W) - torch.clamp(w0, min=0)).clamp(min=0)
            cnt = (nh[:, None] * nw[None, :]).float()
            scale = torch.where(cnt > 0, float(self.pool_h * self.pool_w) / cnt.clamp(min=1), torch.zeros_like(cnt))
            tab = (key, scale.reshape(1, OH, OW, 1).to(device=device, dtype=dtype))
            self._border_tab = tab
        return tab[1]

    def forward_impl(self, x, mb_id):
        d = self._dims(x.shape)
        if self.training:
            self.cache(mb_id)["d"] = d
        if x.is_cuda and x.shape[3] % 8 == 0:
            Cx = native.require()
            x = to_bf16(x).contiguous()
            y = torch.empty((d[0], d[10], d[11], d[3]), dtype=torch.bfloat16, device=x.device)
            Cx.avgpool_fwd(x, y, d)
            sc = self._border_scale(d, x.device, torch.float32)
            return y if sc is None else (y.float() * sc).to(torch.bfloat16)
        y = F.avg_pool2d(x.permute(0, 3, 1, 2).float(), (self.pool_h, self.pool_w), (self.stride_h, self.stride_w),
                         (self.pad_h, self.pad_w), count_include_pad=self.count_include_pad)
        return y.permute(0, 2, 3, 1).contiguous().to(x.dtype)

    def backward_impl(self, dy, mb_id):
        d = self.cache(mb_id)["d"]
        if dy.is_cuda and d[3] % 8 == 0:
            Cx = native.require()
            sc = self._border_scale(d, dy.device, torch.float32)
            if sc is not None:
                dy = dy.float() * sc
            dy = to_bf16(dy).contiguous()
            dx = torch.empty((d[0], d[1], d[2], d[3]), dtype=torch.bfloat16, device=dy.device)
            Cx.avgpool_bwd(dy, dx, d)
            return dx
        x = torch.zeros((d[0], d[3], d[1], d[2]), requires_grad=True)
        y = F.avg_pool2d(x, (self.pool_h, self.pool_w), (self.stride_h, self.stride_w), (self.pad_h, self.pad_w),
                         count_include_pad=self.count_include_pad)
        (gx,) = torch.autograd.grad(y, x, dy.permute(0, 3, 1, 2).float().cpu())
        return gx.permute(0, 2, 3, 1).contiguous().to(dy.dtype).to(dy.device)


class _AvgPoolFullWindow(AvgPool2DLayer):
    """Divisor = pool_h*pool_w whatever the padding: the legacy NCHW semantics."""
    count_include_pad = True


# =============================================================================== Activations
def _act_ref(x, kind, slope):
    if kind == "relu":
        return torch.relu(x)
    if kind == "leaky_relu":
        return F.leaky_relu(x, slope)
    if kind == "sigmoid":
        return torch.sigmoid(x)
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "elu":
        return F.elu(x, slope)
    if kind == "gelu":
        return F.gelu(x, approximate="tanh")
    if kind == "softmax":
        return torch.softmax(x, -1)
    return x


@register()
class ActivationLayer(Layer):
    """none|relu|leaky_relu(0.01)|sigmoid|linear|tanh|elu|gelu(tanh)|softmax (activations.hpp:27-58)."""
    TYPE_NAME = "activation"

    def __init__(self, activation="relu", name="activation", slope=None):
        super().__init__(name)
        self.activation = activation.lower()
        if self.activation not in ACT_KIND and self.activation != "softmax":
            raise ValueError(f"unknown activation {activation}")
        self.slope = slope if slope is not None else (0.01 if self.activation == "leaky_relu" else 1.0)

    def forward_impl(self, x, mb_id):
        if self.activation in ("none", "linear"):
            return x
        if self.training:
            self.cache(mb_id)["x"] = x
        if self.activation == "softmax":
            y = tensor_ops.softmax(x.contiguous()) if x.is_cuda else torch.softmax(x.float(), -1).to(x.dtype)
            if self.training:
                self.cache(mb_id)["y"] = y
            return y
        if x.is_cuda:
            x = x.contiguous()
            y = torch.empty_like(x)
            native.require().act_fwd(x, y, ACT_KIND[self.activation], self.slope)
            return y
        return _act_ref(x.float(), self.activation, self.slope).to(x.dtype)

    def backward_impl(self, dy, mb_id):
        if self.activation in ("none", "linear"):
            return dy
        c = self.cache(mb_id)
        x = c["x"]
        if self.activation == "softmax":
            if dy.is_cuda:
                return tensor_ops.softmax_backward(c["y"], to_dtype(dy, c["y"].dtype).contiguous())
            y = c["y"].float()
            g = dy.float()
            return (y * (g - (g * y).sum(-1, keepdim=True))).to(dy.dtype)
        if dy.is_cuda:
            dy = to_dtype(dy, x.dtype).contiguous()
            dx = torch.empty_like(x)
            native.require().act_bwd(x, dy, dx, ACT_KIND[self.activation], self.slope)
            return dx
        xx = x.float().clone().requires_grad_(True)
        y = _act_ref(xx, self.activation, self.slope)
        (gx,) = torch.autograd.grad(y, xx, dy.float())
        return gx.to(dy.dtype)

    def config_params(self):
        return dict(activation=self.activation)

    @classmethod
    def create_from_config(cls, c):
        return cls(c.get("activation", "relu"), c.name)


class ActivationFactory:
    @staticmethod
    def create(name: str) -> ActivationLayer:
        return ActivationLayer(name, name)


# =============================================================================== Dropout
def _residual_add(x, residual):
    if x.is_cuda and x.dtype == torch.bfloat16 and residual.dtype == torch.bfloat16 and x.shape == residual.shape:
        o = torch.empty_like(x)
        native.require().add_bf16(x.contiguous(), residual.contiguous(), o, False)
        return o
    if x.is_cuda and x.shape == residual.shape and x.dtype in (torch.float16, torch.float32, torch.float64, torch.bfloat16):
        return tensor_ops.add(x.contiguous(), to_dtype(residual, x.dtype).contiguous())
    return x + residual.to(x.dtype)


@register()
class DropoutLayer(Layer):
    """Inverted dropout, scale 1/(1-p); the GPU mask is a pure function of (seed, index) and is
    recomputed in backward (reference stores a bool mask: cuda/dropout.cu:33-177)."""
    TYPE_NAME = "dropout"

    def __init__(self, dropout_rate=0.5, name="dropout"):
        super().__init__(name)
        self.dropout_rate = float(dropout_rate)
        self.fused_into_prev = False   # set by Sequential when the preceding BatchNorm applies it

    def forward_impl(self, x, mb_id, residual=None):
        """``residual`` (the shortcut of the enclosing residual block) is added in the same kernel."""
        if not self.training or self.dropout_rate <= 0 or self.fused_into_prev:
            return x if residual is None else _residual_add(x, residual)
        self._step += 1
        seed = (self.seed * 1000003 + self._step * 7919 + mb_id) & 0x7FFFFFFFFFFFFFFF
        if x.is_cuda:
            seed = (self.seed * 1000003 + (zlib.crc32(self.name.encode()) & 0xFFFFFF) * 7919 + mb_id) & 0x7FFFFFFFFFFFFFFF
            x = to_bf16(x).contiguous()
            y = torch.empty_like(x)
            res = None
            if residual is not None:
                res = to_bf16(residual).contiguous()
                if res.shape != x.shape:
                    res = res.expand_as(x).contiguous()
            native.require().dropout_fwd(x, y, self.dropout_rate, seed, stepstate.get(x.device).dev, res)
            self.cache(mb_id)["seed"] = seed
            return y
        g = torch.Generator().manual_seed(seed & 0x7FFFFFFF)
        keep = (torch.rand(x.shape, generator=g) >= self.dropout_rate).to(x.dtype) / (1 - self.dropout_rate)
        self.cache(mb_id)["keep"] = keep
        y = x * keep
        return y if residual is None else y + residual.to(y.dtype)

    def backward_impl(self, dy, mb_id):
        if not self.training or self.dropout_rate <= 0 or self.fused_into_prev:
            return dy
        c = self.cache(mb_id)
        if dy.is_cuda:
            dy = to_bf16(dy).contiguous()
            dx = torch.empty_like(dy)
            native.require().dropout_fwd(dy, dx, self.dropout_rate, c["seed"], stepstate.get(dy.device).dev)
            return dx
        return dy * c["keep"]

    def config_params(self):
        return dict(dropout_rate=self.dropout_rate)

    @classmethod
    def create_from_config(cls, c):
        return cls(c.get("dropout_rate", 0.5), c.name)


# =============================================================================== shape layers
@register()
class FlattenLayer(Layer):
    TYPE_NAME = "flatten"

    def __init__(self, start_dim=1, end_dim=-1, name="flatten"):
        super().__init__(name)
        self.start_dim, self.end_dim = int(start_dim), int(end_dim)

    def forward_impl(self, x, mb_id):
        self.cache(mb_id)["shape"] = x.shape
        return x.flatten(self.start_dim, self.end_dim)

    def backward_impl(self, dy, mb_id):
        return dy.reshape(self.cache(mb_id)["shape"])

    def output_shape(self, s):
        return tuple(torch.empty(s, device="meta").flatten(self.start_dim, self.end_dim).shape)

    def config_params(self):
        return dict(start_dim=self.start_dim, end_dim=self.end_dim)

    @classmethod
    def create_from_config(cls, c):
        return cls(c.get("start_dim", 1), c.get("end_dim", -1), c.name)


@register()
class IdentityLayer(Layer):
    TYPE_NAME = "identity"

    def forward_impl(self, x, mb_id):
        return x

    def backward_impl(self, dy, mb_id):
        return dy

    @classmethod
    def create_from_config(cls, c):
        return cls(c.name)


@register()
class TransposeLayer(Layer):
    """Swaps the last two dims ([B, L, D] <-> [B, D, L])."""
    TYPE_NAME = "transpose"

    def forward_impl(self, x, mb_id):
        return x.transpose(-1, -2).contiguous()

    def backward_impl(self, dy, mb_id):
        return dy.transpose(-1, -2).contiguous()

    def output_shape(self, s):
        return (*s[:-2], s[-1], s[-2])

    @classmethod
    def create_from_config(cls, c):
        return cls(c.name)


@register()
class SliceLayer(Layer):
    TYPE_NAME = "slice"

    def __init__(self, axis, start, length, name="slice"):
        super().__init__(name)
        self.axis, self.start, self.length = int(axis), int(start), int(length)

    def forward_impl(self, x, mb_id):
        self.cache(mb_id)["shape"] = x.shape
        return x.narrow(self.axis, self.start, self.length).contiguous()

    def backward_impl(self, dy, mb_id):
        dx = torch.zeros(self.cache(mb_id)["shape"], dtype=dy.dtype, device=dy.device)
        dx.narrow(self.axis, self.start, self.length).copy_(dy)
        return dx

    def output_shape(self, s):
        s = list(s)
        s[self.axis] = self.length
        return tuple(s)

    def config_params(self):
        return dict(axis=self.axis, start=self.start, length=self.length)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["axis"], c["start"], c["length"], c.name)


class _NAry(Layer):
    MULTI_INPUT = True
    OP = 0

    def forward_impl(self, xs, mb_id):
        if not isinstance(xs, (list, tuple)):
            xs = [xs]
        if self.training and self.OP == 2:
            self.cache(mb_id)["xs"] = list(xs)
        self.cache(mb_id)["n"] = len(xs)
        out = xs[0]
        for t in xs[1:]:
            if out.is_cuda and out.dtype == torch.bfloat16 and t.dtype == torch.bfloat16 and out.shape == t.shape:
                o = torch.empty_like(out)
                native.require().binary_bf16(out.contiguous(), t.contiguous(), o, self.OP)
                out = o
            elif out.is_cuda and out.shape == t.shape and out.dtype in (torch.float16, torch.float32, torch.float64, torch.bfloat16):
                f = (tensor_ops.add, tensor_ops.sub, tensor_ops.mul)[self.OP]
                out = f(out.contiguous(), to_dtype(t, out.dtype).contiguous())
            else:
                out = out + t if self.OP == 0 else (out - t if self.OP == 1 else out * t)
        return out

    def backward_impl(self, dy, mb_id):
        c = self.cache(mb_id)
        n = c["n"]
        if self.OP == 0:
            return [dy] * n
        if self.OP == 1:
            return [dy] + [-dy] * (n - 1)
        xs = c["xs"]
        outs = []
        for i in range(n):
            g = dy
            for j in range(n):
                if j != i:
                    g = g * xs[j]
            outs.append(g)
        return outs

    @classmethod
    def create_from_config(cls, c):
        return cls(c.name)


@register()
class AddLayer(_NAry):
    TYPE_NAME = "add"
    OP = 0


@register()
class SubLayer(_NAry):
    TYPE_NAME = "sub"
    OP = 1


@register()
class MulLayer(_NAry):
    TYPE_NAME = "mul"
    OP = 2


@register()
class MBroadcastLayer(Layer):
    """Fan one tensor out to M branches; backward sums the branch gradients."""
    TYPE_NAME = "mbroadcast"
    MULTI_OUTPUT = True

    def __init__(self, num_outputs=2, name="mbroadcast"):
        super().__init__(name)
        self.num_outputs = int(num_outputs)

    def forward_impl(self, x, mb_id):
        return [x] * self.num_outputs

    def backward_impl(self, dys, mb_id):
        if not isinstance(dys, (list, tuple)):
            return dys
        out = dys[0]
        for g in dys[1:]:
            out = out + g
        return out

    def config_params(self):
        return dict(num_outputs=self.num_outputs)

    @classmethod
    def create_from_config(cls, c):
        return cls(c.get("num_outputs", 2), c.name)


# =============================================================================== LayerNorm / GroupNorm
@register("layernorm")
class LayerNormLayer(Layer):
    TYPE_NAME = "layer_norm"

    def __init__(self, normalized_shape, epsilon=1e-5, affine=True, name="layer_norm"):
        super().__init__(name)
        self.normalized_shape, self.epsilon, self.affine = int(normalized_shape), float(epsilon), bool(affine)

    def param_descriptors(self):
        if not self.affine:
            return []
        C = self.normalized_shape
        return [ParamDescriptor("gamma", (C,), _const_init(1.0)), ParamDescriptor("beta", (C,), _const_init(0.0))]

    def forward_impl(self, x, mb_id, residual=None, return_sum=False):
        C = self.normalized_shape
        gamma, beta = self.params.get("gamma"), self.params.get("beta")
        if x.is_cuda:
            Cx = native.require()
            # explicit fp16 / fp32 / fp64 io dtype: the any-dtype warp-per-row kernel (gemm_xt.cu); default bf16: norm.cu
            xt = self.io_dtype in gemm_ops.XT_DTYPES
            x = (to_dtype(x, self.io_dtype) if xt else to_bf16(x)).contiguous()
            rows = x.numel() // C
            y = torch.empty_like(x)
            s = torch.empty_like(x) if residual is not None else None
            mean = torch.empty(rows, dtype=torch.float32, device=x.device)
            rstd = torch.empty_like(mean)
            if xt:
                Cx.layernorm_fwd_any(x, to_dtype(residual, x.dtype).contiguous() if residual is not None else None, y, s, rows, C, gamma, beta,
                                     self.epsilon, mean, rstd)
            else:
                Cx.layernorm_fwd(x, residual.contiguous() if residual is not None else None, y, s, rows, C, gamma, beta,
                                 self.epsilon, mean, rstd)
            xin = s if s is not None else x
            if self.training:
                self.cache(mb_id).update(x=xin, mean=mean, rstd=rstd)
            return (y, xin) if return_sum else y
        xf = x.float()
        if residual is not None:
            xf = xf + residual.float()
        mean = xf.mean(-1, keepdim=True)
        var = xf.var(-1, unbiased=False, keepdim=True)
        rstd = torch.rsqrt(var + self.epsilon)
        y = (xf - mean) * rstd
        if self.affine:
            y = y * gamma + beta
        if self.training:
            self.cache(mb_id).update(x=xf, mean=mean, rstd=rstd)
        y = y.to(x.dtype)
        return (y, xf.to(x.dtype)) if return_sum else y

    def backward_impl(self, dy, mb_id, extra=None):
        """``extra``: a gradient to ADD to dx (the residual-branch gradient of the enclosing block), fused
        into the dx kernel on the GPU."""
        c = self.cache(mb_id)
        C = self.normalized_shape
        gamma = self.params.get("gamma")
        if dy.is_cuda and c["x"].dtype in gemm_ops.XT_DTYPES:
            Cx = native.require()
            dy = to_dtype(dy, c["x"].dtype).contiguous()
            dx = torch.empty_like(dy)
            ex = to_dtype(extra, dy.dtype).contiguous() if extra is not None and extra.shape == dy.shape else None
            Cx.layernorm_bwd_any(dy, c["x"], dy.numel() // C, C, gamma, c["mean"], c["rstd"], dx, self.grads.get("gamma"),
                                 self.grads.get("beta"), ex)
            if extra is not None and ex is None:
                dx = dx + extra.to(dx.dtype)
            return dx
        if dy.is_cuda:
            Cx = native.require()
            dy = to_bf16(dy).contiguous()
            dx = torch.empty_like(dy)
            fuse = extra is not None and C % 8 == 0 and extra.dtype == torch.bfloat16 and extra.is_contiguous() and \
                extra.shape == dy.shape
            Cx.layernorm_bwd(dy, c["x"], dy.numel() // C, C, gamma, c["mean"], c["rstd"], dx, self.grads.get("gamma"),
                             self.grads.get("beta"), extra if fuse else None)
            if extra is not None and not fuse:
                dx = dx + extra.to(dx.dtype)
            return dx
        g = dy.float()
        xh = (c["x"] - c["mean"]) * c["rstd"]
        if self.affine:
            self.grads["gamma"].add_((g * xh).reshape(-1, C).sum(0))
            self.grads["beta"].add_(g.reshape(-1, C).sum(0))
            g = g * gamma
        dx = c["rstd"] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        if extra is not None:
            dx = dx + extra.float()
        return dx.to(dy.dtype)

    def config_params(self):
        return dict(normalized_shape=self.normalized_shape, epsilon=self.epsilon, affine=self.affine)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["normalized_shape"], c.get("epsilon", 1e-5), c.get("affine", True), c.name)


@register()
class GroupNormLayer(Layer):
    """GroupNorm on NHWC tensors (reference kernels index NCHW: cuda/groupnorm_ops.cu:46-262)."""
    TYPE_NAME = "groupnorm"

    def __init__(self, num_groups, num_channels, epsilon=1e-5, affine=True, name="groupnorm"):
        super().__init__(name)
        self.num_groups, self.num_channels = int(num_groups), int(num_channels)
        self.epsilon, self.affine = float(epsilon), bool(affine)
        assert self.num_channels % self.num_groups == 0

    def param_descriptors(self):
        if not self.affine:
            return []
        C = self.num_channels
        return [ParamDescriptor("gamma", (C,), _const_init(1.0)), ParamDescriptor("beta", (C,), _const_init(0.0))]

    def forward_impl(self, x, mb_id):
        C, G = self.num_channels, self.num_groups
        N = x.shape[0]
        HW = x.numel() // (N * C)
        gamma, beta = self.params.get("gamma"), self.params.get("beta")
        if x.is_cuda:
            Cx = native.require()
            x = to_bf16(x).contiguous()
            y = torch.empty_like(x)
            mean = torch.empty(N * G, dtype=torch.float32, device=x.device)
            rstd = torch.empty_like(mean)
            Cx.groupnorm_fwd(x, y, N, HW, C, G, gamma, beta, self.epsilon, mean, rstd)
            if self.training:
                self.cache(mb_id).update(x=x, mean=mean, rstd=rstd, dims=(N, HW))
            return y
        xf = x.float().reshape(N, HW, G, C // G)
        mean = xf.mean((1, 3), keepdim=True)
        var = xf.var((1, 3), unbiased=False, keepdim=True)
        rstd = torch.rsqrt(var + self.epsilon)
        xh = ((xf - mean) * rstd).reshape(N, HW, C)
        y = xh * gamma + beta if self.affine else xh
        if self.training:
            self.cache(mb_id).update(xh=xh, rstd=rstd, dims=(N, HW))
        return y.reshape(x.shape).to(x.dtype)

    def backward_impl(self, dy, mb_id):
        c = self.cache(mb_id)
        C, G = self.num_channels, self.num_groups
        N, HW = c["dims"]
        gamma = self.params.get("gamma")
        if dy.is_cuda:
            Cx = native.require()
            dy = to_bf16(dy).contiguous()
            dx = torch.empty_like(dy)
            Cx.groupnorm_bwd(dy, c["x"], N, HW, C, G, gamma, c["mean"], c["rstd"], dx, self.grads.get("gamma"),
                             self.grads.get("beta"))
            return dx
        g = dy.float().reshape(N, HW, C)
        xh = c["xh"]
        if self.affine:
            self.grads["gamma"].add_((g * xh).sum((0, 1)))
            self.grads["beta"].add_(g.sum((0, 1)))
            g = g * gamma
        g4, xh4 = g.reshape(N, HW, G, C // G), xh.reshape(N, HW, G, C // G)
        dx = c["rstd"] * (g4 - g4.mean((1, 3), keepdim=True) - xh4 * (g4 * xh4).mean((1, 3), keepdim=True))
        return dx.reshape(dy.shape).to(dy.dtype)

    def config_params(self):
        return dict(num_groups=self.num_groups, num_channels=self.num_channels, epsilon=self.epsilon, affine=self.affine)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["num_groups"], c["num_channels"], c.get("epsilon", 1e-5), c.get("affine", True), c.name)


# =============================================================================== Embedding family
@register()
class EmbeddingLayer(Layer):
    """Token embedding.  Ids are integers ([B, L] int32/int64); float ids (the reference's wire
    format, embedding_layer.cpp:34-79) are accepted and cast."""
    TYPE_NAME = "embedding"

    def __init__(self, vocab_size, embed_dim, padding_idx=-1, name="embedding"):
        super().__init__(name)
        self.vocab_size, self.embed_dim, self.padding_idx = int(vocab_size), int(embed_dim), int(padding_idx)

    def param_descriptors(self):
        return [ParamDescriptor("weight", (self.vocab_size, self.embed_dim), _normal_init(0.02))]

    def forward_impl(self, ids, mb_id, pos_table=None):
        ids = ids.to(torch.int32) if ids.dtype != torch.int32 else ids
        B, L = ids.shape[0], ids.shape[-1]
        D = self.embed_dim
        if self.training:
            self.cache(mb_id).update(ids=ids, fused_pos=pos_table is not None)
        if ids.is_cuda:
            Cx = native.require()
            y = torch.empty((*ids.shape, D), dtype=torch.bfloat16, device=ids.device)
            Cx.embedding_fwd(ids.contiguous(), self.params["weight"], pos_table, y, ids.numel(), L, D, self.vocab_size)
            return y
        y = self.params["weight"][ids.long().clamp(0, self.vocab_size - 1)]
        if pos_table is not None:
            y = y + pos_table[:L]
        return y

    def backward_impl(self, dy, mb_id, dpos=None):
        c = self.cache(mb_id)
        ids = c["ids"]
        D = self.embed_dim
        if dy.is_cuda:
            Cx = native.require()
            Cx.embedding_bwd(ids.contiguous(), dy.to(torch.bfloat16).contiguous(), self.grads["weight"], dpos,
                             ids.numel(), ids.shape[-1], D, self.vocab_size)
        else:
            self.grads["weight"].index_add_(0, ids.long().reshape(-1).clamp(0, self.vocab_size - 1), dy.float().reshape(-1, D))
            if dpos is not None:
                dpos[:ids.shape[-1]].add_(dy.float().reshape(-1, ids.shape[-1], D).sum(0))
        return None

    def output_shape(self, s):
        return (*s, self.embed_dim)

    def config_params(self):
        return dict(vocab_size=self.vocab_size, embed_dim=self.embed_dim, padding_idx=self.padding_idx)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["vocab_size"], c["embed_dim"], c.get("padding_idx", -1), c.name)


@register("positional_embedding")
class PositionalEmbeddingLayer(Layer):
    """Learned positional embedding added to [B, L, D] (positional_embedding_layer.cpp)."""
    TYPE_NAME = "pos_embedding"

    def __init__(self, embed_dim, seq_len, name="pos_embedding"):
        super().__init__(name)
        self.embed_dim, self.seq_len = int(embed_dim), int(seq_len)

    def param_descriptors(self):
        return [ParamDescriptor("weight", (self.seq_len, self.embed_dim), _normal_init(0.02))]

    def forward_impl(self, x, mb_id):
        L = x.shape[-2]
        return (x.float() + self.params["weight"][:L]).to(x.dtype)

    def backward_impl(self, dy, mb_id):
        L = dy.shape[-2]
        self.grads["weight"][:L].add_(dy.float().reshape(-1, L, self.embed_dim).sum(0))
        return dy

    def config_params(self):
        return dict(embed_dim=self.embed_dim, seq_len=self.seq_len)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["embed_dim"], c["seq_len"], c.name)


@register()
class ClassTokenLayer(Layer):
    """Prepends a learned class token: [B, L, D] -> [B, L+1, D] (cuda/class_token_ops.cu)."""
    TYPE_NAME = "class_token"

    def __init__(self, embed_dim, name="class_token"):
        super().__init__(name)
        self.embed_dim = int(embed_dim)

    def param_descriptors(self):
        return [ParamDescriptor("token", (self.embed_dim,), _normal_init(0.02))]

    def forward_impl(self, x, mb_id):
        B = x.shape[0]
        tok = self.params["token"].to(x.dtype).reshape(1, 1, -1).expand(B, 1, -1)
        return torch.cat([tok, x], 1)

    def backward_impl(self, dy, mb_id):
        self.grads["token"].add_(dy[:, 0].float().sum(0))
        return dy[:, 1:].contiguous()

    def output_shape(self, s):
        return (s[0], s[1] + 1, s[2])

    def config_params(self):
        return dict(embed_dim=self.embed_dim)

    @classmethod
    def create_from_config(cls, c):
        return cls(c["embed_dim"], c.name)


# =============================================================================== legacy NCHW facades
class _NCHWFacade(Layer):
    """Runs an NHWC layer on NCHW tensors (legacy API surface): layout changes go
    through the tiled transpose kernels, the math through the same NHWC kernels."""
    INNER = None

    def __init__(self, *a, **k):
        name = k.get("name", self.TYPE_NAME)
        super().__init__(name)
        self.inner = self.INNER(*a, **k)

    def children(self):
        return [self.inner]

    @staticmethod
    def _to_nhwc(x):
        if x.dim() != 4:
            return x
        if x.is_cuda and x.dtype in (torch.bfloat16, torch.float32):
            N, C, H, W = x.shape
            y = torch.empty((N, H, W, C), dtype=x.dtype, device=x.device)
            native.require().nchw_to_nhwc(x.contiguous(), y, N, C, H, W)
            return y
        return x.permute(0, 2, 3, 1).contiguous()

    @staticmethod
    def _to_nchw(x):
        if x.dim() != 4:
            return x
        if x.is_cuda and x.dtype in (torch.bfloat16, torch.float32):
            N, H, W, C = x.shape
            y = torch.empty((N, C, H, W), dtype=x.dtype, device=x.device)
            native.require().nhwc_to_nchw(x.contiguous(), y, N, C, H, W)
            return y
        return x.permute(0, 3, 1, 2).contiguous()

    def forward_impl(self, x, mb_id):
        return self._to_nchw(self.inner.forward(self._to_nhwc(x), mb_id))

    def backward_impl(self, dy, mb_id):
        g = self.inner.backward(self._to_nhwc(dy), mb_id)
        return self._to_nchw(g) if g is not None else None

    def set_training(self, flag):
        self.training = flag
        self.inner.set_training(flag)
        return self

    def output_shape(self, s):
        if len(s) == 4:
            o = self.inner.output_shape((s[0], s[2], s[3], s[1]))
            return (o[0], o[3], o[1], o[2])
        return self.inner.output_shape(s)

    def config_params(self):
        return self.inner.config_params()

    @classmethod
    def create_from_config(cls, c):
        inner = cls.INNER.create_from_config(c)
        obj = cls.__new__(cls)
        Layer.__init__(obj, c.name)
        obj.inner = inner
        return obj


@register()
class LegacyConv2DLayer(_NCHWFacade):
    TYPE_NAME = "legacy_conv2d"
    INNER = Conv2DLayer


@register()
class LegacyBatchNormLayer(_NCHWFacade):
    """BatchNorm over [N, C, H, W] (legacy_batchnorm_layer.cpp, cuda/batchnorm_nchw_ops.cu:119-550).  On the GPU the
    NCHW tensor is normalised IN ITS OWN LAYOUT by csrc/kernels/norm_nchw.cu (per-channel plane reductions, plane-wise
    apply; no NHWC round trip); parameters, running statistics, config and the CPU path are the inner BatchNormLayer's."""
    TYPE_NAME = "legacy_batchnorm"
    INNER = BatchNormLayer

    def _native(self, x) -> bool:
        return x.is_cuda and x.dim() == 4 and x.shape[1] == self.inner.num_features and self.inner.fused_dropout_p == 0.0 and \
            os.environ.get("TNN_LEGACY_BN_NATIVE", "1") == "1"

    def _nchw_bufs(self, dev):
        b = getattr(self, "_nchw", None)
        if b is None or b["ws"].device != dev:
            C = self.inner.num_features
            b = self._nchw = dict(ws=torch.zeros(2 * C, dtype=torch.float64, device=dev), ss=torch.zeros(2 * C, dtype=torch.float32, device=dev),
                                  coef=torch.zeros(3 * C, dtype=torch.float32, device=dev))
        return b

    def forward_impl(self, x, mb_id):
        if not self._native(x):
            return super().forward_impl(x, mb_id)
        bn, Cx = self.inner, native.require()
        C = bn.num_features
        x = (x if x.dtype in (torch.bfloat16, torch.float32) else to_bf16(x)).contiguous()
        b = self._nchw_bufs(x.device)
        gamma, beta = bn.params.get("gamma"), bn.params.get("beta")
        if self.training:
            rec = torch.empty(4 * C, dtype=torch.float32, device=x.device)        # mean | invstd | (scale, shift) pairs
            ss = rec[2 * C:]
            Cx.bn_nchw_train_stats(x, gamma, beta, bn.epsilon, bn.momentum, bn.params["running_mean"], bn.params["running_var"],
                                   rec[:C], rec[C:2 * C], ss, b["ws"])
            self.cache(mb_id).update(x=x, rec=rec)
        else:
            ss = b["ss"]
            Cx.bn_nchw_eval_scale_shift(C, gamma, beta, bn.params["running_mean"], bn.params["running_var"], bn.epsilon, ss)
        y = torch.empty_like(x)
        Cx.bn_nchw_apply(x, y, ss, bn.use_relu)
        return y

    def backward_impl(self, dy, mb_id):
        c = self.cache(mb_id)
        if "rec" not in c:
            return super().backward_impl(dy, mb_id)
        bn, Cx = self.inner, native.require()
        C = bn.num_features
        x, rec = c["x"], c["rec"]
        dy = to_dtype(dy, x.dtype).contiguous()
        b = self._nchw_bufs(dy.device)
        dx = torch.empty_like(x)
        Cx.bn_nchw_backward(dy, x, dx, bn.params.get("gamma"), rec[:C], rec[C:2 * C], rec[2 * C:], bn.use_relu, bn.grads.get("gamma"),
                            bn.grads.get("beta"), b["coef"], b["ws"])
        return dx


@register()
class LegacyMaxPool2DLayer(_NCHWFacade):
    TYPE_NAME = "legacy_maxpool2d"
    INNER = MaxPool2DLayer


@register()
class LegacyAvgPool2DLayer(_NCHWFacade):
    TYPE_NAME = "legacy_avgpool2d"
    INNER = _AvgPoolFullWindow


@register()
class LegacyDenseLayer(DenseLayer):
    TYPE_NAME = "legacy_dense"
