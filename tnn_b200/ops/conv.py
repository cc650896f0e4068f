"""NHWC Conv2d fprop / dgrad / wgrad.

GPU paths (all hand-written, csrc/kernels/umma_gemm.cu + conv_simt.cu):
  * ``tma``     implicit GEMM on tcgen05; the activation operand is fetched as 4-D TMA boxes
                shifted per filter tap (zero halo = TMA OOB fill, no im2col buffer).  Stride 2
                runs on a phase-split copy of the input so every tap is again a unit-stride
                window; stride-2 dgrad is four per-phase gathers with strided epilogue stores.
  * ``direct``  SIMT kernel for C_in < 8 (first layer).
  * ``im2col``  SIMT im2col/col2im + the tcgen05 GEMM for geometries outside the window path.

Reference: cuDNN-frontend conv graphs (src/nn/layers_impl/cuda/cudnn_conv2d_ops.cu:73-414) and
im2col + cuBLAS (src/nn/layers_impl/legacy_conv2d_layer.cpp:137-227,261-410).
CPU path: torch fp32 reference (the oracle used by the tests).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from . import native
from .gemm import ACT, OUT_BF16, OUT_F32_ATOMIC, bias_grad


def out_size(h: int, k: int, s: int, p: int) -> int:
    return (h + 2 * p - k) // s + 1


@dataclass
class ConvPlan:
    N: int
    H: int
    W: int
    C: int
    K: int
    R: int
    S: int
    stride: int
    pad: int
    OH: int = 0
    OW: int = 0
    path: str = "ref"
    stride_w: int = 0        # anisotropic geometry: width stride / pad (0 / -1 = same as stride / pad)
    pad_w: int = -1
    # fprop / wgrad description
    f_geom: list = field(default_factory=list)
    f_taps: list = field(default_factory=list)
    # dgrad: list of (geom, taps, grid-spec) launches
    d_launches: list = field(default_factory=list)
    d_needs_zero: bool = False


def _fprop_taps(R, S, C, stride, pad, N):
    taps = []
    for r in range(R):
        for s in range(S):
            if stride == 1:
                taps += [r - pad, s - pad, 0, (r * S + s) * C]
            else:
                eh, ew = r - pad, s - pad
                ph, pw = eh % 2, ew % 2
                taps += [(eh - ph) // 2, (ew - pw) // 2, (ph * 2 + pw) * N, (r * S + s) * C]
    return taps


def make_plan(N, H, W, C, K, R, S, stride, pad, device, stride_w=None, pad_w=None) -> ConvPlan:
    p = ConvPlan(N, H, W, C, K, R, S, stride, pad)
    aniso = (stride_w is not None and stride_w != stride) or (pad_w is not None and pad_w != pad)
    if aniso:
        p.stride_w, p.pad_w = int(stride_w if stride_w is not None else stride), int(pad_w if pad_w is not None else pad)
    p.OH, p.OW = out_size(H, R, stride, pad), out_size(W, S, _sw(p), _pw(p))
    if device.type != "cuda":
        p.path = "ref"
        return p
    Cx = native.require()
    if aniso:                       # height / width stride or padding differ: the im2col + tcgen05 GEMM path of the same kernels
        p.path = "im2col"
        return p
    if C < 8 and K % 8 == 0 and K * R * S * C <= 2048:
        p.path = "direct"
        return p
    p.path = "im2col"
    if R * S <= 9 and stride in (1, 2) and C % 8 == 0 and K % 8 == 0 and (R * S == 1 or R * S % 3 == 0):
        if stride == 2 and (H % 2 or W % 2):
            return p
        if stride == 1:
            geom = [N, H, W, C, N, p.OH, p.OW]
        else:
            geom = [4 * N, H // 2, W // 2, C, N, p.OH, p.OW]
        taps = _fprop_taps(R, S, C, stride, pad, N)
        if not Cx.conv_geom_supported(geom, taps):
            return p
        # dgrad launches (input of the GEMM = dY grid, output = dX grid or one phase of it)
        launches = []
        needs_zero = False
        if stride == 1:
            dpad = R - 1 - pad
            dtaps = []
            for r in range(R):
                for s in range(S):
                    dtaps += [r - dpad, s - (S - 1 - pad), 0, (r * S + s) * K]
            dgeom = [N, p.OH, p.OW, K, N, H, W]
            if not Cx.conv_geom_supported(dgeom, dtaps):
                return p
            launches.append((dgeom, dtaps, []))
        else:
            for ph in range(2):
                for pw in range(2):
                    dtaps = []
                    for r in range(R):
                        if (ph + pad - r) % 2:
                            continue
                        for s in range(S):
                            if (pw + pad - s) % 2:
                                continue
                            rf, sf = R - 1 - r, S - 1 - s
                            dtaps += [(ph + pad - r) // 2, (pw + pad - s) // 2, 0, (rf * S + sf) * K]
                    if not dtaps:
                        needs_zero = True
                        continue
                    dgeom = [N, p.OH, p.OW, K, N, H // 2, W // 2]
                    if not Cx.conv_geom_supported(dgeom, dtaps):
                        return p
                    grid = [H // 2, W // 2, H * W * C, 2 * W * C, 2 * C, (ph * W + pw) * C]
                    launches.append((dgeom, dtaps, grid))
        p.path = "tma"
        p.f_geom, p.f_taps = geom, taps
        p.d_launches, p.d_needs_zero = launches, needs_zero
    return p


def _sw(plan):
    return plan.stride_w or plan.stride


def _pw(plan):
    return plan.pad if plan.pad_w < 0 else plan.pad_w


def _geo12(plan, N, OH, OW):
    """{N,H,W,C,R,S,sh,sw,ph,pw,OH,OW} for the im2col / col2im kernels (height / width stride and padding)."""
    return [N, plan.H, plan.W, plan.C, plan.R, plan.S, plan.stride, _sw(plan), plan.pad, _pw(plan), OH, OW]


# ------------------------------------------------------------------------------- reference
def conv2d_fwd_ref(x, w, bias, stride, pad):
    y = F.conv2d(x.permute(0, 3, 1, 2).float(), w.permute(0, 3, 1, 2).float(), bias.float() if bias is not None else None,
                 stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


def conv2d_dgrad_ref(dy, w, x_shape, stride, pad):
    N, H, W, C = x_shape
    dx = torch.nn.grad.conv2d_input((N, C, H, W), w.permute(0, 3, 1, 2).float(), dy.permute(0, 3, 1, 2).float(),
                                    stride=stride, padding=pad)
    return dx.permute(0, 2, 3, 1).contiguous()


def conv2d_wgrad_ref(dy, x, w_shape, stride, pad):
    K, R, S, C = w_shape
    dw = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).float(), (K, C, R, S), dy.permute(0, 3, 1, 2).float(),
                                     stride=stride, padding=pad)
    return dw.permute(0, 2, 3, 1).contiguous()


# ------------------------------------------------------------------------------- GPU entry points
class ConvState:
    """Per-layer device state: bf16 weight views, dgrad (flipped/transposed) weights, caches."""

    def __init__(self):
        self.wt = None            # bf16 [C, R, S, K] flipped
        self.wt_version = -1
        self.w_pad = None         # padded bf16 weight matrix for the im2col path


def conv2d_fwd(plan: ConvPlan, x, w, w_bf16, bias, residual=None, act=None, cache: dict | None = None,
               stats_ws: torch.Tensor | None = None):
    """x [N,H,W,C]; w fp32 [K,R,S,C]; w_bf16 bf16 view of w (or None).  Returns y [N,OH,OW,K].
    ``stats_ws`` (tma path only): fp32 workspace the epilogue accumulates per-channel sum / sum of
    squares of y into (BatchNorm statistics fusion, finalised by ``bn_finalize_stats``)."""
    if plan.path == "ref":
        y = conv2d_fwd_ref(x, w, bias, (plan.stride, _sw(plan)), (plan.pad, _pw(plan)))
        if residual is not None:
            y = y + residual.float()
        if act == "relu":
            y = torch.relu(y)
        return y.to(x.dtype)
    Cx = native.require()
    N, OH, OW, K = plan.N, plan.OH, plan.OW, plan.K
    y = torch.empty((N, OH, OW, K), device=x.device, dtype=torch.bfloat16)
    if plan.path == "direct":
        assert residual is None and act in (None, "none", "linear")
        Cx.conv_direct_fwd(x, w, bias, y, [N, plan.H, plan.W, plan.C, K, plan.R, plan.S, plan.stride, plan.pad, OH, OW])
        return y
    if w_bf16 is None:
        w_bf16 = w.to(torch.bfloat16)
    rsc = plan.R * plan.S * plan.C
    if plan.path == "tma":
        xin = x
        if plan.stride == 2:
            xin = torch.empty((4 * N, plan.H // 2, plan.W // 2, plan.C), device=x.device, dtype=torch.bfloat16)
            Cx.phase_split(x, xin, N, plan.H, plan.W, plan.C)
            if cache is not None:
                cache["x_ps"] = xin
        Cx.conv_fprop(xin, w_bf16, rsc, K, plan.f_geom, plan.f_taps, y, K, [], OUT_BF16, ACT[act], 1.0, bias, residual,
                      stats_ws)
        return y
    # im2col fallback
    ld = (rsc + 7) // 8 * 8
    col = torch.empty((N * OH * OW, ld), device=x.device, dtype=torch.bfloat16)
    if ld != rsc:
        col.zero_()
    Cx.im2col_nhwc(x, col, ld, _geo12(plan, N, OH, OW))
    wm = w_bf16.reshape(K, rsc)
    if ld != rsc:
        wp = torch.zeros((K, ld), device=x.device, dtype=torch.bfloat16)
        wp[:, :rsc] = wm
        wm = wp
    if cache is not None:
        cache["col"] = col
    Cx.gemm(col, wm, N * OH * OW, K, ld, ld, ld, False, False, y, K, OUT_BF16, ACT[act], 1.0, bias,
            residual.reshape(-1, K) if residual is not None else None, 1)
    return y


def conv2d_dgrad(plan: ConvPlan, dy, w, state: ConvState, w_version: int = 0, residual=None, bn_bwd: dict | None = None):
    """dx [N,H,W,C] from dy [N,OH,OW,K].  `residual` (same shape as dx) is added in the epilogue.
    ``bn_bwd`` (tma path, stride 1): fuse the backward reduction of the BatchNorm(+ReLU+dropout) that
    produced this convolution's input into the epilogue -- the result is then g = mask * dx / (1-p) and
    ``bn_bwd["ws"]`` receives per-channel (sum g, sum g * xhat); keys x, pack, relu, p, seed, state, ws."""
    if plan.path == "ref":
        dx = conv2d_dgrad_ref(dy, w, (plan.N, plan.H, plan.W, plan.C), (plan.stride, _sw(plan)), (plan.pad, _pw(plan)))
        if residual is not None:
            dx = dx + residual.float()
        return dx.to(dy.dtype)
    Cx = native.require()
    N, H, W, C, K = plan.N, plan.H, plan.W, plan.C, plan.K
    rsk = plan.R * plan.S * K
    if plan.path == "tma":
        if state.wt is None or state.wt_version != w_version:
            if state.wt is None:
                state.wt = torch.empty((C, plan.R, plan.S, K), device=dy.device, dtype=torch.bfloat16)
            Cx.conv_weight_prep(w, None, state.wt, K, plan.R, plan.S, C)
            state.wt_version = w_version
        if plan.d_needs_zero:
            dx = torch.zeros((N, H, W, C), device=dy.device, dtype=torch.bfloat16)
            if residual is not None:
                dx.copy_(residual)   # phases without taps keep the residual; others overwrite
        else:
            dx = torch.empty((N, H, W, C), device=dy.device, dtype=torch.bfloat16)
        if bn_bwd is not None:
            assert residual is None and len(plan.d_launches) == 1 and not plan.d_needs_zero
            geom, taps, grid = plan.d_launches[0]
            Cx.conv_fprop(dy, state.wt, rsk, C, geom, taps, dx, C, grid, OUT_BF16, 0, 1.0, None, None, bn_bwd["ws"],
                          bn_bwd["x"], bn_bwd["pack"], bool(bn_bwd["relu"]), float(bn_bwd["p"]), int(bn_bwd["seed"]),
                          bn_bwd["state"])
            return dx
        for geom, taps, grid in plan.d_launches:
            Cx.conv_fprop(dy, state.wt, rsk, C, geom, taps, dx, C, grid, OUT_BF16, 0, 1.0, None, residual)
        return dx
    # im2col fallback: dcol = dy @ W  then col2im
    rsc = plan.R * plan.S * C
    ld = (rsc + 7) // 8 * 8
    wm = w.to(torch.bfloat16).reshape(K, rsc)
    if ld != rsc:
        wp = torch.zeros((K, ld), device=dy.device, dtype=torch.bfloat16)
        wp[:, :rsc] = wm
        wm = wp
    M = N * plan.OH * plan.OW
    dcol = torch.empty((M, ld), device=dy.device, dtype=torch.bfloat16)
    Cx.gemm(dy.reshape(M, K), wm, M, ld, K, K, ld, False, True, dcol, ld, OUT_BF16, 0, 1.0, None, None, 1)
    dx = torch.empty((N, H, W, C), device=dy.device, dtype=torch.bfloat16)
    Cx.col2im_nhwc(dcol, ld, dx, _geo12(plan, N, plan.OH, plan.OW))
    if residual is not None:
        Cx.add_bf16(dx, residual, dx, False)
    return dx


def conv2d_wgrad(plan: ConvPlan, dy, x, dw, db=None, cache: dict | None = None, alpha: float = 1.0):
    """dw[K,R,S,C] (fp32) += alpha * wgrad ; db[K] += alpha * sum(dy)."""
    if plan.path == "ref":
        dw.add_(alpha * conv2d_wgrad_ref(dy, x, (plan.K, plan.R, plan.S, plan.C), (plan.stride, _sw(plan)), (plan.pad, _pw(plan))))
        if db is not None:
            db.add_(alpha * dy.float().sum((0, 1, 2)))
        return
    Cx = native.require()
    N, OH, OW, K = plan.N, plan.OH, plan.OW, plan.K
    M = N * OH * OW
    rsc = plan.R * plan.S * plan.C
    if plan.path == "direct":
        assert alpha == 1.0
        Cx.conv_direct_wgrad(x, dy, dw, db, [N, plan.H, plan.W, plan.C, K, plan.R, plan.S, plan.stride, plan.pad, OH, OW])
        return
    if plan.path == "tma":
        xin = x
        if plan.stride == 2:
            xin = cache.get("x_ps") if cache else None
            if xin is None:
                xin = torch.empty((4 * N, plan.H // 2, plan.W // 2, plan.C), device=x.device, dtype=torch.bfloat16)
                Cx.phase_split(x, xin, N, plan.H, plan.W, plan.C)
        Cx.conv_wgrad(dy, K, K, xin, plan.f_geom, plan.f_taps, dw, rsc, float(alpha))
    else:
        ld = (rsc + 7) // 8 * 8
        col = cache.get("col") if cache else None
        if col is None:
            col = torch.empty((M, ld), device=x.device, dtype=torch.bfloat16)
            if ld != rsc:
                col.zero_()
            Cx.im2col_nhwc(x, col, ld, _geo12(plan, N, OH, OW))
        tiles = ((K + 127) // 128) * ((rsc + 127) // 128)
        splits = max(1, min((M + 63) // 64, (2 * 148 + tiles - 1) // tiles))
        # dW[K, rsc] += dy^T[K, M] @ col[M, rsc]; N dimension limited to the real rsc columns
        Cx.gemm(dy.reshape(M, K), col, K, rsc, M, K, ld, True, True, dw, rsc, OUT_F32_ATOMIC, 0, float(alpha), None,
                None, splits)
    if db is not None:
        bias_grad(dy.reshape(M, K), db, alpha)
