"""Bring-up probe: runs each kernel case in its own subprocess (a faulting kernel kills only its
case), prints relative errors.  Usage: python tests/gpu_probe.py [filter]"""
import math
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _gemm(M, N, K, a_mn, b_mn):
    import torch
    from tnn_b200.ops import native
    C = native.require()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    ref = A.float() @ B.float().t()
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    C.gemm(Ain, Bin, M, N, K, Ain.stride(0), Bin.stride(0), a_mn, b_mn, out, N, 0, 0, 1.0, None, None, 1)
    torch.cuda.synchronize()
    return rel(out, ref)


@case
def gemm_kk_128():
    return _gemm(128, 128, 64, False, False)


@case
def gemm_kk_big():
    return _gemm(1024, 512, 768, False, False)


@case
def gemm_kk_odd():
    return _gemm(200, 104, 72, False, False)


@case
def gemm_k_mn():
    return _gemm(256, 256, 256, False, True)


@case
def gemm_mn_k():
    return _gemm(256, 256, 256, True, False)


@case
def gemm_mn_mn():
    return _gemm(256, 256, 256, True, True)


@case
def gemm_bn64():
    return _gemm(256, 64, 256, False, False)


@case
def gemm_bn256():
    os.environ["TNN_GEMM_BLOCK_N"] = "256"
    return _gemm(512, 512, 512, False, False)


@case
def gemm_splitk_atomic():
    import torch
    from tnn_b200.ops import gemm
    dev = torch.device("cuda")
    dy = torch.randn(4096, 256, device=dev).bfloat16()
    x = torch.randn(4096, 128, device=dev).bfloat16()
    dw = torch.zeros(256, 128, device=dev)
    gemm.linear_wgrad(dy, x, dw)
    torch.cuda.synchronize()
    return rel(dw, dy.float().t() @ x.float())


def _conv(N, H, W, C, K, R, stride, pad, what):
    import torch
    from tnn_b200.ops import conv
    dev = torch.device("cuda")
    torch.manual_seed(2)
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = (torch.randn(K, R, R, C, device=dev) / math.sqrt(R * R * C)).bfloat16().float()
    b = torch.randn(K, device=dev)
    plan = conv.make_plan(N, H, W, C, K, R, R, stride, pad, dev)
    st = conv.ConvState()
    cache = {}
    y = conv.conv2d_fwd(plan, x, w, w.bfloat16(), b, cache=cache)
    ref = conv.conv2d_fwd_ref(x, w, b, stride, pad)
    torch.cuda.synchronize()
    out = {"path": plan.path, "fwd": rel(y, ref)}
    dy = torch.randn_like(ref).bfloat16()
    if "d" in what and plan.path != "direct":
        dx = conv.conv2d_dgrad(plan, dy, w, st)
        torch.cuda.synchronize()
        out["dgrad"] = rel(dx, conv.conv2d_dgrad_ref(dy, w, (N, H, W, C), stride, pad))
    if "w" in what:
        dw = torch.zeros(K, R, R, C, device=dev)
        db = torch.zeros(K, device=dev)
        conv.conv2d_wgrad(plan, dy, x, dw, db, cache=cache)
        torch.cuda.synchronize()
        out["wgrad"] = rel(dw, conv.conv2d_wgrad_ref(dy, x, (K, R, R, C), stride, pad))
        out["bgrad"] = rel(db, dy.float().sum((0, 1, 2)))
    return out


@case
def conv_3x3_c128_32():
    return _conv(8, 32, 32, 128, 128, 3, 1, 1, "dw")


@case
def conv_3x3_c16_32():
    return _conv(8, 32, 32, 16, 128, 3, 1, 1, "dw")


@case
def conv_3x3_c256_16():
    return _conv(8, 16, 16, 256, 256, 3, 1, 1, "dw")


@case
def conv_3x3_c512_8():
    return _conv(8, 8, 8, 512, 512, 3, 1, 1, "dw")


@case
def conv_3x3_s2():
    return _conv(8, 32, 32, 128, 256, 3, 2, 1, "dw")


@case
def conv_1x1_s1():
    return _conv(8, 32, 32, 16, 128, 1, 1, 0, "dw")


@case
def conv_1x1_s2():
    return _conv(8, 32, 32, 128, 256, 1, 2, 0, "dw")


@case
def conv_direct():
    return _conv(4, 32, 32, 3, 16, 3, 1, 1, "w")


@case
def conv_im2col():
    return _conv(4, 28, 28, 8, 16, 5, 1, 0, "dw")


@case
def conv_4x4():
    return _conv(16, 4, 4, 512, 512, 3, 1, 1, "dw")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        r = CASES[sys.argv[2]]()
        print(f"RESULT {sys.argv[2]} {r}")
        sys.exit(0)
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    for name in CASES:
        if filt and filt not in name:
            continue
        try:
            p = subprocess.run([sys.executable, __file__, "--run", name], capture_output=True, text=True, timeout=120)
            lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
            if lines:
                print(lines[-1], flush=True)
            else:
                tail = (p.stderr or "").strip().splitlines()[-6:]
                print(f"FAILED {name} rc={p.returncode}: " + " | ".join(tail), flush=True)
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT {name}", flush=True)
