"""CTA-pair (cta_group::2) GEMM / conv path: numerics vs fp32 reference and speed vs the single-CTA
kernel.  Run as  TNN_GEMM_2CTA=1 python tests/gemm2_check.py  (the switch is read once per process)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tnn_b200.nn.layers import Conv2DLayer
from tnn_b200.ops import gemm


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def timeit(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    dev = torch.device("cuda", 0)
    mode = os.environ.get("TNN_GEMM_2CTA", "0")
    ok = True
    torch.manual_seed(0)
    for (M, N, K) in [(256, 128, 64), (256, 256, 128), (512, 256, 576), (384, 128, 192), (1000, 520, 136),
                      (4096, 768, 3072), (32768, 2304, 768), (300, 200, 72)]:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev).bfloat16()
        y = gemm.linear_fwd(x, w, b, act="relu", residual=res)
        ref = torch.relu(x.float() @ w.float().t() + b + res.float())
        e = rel(y, ref)
        good = e < 1e-2
        ok &= good
        t = timeit(lambda: gemm.linear_fwd(x, w, b))
        print(f"[2cta={mode}] gemm {M}x{N}x{K}: rel {e:.2e} {'OK' if good else 'FAIL'}  {t * 1e3:.1f} us  {2.0 * M * N * K / t / 1e9:.0f} TFLOP/s", flush=True)
    for (Nb, H, C, K, ks, st) in [(4, 32, 128, 128, 3, 1), (8, 16, 256, 256, 3, 1), (256, 8, 512, 512, 3, 1), (256, 32, 128, 128, 3, 1),
                                  (256, 16, 256, 256, 3, 1), (256, 32, 128, 256, 3, 2), (6, 16, 64, 128, 1, 1), (256, 32, 16, 128, 3, 1)]:
        pad = ks // 2
        l = Conv2DLayer(C, K, ks, ks, st, st, pad, pad, True, "c").init(dev, 1)
        x = torch.randn(Nb, H, H, C, device=dev).bfloat16()
        y = l.forward(x, 0)
        wr = l.params["weight"].bfloat16().float()
        ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wr.permute(0, 3, 1, 2), l.params["bias"], st, pad).permute(0, 2, 3, 1)
        e = rel(y, ref)
        dy = torch.randn_like(y)
        dx = l.backward(dy, 0)
        refdx = torch.nn.grad.conv2d_input((Nb, C, H, H), wr.permute(0, 3, 1, 2), dy.float().permute(0, 3, 1, 2), st, pad).permute(0, 2, 3, 1)
        e2 = rel(dx, refdx)
        good = e < 1e-2 and e2 < 1e-2
        ok &= good
        t = timeit(lambda: l.forward(x, 0))
        fl = 2.0 * y.numel() * ks * ks * C
        print(f"[2cta={mode}] conv N{Nb} {H}x{H} {C}->{K} k{ks}s{st}: fwd rel {e:.2e} dgrad rel {e2:.2e} {'OK' if good else 'FAIL'}  "
              f"fwd {t * 1e3:.1f} us {fl / t / 1e9:.0f} TFLOP/s", flush=True)
    print("GEMM2_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
