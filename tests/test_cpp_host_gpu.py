"""The C++ host on a B200: the compiled GPU unit tests (layer parity with the C++ CPU oracle, CUDA-graph step vs eager
step), numpy-level parity of GPU layers with PyTorch fp32 oracles, and the in-process flagship benchmark."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_cpp_host_cpu import _host, cfg, t  # noqa: E402


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def test_tnn_tests_binary_gpu_cases():
    _need_gpu()
    _host()
    out = subprocess.run([os.path.join(ROOT, "cpp", "bin", "tnn_tests"), "gpu"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:]
    assert " 0 failed" in out.stdout and "0 skipped" in out.stdout


def test_gpu_conv_and_batchnorm_layers_match_torch_fp32():
    _need_gpu()
    H = _host()
    rs = np.random.RandomState(0)
    for (C, K, stride, hw) in ((16, 32, 1, 16), (32, 64, 2, 16), (64, 64, 1, 8)):
        l = H.Layer(cfg("conv2d", in_channels=C, out_channels=K, kernel_h=3, kernel_w=3, stride_h=stride, stride_w=stride, pad_h=1, pad_w=1, use_bias=True), "GPU:0")
        p = l.params()
        x = rs.randn(8, hw, hw, C).astype("float32")
        xb = torch.from_numpy(x).bfloat16().float()                     # the GPU path computes on bf16 activations
        w = t(p["l.weight"]).bfloat16().float().permute(0, 3, 1, 2).requires_grad_(True)
        b = t(p["l.bias"]).requires_grad_(True)
        xt = xb.permute(0, 3, 1, 2).requires_grad_(True)
        yt = F.conv2d(xt, w, b, stride, 1)
        y = l.forward(xb.numpy())
        ref = yt.permute(0, 2, 3, 1).detach().numpy()
        assert np.abs(y - ref).max() <= 2e-2 * max(1.0, np.abs(ref).max())
        dy = torch.from_numpy(rs.randn(*y.shape).astype("float32")).bfloat16().float()
        yt.backward(dy.permute(0, 3, 1, 2))
        dx = l.backward(dy.numpy())
        g = l.grads()
        rdx = xt.grad.permute(0, 2, 3, 1).numpy()
        assert np.abs(dx - rdx).max() <= 2e-2 * max(1.0, np.abs(rdx).max())
        rw = w.grad.permute(0, 2, 3, 1).numpy()
        assert np.abs(g["l.weight"] - rw).max() <= 2e-2 * max(1.0, np.abs(rw).max())
        assert np.abs(g["l.bias"] - b.grad.numpy()).max() <= 2e-2 * max(1.0, np.abs(b.grad.numpy()).max())
    l = H.Layer(cfg("batchnorm", num_features=64, epsilon=1e-5, momentum=0.1, affine=True, use_relu=True), "GPU:0")
    x = torch.from_numpy(rs.randn(8, 8, 8, 64).astype("float32")).bfloat16().float()
    bn = torch.nn.BatchNorm2d(64, eps=1e-5, momentum=0.1)
    xt = x.permute(0, 3, 1, 2).requires_grad_(True)
    yt = torch.relu(bn(xt))
    y = l.forward(x.numpy())
    assert np.abs(y - yt.permute(0, 2, 3, 1).detach().numpy()).max() < 4e-2
    dy = torch.from_numpy(rs.randn(*y.shape).astype("float32")).bfloat16().float()
    yt.backward(dy.permute(0, 3, 1, 2))
    dx = l.backward(dy.numpy())
    # relative L2: a ReLU mask flips for pre-activations within bf16 rounding of zero
    r = xt.grad.permute(0, 2, 3, 1).numpy()
    assert np.linalg.norm(dx - r) / np.linalg.norm(r) < 3e-2
    assert np.abs(l.params()["l.running_var"] - bn.running_var.numpy()).max() < 2e-2


def test_cpp_trainer_learns_and_bench_train_reports_a_graph_captured_step():
    _need_gpu()
    H = _host()
    tr = H.Trainer("cifar10_resnet9", "GPU:0", 3)
    rs = np.random.RandomState(1)
    xs = rs.rand(64, 32, 32, 3).astype("float32")
    ys = (np.arange(64) % 10).astype("int32")
    losses = [tr.step(xs, ys)[0] for _ in range(12)]
    assert tr.captured() and losses[-1] < 0.7 * losses[0], losses
    out = json.loads(H.bench_train("cifar100_wrn16_8", 64, 4, 3, 0, True))
    assert out["cuda_graph"] and out["gpu_launches"] > 100 and out["value"] > 0 and out["e2e"]["value"] > 0
    assert out["e2e"]["h2d_bytes_per_step"] == 64 * 32 * 32 * 3 * 4 + 64 * 4


def test_cpp_gpt2_engine_on_gpu_bf16_fp8_graph_decode():
    """C++ GPT2Engine on the GPU against its own fp32 CPU oracle (same weights): bf16 prefill logits, graph decode ==
    eager decode, block-scaled fp8 prefill within the stated tolerance of bf16."""
    _need_gpu()
    H = _host()
    from tnn_b200.models.zoo import _gpt2
    seq = _gpt2(128, 2, 2, False, "tiny_gpt", None, seq_len=192, vocab=211, dropout=0.0)
    cj = json.dumps(seq.get_config().to_json())
    cpu = H.GPT2Engine(cj, "CPU", 5, 4)
    gpu = H.GPT2Engine(cj, "GPU:0", 6, 4)
    for k, v in cpu.params().items():
        gpu.set_param(k, v)
    ids = np.random.RandomState(0).randint(0, 211, (4, 160)).astype(np.int32)
    cn, cl = cpu.prefill(ids, True)
    gn, gl = gpu.prefill(ids, True)
    rel = np.linalg.norm(gl - cl) / np.linalg.norm(cl)
    assert rel < 4e-2, rel
    eager = []
    tok = gn
    eager_l1 = None
    for i in range(6):
        tok, lg = gpu.decode(tok, i == 0)
        eager_l1 = lg if i == 0 else eager_l1
        eager.append(tok.copy())
    os.environ["TNN_DECODE_MEGA"] = "0"              # multi-kernel graph: the same kernels as the eager path -> identical tokens
    try:
        gpu.prefill(ids)
        graph = gpu.decode_graph(gn, 6)
        assert not gpu.uses_persistent_decode()
    finally:
        del os.environ["TNN_DECODE_MEGA"]
    assert (graph == np.stack(eager, 1)).all()
    # persistent whole-model decode kernel (one launch per token): same math, different reduction order
    mega = H.GPT2Engine(cj, "GPU:0", 6, 4)
    for k, v in cpu.params().items():
        mega.set_param(k, v)
    mega.prefill(ids)
    m1 = mega.decode_graph(gn, 1)
    assert mega.uses_persistent_decode() and mega.cur_len() == 161
    ml = mega.last_logits()
    assert np.linalg.norm(ml - eager_l1) / np.linalg.norm(eager_l1) < 2e-2
    assert (m1[:, 0] == eager[0]).mean() >= 0.75
    mega.prefill(ids)
    m6 = mega.decode_graph(gn, 6)
    assert (m6[:, 0] == m1[:, 0]).all() and mega.cur_len() == 166
    mega.prefill(ids)
    assert (mega.decode_graph(gn, 6) == m6).all()
    f8 = H.GPT2Engine(cj, "GPU:0", 6, 4, 0, "fp8")
    for k, v in cpu.params().items():
        f8.set_param(k, v)
    f8.enable_fp8()                                   # re-quantise after the weights changed
    _, l8 = f8.prefill(ids, True)
    rel8 = np.linalg.norm(l8 - gl) / np.linalg.norm(gl)
    assert rel8 < 0.12, rel8
    assert f8.precision() == "fp8"
