"""CPU tests of the host runtime: dtype/config, archiver wire format, tensor records, allocators,
graph IR + checkpoints, optimizers vs torch.optim, schedulers, losses, data path, tokenizer,
profiler, and an end-to-end mnist_cnn training run (BASELINE config 1: plumbing on CPU).

Reference tests mirrored: unit_tests/archiver_test.cpp, tensor_test.cpp (save/load),
bf16_test.cpp, device_manager_test.cpp; plus what the reference lacks:
optimizer/scheduler oracles and checkpoint round trips including optimizer state.
"""
import io
import math
import os
import struct

import numpy as np
import pytest
import torch

from tnn_b200.core import tensor_io
from tnn_b200.core.config import TConfig
from tnn_b200.core.dtype import DType, dtype_eps, dtype_size, dtype_to_string, from_torch, string_to_dtype, to_torch
from tnn_b200.models import ExampleModels, load_or_create_model
from tnn_b200.nn import (Graph, GraphBuilder, GraphExecutor, LossFactory, OptimizerFactory, SchedulerFactory,
                         compute_class_corrects, graph_from_sequential)
from tnn_b200.nn.layers import ActivationLayer, AddLayer, DenseLayer, MBroadcastLayer
from tnn_b200.ops import native


# ------------------------------------------------------------------------------- dtype / config
def test_dtype_table():
    assert string_to_dtype("FP32") == string_to_dtype("FLOAT") == DType.FP32
    assert string_to_dtype("BF16") == DType.BF16 and string_to_dtype("DOUBLE") == DType.FP64
    assert dtype_size(DType.BF16) == 2 and dtype_size(DType.FP64) == 8 and dtype_size(DType.INT64_T) == 8
    assert to_torch(DType.BF16) == torch.bfloat16 and from_torch(torch.float16) == DType.FP16
    assert dtype_eps(DType.FP16) == 1e-3 and dtype_eps(DType.BF16) == 1e-2 and dtype_eps(DType.FP32) == 1e-5
    assert dtype_to_string(DType.FP32) in ("FP32", "FLOAT")


def test_bf16_rounding_is_rne():
    # 1 + 2^-8 is exactly between two bf16 values -> ties to even (1.0); 1 + 3*2^-8 -> 1 + 2^-6
    x = torch.tensor([1.0 + 2 ** -8, 1.0 + 3 * 2 ** -8, 1.0 + 2 ** -7])
    y = x.to(torch.bfloat16).float()
    assert y.tolist() == [1.0, 1.0 + 2 ** -6, 1.0 + 2 ** -7]


def test_tconfig_typed_json_round_trip():
    c = TConfig("conv", "conv2d", in_channels=3, use_bias=True, eps=1e-5, act="relu", dims=[1, 2, 3],
                nested=TConfig("inner", "dense", n=4))
    j = c.to_json()
    assert j["name"] == "conv" and j["type"] == "conv2d"
    c2 = TConfig.from_json(j)
    assert c2 == c and c2["in_channels"] == 3 and c2["use_bias"] is True and c2.get("missing", 7) == 7
    assert isinstance(c2["eps"], float) and isinstance(c2["in_channels"], int)
    assert TConfig.from_json(c.dump()) == c


# ------------------------------------------------------------------------------- archiver
needs_native = pytest.mark.skipif(native.load() is None, reason="native extension not built")


@needs_native
def test_archiver_sizer_writer_reader_agree():
    Cx = native.require()
    blob = bytes(range(256)) * 64
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)

    def emit(a):
        a.u16(7)
        a.u64(1 << 40)
        a.f32(1.5)
        a.boolean(True)
        a.string("hello archive")
        a.sizes([3, 1, 4, 1, 5])
        return a

    s = emit(Cx.Sizer())
    s.raw(blob)
    s.tensor(int(DType.FP32), list(t.shape), t.numel() * 4)
    w = emit(Cx.Writer())
    w.raw(blob)
    w.tensor(int(DType.FP32), list(t.shape), t.data_ptr(), t.numel() * 4)
    data = w.getvalue()
    assert s.size() == w.size() == len(data)
    # exact layout: u16, u64, f32, u8, u64 len + bytes, u64 n + n*u64
    assert data[:2] == struct.pack("<H", 7) and data[2:10] == struct.pack("<Q", 1 << 40)
    assert data[15:23] == struct.pack("<Q", len("hello archive"))
    r = Cx.Reader(data, False)
    assert r.u16() == 7 and r.u64() == 1 << 40 and r.f32() == 1.5 and r.boolean() is True
    assert r.string() == b"hello archive" and r.sizes() == [3, 1, 4, 1, 5]
    assert r.raw(len(blob)) == blob
    assert r.u32() == int(DType.FP32) and r.sizes() == [2, 3, 4]
    out = torch.empty_like(t)
    r.raw_into(out.data_ptr(), t.numel() * 4, 4)
    assert torch.equal(out, t) and r.remaining() == 0


@needs_native
def test_archiver_endianness_swap():
    Cx = native.require()
    w = Cx.Writer()
    w.u32(0x01020304)
    w.f32(2.5)
    be = w.getvalue()[:4][::-1] + w.getvalue()[4:][::-1]       # the same values written big-endian
    r = Cx.Reader(be, True)
    assert r.u32() == 0x01020304 and r.f32() == 2.5
    assert Cx.host_is_little_endian() is True


def test_tensor_record_format_and_trailer():
    t = torch.randn(3, 5)
    f = io.BytesIO()
    tensor_io.save_tensor(f, t)
    raw = f.getvalue()
    # {u32 dtype, u64 ndims, u64 dims[], raw}  (include/tensor/tensor.hpp:585-606)
    assert struct.unpack("<I", raw[:4])[0] == int(DType.FP32)
    assert struct.unpack("<QQQ", raw[4:28]) == (2, 3, 5)
    assert len(raw) == 28 + 60
    f.seek(0)
    assert torch.equal(tensor_io.load_tensor(f), t)
    f = io.BytesIO()
    tensor_io.save_tensor(f, t.to(torch.bfloat16))
    tensor_io.write_trailer(f, {"step": 12, "m": torch.ones(4), "cfg": {"a": 1}})
    f.seek(0)
    assert tensor_io.load_tensor(f).dtype == torch.bfloat16
    ex = tensor_io.read_trailer(f)
    assert ex["step"] == 12 and torch.equal(ex["m"], torch.ones(4)) and ex["cfg"] == {"a": 1}
    dst = torch.zeros(3, 4)
    f = io.BytesIO()
    tensor_io.save_tensor(f, t)
    f.seek(0)
    with pytest.raises(Exception):
        tensor_io.load_into(f, dst)                              # shape mismatch is an error


# ------------------------------------------------------------------------------- allocators
@needs_native
def test_del_allocator_double_ended_and_pool_reuse():
    from tnn_b200.core.device import DELAllocator, DeviceManager, PoolAllocator, getHost
    a = DELAllocator("cpu", slab_bytes=1 << 20)
    t1 = a.allocate((256, 256), torch.float32)
    a.flip()
    t2 = a.allocate((128,), torch.float32)
    assert t1.data_ptr() % 256 == 0 and t2.data_ptr() % 256 == 0
    assert a.total_allocated() >= t1.numel() * 4 + t2.numel() * 4 and a.side() is True
    assert t2.data_ptr() > t1.data_ptr()                         # right side of the same slab
    a.free(t1)
    a.free(t2)
    a.clear()
    assert a.total_allocated() == 0 and a.peak_allocated() >= 256 * 256 * 4
    p = PoolAllocator("cpu")
    x = p.allocate((1000,), torch.float32)
    ptr = x.data_ptr()
    p.free(x)
    y = p.allocate((900,), torch.float32)                       # block <= 4x the request is reused
    assert y.data_ptr() == ptr and p.cached_bytes() == 0
    dm = DeviceManager.getInstance()
    assert dm.getDevice("CPU").type.name == "CPU" and getHost().total_memory() > 0


# ------------------------------------------------------------------------------- graph IR
def test_graph_builder_topo_sort_cycle_and_multi_input():
    b = GraphBuilder("g")
    n1 = b.add_layer(DenseLayer(4, 4, name="a"), "a")
    n2 = b.add_layer(MBroadcastLayer(2, "bc"), "bc")
    n3 = b.add_layer(ActivationLayer("tanh", "t"), "t")
    n4 = b.add_layer(AddLayer("add"), "add")
    # declared out of order on purpose
    b.add_edge(["u", "v2"], ["out"], n4)
    b.add_edge(["v"], ["v2"], n3)
    b.add_edge(["h"], ["u", "v"], n2)
    b.add_edge(["in"], ["h"], n1)
    g = b.compile("cpu", seed=3)
    assert [e.op for e in g.edges] == ["a", "bc", "t", "add"]
    ex = GraphExecutor(g)
    x = torch.randn(3, 4)
    out = ex.forward({"in": x})["out"]
    h = x @ n1.layer.params["weight"].t() + n1.layer.params["bias"]
    torch.testing.assert_close(out, h + torch.tanh(h))
    gin = ex.backward({"out": torch.ones_like(out)})["in"]
    assert gin.shape == x.shape
    bad = GraphBuilder("cyc")
    a = bad.add_layer(DenseLayer(2, 2, name="a"), "a")
    c = bad.add_layer(DenseLayer(2, 2, name="c"), "c")
    bad.add_edge(["x"], ["y"], a)
    bad.add_edge(["y"], ["x"], c)
    with pytest.raises(ValueError):
        bad.compile("cpu")


def test_graph_context_packs_params_into_aligned_slabs():
    g = load_or_create_model("mnist_cnn", device="cpu", seed=1)
    ctx = g.context()
    assert ctx.param_slab.numel() == ctx.grad_slab.numel() >= ctx.num_parameters()
    for layer, d, off, n in ctx.entries:
        assert off % 64 == 0                                     # 256-byte aligned sub-spans
        assert layer.params[d.name].data_ptr() == ctx.param_slab.data_ptr() + off * 4
        assert layer.grads[d.name].data_ptr() == ctx.grad_slab.data_ptr() + off * 4
    ctx.grad_slab.fill_(1.0)
    ctx.zero_grads()
    assert all(float(l.grads[d.name].abs().sum()) == 0 for l, d, _, _ in ctx.entries)


def test_checkpoint_round_trip_including_optimizer_state(tmp_path):
    from tnn_b200.nn.train import TrainStep, load_checkpoint, save_checkpoint
    g = load_or_create_model("mnist_cnn", device="cpu", seed=5)
    opt = OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4)
    sched = SchedulerFactory.create_step_lr(opt, 3, 0.5)
    step = TrainStep(g, LossFactory.create_logsoftmax_crossentropy(), opt, sched)
    x, y = torch.rand(8, 28, 28, 1), torch.randint(0, 10, (8,), dtype=torch.int32)
    for _ in range(4):
        step(x, y)
    path = str(tmp_path / "ckpt")
    save_checkpoint(g, path, opt, sched, 4)
    head = open(path, "rb").read(64)
    assert head.lstrip().startswith(b"{")                        # JSON graph config first (graph.hpp:119-126)
    opt2 = OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4)
    sched2 = SchedulerFactory.create_step_lr(opt2, 3, 0.5)
    g2, extra = load_checkpoint(path, "cpu", opt2, sched2)
    assert extra["step"] == 4 and opt2.step_count == opt.step_count
    assert abs(opt2.get_learning_rate() - opt.get_learning_rate()) < 1e-12
    for (n1, p1), (n2, p2) in zip(g.ctx.parameters(), g2.ctx.parameters()):
        assert n1 == n2 and torch.equal(p1, p2)
    # resumed optimisation continues bit-identically
    s2 = TrainStep(g2, LossFactory.create_logsoftmax_crossentropy(), opt2, sched2)
    a, b = step(x, y), s2(x, y)
    torch.testing.assert_close(a, b)
    torch.testing.assert_close(g.ctx.param_slab, g2.ctx.param_slab)


# ------------------------------------------------------------------------------- optimizers
def _toy_graph():
    from tnn_b200.nn import Sequential
    seq = Sequential([DenseLayer(6, 5, name="a"), ActivationLayer("tanh"), DenseLayer(5, 3, name="b")], "toy")
    return graph_from_sequential(seq, "cpu", seed=2, name="toy")


@pytest.mark.parametrize("kind", ["adam_coupled", "adamw", "sgd", "sgd_momentum"])
def test_optimizers_match_reference_formulas(kind):
    g = _toy_graph()
    ctx = g.context()
    if kind == "adam_coupled":
        opt = OptimizerFactory.create_adam(1e-2, 0.9, 0.999, 1e-3, 3e-4, False)
    elif kind == "adamw":
        opt = OptimizerFactory.create_adam(1e-2, 0.9, 0.999, 1e-8, 1e-2, True)
    elif kind == "sgd":
        opt = OptimizerFactory.create_sgd(0.1, 0.0)
    else:
        opt = OptimizerFactory.create_sgd(0.1, 0.9)
    opt.attach(ctx)
    n = ctx.numel
    p = ctx.param_slab[:n].clone()
    m = torch.zeros(n)
    v = torch.zeros(n)
    gen = torch.Generator().manual_seed(0)
    for t in range(1, 5):
        gr = torch.randn(n, generator=gen)
        ctx.grad_slab[:n].copy_(gr)
        opt.update()
        if kind.startswith("adam"):
            lr, wd = 1e-2, (3e-4 if kind == "adam_coupled" else 1e-2)
            eps = 1e-3 if kind == "adam_coupled" else 1e-8
            if kind == "adamw":
                p = p - wd * lr * p
            m = 0.9 * m + 0.1 * gr
            v = 0.999 * v + 0.001 * gr * gr
            upd = lr * (m / (1 - 0.9 ** t)) / ((v / (1 - 0.999 ** t)).sqrt() + eps)
            if kind == "adam_coupled":
                upd = upd + wd * lr * p
            p = p - upd
        elif kind == "sgd":
            p = p - 0.1 * gr
        else:
            m = 0.9 * m - 0.1 * gr          # v <- mu v - lr g ; p <- p + v  (sgd_kernels.cu:26-35)
            p = p + m
        torch.testing.assert_close(ctx.param_slab[:n], p, atol=1e-6, rtol=1e-5)
        assert float(ctx.grad_slab.abs().sum()) == 0.0 or not opt.fused_zero_grad
    cfg = opt.get_config()
    assert OptimizerFactory.create_from_config(cfg.to_json()).get_config() == cfg


def test_schedulers_follow_torch():
    import torch.optim.lr_scheduler as S

    def pair():
        o = OptimizerFactory.create_sgd(0.1)
        w = torch.nn.Parameter(torch.zeros(1))
        return o, torch.optim.SGD([w], lr=0.1)

    cases = [
        ("step_lr", dict(step_size=3, gamma=0.5), lambda to: S.StepLR(to, 3, 0.5)),
        ("multi_step_lr", dict(milestones=[2, 5], gamma=0.1), lambda to: S.MultiStepLR(to, [2, 5], 0.1)),
        ("exponential_lr", dict(gamma=0.9), lambda to: S.ExponentialLR(to, 0.9)),
        ("cosine_annealing_lr", dict(T_max=10, eta_min=0.01), lambda to: S.CosineAnnealingLR(to, 10, 0.01)),
        ("polynomial_lr", dict(total_steps=8, power=2.0), lambda to: S.PolynomialLR(to, 8, 2.0)),
    ]
    for name, kw, mk in cases:
        o, to = pair()
        try:
            s = SchedulerFactory.create(name, o, **kw)
        except KeyError:
            s = SchedulerFactory.create(name.replace("_lr", ""), o, **kw)
        ts = mk(to)
        for _ in range(9):
            s.step()
            to.step()
            ts.step()
            assert abs(o.get_learning_rate() - to.param_groups[0]["lr"]) < 1e-9, name
        cfg = s.get_config()
        s2 = SchedulerFactory.create_from_config(cfg.to_json() if hasattr(cfg, "to_json") else cfg, OptimizerFactory.create_sgd(0.1))
        assert type(s2) is type(s)
    # warm-up + cosine and plateau: qualitative checks
    o = OptimizerFactory.create_sgd(1.0)
    s = SchedulerFactory.create_warmup_cosine(o, 4, 20, 0.0, 0.0)
    lrs = []
    for _ in range(20):
        s.step()
        lrs.append(o.get_learning_rate())
    assert lrs[0] < lrs[3] <= 1.0 and lrs[-1] < 0.05 and max(lrs) <= 1.0 + 1e-9
    o = OptimizerFactory.create_sgd(1.0)
    s = SchedulerFactory.create("reduce_lr_on_plateau", o, factor=0.5, patience=1)
    for metric in (1.0, 1.0, 1.0, 1.0):
        s.step(metric)
    assert o.get_learning_rate() < 1.0
    sd = s.state_dict()
    s.load_state_dict(sd)


# ------------------------------------------------------------------------------- losses
def test_losses_against_torch():
    import torch.nn.functional as F
    logits = torch.randn(7, 5)
    y = torch.randint(0, 5, (7,), dtype=torch.int32)
    ce = LossFactory.create("logsoftmax_crossentropy")
    stats, grad = ce.loss_and_grad(logits, y)
    ref = F.cross_entropy(logits, y.long())
    assert abs(float(stats[0]) - float(ref)) < 1e-6
    lg = logits.clone().requires_grad_(True)
    F.cross_entropy(lg, y.long()).backward()
    torch.testing.assert_close(grad, lg.grad, atol=1e-6, rtol=1e-5)
    assert int(stats[1]) == int((logits.argmax(-1) == y).sum()) == compute_class_corrects(logits, y)
    for name, fn in (("mse", F.mse_loss), ("mae", F.l1_loss), ("huber", lambda a, b: F.huber_loss(a, b, delta=1.0))):
        l = LossFactory.create(name)
        p, t = torch.randn(4, 3), torch.randn(4, 3)
        pp = p.clone().requires_grad_(True)
        r = fn(pp, t)
        r.backward()
        assert abs(l.compute_loss(p, t) - float(r)) < 1e-6
        torch.testing.assert_close(l.compute_gradient(p, t), pp.grad, atol=1e-6, rtol=1e-5)
    for alias in ("ce", "crossentropy", "mean_squared_error", "mean_absolute_error", "logsoftmax_ce"):
        assert LossFactory.create(alias) is not None


# ------------------------------------------------------------------------------- data path
@needs_native
def test_cifar_binary_loader_and_augmentation(tmp_path):
    from tnn_b200.data import AugmentationBuilder, CIFAR10DataLoader, DataLoaderFactory
    rng = np.random.default_rng(0)
    n = 50
    labels = rng.integers(0, 10, n, dtype=np.uint8)
    pix = rng.integers(0, 256, (n, 3, 32, 32), dtype=np.uint8)
    rec = np.concatenate([labels[:, None], pix.reshape(n, -1)], 1)
    os.makedirs(tmp_path / "cifar-10-batches-bin", exist_ok=True)
    for i in range(1, 6):
        rec.tofile(tmp_path / "cifar-10-batches-bin" / f"data_batch_{i}.bin")
    rec.tofile(tmp_path / "cifar-10-batches-bin" / "test_batch.bin")
    ld = CIFAR10DataLoader()
    ld.load_data(str(tmp_path), train=False)
    assert ld.size() == n and tuple(ld.get_data_shape()) == (32, 32, 3)
    x, y = ld.get_batch(16)
    assert x.shape == (16, 32, 32, 3) and y.dtype == torch.int32
    # CHW -> NHWC transpose and /255 (cifar10_data_loader.hpp:152-165)
    torch.testing.assert_close(x[0], torch.from_numpy(pix[0]).permute(1, 2, 0).float() / 255)
    assert y.tolist() == labels[:16].tolist()
    seen = 16
    while True:
        b = ld.get_batch(16)
        if b is None:
            break
        seen += b[0].shape[0]
    assert seen == n
    ld.set_seed(3)
    ld.shuffle()
    ld.reset()
    x2, y2 = ld.get_batch(n)
    assert sorted(y2.tolist()) == sorted(labels.tolist()) and y2.tolist() != labels.tolist()
    aug = AugmentationBuilder(seed=1).horizontal_flip(1.0).build()
    ld.set_augmentation(aug)
    ld.reset()
    tr, va = DataLoaderFactory.create("cifar10", str(tmp_path))
    assert tr.size() == 5 * n and va.size() == n
    out = aug.apply(x.clone())
    torch.testing.assert_close(out, x.flip(2))
    full = (AugmentationBuilder(seed=2).random_crop(1.0, 4).cutout(1.0, 8).brightness(1.0, 0.2).contrast(1.0, 0.2)
            .gaussian_noise(1.0, 0.01).rotation(1.0, 10).vertical_flip(0.5).normalize([0.5] * 3, [0.25] * 3).build())
    z = full.apply_python(x.clone(), full=True)
    assert z.shape == x.shape and torch.isfinite(z).all()


@needs_native
def test_token_dataset_and_tokenizer(tmp_path):
    from tnn_b200.data import OpenWebTextDataLoader, Tokenizer
    toks = (np.arange(5000) % 997).astype(np.uint16)
    toks.tofile(tmp_path / "train.bin")
    toks[:2000].tofile(tmp_path / "val.bin")
    ld = OpenWebTextDataLoader(seq_len=32)
    ld.load_data(str(tmp_path), train=True)
    x, y = ld.get_batch(4)
    assert x.shape == (4, 32) and y.shape == (4, 32)
    assert torch.equal(x[:, 1:].long(), y[:, :-1].long())        # next-token labels, integer class ids
    w = ld.window(10, 8)
    assert w.reshape(-1).tolist() == list(range(10, 18))
    tok = Tokenizer.byte_fallback([b"he", b"llo", b" wor", b"ld"])
    ids = tok.encode("hello world")
    assert tok.decode(ids) == "hello world" and len(ids) < len("hello world")
    path = str(tmp_path / "vocab.bin")
    tok.save(path)
    raw = open(path, "rb").read()
    assert struct.unpack("<I", raw[:4])[0] == tok.vocab_size()   # u32 n; {u32 len; bytes} x n
    t2 = Tokenizer(path)
    assert t2.decode(ids) == "hello world" and t2.vocab_size() == tok.vocab_size()


@needs_native
def test_profiler_merge_rebases_and_serialises():
    from tnn_b200.utils.profiler import EventType, Profiler
    a, b = Profiler(), Profiler()
    a.init_start_time()
    b.init_start_time()
    t = Profiler.now_ns()
    a.add_event(EventType.COMPUTE, t, t + 1000, "fwd", "stage0")
    b.add_event(EventType.COMMUNICATION, t + 500, t + 700, "send", "stage1")
    b2 = Profiler.deserialize(b.serialize())
    a.merge(b2)
    a.sort()
    ev = a.events()
    assert len(ev) == 2 and {e["name"] for e in ev} == {"fwd", "send"}
    with a.span("host-span"):
        pass
    assert len(a.events()) == 3


def test_env_file_and_training_config(tmp_path):
    from tnn_b200.nn.train import TrainingConfig
    from tnn_b200.utils.env import Env
    p = tmp_path / ".env"
    p.write_text("EPOCHS=3\nBATCH_SIZE = 64\n# comment\nMODEL_NAME=\"cifar100_wrn16_8\"\nLR_INITIAL=0.01\nDEVICE_TYPE=CPU\n")
    e = Env(str(p))
    assert e.get("EPOCHS", 1, int) == 3 and e.get("MISSING", 5, int) == 5
    c = TrainingConfig.load_from_env(str(p))
    assert c.epochs == 3 and c.batch_size == 64 and c.model_name == "cifar100_wrn16_8" and c.lr_initial == 0.01
    assert c.device() == torch.device("cpu")
    import json
    j = tmp_path / "cfg.json"
    j.write_text(json.dumps({"epochs": 2, "batch_size": 16, "model_name": "mnist_cnn"}))
    c2 = TrainingConfig.load_from_json(str(j))
    assert c2.epochs == 2 and c2.model_name == "mnist_cnn"


# ------------------------------------------------------------------------------- model zoo + trainer
def test_every_zoo_model_builds_and_reports_shape():
    names = ExampleModels.available_models()
    for required in ["mnist_cnn", "cifar10_vgg", "cifar10_resnet9", "cifar100_resnet18", "cifar100_wrn16_8",
                     "tiny_imagenet_resnet18", "tiny_imagenet_wrn16_8", "tiny_imagenet_resnet50", "tiny_imagenet_vit",
                     "tiny_imagenet_flash_vit", "imagenet_resnet50", "gpt2_small", "gpt2_medium", "gpt2_large",
                     "flash_gpt2_small", "flash_gpt2_medium", "flash_gpt2_large"]:
        assert required in names, required
    shapes = {"mnist_cnn": ((2, 28, 28, 1), (2, 10)), "cifar10_resnet9": ((2, 32, 32, 3), (2, 10)),
              "cifar100_wrn16_8": ((2, 32, 32, 3), (2, 100)), "cifar10_vgg": ((2, 32, 32, 3), (2, 10)),
              "cifar100_resnet18": ((2, 32, 32, 3), (2, 100))}
    for n, (i, o) in shapes.items():
        m = ExampleModels.create(n)
        assert tuple(m.output_shape(i)) == o, n
        assert m.forward_flops(i) > 0
    wrn = ExampleModels.create("cifar100_wrn16_8")
    nparams = sum(math.prod(d.shape) for l in wrn.all_layers() for d in l.param_descriptors() if d.trainable)
    assert 10.5e6 < nparams < 11.5e6                              # ~11.0 M
    g = ExampleModels.create("gpt2_small", seq_len=16, vocab=64)
    assert tuple(g.output_shape((2, 16))) == (2, 16, 64)


def test_small_gpt_trains_on_cpu():
    from tnn_b200.nn.train import TrainStep
    seq = ExampleModels.create("gpt2_small", seq_len=16, vocab=48)
    # shrink: keep 2 blocks so the CPU test stays fast
    g = graph_from_sequential(seq, "cpu", seed=3, name="tiny_gpt")
    opt = OptimizerFactory.create_adam(3e-4)
    step = TrainStep(g, LossFactory.create_logsoftmax_crossentropy(), opt)
    gen = torch.Generator().manual_seed(0)
    x = torch.randint(0, 48, (2, 16), generator=gen, dtype=torch.int32)
    y = torch.roll(x, -1, 1)
    losses = [float(step(x, y)[0]) for _ in range(6)]
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0]


def test_mnist_cnn_trainer_cpu_world1(tmp_path):
    """BASELINE config 1: mnist_cnn Sequential fp32 on CPU through train_model (plumbing)."""
    from tnn_b200.data import SyntheticDataLoader
    from tnn_b200.nn.train import TrainingConfig, train_model, validate_model
    torch.manual_seed(0)
    g = load_or_create_model("mnist_cnn", device="cpu", seed=1)
    train = SyntheticDataLoader((28, 28, 1), 10, n=256, seed=1)
    val = SyntheticDataLoader((28, 28, 1), 10, n=64, seed=1)
    train.load_data()
    val.load_data()
    cfg = TrainingConfig(epochs=3, batch_size=32, progress_print_interval=0, snapshot_dir=str(tmp_path), device_type="CPU")
    opt = OptimizerFactory.create_adam(2e-3, 0.9, 0.999, 1e-3, 0.0)
    crit = LossFactory.create_logsoftmax_crossentropy()
    sched = SchedulerFactory.create_step_lr(opt, 1000, 0.1)
    lines = []
    hist = train_model(g, train, val, opt, crit, sched, cfg, log=lines.append)
    assert len(hist["train_loss"]) == 3 and hist["train_loss"][-1] < hist["train_loss"][0]
    assert os.path.exists(os.path.join(str(tmp_path), g.name))   # best-validation snapshot
    vl, va = validate_model(g, val, crit, cfg)
    assert math.isfinite(vl) and 0.0 <= va <= 1.0
    g2 = Graph.load_state(os.path.join(str(tmp_path), g.name), "cpu")
    assert [n for n, _ in g2.ctx.parameters()] == [n for n, _ in g.ctx.parameters()]


@needs_native
def test_native_cpu_gemm_and_thread_pool():
    from tnn_b200.utils.threading import ThreadWrapper, cpu_gemm, parallel_for, parallel_for_2d
    a, b = torch.randn(130, 70), torch.randn(70, 45)
    ref = a @ b
    for ta, tb, aa, bb in ((False, False, a, b), (True, False, a.t().contiguous(), b), (False, True, a, b.t().contiguous()),
                           (True, True, a.t().contiguous(), b.t().contiguous())):
        torch.testing.assert_close(cpu_gemm(aa, bb, ta, tb), ref, atol=1e-4, rtol=1e-4)
    c = torch.ones(130, 45)
    torch.testing.assert_close(cpu_gemm(a, b, alpha=2.0, beta=0.5, out=c), 2 * ref + 0.5, atol=1e-4, rtol=1e-4)   # beta accumulate
    # shapes that cross the 96-row / 256-deep / 16-column blocking and leave ragged 6x16 edge tiles
    for M, N, K in ((1, 1, 1), (5, 17, 3), (97, 33, 257), (200, 130, 600)):
        x, y = torch.randn(M, K), torch.randn(K, N)
        torch.testing.assert_close(cpu_gemm(x, y), x @ y, atol=2e-4, rtol=1e-4)
        torch.testing.assert_close(cpu_gemm(x, y, threads=1), cpu_gemm(x, y, threads=4))       # same result on any pool size
    hits = [0] * 11
    parallel_for(11, lambda s, e: [hits.__setitem__(i, hits[i] + 1) for i in range(s, e)], grain=3)
    assert hits == [1] * 11
    tiles = []
    parallel_for_2d(70, 40, lambda r0, r1, c0, c1: tiles.append((r1 - r0) * (c1 - c0)))
    assert sum(tiles) == 70 * 40
    assert ThreadWrapper(2).execute(torch.get_num_threads) == 2


def test_legacy_matrix_value_semantics():
    """matrix.hpp surface: element-wise +,-, scalar *,/, product, clone/reshape/resize, fills."""
    import pytest
    import torch
    from tnn_b200.core import Matrix
    a = Matrix(3, 4).fill(2.0)
    b = Matrix(3, 4, torch.arange(12, dtype=torch.float32))
    c = a + b
    assert torch.equal(c.data(), b.data() + 2)
    c -= a
    assert torch.equal(c.data(), b.data())
    assert torch.allclose((b * 0.5).data(), b.data() * 0.5) and torch.allclose((b / 4).data(), b.data() / 4)
    with pytest.raises(ZeroDivisionError):
        b / 0
    with pytest.raises(ValueError):
        a + Matrix(4, 3)
    w = Matrix(4, 2).fill_random_normal(0.0, 1.0, seed=5)
    assert torch.allclose((b * w).data(), b.data() @ w.data(), atol=1e-5)
    assert (b @ w).rows() == 3 and (b @ w).cols() == 2
    with pytest.raises(ValueError):
        a * Matrix(3, 4)
    d = b.clone()
    d[0, 0] = 100.0
    assert b[0, 0] == 0.0                                     # clone owns its storage
    assert b.reshape(2, 6).data().shape == (2, 6)
    assert torch.equal(b.transpose().data(), b.data().t())
    d.resize(5, 5)
    assert d.size() == 25 and float(d.data().abs().sum()) == 0.0
    u = Matrix(64, 64).fill_random_uniform(0.25, seed=9)
    assert float(u.data().abs().max()) <= 0.25 and float(u.data().std()) > 0.05


def test_step_state_ring_gives_every_step_its_own_source_row():
    """The async H2D copy of the step state must not share its pinned source with later steps."""
    import torch
    from tnn_b200.core import stepstate
    st = stepstate.StepState(torch.device("cpu"))

    class _Opt:
        step_count, lr, beta1, beta2, epsilon, weight_decay, decoupled, grad_scale, momentum = 0, 0.1, 0.9, 0.999, 1e-8, 0.0, False, 1.0, 0.0
    rows = set()
    for i in range(5):
        _Opt.step_count = i
        st.advance(_Opt)
        rows.add(st.host.data_ptr())
        assert int(st.host[12:14].view(torch.int64)[0]) == i + 1            # seed counter
        assert abs(float(st.dev[8]) - (1.0 - 0.9 ** (i + 1))) < 1e-6        # Adam bias correction of THIS step
    assert len(rows) == 5
    st.advance(None)                                                        # scalars persist when not refreshed
    assert abs(float(st.dev[0]) - 0.1) < 1e-7 and int(st.host[12:14].view(torch.int64)[0]) == 6


def test_side_stream_bookkeeping_is_a_noop_without_pending_work():
    from tnn_b200.core import side_stream
    side_stream.enter()
    side_stream.enter()
    side_stream.exit_()
    side_stream.exit_()
    side_stream.exit_()          # unbalanced exit must not underflow the depth counter
    assert side_stream._depth == 0
    from tnn_b200.ops import gemm
    assert gemm._STATIC_WEIGHTS is False
    with gemm.static_weights():
        assert gemm._STATIC_WEIGHTS is True
        with gemm.static_weights():
            pass
        assert gemm._STATIC_WEIGHTS is True
    assert gemm._STATIC_WEIGHTS is False


def test_logger_step_log_and_hardware_info(tmp_path):
    """utils: console+file logger, JSON-lines step log, hardware / memory / affinity probes
    (reference: include/logging/logger.hpp, include/utils/hardware_info.hpp, thread_affinity.hpp)."""
    import json
    from tnn_b200.utils.hardware_info import HardwareInfo, ThreadAffinity, get_memory_usage_kb, get_shape_hash
    from tnn_b200.utils.logger import GlobalLogger, Logger, StepLog
    lf = tmp_path / "run.log"
    log = Logger("unit", str(lf), "debug")
    log.info("hello", 1)
    log.debug("dbg")
    log.warn("careful")
    assert GlobalLogger.get() is GlobalLogger.get()
    text = lf.read_text()
    assert "hello" in text and "careful" in text and "dbg" in text
    sl = StepLog(str(tmp_path / "steps.jsonl"))
    sl.write(step=1, loss=2.5)
    sl.write(step=2, loss=2.0)
    sl.close()
    rows = [json.loads(l) for l in (tmp_path / "steps.jsonl").read_text().splitlines()]
    assert [r["step"] for r in rows] == [1, 2] and rows[1]["loss"] == 2.0
    hw = HardwareInfo()
    s = hw.summary()
    assert "CPU" in s and "RAM" in s
    assert get_memory_usage_kb() > 1000
    assert get_shape_hash((2, 3, 4)) == get_shape_hash([2, 3, 4]) != get_shape_hash((4, 3, 2))
    cur = ThreadAffinity.current()
    assert len(cur) >= 1
    one = sorted(cur)[:1]
    assert ThreadAffinity.pin(one) and ThreadAffinity.current() == set(one)
    ThreadAffinity.pin(sorted(cur))


def test_native_simd_dot_product():
    """AVX2/FMA dot product of the C++ runtime (reference: the hand-written assembly in include/utils/simd_asm.hpp)."""
    import torch
    from tnn_b200.utils.threading import cpu_dot
    g = torch.Generator().manual_seed(0)
    for n in (1, 7, 8, 31, 32, 33, 257, 4096, 100003):
        a, b = torch.randn(n, generator=g), torch.randn(n, generator=g)
        ref = float((a.double() * b.double()).sum())
        assert abs(cpu_dot(a, b) - ref) <= 1e-4 * max(1.0, abs(ref)) + 1e-3 * (n ** 0.5) * 1e-2


def test_misc_helpers():
    """misc.hpp / parser.hpp counterparts."""
    import enum

    import pytest
    from tnn_b200.utils import misc

    class Mode(enum.Enum):
        SYNC = 0
        SEMI_ASYNC = 1
    assert misc.from_str(" 12 ", int) == 12 and misc.from_str("1e-3", float) == 1e-3
    assert misc.from_str("Yes", bool) is True and misc.from_str("off", bool) is False
    with pytest.raises(ValueError):
        misc.from_str("maybe", bool)
    assert misc.from_str("semi_async", Mode) is Mode.SEMI_ASYNC and misc.from_str("0", Mode) is Mode.SYNC
    assert misc.get_enum_vector(Mode) == [Mode.SYNC, Mode.SEMI_ASYNC]
    h1 = misc.hash_combine(misc.hash_combine(0, 3), 5)
    assert h1 == misc.hash_combine(misc.hash_combine(0, 3), 5) != misc.hash_combine(misc.hash_combine(0, 5), 3)
    calls = []
    ms = misc.benchmark(lambda: calls.append(1), iters=5, warmup=2)
    assert len(calls) == 7 and ms >= 0.0


def test_image_folder_loaders_decode_jpeg_trees_and_cache(tmp_path):
    """TinyImageNet / ImageNet-100 layouts: lazy threaded JPEG decoding, labels from the directory structure /
    annotation file, optional one-off decoded cache (tiny_imagenet_data_loader.hpp, imagenet100_data_loader.hpp)."""
    import numpy as np
    import torch
    from PIL import Image
    from tnn_b200.data import DataLoaderFactory, ImageNet100DataLoader, TinyImageNetDataLoader
    rng = np.random.default_rng(0)

    def jpeg(path, value, hw=(16, 16)):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        arr = np.full((*hw, 3), value, dtype=np.uint8)
        Image.fromarray(arr).save(path, quality=95)

    import os
    # ---- TinyImageNet layout
    root = tmp_path / "tiny"
    wnids = ["n01", "n02", "n03"]
    os.makedirs(root, exist_ok=True)
    (root / "wnids.txt").write_text("\n".join(wnids) + "\n")
    for ci, w in enumerate(wnids):
        for k in range(4):
            jpeg(str(root / "train" / w / "images" / f"{w}_{k}.JPEG"), 40 + 80 * ci)
    ann = []
    for k in range(6):
        jpeg(str(root / "val" / "images" / f"val_{k}.JPEG"), 40 + 80 * (k % 3))
        ann.append(f"val_{k}.JPEG\t{wnids[k % 3]}\t0\t0\t1\t1")
    (root / "val" / "val_annotations.txt").write_text("\n".join(ann) + "\n")
    tr, va = DataLoaderFactory.create("tiny_imagenet", str(root))
    assert tr.size() == 12 and va.size() == 6 and tr.get_data_shape() == (64, 64, 3)
    x, y = tr.get_batch(12)
    assert x.shape == (12, 64, 64, 3) and x.dtype == torch.float32 and y.dtype == torch.int32
    for i in range(12):                                   # the pixel level encodes the class
        assert abs(float(x[i].mean()) * 255 - (40 + 80 * int(y[i]))) < 3.0
    xv, yv = va.get_batch(6)
    assert sorted(yv.tolist()) == [0, 0, 1, 1, 2, 2]
    # decoded cache: same content through the memory-mapped path
    cached = TinyImageNetDataLoader().cache(str(root), True)
    assert cached.x is not None and cached.size() == 12
    xc, yc = cached.get_batch(12)
    assert torch.equal(yc.sort().values, y.sort().values) and abs(float(xc.mean()) - float(x.mean())) < 1e-3
    # ---- class-per-directory layout (ImageNet-100), resized to the requested edge
    root2 = tmp_path / "in100"
    for split, n in (("train", 3), ("val", 2)):
        for ci, cname in enumerate(["cat", "dog"]):
            for k in range(n):
                jpeg(str(root2 / split / cname / f"{k}.jpg"), 60 + 100 * ci, hw=(20, 30))
    ld = ImageNet100DataLoader(size=32).load_data(str(root2), True)
    x2, y2 = ld.get_batch(6)
    assert x2.shape == (6, 32, 32, 3) and sorted(y2.tolist()) == [0, 0, 0, 1, 1, 1] and ld.class_names == ["cat", "dog"]


def test_wifi_loader_columns_normalisation_and_modes(tmp_path):
    """wifi_data_loader.hpp: column ranges, header handling, z-score normalisation + de-normalisation,
    regression and classification targets."""
    import numpy as np
    import torch
    from tnn_b200.data import WiFiDataLoader
    rng = np.random.default_rng(1)
    feats = rng.normal(-70, 10, size=(50, 6)).astype(np.float32)
    xy = rng.normal(100, 20, size=(50, 2)).astype(np.float32)
    floor = rng.integers(0, 4, size=(50, 1)).astype(np.float32)
    table = np.concatenate([feats, xy, floor], 1)
    f = tmp_path / "wifi.csv"
    np.savetxt(f, table, delimiter=",", header=",".join(f"c{i}" for i in range(9)), comments="")
    reg = WiFiDataLoader(True).load_data(str(f), 0, 6, 6, 8, True)
    assert reg.size() == 50 and reg.num_features() == 6 and reg.num_outputs() == 2
    raw_y = reg.y.clone()
    reg.normalize_data()
    assert abs(float(reg.x.mean())) < 1e-4 and abs(float(reg.x.std(unbiased=False)) - 1.0) < 5e-2
    assert torch.allclose(reg.denormalize_targets(reg.y), raw_y, atol=1e-3)
    xb, yb = reg.get_batch(16)
    assert xb.shape == (16, 6) and yb.shape == (16, 2)
    cls = WiFiDataLoader(False).load_data(str(f), 0, 6, 8, 9, True)
    assert cls.y.dtype == torch.int32 and cls.num_outputs() == int(floor.max()) + 1
    assert torch.equal(cls.y, torch.from_numpy(floor[:, 0].astype(np.int32)))


def test_concurrent_map_semantics_and_thread_safety():
    """concurrent_map.hpp: emplace / try_get / erase / contains / operator[] / size under concurrent writers."""
    import threading
    from tnn_b200.utils.concurrent import ConcurrentMap
    m = ConcurrentMap(default_factory=int)
    assert m.emplace("a", 1) and not m.emplace("a", 2)            # emplace keeps the first value
    assert m.try_get("a") == (True, 1) and m.try_get("zz", -1) == (False, -1)
    assert m["fresh"] == 0 and m.contains("fresh") and "fresh" in m and m.size() == 2
    assert m.erase("fresh") and not m.erase("fresh") and len(m) == 1

    def work(tid):
        for i in range(2000):
            m.update_with("ctr", lambda v: (v or 0) + 1)
            m.emplace((tid, i % 50), i)
    ts = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert m["ctr"] == 8 * 2000 and m.size() == 2 + 8 * 50
    strict = ConcurrentMap()
    try:
        strict["missing"]
        assert False
    except KeyError:
        pass


def test_tasks_on_host_threads_and_flows():
    """task.hpp: ready() / sync() with an error status, CPU tasks on worker threads, flow-bound tasks."""
    import time
    from tnn_b200.core.device import DeviceManager, create_cpu_task, create_cuda_task
    t = create_cpu_task(lambda a, b: (time.sleep(0.05), a + b)[1], 2, 3)
    assert t.sync() is None and t.ready() and t.result() == 5
    bad = create_cpu_task(lambda: 1 / 0)
    err = bad.sync()
    assert isinstance(err, ZeroDivisionError) and bad.ready()
    flow = DeviceManager.getInstance().getHost().getFlow()
    ft = create_cuda_task(flow, lambda x: x * 2, 21)          # a host flow runs the work inline
    assert ft.ready() and ft.sync() is None and ft.result() == 42
    assert isinstance(create_cuda_task(flow, lambda: [][1]).sync(), IndexError)


def test_native_planner_matches_python_oracles():
    """csrc/runtime/planner.hpp: graph order, slab layout, buckets, pipeline schedules and partitions, each checked
    against the straightforward Python twin that stays in the tree as the oracle."""
    import random
    from tnn_b200.nn.graph import ALIGN_ELEMS, _plan_slab
    from tnn_b200.ops import native
    from tnn_b200.parallel.partitioner import (NaiveDataPartitioner, NaivePipelinePartitioner, SeqPartition,
                                               _min_bottleneck_split, _min_bottleneck_split_py)
    C = native.require()
    rnd = random.Random(7)
    # --- topological order on random layered DAGs declared in shuffled order
    for trial in range(20):
        b = GraphBuilder(f"g{trial}")
        n_ops = rnd.randint(1, 12)
        made = ["in"]
        spec = []
        for i in range(n_ops):
            k = rnd.randint(1, min(2, len(made)))
            ins = rnd.sample(made, k)
            out = f"t{i}"
            made.append(out)
            spec.append((ins, out, k))
        rnd.shuffle(spec)
        for ins, out, k in spec:
            node = b.add_layer(AddLayer(f"add_{out}") if k > 1 else ActivationLayer("tanh", f"act_{out}"), f"op_{out}")
            b.add_edge(ins, [out], node)
        fast, slow = b.topo_sort(), b._topo_sort_py()
        assert [e.op for e in fast] == [e.op for e in slow]
        pos = {e.outputs[0]: i for i, e in enumerate(fast)}
        assert all(pos[i] < pos[e.outputs[0]] for e in fast for i in e.inputs if i in pos)
    with pytest.raises(ValueError, match="cycle"):
        C.plan_topo_order([[1], [0]], [[0], [1]])
    two = GraphBuilder("dup")
    two.add_edge(["a"], ["x"], two.add_layer(ActivationLayer("relu", "r1"), "r1"))
    two.add_edge(["b"], ["x"], two.add_layer(ActivationLayer("relu", "r2"), "r2"))
    with pytest.raises(ValueError, match="io node x has two producers"):
        two.topo_sort()
    # --- slab layout and buckets
    sizes = [rnd.randint(1, 5000) for _ in range(50)]
    off = _plan_slab(sizes)
    cur = 0
    for n, o in zip(sizes, off):
        assert o == cur and o % ALIGN_ELEMS == 0
        cur += (n + ALIGN_ELEMS - 1) // ALIGN_ELEMS * ALIGN_ELEMS
    assert off[-1] == cur and _plan_slab([]) == [0]
    buckets = C.plan_buckets(off, 20000)
    assert buckets[0][1] == off[-1] and buckets[-1][0] == 0                 # last-produced gradients first, full cover
    assert all(a[0] == b[1] for a, b in zip(buckets, buckets[1:])) and all(lo in off for lo, _ in buckets)
    assert all(hi - lo >= 20000 for lo, hi in buckets[:-1])
    assert C.plan_buckets(off, 0) == [(0, off[-1])]
    # --- schedules: what parallel/pipeline.PipelineEngine.train_batch walks (warm-up, 1F1B, cool-down)
    for S in (1, 2, 4, 8):
        for M in (1, 2, 3, 8):
            for r in range(S):
                sch = C.plan_schedule_1f1b(S, r, M)
                warm = min(S - r - 1, M)
                assert [m for o, m in sch if o == "F"] == list(range(M)) == [m for o, m in sch if o == "B"]
                assert [o for o, _ in sch[:warm]] == ["F"] * warm and [o for o, _ in sch[len(sch) - warm:]] == ["B"] * warm
                assert C.plan_peak_in_flight(sch) == min(S - r, M)
                done = set()
                for o, m in sch:                                           # a backward never precedes its forward
                    assert o == "F" or m in done
                    done.add(m) if o == "F" else None
    assert C.plan_peak_in_flight(C.plan_schedule_gpipe(6)) == 6
    # --- partitions
    layers = [DenseLayer(4, 4, name=f"l{i}") for i in range(11)]
    assert [(p.start_offset, p.length) for p in NaivePipelinePartitioner([1, 2, 1]).partition_model(layers)] == [(0, 2), (2, 5), (7, 4)]
    assert [(p.start_offset, p.length) for p in NaiveDataPartitioner([1, 3]).partition_input(torch.zeros(10, 2))] == [(0, 2), (2, 8)]
    with pytest.raises(ValueError):
        NaivePipelinePartitioner([1, 1, 1]).partition_model(layers[:2])
    for _ in range(30):
        L, n = rnd.randint(1, 14), rnd.randint(1, 5)
        cost = [rnd.uniform(0.1, 10) for _ in range(L)]
        bnd = [rnd.uniform(0, 3) for _ in range(L)] if rnd.random() < 0.5 else None
        a, bpy = _min_bottleneck_split(cost, n, bnd), _min_bottleneck_split_py(cost, n, bnd)
        assert a == bpy and all(isinstance(p, SeqPartition) for p in a)
        assert a[0].start_offset == 0 and a[-1].end == L and all(x.end == y.start_offset for x, y in zip(a, a[1:]))


def test_layer_profiler_and_trainer_switches(capsys):
    """PROFILER_TYPE / PRINT_LAYER_PROFILING / PRINT_LAYER_MEMORY_USAGE: per-layer fwd/bwd table, restored on detach."""
    from tnn_b200.data import SyntheticDataLoader
    from tnn_b200.nn.train import TrainingConfig, train_model
    from tnn_b200.utils.profiler import LayerProfiler
    g = load_or_create_model("mnist_cnn", device="cpu", seed=2)
    model = g.get_layers()[0]
    conv1 = model.layers[0]
    prof = LayerProfiler("CUMULATIVE").attach(model)
    assert "forward_impl" in conv1.__dict__
    g.set_training(True)
    x = torch.rand(4, 28, 28, 1)
    for _ in range(2):
        y = model.forward(x, 0)
        model.backward(torch.ones_like(y), 0)
    rows = {r[0]: r for r in prof.rows()}
    assert rows["mnist_cnn/conv1"][3] == 2 and rows["mnist_cnn/conv1"][4] == 2 and rows["mnist_cnn/conv1"][1] > 0
    txt = prof.report()
    assert "mnist_cnn/output" in txt and "share" in txt and prof.rows()[0][3] == 2         # cumulative keeps the counts
    assert "params KiB" in prof.memory_report()
    prof.detach()
    assert "forward_impl" not in conv1.__dict__ and "backward_impl" not in conv1.__dict__
    cfg = TrainingConfig(epochs=1, batch_size=16, progress_print_interval=0, device_type="CPU", profiler_type="NORMAL",
                         print_layer_memory_usage=True)
    train = SyntheticDataLoader((28, 28, 1), 10, 32)
    train.load_data()
    import tempfile
    cfg.snapshot_dir = tempfile.mkdtemp()
    lines = []
    hist = train_model(g, train, None, OptimizerFactory.create_sgd(0.01), LossFactory.create_logsoftmax_crossentropy(), None, cfg,
                       log=lines.append)
    out = "\n".join(lines)
    assert "layer profile, epoch 1 (normal)" in out and "layer memory, epoch 1" in out and "mnist_cnn/bn1" in out
    assert hist["layer_profile"] and "forward_impl" not in conv1.__dict__


def test_env_static_and_instance_lookups(tmp_path, monkeypatch):
    """env.hpp: Env::get<T>(key, default) is static upstream; both spellings work here and the default's type decides."""
    from tnn_b200.utils.env import Env
    f = tmp_path / ".env"
    f.write_text(chr(10).join(["# comment", "EPOCHS=7", "NAME='wrn'", "FLAG=true", "LR=0.5", ""]))
    e = Env(str(f))
    assert e.get("EPOCHS", 1) == 7 and e.get("LR", 0.1) == 0.5 and e.get("NAME", "x") == "wrn" and e.get("FLAG", False) is True
    assert e.get("MISSING", 3) == 3 and e.get("EPOCHS", typ=str) == "7" and e.has("NAME") and not e.has("NOPE_XYZ")
    monkeypatch.setenv("EPOCHS", "9")                       # the process environment wins over the file
    assert e.get("EPOCHS", 1) == 9
    monkeypatch.chdir(tmp_path)
    Env.reload(".env")
    assert Env.get("NAME", "") == "wrn" and Env.get("EPOCHS", 0) == 9 and Env.has("LR")
    monkeypatch.setenv("ONLY_IN_PROCESS", "12")
    assert Env.get("ONLY_IN_PROCESS", 0) == 12
    Env.reload("/nonexistent/.env")


def test_checkpoint_header_is_the_reference_graph_schema(tmp_path):
    """Graph::save_state writes GraphConfig as a TConfig (include/nn/graph.hpp:78-126, src/nn/op_node.cpp:6-15,
    src/common/config.cpp:225-237): {name, type:"graph", parameters:{op_nodes:{type:"json", value:[{type:"op_node",
    parameters:{uid,name,layer_config}}]}, io_nodes, edges:[{producers, consumers, op_node}]}}.  A header hand-written in
    exactly that schema (what a reference-produced file contains) followed by tensor records must load."""
    import io
    import json
    import struct

    from tnn_b200.core import tensor_io
    from tnn_b200.models import load_or_create_model
    from tnn_b200.nn.graph import Graph

    g = load_or_create_model("mnist_cnn", device="cpu", seed=3)
    buf = io.BytesIO()
    g.save_state(buf)
    raw = buf.getvalue()
    hdr, end = json.JSONDecoder().raw_decode(raw.decode("latin-1"))
    assert hdr["type"] == "graph" and set(hdr["parameters"]) == {"op_nodes", "io_nodes", "edges"}
    ops = hdr["parameters"]["op_nodes"]
    assert ops["type"] == "json" and ops["value"][0]["type"] == "op_node"
    assert set(ops["value"][0]["parameters"]) == {"uid", "name", "layer_config"}
    assert ops["value"][0]["parameters"]["layer_config"]["type"] == "json"
    assert hdr["parameters"]["io_nodes"]["value"][0] == {"name": "", "type": "io_node", "parameters": {"uid": {"type": "string", "value": "input"}}}
    e0 = hdr["parameters"]["edges"]["value"][0]
    assert set(e0) == {"producers", "consumers", "op_node"}
    seq_cfg = ops["value"][0]["parameters"]["layer_config"]["value"]
    assert seq_cfg["type"] == "sequential" and seq_cfg["parameters"]["layers"]["type"] == "json"

    # golden: a reference-style file for a one-dense-layer graph, built byte by byte without our writer
    dense = {"name": "fc", "type": "dense", "parameters": {"input_features": {"type": "size_t", "value": 3},
                                                            "output_features": {"type": "size_t", "value": 2},
                                                            "use_bias": {"type": "bool", "value": True}}}
    seq = {"name": "m", "type": "sequential", "parameters": {"layers": {"type": "json", "value": [dense]}}}
    golden = {"name": "tiny", "type": "graph", "parameters": {
        "op_nodes": {"type": "json", "value": [{"name": "", "type": "op_node", "parameters": {
            "uid": {"type": "string", "value": "model"}, "name": {"type": "string", "value": "m"},
            "layer_config": {"type": "json", "value": seq}}}]},
        "io_nodes": {"type": "json", "value": [{"name": "", "type": "io_node", "parameters": {"uid": {"type": "string", "value": u}}}
                                               for u in ("input", "output")]},
        "edges": {"type": "json", "value": [{"producers": ["input"], "consumers": ["output"], "op_node": "model"}]}}}
    W = torch.arange(6, dtype=torch.float32).reshape(2, 3)
    b = torch.tensor([0.5, -0.5])
    blob = json.dumps(golden, indent=4).encode()
    for t in (W, b):                               # TensorImpl::save: u32 dtype (FP32 = 5), u64 ndims, u64 dims[], raw
        blob += struct.pack("<IQ", 5, t.dim()) + struct.pack(f"<{t.dim()}Q", *t.shape) + t.numpy().tobytes()
    p = tmp_path / "tiny"
    p.write_bytes(blob)
    g2 = Graph.load_state(str(p), "cpu")
    assert g2.name == "tiny"
    x = torch.tensor([[1.0, 2.0, 3.0]])
    from tnn_b200.nn.graph import GraphExecutor
    y = GraphExecutor(g2).forward({"input": x})["output"]
    torch.testing.assert_close(y, x @ W.t() + b)
    # and the round-1 flat header still loads
    flat = {"name": "old", "op_nodes": [{"uid": "model", "name": "m", "layer_config": seq}], "io_nodes": [{"uid": "input"}, {"uid": "output"}],
            "edges": [{"inputs": ["input"], "outputs": ["output"], "op": "model"}]}
    p2 = tmp_path / "old"
    p2.write_bytes(json.dumps(flat).encode() + blob[len(json.dumps(golden, indent=4).encode()):])
    g3 = Graph.load_state(str(p2), "cpu")
    torch.testing.assert_close(GraphExecutor(g3).forward({"input": x})["output"], y)


def test_symmetric_memory_host_side_pieces(tmp_path):
    """csrc/runtime/symmetric_memory.hpp without a GPU: the offset free list (best fit, coalescing, refused double frees),
    descriptors crossing a process boundary through FdMailbox (the carrier of cuMemExportToShareableHandle's POSIX fds),
    and probe() naming the reason the VMM group is unavailable on a CPU box."""
    import os

    from tnn_b200.ops import native
    Cx = native.require()
    fl = Cx.OffsetFreeList(1 << 16, 256)
    a, b, c = fl.allocate(100), fl.allocate(1000), fl.allocate(100)
    assert (a, b, c) == (0, 256, 1280) and fl.block_size(b) == 1024 and fl.in_use == 256 + 1024 + 256
    assert fl.release(b) and not fl.release(b)
    assert fl.allocate(512) == 256                      # best fit: the 1 KB hole, not the tail
    assert fl.allocate(1 << 20) is None
    for off in (a, 256, c):
        assert fl.release(off)
    assert fl.num_free_ranges == 1 and fl.largest_free == 1 << 16 and fl.peak >= 1536

    d = str(tmp_path)
    pid = os.fork()
    if pid == 0:                                        # child: gets the write end of a pipe through the mailbox
        rc = 1
        try:
            box = Cx.FdMailbox(d, "child.sock")
            tag, fd = box.recv(20000)
            if tag == 42 and fd >= 0:
                os.write(fd, b"via scm_rights")
                os.close(fd)
                rc = 0
        finally:
            os._exit(rc)
    r, w = os.pipe()
    box = Cx.FdMailbox(d, "parent.sock")
    box.send(os.path.join(d, "child.sock"), 42, w, 20000)   # retried until the child has bound
    os.close(w)
    assert os.read(r, 64) == b"via scm_rights"
    os.close(r)
    assert os.waitpid(pid, 0)[1] == 0
    with pytest.raises(RuntimeError, match="timed out"):
        box.recv(50)

    if not torch.cuda.is_available():
        sup = Cx.SymmetricMemory.probe([0, 1])
        assert not sup["vmm"] and not sup["multicast"] and sup["why"]
        with pytest.raises(RuntimeError, match="SymmetricMemory"):
            Cx.SymmetricMemory([0, 1], 1 << 20)
