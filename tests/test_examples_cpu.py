"""The example CLIs (reference: examples/*.cpp) must at least import, and the single-process trainer must run
an epoch on synthetic data end to end on the CPU path."""
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = sorted(f for f in os.listdir(os.path.join(ROOT, "examples")) if f.endswith(".py"))


@pytest.mark.parametrize("name", EXAMPLES)
def test_example_module_imports(name):
    spec = importlib.util.spec_from_file_location("example_" + name[:-3], os.path.join(ROOT, "examples", name))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                 # top level only: every example guards its entry point with __main__


def test_trainer_cli_runs_one_synthetic_epoch_on_cpu(tmp_path):
    env = dict(os.environ, MODEL_NAME="mnist_cnn", DEVICE_TYPE="CPU", BATCH_SIZE="64", EPOCHS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "trainer.py"), "--synthetic"], cwd=tmp_path,
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "epoch 1 done" in out.stdout and "[trainer] done" in out.stdout


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` must exit 0 with exactly one JSON line.  On a GPU box it times the unmodified
    reference torch WRN-16-8 (baseline/_ref); without CUDA it says so.  The reference module itself must import
    and build the model without touching tnn_b200."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" in d
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_arm
    assert reference_arm.ensure_installed() is None
    ref = reference_arm.load_reference_module()
    import torch
    m = ref.WRN16_8CIFAR100(num_classes=100)
    assert m(torch.rand(2, 3, 32, 32)).shape == (2, 100)
    assert sum(p.numel() for p in m.parameters()) > 10_000_000


def _run(rel, *args, cwd=None, env=None, timeout=300):
    out = subprocess.run([sys.executable, os.path.join(ROOT, rel), *args], cwd=cwd, env=env, capture_output=True, text=True,
                         timeout=timeout)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-2500:])
    return out.stdout


def test_small_example_apps_run_on_cpu(tmp_path):
    """device_manager_example / hardware_info_example / msequential_example / test_del / graph_test twins."""
    import json
    assert "Memory access test passed" in _run("examples/device_manager_example.py")
    out = _run("examples/hardware_info_example.py", "--samples", "1", "--interval", "0.01", "--out-dir", str(tmp_path))
    assert "NUMA Node 0" in out
    info = json.load(open(tmp_path / "temp" / "hardware_info.json"))
    assert info["logical_cores"] >= 1 and "numa" in info
    assert (tmp_path / "logs" / "cpu_status.csv").read_text().count("\n") == 2
    out = _run("examples/msequential_example.py")
    assert "Output shape: (4, 5)" in out and "Output shape: (8, 50)" in out and out.count("Backward pass completed") == 2
    assert "ptrX: offset    10240" in _run("examples/del_allocator_example.py")
    out = _run("examples/graph_example.py", "--steps", "2", "--batch", "16", cwd=tmp_path,
               env=dict(os.environ, DEVICE_TYPE="CPU", MODEL_NAME="mnist_cnn"))
    assert "op nodes reduced to 1" in out and out.count("Loss:") == 2


def test_legacy_zoo_matches_nhwc_zoo_and_legacy_trainer_runs(tmp_path):
    """legacy::ExampleModels: same network on NCHW tensors built from legacy_* layers (same seed -> same output)."""
    import torch
    from tnn_b200.models import ExampleModels, legacy
    assert "cifar100_wrn16_8" in legacy.ExampleModels.available_models() and "gpt2_small" in legacy.ExampleModels.available_models()
    a = ExampleModels.create("cifar10_resnet9").init("cpu", seed=1).set_training(False)
    b = legacy.ExampleModels.create("cifar10_resnet9").init("cpu", seed=1).set_training(False)
    types = {type(l).__name__ for l in b.layers}
    assert "LegacyConv2DLayer" in types and "Conv2DLayer" not in types
    x = torch.randn(2, 32, 32, 3)
    torch.testing.assert_close(a.forward(x, 0), b.forward(x.permute(0, 3, 1, 2).contiguous(), 0))
    vit = legacy.ExampleModels.create("tiny_imagenet_vit").init("cpu", seed=2).set_training(True)
    y = vit.forward(torch.randn(1, 3, 64, 64), 0)
    assert y.shape == (1, 200) and vit.backward(torch.ones_like(y), 0).shape == (1, 3, 64, 64)
    env = dict(os.environ, MODEL_NAME="mnist_cnn", DEVICE_TYPE="CPU", BATCH_SIZE="16", EPOCHS="1")
    out = _run("examples/legacy_trainer.py", "--synthetic", "--samples", "64", cwd=tmp_path, env=env)
    assert "epoch 1 done" in out and "[legacy_trainer] done" in out


def test_cpu_runnable_benchmarks(tmp_path):
    """dense_benchmark / grad_clear_benchmark / tcp_communicator_benchmark twins on their CPU paths."""
    import json
    last = lambda s: json.loads([l for l in s.splitlines() if l.startswith("{")][-1])  # noqa: E731
    d = last(_run("benchmarks/dense_bench.py", "--in", "512", "--out", "64", "--batch", "8", "--passes", "2"))
    assert d["max_diff"] == 0.0 and d["max_grad_diff"] == 0.0 and d["max_weight_grad_diff"] == 0.0
    g = last(_run("benchmarks/grad_clear_bench.py", "mnist_cnn"))
    assert g["tensors"] > 0 and g["bulk_ms"] > 0
    for transport in ("tcp", "roce"):
        t = last(_run("benchmarks/tcp_communicator_bench.py", "--self-test", "--mb", "1", "--seconds", "1", "--transport", transport))
        assert t["bench"] == f"{transport}_communicator" and t["messages"] > 0 and t["MBps"] > 0


def test_torch_pipeline_baseline_matches_single_stage(tmp_path):
    """baseline/torch_pipeline.py (the library arm for the pipeline configs): two gloo stages follow the one-stage loss."""
    import json
    script = os.path.join(ROOT, "baseline", "torch_pipeline.py")
    common = ["--model", "tiny", "--batch", "16", "--microbatches", "4", "--steps", "2", "--warmup", "1", "--print-losses"]
    one = subprocess.run([sys.executable, script, *common], capture_output=True, text=True, timeout=300, cwd=tmp_path,
                         env=dict(os.environ, MASTER_PORT="29561"))
    assert one.returncode == 0, one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29562", script, *common, "--proportions", "1,1"], capture_output=True,
                         text=True, timeout=300, cwd=tmp_path)
    assert two.returncode == 0, two.stderr[-2000:]
    l1 = json.loads([l for l in one.stdout.splitlines() if l.startswith("{")][-1])
    l2 = json.loads([l for l in two.stdout.splitlines() if l.startswith("{")][-1])
    assert l2["n_gpus"] == 2 and l2["config"]["parallelism"] == "pp2" and len(l1["losses"]) == 3
    assert all(abs(a - b) < 1e-4 for a, b in zip(l1["losses"], l2["losses"]))


def test_usage_tours_run():
    """docs/USAGE.md quotes these two programs: tensors, LayerBuilder / TrainStep, augmentation (custom strategy, removal by
    name / index) and an in-process two-stage pipeline, in both hosts."""
    out = _run("examples/usage_tour.py", "--device", "cpu")
    assert "usage tour complete" in out and "after removals: 3 strategies" in out and "pipeline batch 2" in out
    from tnn_b200 import build
    build.build_host(verbose=False)
    cpp = subprocess.run([os.path.join(ROOT, "cpp", "bin", "usage_tour"), "--device", "CPU"], capture_output=True, text=True, timeout=300)
    assert cpp.returncode == 0 and "usage tour complete" in cpp.stdout and "after removals: 3 strategies" in cpp.stdout, cpp.stdout[-1500:] + cpp.stderr[-1500:]


def test_vocab_export_and_corpus_preparation_tools(tmp_path):
    """tools/export_vocab.py + tools/prepare_openwebtext.py (the reference's python/export_vocab.py and python/openwebtext.py
    without tiktoken / datasets): a Hugging-Face style byte-level vocab.json becomes vocab.bin, text + jsonl documents become
    uint16 train.bin / val.bin with an end-of-text token after every document, and the OpenWebText loader reads them."""
    import json

    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        from export_vocab import bytes_to_unicode
    finally:
        sys.path.pop(0)
    u = bytes_to_unicode()
    assert len(u) == 256 and len(set(u.values())) == 256 and u[ord("A")] == "A" and u[ord(" ")] == "Ġ"
    toks = [u[b] for b in range(256)] + [u[ord("t")] + u[ord("h")], u[ord(" ")] + u[ord("t")] + u[ord("h")], u[ord("e")] + u[ord("r")], "<|endoftext|>"]
    (tmp_path / "vocab.json").write_text(json.dumps({t: i for i, t in enumerate(toks)}))
    (tmp_path / "a.txt").write_text("the other weather, then there")
    (tmp_path / "b.jsonl").write_text(json.dumps({"text": "hello the world"}) + "\n" + json.dumps({"text": "another one"}) + "\n")
    out = tmp_path / "out"
    assert "wrote 260 tokens" in _run("tools/export_vocab.py", "--from", "vocab-json", str(tmp_path / "vocab.json"), "--out", str(out / "vocab.bin"))
    rep = json.loads(_run("tools/prepare_openwebtext.py", str(tmp_path / "a.txt"), str(tmp_path / "b.jsonl"), "--vocab", str(out / "vocab.bin"),
                          "--out", str(out), "--val-fraction", "0.5").splitlines()[-1])
    assert rep["documents"] == 3 and rep["eot"] == 259 and rep["val_tokens"] > 0 and rep["train_tokens"] > 0
    from tnn_b200.data import Tokenizer
    from tnn_b200.data.loaders import OpenWebTextDataLoader
    tok = Tokenizer(str(out / "vocab.bin"))
    ids = np.concatenate([np.fromfile(out / f"{s}.bin", dtype=np.uint16) for s in ("train", "val")]).tolist()
    assert ids.count(259) == 3                                         # one end-of-text per document
    text = tok.decode(ids)
    for doc in ("the other weather, then there", "hello the world", "another one"):
        assert doc + "<|endoftext|>" in text
    assert 257 in ids and 258 in ids                                   # the merges were used (" th", "er")
    dl = OpenWebTextDataLoader(seq_len=4).load_data(str(out))
    x, y = dl.get_batch(2)
    assert x.shape == (2, 4) and bool((x[:, 1:] == y[:, :-1]).all())
    assert "wrote 257 tokens" in _run("tools/export_vocab.py", "--from", "byte-fallback", "--out", str(out / "bytes.bin"))
