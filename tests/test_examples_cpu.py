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
    """`bench.py --impl reference` must exit 0 with exactly one JSON line (the reference is a C++/CMake tree
    that cannot be pip-installed offline: the line says so)."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" in d
