#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/ -x -q -m gpu > gpurun_out/b38_pytest_gpu.log 2>&1; tail -4 gpurun_out/b38_pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b38_smoke.log 2>&1; tail -1 gpurun_out/b38_smoke.log | cut -c1-200
timeout 300 python bench.py > gpurun_out/b38_bench.json 2>/dev/null; cut -c1-420 gpurun_out/b38_bench.json
