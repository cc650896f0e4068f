#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/ -x -q -m gpu > gpurun_out/b41_pytest_gpu.log 2>&1; tail -3 gpurun_out/b41_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 200 python bench.py > gpurun_out/b41_bench.json 2>/dev/null; cut -c1-330 gpurun_out/b41_bench.json
