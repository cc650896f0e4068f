#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/ -x -q -m gpu > gpurun_out/b39_pytest_gpu.log 2>&1; tail -4 gpurun_out/b39_pytest_gpu.log | cut -c1-250
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/b39_bench.json 2>/dev/null; python -c "import json;d=json.loads(open('gpurun_out/b39_bench.json').read().strip().splitlines()[-1]);print('bench',d['ms_per_step'],d['value'],d['e2e']['value'],d['final_loss'])"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train_b39.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/launches_train_b39.log 2>&1
