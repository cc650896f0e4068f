#!/bin/bash
# GPU batch 30: tall BN backward apply
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_training_gpu.py -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b30.log 2>&1
tail -4 gpurun_out/pytest_gpu_b30.log | cut -c1-250
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/b30_bench.json 2> gpurun_out/b30_bench.err
python -c "import json;d=json.loads(open('gpurun_out/b30_bench.json').read().strip().splitlines()[-1]);print('bench',d['ms_per_step'],d['value'],d.get('e2e',{}).get('value'),d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b30_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train_b30.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/launches_train_b30.log 2>&1
