#!/bin/bash
# GPU batch 37: cp.async-ring BN kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/b37_pytest_gpu.log 2>&1; tail -4 gpurun_out/b37_pytest_gpu.log | cut -c1-250
for v in async noasync async2; do
  case $v in async|async2) env="TNN_BN_ASYNC=1";; noasync) env="TNN_BN_ASYNC=0";; esac
  env $env timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e > gpurun_out/b37_bench_$v.json 2> gpurun_out/b37_bench_$v.err
  python -c "import json;d=json.loads(open('gpurun_out/b37_bench_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['value'],d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b37_bench_$v.err
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train_b37.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/launches_train_b37.log 2>&1
