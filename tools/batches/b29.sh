#!/bin/bash
# GPU batch 29: tall BN statistics / backward-reduction kernels A/B
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b29.log 2>&1
tail -5 gpurun_out/pytest_gpu_b29.log | cut -c1-250
for v in tall notall tall2 notall2; do
  case $v in
    tall|tall2) env="TNN_BN_STATS_TALL=1";;
    notall|notall2) env="TNN_BN_STATS_TALL=0";;
  esac
  env $env timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e > gpurun_out/b29_bench_$v.json 2> gpurun_out/b29_bench_$v.err
  python -c "import json;d=json.loads(open('gpurun_out/b29_bench_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['value'],d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b29_bench_$v.err
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train_b29.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/launches_train_b29.log 2>&1
