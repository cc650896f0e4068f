#!/bin/bash
# GPU batch 16: side-stream wgrad A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_training_gpu.py -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b16.log 2>&1
tail -5 gpurun_out/pytest_gpu_b16.log | cut -c1-250
for v in side noside side2 noside2; do
  case $v in
    side|side2) env="TNN_WGRAD_STREAM=1";;
    noside|noside2) env="TNN_WGRAD_STREAM=0";;
  esac
  env $env timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/b16_bench_$v.json 2> gpurun_out/b16_bench_$v.err
  python -c "import json;d=json.loads(open('gpurun_out/b16_bench_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['value'],d.get('e2e',{}).get('value'),d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b16_bench_$v.err
done
