#!/bin/bash
# GPU batch 20: PDL decode chain A/B, 2-CTA thresholds A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_training_gpu.py -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b20.log 2>&1
tail -5 gpurun_out/pytest_gpu_b20.log | cut -c1-250
for v in pdl1 pdl0; do
  case $v in pdl1) env="TNN_PDL=1";; pdl0) env="TNN_PDL=0";; esac
  env $env timeout 300 python benchmarks/gpt2_inference_bench.py --impl tnn > gpurun_out/b20_gpt2_inf_$v.jsonl 2>&1
  echo $v; tail -1 gpurun_out/b20_gpt2_inf_$v.jsonl | cut -c1-330
done
for v in base n128 n128kb8 base2; do
  case $v in
    base|base2) env="A=1";;
    n128) env="TNN_GEMM_2CTA_MIN_N=128 TNN_GEMM_2CTA_MIN_KB=16";;
    n128kb8) env="TNN_GEMM_2CTA_MIN_N=128 TNN_GEMM_2CTA_MIN_KB=8";;
  esac
  env $env timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e > gpurun_out/b20_bench_$v.json 2> gpurun_out/b20_bench_$v.err
  python -c "import json;d=json.loads(open('gpurun_out/b20_bench_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['value'],d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b20_bench_$v.err
done
