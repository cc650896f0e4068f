#!/bin/bash
# GPU batch 15: tests, A/B after the BN-backward kernel rewrite, GPT-2 inference with cluster split-K decode GEMMs
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b15.log 2>&1
tail -8 gpurun_out/pytest_gpu_b15.log | cut -c1-250
for v in base recompute fusebwd base2; do
  case $v in
    base|base2) env="A=1";;
    fusebwd) env="TNN_BN_FUSE_BWD=1";;
    recompute) env="TNN_BN_RECOMPUTE_MASK=1";;
  esac
  env $env timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/b15_bench_$v.json 2> gpurun_out/b15_bench_$v.err
  python -c "import json;d=json.loads(open('gpurun_out/b15_bench_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['value'],d.get('e2e',{}).get('value'),d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b15_bench_$v.err
done
timeout 300 python benchmarks/gpt2_inference_bench.py --impl tnn > gpurun_out/b15_gpt2_inf.jsonl 2>&1
TNN_SKINNY_KSPLIT=1 timeout 300 python benchmarks/gpt2_inference_bench.py --impl tnn > gpurun_out/b15_gpt2_inf_ksplit1.jsonl 2>&1
tail -2 gpurun_out/b15_gpt2_inf.jsonl | cut -c1-400; tail -1 gpurun_out/b15_gpt2_inf_ksplit1.jsonl | cut -c1-400
timeout 200 python benchmarks/batchnorm_bench.py > gpurun_out/b15_bn_bench.jsonl 2>&1; tail -6 gpurun_out/b15_bn_bench.jsonl | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train_b15.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/launches_train_b15.log 2>&1
ls -la gpurun_out | tail -5
