#!/bin/bash
# GPU batch 23: validate defaults, ncu full on BN kernels, GPT-2 train launch list
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b23.log 2>&1
tail -5 gpurun_out/pytest_gpu_b23.log | cut -c1-250
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/b23_bench.json 2> gpurun_out/b23_bench.err
python -c "import json;d=json.loads(open('gpurun_out/b23_bench.json').read().strip().splitlines()[-1]);print('bench',d['ms_per_step'],d['value'],d.get('e2e',{}).get('value'),d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b23_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_bn_stats|k_bn_bwd_reduce|k_bn_bwd_apply|k_bn_apply" -s 60 -c 16 -o gpurun_out/prof_bn -f python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/prof_bn.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_gpt2_train.csv python benchmarks/gpt2_train_bench.py --impl tnn --steps 1 --warmup 3 > gpurun_out/gpt2_train_ncu.log 2>&1
ls -la gpurun_out | tail -4
