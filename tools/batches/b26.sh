#!/bin/bash
# GPU batch 26: vectorised LayerNorm kernels + fused shortcut gradient
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b26.log 2>&1
tail -5 gpurun_out/pytest_gpu_b26.log | cut -c1-250
timeout 300 python benchmarks/gpt2_train_bench.py --impl tnn --steps 8 --warmup 3 > gpurun_out/b26_gpt2_train.jsonl 2>&1; tail -1 gpurun_out/b26_gpt2_train.jsonl | cut -c1-300
timeout 300 python benchmarks/gpt2_inference_bench.py --impl tnn > gpurun_out/b26_gpt2_inf.jsonl 2>&1; tail -1 gpurun_out/b26_gpt2_inf.jsonl | cut -c1-330
timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/b26_bench.json 2> gpurun_out/b26_bench.err
python -c "import json;d=json.loads(open('gpurun_out/b26_bench.json').read().strip().splitlines()[-1]);print('bench',d['ms_per_step'],d['value'],d.get('e2e',{}).get('value'),d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b26_bench.err
