#!/bin/bash
# GPU batch 32: automatic 128x256 tiles for wide K-major GEMMs
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "linear or dense or skinny" --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b32.log 2>&1; tail -2 gpurun_out/pytest_gpu_b32.log | cut -c1-200
for v in on off on2 off2; do
  case $v in on|on2) env="TNN_GEMM_AUTO_BN256=1";; off|off2) env="TNN_GEMM_AUTO_BN256=0";; esac
  env $env timeout 300 python benchmarks/gpt2_inference_bench.py --impl tnn > gpurun_out/b32_gpt2_inf_$v.jsonl 2>/dev/null
  env $env timeout 300 python benchmarks/gpt2_train_bench.py --impl tnn --steps 8 --warmup 3 > gpurun_out/b32_gpt2_train_$v.jsonl 2>/dev/null
  python - <<PY
import json
a=json.loads(open('gpurun_out/b32_gpt2_inf_$v.jsonl').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/b32_gpt2_train_$v.jsonl').read().strip().splitlines()[-1])
print('$v prefill_ms', a['prefill_ms'], 'decode_graph_ms', a['decode_graph_ms_per_step'], 'train_ms', b['ms_per_step'])
PY
done
