#!/bin/bash
# GPU batch 40: 8-softmax-warp attention forward (TNN_ATTN_SPLIT=1) vs the 4-warp kernel
mkdir -p gpurun_out
TNN_ATTN_SPLIT=1 timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_training_gpu.py tests/test_layers_gpu.py -m gpu -q -k "attention or gpt2" --timeout=60 --timeout-method=thread > gpurun_out/b40_pytest_split.log 2>&1; tail -3 gpurun_out/b40_pytest_split.log | cut -c1-200
for v in 1 0; do
  TNN_ATTN_SPLIT=$v timeout 150 python benchmarks/attention_bench.py > gpurun_out/b40_attention_split$v.jsonl 2>&1; echo "split=$v"; grep attention_fwd gpurun_out/b40_attention_split$v.jsonl | cut -c1-140
  TNN_ATTN_SPLIT=$v timeout 200 python benchmarks/gpt2_inference_bench.py --impl tnn > gpurun_out/b40_gpt2_inf_split$v.jsonl 2>/dev/null; cut -c1-200 gpurun_out/b40_gpt2_inf_split$v.jsonl
done
