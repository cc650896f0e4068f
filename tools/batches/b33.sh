#!/bin/bash
# GPU batch 33: attention forward with single-sweep softmax
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_training_gpu.py tests/test_layers_gpu.py -m gpu -q -k "attention or gpt2" --timeout=60 --timeout-method=thread > gpurun_out/pytest_gpu_b33.log 2>&1; tail -3 gpurun_out/pytest_gpu_b33.log | cut -c1-200
timeout 200 python benchmarks/attention_bench.py > gpurun_out/b33_attention_bench.jsonl 2>&1; grep attention_fwd gpurun_out/b33_attention_bench.jsonl | cut -c1-260
timeout 300 python benchmarks/gpt2_inference_bench.py --impl tnn > gpurun_out/b33_gpt2_inf.jsonl 2>/dev/null; cut -c1-330 gpurun_out/b33_gpt2_inf.jsonl
timeout 300 python benchmarks/gpt2_train_bench.py --impl tnn --steps 8 --warmup 3 > gpurun_out/b33_gpt2_train.jsonl 2>/dev/null; cut -c1-200 gpurun_out/b33_gpt2_train.jsonl
