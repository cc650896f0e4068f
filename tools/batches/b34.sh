#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/ -x -q -m gpu > gpurun_out/b34_pytest_gpu.log 2>&1; tail -3 gpurun_out/b34_pytest_gpu.log | cut -c1-200
timeout 200 python benchmarks/attention_bench.py > gpurun_out/b34_attention_bench.jsonl 2>&1; grep attention_fwd gpurun_out/b34_attention_bench.jsonl | cut -c1-200
timeout 300 python benchmarks/gpt2_inference_bench.py --impl tnn > gpurun_out/b34_gpt2_inf.jsonl 2>/dev/null; cut -c1-330 gpurun_out/b34_gpt2_inf.jsonl
timeout 300 python bench.py > gpurun_out/b34_bench.json 2>/dev/null; cut -c1-330 gpurun_out/b34_bench.json
