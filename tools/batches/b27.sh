#!/bin/bash
# GPU batch 27: dropout + residual add fusion, Dense-tail residual fusion
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b27.log 2>&1
tail -5 gpurun_out/pytest_gpu_b27.log | cut -c1-250
timeout 300 python benchmarks/gpt2_train_bench.py --impl tnn --steps 8 --warmup 3 > gpurun_out/b27_gpt2_train.jsonl 2>&1; tail -1 gpurun_out/b27_gpt2_train.jsonl | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_gpt2_train_b27.csv python benchmarks/gpt2_train_bench.py --impl tnn --steps 1 --warmup 3 > gpurun_out/gpt2_train_ncu_b27.log 2>&1
