#!/bin/bash
# final validation of the round: the driver's own commands
mkdir -p gpurun_out
timeout 700 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; cut -c1-900 gpurun_out/final_bench_default.json
timeout 200 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_reference.json
timeout 300 python bench.py --impl torch --steps 30 --warmup 5 > gpurun_out/final_bench_torch.json 2>/dev/null; cut -c1-300 gpurun_out/final_bench_torch.json
timeout 300 python benchmarks/gpt2_train_bench.py --steps 8 --warmup 3 > gpurun_out/final_gpt2_train.jsonl 2>/dev/null; cut -c1-250 gpurun_out/final_gpt2_train.jsonl
timeout 300 python benchmarks/gpt2_inference_bench.py > gpurun_out/final_gpt2_inf.jsonl 2>/dev/null; cut -c1-330 gpurun_out/final_gpt2_inf.jsonl
