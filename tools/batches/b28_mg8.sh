#!/bin/bash
# 8-GPU batch: fused DP bench with the end-of-round build
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29555 bench.py --gpus 8 --steps 40 --warmup 8 > gpurun_out/b28_mg8_bench_tnn.json 2> gpurun_out/b28_mg8_bench_tnn.err
cut -c1-600 gpurun_out/b28_mg8_bench_tnn.json; tail -2 gpurun_out/b28_mg8_bench_tnn.err | cut -c1-200
