#!/bin/bash
# ncu --set full of the final build's top kernels (WRN step + GPT-2 attention kernels)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"umma_gemm2_persist|umma_gemm_persist|umma_wgrad|tall" -s 150 -c 28 -o gpurun_out/prof_final_wrn -f python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/prof_final_wrn.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd|attn_bwd|k_skinny|k_dp_step|k_adam" -c 6 -o gpurun_out/prof_final_gpt2 -f python benchmarks/gpt2_train_bench.py --impl tnn --steps 1 --warmup 1 > gpurun_out/prof_final_gpt2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
