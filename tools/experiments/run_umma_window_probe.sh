#!/usr/bin/env bash
# First GPU call of the next round (≈1 GPU-minute): settles how the 9 conv taps can address one halo'd shared-memory
# window (docs/NEXT_STEPS.md item 1).  Usage:
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash tools/experiments/run_umma_window_probe.sh'
set -euo pipefail
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -I tnn_b200/csrc/kernels \
     tools/experiments/umma_window_probe.cu -o gpurun_out/umma_window_probe
timeout 60 gpurun_out/umma_window_probe | tee gpurun_out/umma_window_probe.txt
