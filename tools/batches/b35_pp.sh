#!/bin/bash
# 2-GPU: pipeline microbatch-count sweep with CUDA-graph units (global batch 256)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for m in 4 2 8; do
  PP_MICRO=$m timeout 200 $T --master-port 2955$m tests/pp_check.py > gpurun_out/b35_pp2_m$m.log 2>&1
  echo "== M=$m rc=$?"; grep -E "losses ref|pipeline step|PP_CHECK" gpurun_out/b35_pp2_m$m.log | cut -c1-260
done
