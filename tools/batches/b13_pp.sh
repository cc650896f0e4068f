#!/bin/bash
# 2-GPU batch: 1F1B pipeline with CUDA-graph units vs eager units (same box, same run)
mkdir -p gpurun_out
for g in 1 0; do
  TNN_PP_GRAPH=$g timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$g tests/pp_check.py > gpurun_out/b13_pp2_graph$g.log 2>&1
  echo "== TNN_PP_GRAPH=$g rc=$?"; grep -E "losses|pipeline|PP_CHECK|Error|error" gpurun_out/b13_pp2_graph$g.log | cut -c1-300 | tail -8
done
