#!/bin/bash
# 2-GPU batch: dp_check, pipeline with CUDA-graph units vs eager, DP bench (fused + nccl)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $T --master-port 29533 tests/dp_check.py > gpurun_out/b17_dp_check.log 2>&1
grep -E "unit step|replicas|train fused|step ms|DP_CHECK|Error" gpurun_out/b17_dp_check.log | cut -c1-260 | head -12
for g in 1 0; do
  TNN_PP_GRAPH=$g timeout 300 $T --master-port 2954$g tests/pp_check.py > gpurun_out/b17_pp2_graph$g.log 2>&1
  echo "== TNN_PP_GRAPH=$g rc=$?"; grep -E "losses|pipeline|PP_CHECK|Error|error" gpurun_out/b17_pp2_graph$g.log | cut -c1-300 | tail -6
done
timeout 150 $T --master-port 29555 bench.py --gpus 2 --steps 40 --warmup 8 > gpurun_out/b17_bench_tnn.json 2> gpurun_out/b17_bench_tnn.err
cut -c1-420 gpurun_out/b17_bench_tnn.json; tail -2 gpurun_out/b17_bench_tnn.err
timeout 150 $T --master-port 29566 bench.py --gpus 2 --dp nccl --steps 40 --warmup 8 --no-e2e > gpurun_out/b17_bench_tnn_nccl.json 2>/dev/null
cut -c1-200 gpurun_out/b17_bench_tnn_nccl.json
