#!/bin/bash
# usage (on a GPU box): tools/multigpu_suite.sh N   -- N = 2, 4 or 8 GPUs; writes gpurun_out/mg<N>_*.{log,json}
N="${1:-2}"
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $T --master-port 29533 tests/dp_check.py > gpurun_out/mg${N}_dp_check.log 2>&1
grep -E "unit step|replicas|train fused|step ms|DP_CHECK|Error" gpurun_out/mg${N}_dp_check.log | cut -c1-260 | head -24
timeout 240 $T --master-port 29544 tests/pp_check.py > gpurun_out/mg${N}_pp_check.log 2>&1
grep -E "pipeline|losses|PP_CHECK|Error" gpurun_out/mg${N}_pp_check.log | cut -c1-260 | head
timeout 150 $T --master-port 29555 bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/mg${N}_bench_tnn.json 2> gpurun_out/mg${N}_bench_tnn.err
cut -c1-420 gpurun_out/mg${N}_bench_tnn.json; tail -2 gpurun_out/mg${N}_bench_tnn.err
timeout 150 $T --master-port 29566 bench.py --gpus $N --dp nccl --steps 40 --warmup 8 --no-e2e > gpurun_out/mg${N}_bench_tnn_nccl.json 2>/dev/null
cut -c1-200 gpurun_out/mg${N}_bench_tnn_nccl.json
timeout 150 $T --master-port 29577 bench.py --gpus $N --impl torch --steps 40 --warmup 8 > gpurun_out/mg${N}_bench_torch.json 2>/dev/null
cut -c1-200 gpurun_out/mg${N}_bench_torch.json
# library arm for the pipeline configs: torch autograd + cuDNN stages, 1F1B over NCCL p2p (baseline/torch_pipeline.py)
timeout 200 $T --master-port 29588 baseline/torch_pipeline.py --model cifar100_wrn16_8 --batch 256 --microbatches 4 --steps 30 --warmup 8 \
  > gpurun_out/mg${N}_torch_pipeline.json 2> gpurun_out/mg${N}_torch_pipeline.err
cut -c1-300 gpurun_out/mg${N}_torch_pipeline.json; tail -2 gpurun_out/mg${N}_torch_pipeline.err
