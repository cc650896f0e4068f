#!/bin/bash
# GPU batch 22: predicated BN loops + residual L2 prefetch; fused vs separate BN statistics again
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_training_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b22.log 2>&1
tail -5 gpurun_out/pytest_gpu_b22.log | cut -c1-250
for v in fuse nofuse fuse2 nofuse2; do
  case $v in
    fuse|fuse2) env="TNN_BN_FUSE_STATS=1";;
    nofuse|nofuse2) env="TNN_BN_FUSE_STATS=0";;
  esac
  env $env timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e > gpurun_out/b22_bench_$v.json 2> gpurun_out/b22_bench_$v.err
  python -c "import json;d=json.loads(open('gpurun_out/b22_bench_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['value'],d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b22_bench_$v.err
done
TNN_BN_FUSE_STATS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train_b22_nofuse.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/launches_train_b22.log 2>&1
