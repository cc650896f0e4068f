#!/bin/bash
# GPU batch 14: full gpu tests, same-run A/B switches, GEMM pair path check, ncu of backward kernels, GPT-2 benches
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout=90 --timeout-method=thread > gpurun_out/pytest_gpu_b14.log 2>&1
tail -12 gpurun_out/pytest_gpu_b14.log | cut -c1-250
for v in base nofusebwd legacy_wgrad recompute base2; do
  case $v in
    base|base2) env="A=1";;
    nofusebwd) env="TNN_BN_FUSE_BWD=0";;
    recompute) env="TNN_BN_FUSE_BWD=0 TNN_BN_RECOMPUTE_MASK=1";;
    legacy_wgrad) env="TNN_WGRAD_LEGACY_SPLITS=1";;
  esac
  env $env timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/b14_bench_$v.json 2> gpurun_out/b14_bench_$v.err
  python -c "import json;d=json.loads(open('gpurun_out/b14_bench_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['value'],d.get('e2e',{}).get('value'),d.get('clocks',{}).get('sm_mhz'),d.get('final_loss'))" || tail -3 gpurun_out/b14_bench_$v.err
done
timeout 200 python benchmarks/gemm_bench.py gpt2_qkv gpt2_proj gpt2_fc1 gpt2_fc2 square_4k square_8k > gpurun_out/b14_gemm_default.jsonl 2>&1
TNN_GEMM_BLOCK_N=256 timeout 200 python benchmarks/gemm_bench.py gpt2_qkv gpt2_fc1 gpt2_fc2 square_4k > gpurun_out/b14_gemm_bn256.jsonl 2>&1
cut -c1-420 gpurun_out/b14_gemm_default.jsonl | tail -8; cut -c1-420 gpurun_out/b14_gemm_bn256.jsonl | tail -6
timeout 300 python benchmarks/gpt2_train_bench.py --steps 8 --warmup 3 > gpurun_out/b14_gpt2_train.jsonl 2>&1
tail -4 gpurun_out/b14_gpt2_train.jsonl | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd|wgrad|bn_apply" -s 40 -c 14 -o gpurun_out/prof_bwd -f python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/prof_bwd.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_gpt2_decode.csv python benchmarks/gpt2_inference_bench.py --impl tnn --iters 1 --decode 8 > gpurun_out/gpt2_decode_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train_b14.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/launches_train_b14.log 2>&1
ls -la gpurun_out | tail -8
