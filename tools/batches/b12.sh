#!/bin/bash
# GPU batch 12: full gpu tests, same-run A/B (BN mask recompute, wgrad), BN=256 GEMM, ncu of bwd kernels
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k attention_bwd --timeout=60 --timeout-method=thread > gpurun_out/pytest_attn_bwd.log 2>&1
tail -25 gpurun_out/pytest_attn_bwd.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q --timeout=90 --timeout-method=thread --deselect tests/test_kernels_gpu.py::test_attention_bwd_matches_reference > gpurun_out/pytest_gpu_b12.log 2>&1
tail -3 gpurun_out/pytest_gpu_b12.log
for v in base recompute legacy_wgrad base2; do
  case $v in
    base|base2) env="";;
    recompute) env="TNN_BN_RECOMPUTE_MASK=1";;
    legacy_wgrad) env="TNN_WGRAD_LEGACY_SPLITS=1";;
  esac
  env $env timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/b12_bench_$v.json 2> gpurun_out/b12_bench_$v.err
  python -c "import json;d=json.loads(open('gpurun_out/b12_bench_$v.json').read().strip().splitlines()[-1]);print('$v',d['ms_per_step'],d['value'],d.get('e2e'),d.get('clocks'))"
done
TNN_GEMM_BLOCK_N=256 timeout 200 python benchmarks/gemm_bench.py gpt2_qkv gpt2_proj gpt2_fc1 gpt2_fc2 square_4k square_8k > gpurun_out/b12_gemm_bn256.jsonl 2>&1
timeout 200 python benchmarks/gemm_bench.py gpt2_qkv gpt2_proj gpt2_fc1 gpt2_fc2 square_4k square_8k > gpurun_out/b12_gemm_default.jsonl 2>&1
cut -c1-400 gpurun_out/b12_gemm_bn256.jsonl; cut -c1-400 gpurun_out/b12_gemm_default.jsonl
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd|wgrad" -s 30 -c 12 -o gpurun_out/prof_bwd -f python bench.py --steps 1 --warmup 3 > gpurun_out/prof_bwd.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_gpt2_decode.csv python benchmarks/gpt2_inference_bench.py --impl tnn --iters 1 --decode 8 > gpurun_out/gpt2_decode_ncu.log 2>&1
timeout 300 python benchmarks/attention_bench.py > gpurun_out/b12_attention_bench.jsonl 2>&1
cut -c1-330 gpurun_out/b12_attention_bench.jsonl
ls -la gpurun_out | tail -12
