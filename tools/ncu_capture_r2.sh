#!/bin/bash
# Round-2 ncu evidence (ONE GPU): `--set full` of the block-scaled fp8 GEMM, the flash-attention forward, one conv GEMM
# + one BatchNorm kernel of the WRN step, and the persistent decode kernel.  Reports land in gpurun_out/; summarise here
# with tools/ncu_summary.py into profiles/r2/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
# fp8 prefill: 2nd qkv-shaped block-scaled GEMM and the 2nd attention forward of a GPT-2 small prefill (bs 32, seq 1024)
timeout 600 $NCU -k regex:umma_gemm_mxfp8 -s 6 -c 1 -o gpurun_out/ncu_r2_fp8_gemm python tools/prof_gpt2.py --mode prefill-fp8 --reps 1 > gpurun_out/ncu_r2_fp8_gemm.log 2>&1
timeout 600 $NCU -k regex:attn_fwd_kernel -s 2 -c 1 -o gpurun_out/ncu_r2_attn_fwd python tools/prof_gpt2.py --mode prefill-bf16 --reps 1 > gpurun_out/ncu_r2_attn_fwd.log 2>&1
# WRN-16-8 step through the C++ host (eager steps before the capture): a CTA-pair conv GEMM and the BatchNorm kernels
timeout 600 $NCU -k regex:'umma_gemm2_persist_kernel|k_bn_' -s 40 -c 6 -o gpurun_out/ncu_r2_wrn_conv_bn cpp/bin/bench_cpp --steps 2 --warmup 1 --no-graph > gpurun_out/ncu_r2_wrn.log 2>&1
# persistent decode kernel (one launch = one token)
( cd /tmp && timeout 900 $NCU -k regex:k_decode_mega -s 3 -c 1 -o /root/repo/gpurun_out/ncu_r2_decode_mega /root/repo/cpp/bin/gpt2_inference --bench 32 1024 4 > /root/repo/gpurun_out/ncu_r2_decode_mega.log 2>&1 )
ls -la gpurun_out/*.ncu-rep
