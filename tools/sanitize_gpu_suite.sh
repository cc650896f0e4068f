#!/bin/bash
# GPU sanitizer evidence for profiles/: memcheck + racecheck + synccheck over the tensor-core kernel tests (tcgen05 GEMM /
# conv / attention / block-scaled fp8) and the C++ host's GPU parity tests.  Usage (GPU box): tools/sanitize_gpu_suite.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 9 --launch-timeout 300"
K='(linear_fwd and 300) or test_conv_residual_relu_epilogue or (attention_fwd and 128) or tall_column_sums'
for tool in memcheck racecheck synccheck; do
  echo "=== $tool: tests/test_kernels_gpu.py -k '$K'" > gpurun_out/sanitizer_${tool}.txt
  timeout 900 $S --tool $tool python -m pytest tests/test_kernels_gpu.py -x -q -k "$K" > gpurun_out/san_${tool}_kernels.log 2>&1; grep -vE "^$|warnings summary|Warning|warn|Host Frame" gpurun_out/san_${tool}_kernels.log | head -60 >> gpurun_out/sanitizer_${tool}.txt
  echo "=== $tool: tests/test_mxfp8_gpu.py (quantise, 128/256-row GEMMs, LayerNorm->MXFP8)" >> gpurun_out/sanitizer_${tool}.txt
  timeout 900 $S --tool $tool python -m pytest tests/test_mxfp8_gpu.py -x -q -k "(quantize and 128-128) or (gemm and (128-128-128 or 256-192-256)) or (layernorm and 256)" > gpurun_out/san_${tool}_mxfp8.log 2>&1; grep -vE "^$|warnings summary|Warning|warn|Host Frame" gpurun_out/san_${tool}_mxfp8.log | head -60 >> gpurun_out/sanitizer_${tool}.txt
  echo "=== $tool: cpp/bin/tnn_tests gpu_parity_dense_conv (C++ host)" >> gpurun_out/sanitizer_${tool}.txt
  timeout 900 $S --tool $tool cpp/bin/tnn_tests gpu_parity_dense_conv 2>&1 | grep -v "Host Frame" | head -60 >> gpurun_out/sanitizer_${tool}.txt
  echo "=== $tool: cpp/bin/tnn_tests gpu_gpt2 (C++ GPT2Engine: bf16 / MXFP8 prefill, graph decode, persistent decode kernel)" >> gpurun_out/sanitizer_${tool}.txt
  timeout 900 $S --tool $tool cpp/bin/tnn_tests gpu_gpt2 2>&1 | grep -v "Host Frame" | head -60 >> gpurun_out/sanitizer_${tool}.txt
done
tail -4 gpurun_out/sanitizer_*.txt
