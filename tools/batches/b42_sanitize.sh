#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in the round (bounded to ~100 s)
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 60 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py -x -q -m gpu -k "attention_bwd_matches_reference and 128-2 or skinny or tall_column or softmax_ce_vectorised or keep_bits or layernorm" > gpurun_out/b42_memcheck.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/b42_memcheck.log | cut -c1-200
