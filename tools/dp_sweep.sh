#!/bin/bash
# usage (GPU box): tools/dp_sweep.sh N  -- data-parallel bench variants at N GPUs (fused kernel knobs), one JSON line each
N="${1:-2}"
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # name, env...
  local name="$1"; shift
  env "$@" timeout 200 $T --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 40 --warmup 8 --no-gpt2 --no-e2e \
      > gpurun_out/dp${N}_${name}.json 2> gpurun_out/dp${N}_${name}.err
  python - "$name" gpurun_out/dp${N}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    dp = d.get("dp", {})
    print(f"{sys.argv[1]:28s} {d['value']:10.0f} img/s {d['ms_per_step']:.3f} ms/step  path={dp.get('path')} buckets={dp.get('buckets')} us={[round(b['us'],1) for b in dp.get('per_bucket', [])]} roof={[round(b['frac_of_roofline'] or 0,3) for b in dp.get('per_bucket', [])]}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default X=1
run ctas16 TNN_DP_CTAS=16
run ctas8 TNN_DP_CTAS=8
run ctas148 TNN_DP_CTAS=148
run no_overlap TNN_DP_OVERLAP=0
run p2p TNN_DP_MULTICAST=0
run one_bucket TNN_DP_BUCKET_ELEMS=100000000
run small_buckets TNN_DP_BUCKET_ELEMS=1000000
run nccl X=1 2>/dev/null
