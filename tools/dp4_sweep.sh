N=4
cd /root/repo; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { local name="$1"; shift
  env "$@" timeout 200 $T --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 40 --warmup 8 --no-gpt2 --no-e2e > gpurun_out/dp${N}_${name}.json 2> gpurun_out/dp${N}_${name}.err
  python - "$name" gpurun_out/dp${N}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); dp = d.get("dp", {})
    print(f"{sys.argv[1]:18s} {d['value']:9.0f} img/s {d['ms_per_step']:.3f} ms path={dp.get('path')} buckets={dp.get('buckets')} us={[round(b['us'],1) for b in dp.get('per_bucket', [])]}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default X=1
run ctas16 TNN_DP_CTAS=16
run p2p TNN_DP_MULTICAST=0
run buckets1m TNN_DP_BUCKET_ELEMS=1000000
run buckets512k TNN_DP_BUCKET_ELEMS=524288
run default_again X=1
