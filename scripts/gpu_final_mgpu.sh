#!/bin/bash
# gpurun --gpus N: bench.py at N ranks for the given workloads (peer windows, default settings)
set -u
N=$1; TAG=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
export DBA_P2P_TIMEOUT_MS=8000
for WL in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29815 \
     bench.py --gpus $N --workload $WL > $OUT/bench_${WL}_${N}gpu.json 2> $OUT/bench_${WL}_${N}gpu.err
  echo "$WL x$N exit $?"
  python - "$OUT/bench_${WL}_${N}gpu.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 2), "it/s  ms/step", round(d["ms_per_step"], 3), " e2e", round(d["e2e"]["value"], 2), {k: round(1e3 * v["total_ms"] / max(v["launches"], 1), 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("no json", e)
PY
done
