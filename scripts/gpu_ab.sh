#!/bin/bash
# A/B runs of bench.py under different env settings: scripts/gpu_ab.sh TAG "ENV1=a" "ENV2=b" ...
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for E in "$@"; do
  N=$(echo "$E" | tr ' =' '__')
  env $E timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_$N.json 2> $OUT/bench_$N.err
  python - "$OUT/bench_$N.json" "$E" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"], 2), "it/s  e2e", round(d["e2e"]["value"], 2), {k: round(1e3 * v["total_ms"] / max(v["launches"], 1), 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("no json", sys.argv[1], e)
PY
done
