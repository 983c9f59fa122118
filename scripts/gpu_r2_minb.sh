#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2minb}
mkdir -p $OUT
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 10 --no-cpu-baseline --no-exact-step > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - $OUT/bench_$tag.json $tag <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["kernels"]
f = lambda n: round(1e3 * k[n]["total_ms"] / k[n]["launches"], 1)
print(sys.argv[2], "value", round(d["value"], 2), "jac", f("jacobian"), "prep", f("point_prepare"), "bsub", f("back_substitute"), "gather", f("camera_gather"))
PY
}
run default X=1
run bsub3 DBA_BSUB_MINB=3
