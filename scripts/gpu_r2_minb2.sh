#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2minb4}
mkdir -p $OUT
run() {
  tag=$1; wl=$2; ls=$3; shift; shift; shift
  env "$@" timeout 600 python bench.py --gpus 1 --workload $wl --linear-solver $ls --steps 8 --no-cpu-baseline --no-exact-step > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - $OUT/bench_$tag.json $tag <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["kernels"]
f = lambda n: round(1e3 * k[n]["total_ms"] / k[n]["launches"], 1) if n in k else None
print(sys.argv[2], "value", round(d["value"], 2), "jac", f("jacobian"), "prep", f("point_prepare"), "bsub", f("back_substitute"), "gather", f("camera_gather"))
PY
}
for mb in 2 3 4; do
run arc_dense_jac$mb arc1m dense DBA_JAC_MINB=$mb
done
run arc_dense_flat arc1m dense DBA_JAC=flat
run tea_dense teabottle dense X=1
