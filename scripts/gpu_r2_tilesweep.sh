#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2tile}
mkdir -p $OUT
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 10 --no-cpu-baseline --no-exact-step > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - $OUT/bench_$tag.json $tag <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["kernels"]["spmv_mf_pcg"]
print(sys.argv[2], "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 1), "spmv us", round(1e3 * k["total_ms"] / k["launches"], 1), "cost", d.get("final_cost"))
PY
}
run default X=1
run t256m3 DBA_TILE=256 DBA_MF_MINB=3
run t256m4 DBA_TILE=256 DBA_MF_MINB=4
run t256m2 DBA_TILE=256 DBA_MF_MINB=2
run t1024 DBA_TILE=1024
