#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2sched}
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dense.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log ); tail -3 $OUT/pytest.log
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --gpus 1 --steps 10 --no-cpu-baseline --no-exact-step > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  grep "tail trace" $OUT/bench_$tag.err | tail -1 | cut -c1-250
  python - $OUT/bench_$tag.json $tag <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["kernels"]["spmv_mf_pcg"]
print(sys.argv[2], "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 1), "spmv us", round(1e3 * k["total_ms"] / k["launches"], 1), "cost", d.get("final_cost"))
PY
}
run dynamic X=1
run static DBA_MF_SCHED=static
run dynamic_trace DBA_TAIL_TRACE=1
run static_trace DBA_MF_SCHED=static DBA_TAIL_TRACE=1
