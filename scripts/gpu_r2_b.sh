#!/bin/bash
# round 2: dense tests + dense benches (quick)
set -u
OUT=gpurun_out/${1:-r2b}
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q -x > $OUT/pytest_dense.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_dense.log )
tail -5 $OUT/pytest_dense.log
for wl in arc1m teabottle; do
    timeout 300 python bench.py --workload $wl --linear-solver dense --no-cpu-baseline > $OUT/bench_${wl}_dense.json 2> $OUT/bench_${wl}_dense.err; echo "bench $wl dense exit $?"
    python - $OUT/bench_${wl}_dense.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 1), "it/s e2e", round(d["e2e"]["value"], 1), {k: (v["launches"], round(v["total_ms"], 3)) for k, v in d["kernels"].items()})
except Exception as e:
    print("no json", e)
PY
done
