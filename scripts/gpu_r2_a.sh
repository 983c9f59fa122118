#!/bin/bash
# round 2, call A: GPU tests (all), smoke, dense-solver benches
set -u
OUT=gpurun_out/r2a
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -30 $OUT/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log ); tail -3 $OUT/smoke.log
for wl in arc1m teabottle; do
  for ls in dense pcg; do
    timeout 300 python bench.py --workload $wl --linear-solver $ls --no-cpu-baseline > $OUT/bench_${wl}_${ls}.json 2> $OUT/bench_${wl}_${ls}.err; echo "bench $wl $ls exit $?"
    python - $OUT/bench_${wl}_${ls}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 1), "it/s e2e", round(d["e2e"]["value"], 1), {k: (v["launches"], round(v["total_ms"], 3)) for k, v in d["kernels"].items()})
except Exception as e:
    print("no json", e)
PY
  done
done
