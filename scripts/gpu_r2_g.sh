#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2g}
mkdir -p $OUT
nproc; free -g | head -2
( timeout 1500 python -m pytest tests/test_gpu_scale.py tests/test_gpu_dense.py -m gpu -q --durations=10 > $OUT/pytest_scale.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_scale.log )
tail -40 $OUT/pytest_scale.log
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
