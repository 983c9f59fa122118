#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2c}
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_schur_dense -s 4 -c 1 -f -o $OUT/schur_dense_arc1m \
   python bench.py --workload arc1m --linear-solver dense --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu1.log 2>&1; echo "ncu1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_cholesky -s 4 -c 1 -f -o $OUT/cholesky_teabottle \
   python bench.py --workload teabottle --linear-solver dense --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu2.log 2>&1; echo "ncu2 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_arc1m.csv \
   python bench.py --workload arc1m --linear-solver dense --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu3.log 2>&1; echo "ncu3 exit $?"
