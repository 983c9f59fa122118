#!/bin/bash
# Short gpurun call: GPU tests + the default bench (+ optional A/B runs given as env prefixes)
set -u
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -4 $OUT/pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d["value"], 2), "it/s  e2e", round(d["e2e"]["value"], 2), " roof", round(d["roofline"]["frac"], 3) if d.get("roofline") else None,
          {k: (v["launches"], round(v["total_ms"], 2)) for k, v in d["kernels"].items()})
except Exception as e:
    print("no json", sys.argv[1], e)
PY
}
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_bal5m.json 2> $OUT/bench_bal5m.err; echo "bench exit $?"; show $OUT/bench_bal5m.json
DBA_MF_TAIL=0 timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_bal5m_notail.json 2> $OUT/bench_bal5m_notail.err; show $OUT/bench_bal5m_notail.json
timeout 300 python bench.py --workload arc1m --no-cpu-baseline > $OUT/bench_arc1m.json 2> $OUT/bench_arc1m.err; show $OUT/bench_arc1m.json
