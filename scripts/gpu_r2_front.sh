#!/bin/bash
# front-half experiment: camera gather reading planes (DBA_MF_FRONT=0) vs recomputing
set -u
OUT=gpurun_out/${1:-r2front}
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dense.py tests/test_gpu_loss.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log ); tail -5 $OUT/pytest.log
for m in 0 1; do
  DBA_MF_FRONT=$m timeout 600 python bench.py --gpus 1 --steps 10 --no-cpu-baseline --no-exact-step > $OUT/bench_front$m.json 2> $OUT/bench_front$m.err
  python - $OUT/bench_front$m.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 1), "final cost", d.get("final_cost"))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["total_ms"])[:9]:
    print("  %-18s %4d launches %8.3f ms  %7.1f us/launch" % (k, v["launches"], v["total_ms"], 1e3 * v["total_ms"] / max(v["launches"], 1)))
PY
done
