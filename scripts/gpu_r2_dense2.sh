#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2dense2}
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_parity.py tests/test_gpu_loss.py tests/test_gpu_host.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log ); tail -5 $OUT/pytest.log
( timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "arc1m" > $OUT/pytest2.log 2>&1; echo "pytest exit $?" >> $OUT/pytest2.log ); tail -3 $OUT/pytest2.log
show() {
python - $1 <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 1), "final cost", d.get("final_cost"))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["total_ms"])[:12]:
    print("  %-18s %4d launches %8.3f ms  %7.1f us/launch" % (k, v["launches"], v["total_ms"], 1e3 * v["total_ms"] / max(v["launches"], 1)))
PY
}
timeout 600 python bench.py --workload arc1m --linear-solver dense --steps 8 --no-cpu-baseline --no-exact-step > $OUT/bench_arc1m_dense.json 2> $OUT/bench_arc1m_dense.err; show $OUT/bench_arc1m_dense.json
timeout 600 python bench.py --workload teabottle --linear-solver dense --steps 8 --no-cpu-baseline --no-exact-step > $OUT/bench_tea_dense.json 2> $OUT/bench_tea_dense.err; show $OUT/bench_tea_dense.json
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-exact-step > $OUT/bench_bal5m.json 2> $OUT/bench_bal5m.err; show $OUT/bench_bal5m.json | head -9
