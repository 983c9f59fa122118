#!/bin/bash
# the driver's own invocations: reference arm, then our arm (both as the driver runs them)
set -u
OUT=gpurun_out/${1:-r2bench}
mkdir -p $OUT
SECONDS=0; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/ref.json 2> $OUT/ref.err; echo "ref exit $?"; echo "wall ${SECONDS}s"
python - $OUT/ref.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("impl", "value", "steps", "final_cost")}, d["config"]["linear_solver"][:60], d["cpu_baseline"]["sample"][:120], d.get("cpu_baseline_dense_schur"))
PY
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/ours.json 2> $OUT/ours.err; echo "ours exit $?"; echo "wall ${SECONDS}s"
python - $OUT/ours.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "final", d["final_cost"])
print("cpu", d["cpu_baseline"]); print("cpu_same", d["cpu_baseline_same_algorithm"]); print("exact", d["exact_step"]); print("clocks", d["clocks"])
PY
