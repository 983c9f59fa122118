#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2m}
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log ); tail -4 $OUT/pytest_gpu.log
DBA_TAIL_TRACE=1 timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-exact-step > $OUT/b1.json 2> $OUT/b1.err; grep "tail trace" $OUT/b1.err | tail -1
python - $OUT/b1.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["kernels"].get("spmv_mf_pcg") or {}
print(d["n_gpus"], "gpus", round(d["value"], 1), "it/s e2e", round(d["e2e"]["value"], 1), "spmv_mf_pcg us/iter", round(1e3 * k.get("total_ms", 0) / max(k.get("launches", 1), 1), 1), "launches", d["gpu_launches"], "roof", d["roofline"]["frac"])
PY
DBA_PCG_PERSIST=0 timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-exact-step > $OUT/b1_np.json 2> $OUT/b1_np.err
python - $OUT/b1_np.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("persist=0:", round(d["value"], 1), "it/s")
PY
