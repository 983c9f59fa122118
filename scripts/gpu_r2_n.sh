#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2n}
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log ); tail -3 $OUT/pytest_gpu.log
for mode in 1 0; do
DBA_PCG_PERSIST=$mode DBA_TAIL_TRACE=1 timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-exact-step > $OUT/b1_$mode.json 2> $OUT/b1_$mode.err; grep "tail trace" $OUT/b1_$mode.err | tail -1
python - $OUT/b1_$mode.json $mode <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["kernels"].get("spmv_mf_pcg") or {}
print("persist", sys.argv[2], round(d["value"], 1), "it/s e2e", round(d["e2e"]["value"], 1), "spmv_mf_pcg us/iter", round(1e3 * k.get("total_ms", 0) / max(k.get("launches", 1), 1), 1), "launches", d["gpu_launches"], "roof", round(d["roofline"]["frac"], 3))
PY
done
