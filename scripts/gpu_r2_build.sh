#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2build}
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log ); tail -25 $OUT/pytest.log
for mode in host device; do
  DBA_BUILD=$mode DBA_TIMING=1 timeout 300 python scripts/e2e_probe.py > $OUT/e2e_$mode.log 2>&1; grep "problem_set " $OUT/e2e_$mode.log | tail -2; grep "device\]" $OUT/e2e_$mode.log | tail -4
done
