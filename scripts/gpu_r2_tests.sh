#!/bin/bash
# all GPU tests (+ optional extra pytest args), smoke
set -u
OUT=gpurun_out/${1:-r2t}
shift || true
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q --durations=5 "$@" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -25 $OUT/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log ); tail -2 $OUT/smoke.log
