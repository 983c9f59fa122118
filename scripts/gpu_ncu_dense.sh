#!/bin/bash
set -u
OUT=gpurun_out/r2ncu_dense
mkdir -p $OUT
WL=${1:-arc1m}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_schur_dense -s 4 -c 1 -f -o $OUT/schur_dense_$WL \
   python bench.py --workload $WL --linear-solver dense --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu1.log 2>&1; echo "ncu1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_cholesky -s 4 -c 1 -f -o $OUT/cholesky_$WL \
   python bench.py --workload $WL --linear-solver dense --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu2.log 2>&1; echo "ncu2 exit $?"
ls -la $OUT
