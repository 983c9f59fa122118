#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2front7}
mkdir -p $OUT
for k in k_back_substitute_mf k_jacobian_tile k_point_prepare; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/$k \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-exact-step > $OUT/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
