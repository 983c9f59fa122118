#!/bin/bash
# ncu --set full captures of the once-per-LM-iteration kernels (one launch each)
set -u
OUT=gpurun_out/${1:-ncu2}
mkdir -p $OUT
for K in k_jacobian_tile k_camera_gather k_point_prepare k_back_substitute; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o $OUT/$K \
     python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/$K.log 2>&1; echo "$K exit $?"
done
ls -la $OUT
