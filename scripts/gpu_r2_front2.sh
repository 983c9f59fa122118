#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2front2}
mkdir -p $OUT
for m in 0 1; do
DBA_MF_FRONT=$m timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_camera_gather|k_point_prepare|k_jacobian|k_back" -c 40 --csv --log-file $OUT/l$m.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-exact-step > $OUT/ncu$m.log 2>&1; echo "ncu list exit $?"
python scripts/launch_summary.py $OUT/l$m.csv | head -12
done
DBA_MF_FRONT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_camera_gather_mf -s 3 -c 1 -f -o $OUT/gather_mf \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-exact-step > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
