#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2dense10}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_dense|k_pair|k_camera_gather|k_jacobian|k_back|k_point" -c 100 --csv --log-file $OUT/l.csv \
   python bench.py --workload arc1m --linear-solver dense --steps 4 --warmup 1 --no-cpu-baseline --no-exact-step > $OUT/ncu.log 2>&1; echo "ncu list exit $?"
python scripts/launch_summary.py $OUT/l.csv | head -16
for k in k_dense_z; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/$k \
   python bench.py --workload arc1m --linear-solver dense --steps 4 --warmup 1 --no-cpu-baseline --no-exact-step > $OUT/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
