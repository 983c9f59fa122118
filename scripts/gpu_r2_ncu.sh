#!/bin/bash
# round 2 profiles: bench (exit 0 first), ncu launch list of the same command, full captures of the dominant kernels
set -u
OUT=gpurun_out/${1:-r2ncu}
mkdir -p $OUT
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_short.json 2> $OUT/bench_short.err; echo "bench exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_mf -s 30 -c 1 -f -o $OUT/spmv_mf \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-exact-step > $OUT/ncu_full1.log 2>&1; echo "ncu spmv_mf exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_schur_dense -s 4 -c 1 -f -o $OUT/schur_dense \
   python bench.py --workload arc1m --linear-solver dense --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full2.log 2>&1; echo "ncu schur_dense exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_ldlt_small -s 4 -c 1 -f -o $OUT/ldlt \
   python bench.py --workload arc1m --linear-solver dense --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full3.log 2>&1; echo "ncu ldlt exit $?"
ls -la $OUT
