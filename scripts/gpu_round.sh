#!/bin/bash
# One gpurun call: GPU tests, the default bench, an ncu launch list of the same bench command and
# one `ncu --set full` capture of the dominant kernel.  Outputs land in gpurun_out/ (scratch);
# summaries worth judging are copied into profiles/ by hand afterwards.
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log )
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_bal5m.json 2> $OUT/bench_bal5m.err; echo "bench exit $?"
tail -c 600 $OUT/bench_bal5m.json
timeout 300 python bench.py --workload arc1m --no-cpu-baseline > $OUT/bench_arc1m.json 2> $OUT/bench_arc1m.err; echo "bench arc exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_mf -s 30 -c 1 -f -o $OUT/spmv_mf \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
