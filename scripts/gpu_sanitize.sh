#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2san}
mkdir -p $OUT
timeout 300 python scripts/sanitize_probe.py > $OUT/plain.log 2>&1; echo "plain exit $?"; tail -3 $OUT/plain.log
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_probe.py > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"; tail -5 $OUT/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_probe.py > $OUT/racecheck.log 2>&1; echo "racecheck exit $?"; tail -5 $OUT/racecheck.log
timeout 1500 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_probe.py > $OUT/synccheck.log 2>&1; echo "synccheck exit $?"; tail -5 $OUT/synccheck.log
