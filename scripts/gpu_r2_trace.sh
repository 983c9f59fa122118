#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2trace}
mkdir -p $OUT
export DBA_TAIL_TRACE=1
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-exact-step > $OUT/b1.json 2> $OUT/b1.err; grep "tail trace" $OUT/b1.err | tail -2
for n in 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus $n --no-cpu-baseline > $OUT/b$n.json 2> $OUT/b$n.err; grep "tail trace" $OUT/b$n.err | sort | tail -$n | cut -c1-260
done
