#!/bin/bash
# Multi-GPU gpurun call (gpurun --gpus N): the 2-GPU parity test, then bench.py at N ranks with the
# fused peer-window PCG tail and, for comparison, with DBA_P2P=0 (ncclAllReduce per PCG iteration).
set -u
N=${1:-2}
TAG=${2:-mgpu}
WL=${3:-bal5m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
export DBA_P2P_TIMEOUT_MS=5000
( timeout 600 python -m pytest tests/test_gpu_host.py -m gpu -x -q -k two_gpu > $OUT/pytest_2gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_2gpu.log )
tail -15 $OUT/pytest_2gpu.log
run() {  # name, extra env
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 \
     bench.py --gpus $N --workload $WL > $OUT/bench_${WL}_${N}gpu_$name.json 2> $OUT/bench_${WL}_${N}gpu_$name.err
  echo "bench $name exit $?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${WL}_${N}gpu_$name.json").read().strip().splitlines()[-1])
    print("$name", d["value"], d["unit"], "ms/step", d["ms_per_step"], {k: (v["launches"], round(v["total_ms"], 2)) for k, v in d["kernels"].items() if k in ("spmv_mf", "pcg_fused", "partials_to_q", "pcg_vector")})
except Exception as e:
    print("no json", e)
PY
}
run p2p DBA_P2P=1
run nccl DBA_P2P=0
tail -5 $OUT/*.err
