#!/bin/bash
# multi-GPU call: 2-GPU parity test (peer windows and DBA_P2P=0), then bench at N GPUs
set -u
N=${1:-2}
OUT=gpurun_out/${2:-r2mgpu}
mkdir -p $OUT
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
( timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q -k two_gpu > $OUT/pytest_mgpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_mgpu.log ); tail -4 $OUT/pytest_mgpu.log
( DBA_P2P=0 timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q -k two_gpu > $OUT/pytest_mgpu_nccl.log 2>&1; echo "pytest(nccl) exit $?" >> $OUT/pytest_mgpu_nccl.log ); tail -2 $OUT/pytest_mgpu_nccl.log
for n in $(seq 1 $N); do
  if [ $n -eq 3 ] || [ $n -eq 5 ] || [ $n -eq 6 ] || [ $n -eq 7 ]; then continue; fi
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-exact-step > $OUT/bench_bal5m_1gpu.json 2> $OUT/bench_bal5m_1gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus $n --no-cpu-baseline > $OUT/bench_bal5m_${n}gpu.json 2> $OUT/bench_bal5m_${n}gpu.err
  fi
  echo "bench $n exit $?"
  python - $OUT/bench_bal5m_${n}gpu.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["kernels"].get("spmv_mf_pcg") or {}
    print(d["n_gpus"], "gpus", round(d["value"], 1), "it/s e2e", round(d["e2e"]["value"], 1), "spmv_mf_pcg us/launch", round(1e3 * k.get("total_ms", 0) / max(k.get("launches", 1), 1), 1), "parity", d.get("parity_vs_1gpu"))
except Exception as e:
    print("no json", e)
PY
done
