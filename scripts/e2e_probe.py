import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
from deeparc_sfm_b200 import capi
p = bench.build_workload('bal5m')
eng = capi.Engine(device=0)
out = None
for i in range(3):
    t0 = time.perf_counter(); eng.problem_set(p); t1 = time.perf_counter()
    s = eng.solve(bench.solve_options(capi, 10, 20)); t2 = time.perf_counter()
    out = eng.params_get(out=out); t3 = time.perf_counter()
    print(f"problem_set {1e3*(t1-t0):.1f} ms  solve {1e3*(t2-t1):.1f} ms  get {1e3*(t3-t2):.1f} ms  nproc {os.cpu_count()}", file=sys.stderr)
eng.close()
