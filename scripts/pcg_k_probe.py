"""GPU vs oracle at a fixed PCG iteration count K: cost after one LM iteration, per problem size and K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deeparc_sfm_b200 import capi, synthetic
from tests import oracle_lib

O = oracle_lib.Oracle()
probs = {
    "bal40": synthetic.bal_like(n_cam=40, n_pts=1500, obs_per_point=5, window=12, seed=12),
    "small200": synthetic.bal_like(n_cam=200, n_pts=50_000, obs_per_point=5, window=50, name="small"),
    "mid600": synthetic.bal_like(n_cam=600, n_pts=300_000, obs_per_point=5, window=50, name="mid"),
}
if len(sys.argv) > 1 and sys.argv[1] == "big":
    probs = {"bal5m": synthetic.bal_like(n_cam=1700, n_pts=1_000_000, obs_per_point=5, window=50, name="bal5m")}
FIXED = dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
for name, p in probs.items():
    for K in (1, 2, 5, 10, 20):
        opts = capi.make_options(max_num_iterations=1, linear_solver=capi.DBA_LS_PCG, pcg_rel_tolerance=0.0,
                                 pcg_max_iterations=K, pcg_min_iterations=0, **FIXED)
        so, _ = O.solve(p, opts)
        row = [f"{name} K={K:2d} oracle {so.final_cost:.9e} mcc {so.trace('model_cost_change')[1]:.6e}"]
        for mode in ("mf", "planes"):
            os.environ["DBA_SPMV"] = mode
            eng = capi.Engine(device=0)
            eng.problem_set(p)
            sg = eng.solve(opts)
            eng.close()
            row.append(f"{mode} {sg.final_cost:.9e} rel {abs(sg.final_cost - so.final_cost) / so.final_cost:.2e} mcc {sg.trace('model_cost_change')[1]:.6e}")
        print(" | ".join(row), flush=True)
