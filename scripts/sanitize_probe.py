"""Small solves through every product path of the engine, for compute-sanitizer (memcheck / racecheck):
matrix-free product with the fused PCG tail, with the tail as its own launch, with the unfused vector
kernels, the plane product, the dense reduced system (both factorisation kernels), points-only,
the filter and the hemisphere fit.  Problems are tiny: the tools slow kernels down 10-100x."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deeparc_sfm_b200 import capi, synthetic

FIXED = dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
bal = synthetic.bal_like(n_cam=24, n_pts=700, obs_per_point=5, window=10, seed=61)
plain = synthetic.bal_like(n_cam=16, n_pts=400, obs_per_point=4, window=8, seed=62, free_intrinsics=0)
rig = synthetic.arc_rig(n_arc=3, n_ring=4, n_pts=300, obs_per_point=6, seed=63)
big_rig = synthetic.arc_rig(n_arc=6, n_ring=24, n_pts=200, obs_per_point=12, seed=64)  # 29 blocks = 174 unknowns: blocked Cholesky
runs = [("mf fused tail", bal, {}, capi.DBA_LS_PCG), ("mf tail off", bal, {"DBA_MF_TAIL": "0"}, capi.DBA_LS_PCG),
        ("mf unfused", plain, {"DBA_PCG_FUSED": "0"}, capi.DBA_LS_PCG), ("planes", bal, {"DBA_SPMV": "planes"}, capi.DBA_LS_PCG),
        ("two-pose planes", rig, {}, capi.DBA_LS_PCG), ("two-pose mf", rig, {"DBA_SPMV": "mf"}, capi.DBA_LS_PCG),
        ("dense ldlt", rig, {}, capi.DBA_LS_DENSE), ("dense cholesky", big_rig, {}, capi.DBA_LS_DENSE), ("dense 9-dof", bal, {}, capi.DBA_LS_DENSE)]
for name, p, env, ls in runs:
    for k in ("DBA_MF_TAIL", "DBA_PCG_FUSED", "DBA_SPMV"):
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = capi.Engine(device=0)
    eng.problem_set(p)
    s = eng.solve(capi.make_options(max_num_iterations=2, linear_solver=ls, pcg_rel_tolerance=0.0, pcg_max_iterations=6, **FIXED))
    eng.eval(residuals=True, jacobians=True)
    print(f"{name:18s} cost {s.initial_cost:.6e} -> {s.final_cost:.6e}  launches {s.kernel_launches}", flush=True)
    eng.close()
for k in ("DBA_MF_TAIL", "DBA_PCG_FUSED", "DBA_SPMV"):
    os.environ.pop(k, None)
eng = capi.Engine(device=0)
q = rig.copy()
q.freeze_camera = 1
eng.problem_set(q)
eng.solve(capi.make_options(max_num_iterations=2, **FIXED))
eng.problem_set(rig)
eng.filter(5.0, np.zeros(3), 1.0)
eng.filter_mse()
eng.fit_hemisphere(np.array([[0.5, 0, 0.5], [-0.5, 0, 0.5], [0, 0.5, 0.5], [0, -0.5, 0.5], [0, 0, 1.0]]))
eng.close()
print("SANITIZE PROBE DONE", flush=True)
