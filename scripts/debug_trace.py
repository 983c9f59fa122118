import sys
sys.path.insert(0, '.')
import numpy as np
from deeparc_sfm_b200 import capi
from tests import oracle_lib
from tests.test_gpu_parity import PROBLEMS
np.set_printoptions(linewidth=200, precision=10)
O = oracle_lib.Oracle()
E = capi.Engine()
for name, freeze in (("rig", 1), ("rig", 0)):
    p = PROBLEMS[name].copy(); p.freeze_camera = freeze
    kw = dict(max_num_iterations=8, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    E.problem_set(p)
    sg = E.solve(capi.make_options(linear_solver=capi.DBA_LS_PCG, pcg_rel_tolerance=1e-13, pcg_max_iterations=2000, **kw))
    so, xo = O.solve(p, capi.make_options(linear_solver=capi.DBA_LS_DENSE, **kw))
    print("==", name, "freeze", freeze, sg.message, "|", so.message)
    for f in ("cost", "step_is_successful", "step_is_valid", "relative_decrease", "trust_region_radius", "model_cost_change", "step_norm", "gradient_max_norm", "gradient_norm", "linear_solver_iterations"):
        print(f, "\n  gpu", sg.trace(f), "\n  cpu", so.trace(f))
