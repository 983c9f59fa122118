"""GPU tests of the robust loss (DBA_LOSS_CAUCHY): the reference builds its problem with a NULL loss
(reference src/sfm.cc:48) and keeps `new ceres::CauchyLoss(0.5)` in a comment (:49).  The engine
applies Ceres' corrector inside the Jacobian / cost kernels; compared here with the CPU oracle's
restatement of loss_function.cc + corrector.cc on problems with gross outliers."""
import numpy as np
import pytest

from deeparc_sfm_b200 import capi, synthetic
from tests.test_gpu_parity import PROBLEMS

pytestmark = pytest.mark.gpu

FIXED = dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)


def _with_outliers(p, frac=0.03, sigma=60.0, seed=5):
    q = p.copy()
    rng = np.random.default_rng(seed)
    bad = rng.choice(q.n_obs, size=max(int(frac * q.n_obs), 1), replace=False)
    q.obs_xy[bad] += rng.normal(0.0, sigma, size=(bad.size, 2))
    return q


@pytest.mark.parametrize("name,ls,a", [("rig", capi.DBA_LS_DENSE, 0.5), ("rig", capi.DBA_LS_PCG, 2.0), ("bal", capi.DBA_LS_DENSE, 1.0),
                                       ("bal", capi.DBA_LS_PCG, 0.5), ("plain", capi.DBA_LS_PCG, 3.0)])
def test_cauchy_loss_matches_oracle(engine, oracle, name, ls, a):
    p = _with_outliers(PROBLEMS[name])
    n_iter = 5
    og = capi.make_options(max_num_iterations=n_iter, linear_solver=ls, pcg_rel_tolerance=1e-13, pcg_max_iterations=3000,
                           loss_type=capi.DBA_LOSS_CAUCHY, loss_scale=a, **FIXED)
    oo = capi.make_options(max_num_iterations=n_iter, linear_solver=capi.DBA_LS_DENSE, loss_type=capi.DBA_LOSS_CAUCHY,
                           loss_scale=a, **FIXED)
    engine.problem_set(p)
    raw_cost = engine.eval(residuals=False)["cost"]
    sg = engine.solve(og)
    xg = engine.params_get()
    so, xo = oracle.solve(p, oo)
    assert sg.initial_cost < raw_cost  # 1/2 sum rho(s) < 1/2 sum s: the loss is really applied
    assert abs(sg.initial_cost - so.initial_cost) <= 1e-10 * so.initial_cost
    assert np.array_equal(sg.trace("step_is_successful"), so.trace("step_is_successful"))
    np.testing.assert_allclose(sg.trace("cost"), so.trace("cost"), rtol=1e-6)
    np.testing.assert_allclose(sg.trace("trust_region_radius"), so.trace("trust_region_radius"), rtol=1e-5)
    for k in ("pts", "ext_rot", "ext_trans", "intr_focal", "intr_dist"):
        scale = max(np.max(np.abs(xo[k])), 1e-300)
        assert np.max(np.abs(xg[k] - xo[k])) <= 1e-6 * scale, k
    # dba_eval keeps reporting the plain functor (filterPoint3d re-evaluates it without a loss)
    assert engine.eval(residuals=False)["cost"] > sg.final_cost


def test_cauchy_loss_downweights_outliers(engine):
    """With 3 % gross outliers the plain least-squares fit is pulled away from the generating parameters;
    the Cauchy fit stays with the inliers (median inlier residual back at the 0.5 px noise level)."""
    p = _with_outliers(synthetic.bal_like(n_cam=60, n_pts=4000, obs_per_point=6, window=15, seed=77, free_intrinsics=0))
    clean = synthetic.bal_like(n_cam=60, n_pts=4000, obs_per_point=6, window=15, seed=77, free_intrinsics=0)
    inlier = np.all(p.obs_xy == clean.obs_xy, axis=1)
    med = {}
    for loss in (capi.DBA_LOSS_NONE, capi.DBA_LOSS_CAUCHY):
        engine.problem_set(p)
        engine.solve(capi.make_options(max_num_iterations=30, loss_type=loss, loss_scale=1.0))
        r = engine.eval(residuals=True)["residuals"]
        med[loss] = float(np.median(np.linalg.norm(r[inlier], axis=1)))
    assert med[capi.DBA_LOSS_CAUCHY] < 0.8 and med[capi.DBA_LOSS_CAUCHY] < 0.7 * med[capi.DBA_LOSS_NONE], med


def test_loss_option_validation(engine):
    engine.problem_set(PROBLEMS["plain"])
    for kw in (dict(loss_type=7), dict(loss_type=capi.DBA_LOSS_CAUCHY, loss_scale=0.0)):
        with pytest.raises(capi.EngineError) as e:
            engine.solve(capi.make_options(max_num_iterations=1, **kw))
        assert e.value.status == capi.DBA_ERR_INVALID_ARGUMENT
