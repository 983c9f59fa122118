"""GPU parity at the sizes BASELINE.json's configs name — the engine through the C ABI against the
CPU oracle on the WHOLE problem, not against size-independent properties only:

  configs[3] bal5m      1.7k cameras, 1M points, 5M observations, 9-dof cameras: every residual vs
                        the oracle; two LM iterations in the very mode bench.py times (implicit Schur,
                        block-Jacobi PCG, exactly K = 20 iterations, tolerance 0) against the oracle's
                        ITERATIVE_SCHUR extension run with the same K (oracle/mini_ceres.cc::SolveImplicit)
  configs[2] arc1m      100 cameras (19 pose blocks), 100k points, 1M observations, hemisphere term on:
                        every residual, three LM iterations with the exact DENSE_SCHUR step on both
                        sides, and the hemisphere fit of the rig's camera centres
  configs[0..1]         teabottle stand-in (the files are absent from the reference mount): residuals,
                        four DENSE_SCHUR iterations
  configs[4] stress50m  10k cameras, 10M points, 50M observations (the oracle would need minutes and
                        ~100 GB): cost vs an independent numpy evaluation, and the shard-sum identity —
                        the costs of the eight point shards of dba_shard_plan, each evaluated as its own
                        problem, add up to the cost of the whole.

Tolerances are those of tests/test_gpu_parity.py (north_star): residuals 1e-10 relative (+ the
rounding of `predicted - observed`), cost traces 1e-6, parameters 1e-6 relative per array.
"""
import numpy as np
import pytest

from deeparc_sfm_b200 import capi, synthetic
from tests.test_gpu_parity import _check_residuals

pytestmark = pytest.mark.gpu

FIXED = dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)


def _compare_runs(sg, xg, so, xo, cost_rtol=1e-6):
    assert sg.num_iterations == so.num_iterations
    assert np.array_equal(sg.trace("step_is_successful"), so.trace("step_is_successful"))
    np.testing.assert_allclose(sg.trace("cost"), so.trace("cost"), rtol=cost_rtol)
    np.testing.assert_allclose(sg.trace("trust_region_radius"), so.trace("trust_region_radius"), rtol=1e-5)
    assert abs(sg.final_cost - so.final_cost) <= 1e-6 * so.final_cost
    for k in ("pts", "ext_rot", "ext_trans", "intr_focal", "intr_dist"):
        scale = max(np.max(np.abs(xo[k])), 1e-300)
        assert np.max(np.abs(xg[k] - xo[k])) <= 1e-6 * scale, f"{k}: {np.max(np.abs(xg[k] - xo[k])) / scale:.3e}"


def test_bal5m_full_size_vs_oracle_same_k(engine, oracle):
    p = synthetic.bal_like(n_cam=1700, n_pts=1_000_000, obs_per_point=5, window=50, name="bal5m")
    engine.problem_set(p)
    g = engine.eval(residuals=True, jacobians=False)
    o = oracle.eval(p, residuals=True, jacobians=False)
    _check_residuals(g["residuals"], o["residuals"], p.obs_xy, p)
    assert abs(g["cost"] - o["cost"]) <= 1e-12 * o["cost"]
    # the timed mode of bench.py: K = 20 PCG iterations per LM iteration, no tolerance, on both sides
    opts = capi.make_options(max_num_iterations=2, linear_solver=capi.DBA_LS_PCG, pcg_rel_tolerance=0.0,
                             pcg_max_iterations=20, pcg_min_iterations=0, **FIXED)
    sg = engine.solve(opts)
    xg = engine.params_get()
    so, xo = oracle.solve(p, opts)
    assert so.linear_solver_used == capi.DBA_LS_PCG == sg.linear_solver_used
    assert np.all(sg.trace("linear_solver_iterations")[1:] == 20) and np.all(so.trace("linear_solver_iterations")[1:] == 20)
    _compare_runs(sg, xg, so, xo)
    # same inexact step: the model cost change of each iteration agrees as well
    np.testing.assert_allclose(sg.trace("model_cost_change")[1:], so.trace("model_cost_change")[1:], rtol=1e-6)


def _camera_centres(p):
    """C = -R_a^T t_a - R_a^T R_b^T t_b per distinct (arc, ring) camera (reference
    src/DeepArcManager.cc:501-518), from the flat problem."""
    cams = np.unique(np.stack([p.obs_pose_a, p.obs_pose_b], axis=1), axis=0)
    out = []
    for a, b in cams:
        Ra = synthetic.rodrigues(p.ext_rot[a][None])[0]
        c = -Ra.T @ p.ext_trans[a]
        if b >= 0:
            Rb = synthetic.rodrigues(p.ext_rot[b][None])[0]
            c = c - Ra.T @ (Rb.T @ p.ext_trans[b])
        out.append(c)
    return np.array(out)


def test_arc1m_full_size_vs_oracle_dense_schur(engine, oracle):
    p = synthetic.arc_rig(n_arc=10, n_ring=10, n_pts=100_000, obs_per_point=10, name="arc1m")
    assert p.n_obs == 1_000_000 and p.n_ext == 19
    engine.problem_set(p)
    g = engine.eval(residuals=True, jacobians=False)
    o = oracle.eval(p, residuals=True, jacobians=False)
    _check_residuals(g["residuals"], o["residuals"], p.obs_xy, p)
    assert abs(g["cost"] - o["cost"]) <= 1e-12 * o["cost"]
    opts = capi.make_options(max_num_iterations=3, linear_solver=capi.DBA_LS_DENSE, **FIXED)
    sg = engine.solve(opts)
    xg = engine.params_get()
    so, xo = oracle.solve(p, opts)
    assert sg.linear_solver_used == capi.DBA_LS_DENSE and sg.linear_solver_failures == 0
    _compare_runs(sg, xg, so, xo, cost_rtol=1e-8)
    # hemisphere term on (reference src/sfm.cc:86-103): sphere through the camera centres
    centres = _camera_centres(p)
    assert centres.shape == (100, 3)
    cg, rg, hg = engine.fit_hemisphere(centres)
    co, ro, ho = oracle.fit_hemisphere(centres)
    assert hg.termination == ho.termination and hg.num_iterations == ho.num_iterations
    np.testing.assert_allclose(cg, co, rtol=1e-8, atol=1e-12)
    assert abs(rg - ro) <= 1e-8 * abs(ro)


def test_teabottle_standin_full_size_vs_oracle_dense_schur(engine, oracle):
    p = synthetic.teabottle_like(n_pts=20_000, obs_per_point=8)
    engine.problem_set(p)
    g = engine.eval(residuals=True, jacobians=False)
    o = oracle.eval(p, residuals=True, jacobians=False)
    _check_residuals(g["residuals"], o["residuals"], p.obs_xy, p)
    opts = capi.make_options(max_num_iterations=4, linear_solver=capi.DBA_LS_DENSE, **FIXED)
    sg = engine.solve(opts)
    xg = engine.params_get()
    so, xo = oracle.solve(p, opts)
    assert sg.linear_solver_used == capi.DBA_LS_DENSE and sg.reduced_system_size == 300
    _compare_runs(sg, xg, so, xo, cost_rtol=1e-8)


def _slice_points(p, lo, hi):
    """The sub-problem of points [lo, hi) with all cameras (what one rank of a sharded job owns)."""
    sel = (p.obs_pt >= lo) & (p.obs_pt < hi)
    return synthetic.Problem(
        obs_xy=p.obs_xy[sel], obs_pt=(p.obs_pt[sel] - lo).astype(np.int32), obs_pose_a=p.obs_pose_a[sel],
        obs_pose_b=p.obs_pose_b[sel], obs_intr=p.obs_intr[sel], pts=p.pts[lo:hi], ext_rot=p.ext_rot,
        ext_trans=p.ext_trans, intr_center=p.intr_center, intr_focal=p.intr_focal, intr_dist=p.intr_dist,
        intr_nf=p.intr_nf, intr_nd=p.intr_nd, ext_const=p.ext_const, freeze_camera=p.freeze_camera,
        free_intrinsics=p.free_intrinsics)


def test_stress50m_cost_and_shard_sum(engine):
    p = synthetic.bal_like(n_cam=10_000, n_pts=10_000_000, obs_per_point=5, window=50, name="stress50m")
    assert p.n_obs == 50_000_000
    engine.problem_set(p)
    c_gpu = engine.eval(residuals=False)["cost"]
    r = synthetic.project(p) - p.obs_xy
    per_obs = np.sum(r.astype(np.longdouble) ** 2, axis=1)
    c_np = 0.5 * float(np.sum(per_obs))
    assert abs(c_gpu - c_np) <= 1e-10 * c_np, (c_gpu, c_np)
    # one LM iteration of the timed mode runs and decreases the cost (50M observations through every kernel)
    s = engine.solve(capi.make_options(max_num_iterations=1, linear_solver=capi.DBA_LS_PCG, pcg_rel_tolerance=0.0,
                                       pcg_max_iterations=20, **FIXED))
    assert s.num_successful_steps == 1 and s.final_cost < 0.1 * s.initial_cost
    assert abs(s.initial_cost - c_np) <= 1e-10 * c_np
    # shard-sum identity over the 8-rank point sharding of dba_shard_plan
    pt_begin, obs_count = capi.shard_plan(p, 8)
    assert pt_begin[0] == 0 and pt_begin[-1] == p.n_pts and int(obs_count.sum()) == p.n_obs
    assert obs_count.max() - obs_count.min() <= 2 * 5  # balanced by observation count to within a point or two
    total = 0.0
    for k in range(8):
        q = _slice_points(p, int(pt_begin[k]), int(pt_begin[k + 1]))
        assert q.n_obs == obs_count[k]
        engine.problem_set(q)
        c_k = engine.eval(residuals=False)["cost"]
        lo = np.searchsorted(p.obs_pt, pt_begin[k], side="left")
        hi = np.searchsorted(p.obs_pt, pt_begin[k + 1], side="left")
        c_k_np = 0.5 * float(np.sum(per_obs[lo:hi]))
        assert abs(c_k - c_k_np) <= 1e-10 * c_k_np
        total += c_k
    assert abs(total - c_gpu) <= 1e-12 * c_gpu, (total, c_gpu)
