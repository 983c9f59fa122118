"""GPU tests of the explicit reduced system + device Cholesky (DBA_LS_DENSE), the engine's
counterpart of the reference's `options.linear_solver_type = ceres::DENSE_SCHUR`
(reference src/sfm.cc:67, :95), against the CPU oracle's exact DENSE_SCHUR step
(oracle/mini_ceres.cc::EliminateAndSolve).  No PCG anywhere in these runs: the step of every
LM iteration is the exact solution of the damped normal equations on both sides."""
import numpy as np
import pytest

from deeparc_sfm_b200 import capi, synthetic
from tests.test_gpu_parity import PROBLEMS, _compare_solve, _perturbed_bal

pytestmark = pytest.mark.gpu

FIXED = dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)


@pytest.mark.parametrize("name,n_iter", [("rig", 5), ("plain", 6), ("bal", 6), ("nf2_nd2", 4), ("small_angle", 4), ("nd1", 4)])
def test_dense_schur_matches_oracle(engine, oracle, name, n_iter):
    """Every problem shape: composed arc o ring poses with the gauge block constant (19 blocks),
    6-dof plain poses, 9-dof cameras, nf = 2 / nd = 2, the small-angle branch."""
    sg, so = _compare_solve(engine, oracle, PROBLEMS[name], n_iter=n_iter, linear_solver=capi.DBA_LS_DENSE)
    assert sg.linear_solver_used == capi.DBA_LS_DENSE == so.linear_solver_used
    assert sg.pcg_iterations_total == n_iter and sg.linear_solver_failures == 0
    assert np.all(sg.trace("linear_solver_iterations")[1:] == 1)
    # an exact step on both sides: the traces agree far inside the 1e-6 budget
    np.testing.assert_allclose(sg.trace("cost"), so.trace("cost"), rtol=1e-9)
    np.testing.assert_allclose(sg.trace("model_cost_change")[1:], so.trace("model_cost_change")[1:], rtol=1e-7)


def test_dense_schur_with_rejected_and_invalid_steps(engine, oracle):
    """Trust-region rejections with the dense solver (40 cameras x 9 = 360 unknowns)."""
    sg, so = _compare_solve(engine, oracle, _perturbed_bal(0.2, 0.05), n_iter=7, linear_solver=capi.DBA_LS_DENSE)
    assert 0 < sg.num_unsuccessful_steps == so.num_unsuccessful_steps


def test_dense_schur_points_only_and_default_tolerances(engine, oracle):
    """freeze_camera: no camera side at all (the reduced system is empty); then the driver's own
    call: 100 iterations, Ceres default tolerances, DBA_LS_AUTO -> DENSE for a 19-block rig."""
    p = PROBLEMS["rig"].copy()
    p.freeze_camera = 1
    _compare_solve(engine, oracle, p, n_iter=3, linear_solver=capi.DBA_LS_DENSE)
    p = PROBLEMS["rig"]
    engine.problem_set(p)
    sg = engine.solve(capi.make_options(max_num_iterations=100, linear_solver=capi.DBA_LS_AUTO))
    so, xo = oracle.solve(p, capi.make_options(max_num_iterations=100, linear_solver=capi.DBA_LS_DENSE))
    assert sg.linear_solver_used == capi.DBA_LS_DENSE
    assert sg.termination == so.termination == capi.DBA_CONVERGENCE
    assert sg.num_iterations == so.num_iterations
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    xg = engine.params_get()
    for k in ("pts", "ext_rot", "ext_trans"):
        assert np.max(np.abs(xg[k] - xo[k])) <= 1e-6 * np.max(np.abs(xo[k])), k


def test_dense_equals_converged_pcg(engine):
    """The same LM trace from the exact factorisation and from the PCG driven to 1e-13."""
    for name in ("rig", "bal"):
        p = PROBLEMS[name]
        runs = {}
        for ls in (capi.DBA_LS_DENSE, capi.DBA_LS_PCG):
            engine.problem_set(p)
            s = engine.solve(capi.make_options(max_num_iterations=5, linear_solver=ls, pcg_rel_tolerance=1e-13,
                                               pcg_max_iterations=3000, **FIXED))
            runs[ls] = (s, engine.params_get())
        (sa, xa), (sb, xb) = runs[capi.DBA_LS_DENSE], runs[capi.DBA_LS_PCG]
        assert sb.linear_solver_used == capi.DBA_LS_PCG and sb.pcg_unconverged_solves == 0
        np.testing.assert_allclose(sa.trace("cost"), sb.trace("cost"), rtol=1e-8)
        for k in ("pts", "ext_rot", "ext_trans"):
            assert np.max(np.abs(xa[k] - xb[k])) <= 1e-7 * np.max(np.abs(xa[k])), (name, k)


def test_dense_limits_are_enforced_not_papered_over(engine):
    """200 cameras x 9 = 1800 reduced unknowns: DENSE is refused (never a silent PCG), AUTO runs
    the PCG and says so; a PCG that cannot reach its tolerance is counted."""
    p = synthetic.bal_like(n_cam=200, n_pts=4000, obs_per_point=5, window=30, seed=41)
    engine.problem_set(p)
    with pytest.raises(capi.EngineError) as e:
        engine.solve(capi.make_options(max_num_iterations=2, linear_solver=capi.DBA_LS_DENSE))
    assert e.value.status == capi.DBA_ERR_UNSUPPORTED
    s = engine.solve(capi.make_options(max_num_iterations=2, linear_solver=capi.DBA_LS_AUTO, **FIXED))
    assert s.linear_solver_used == capi.DBA_LS_PCG and s.reduced_system_size == 1800
    engine.params_reset()
    s = engine.solve(capi.make_options(max_num_iterations=2, linear_solver=capi.DBA_LS_PCG, pcg_rel_tolerance=1e-14,
                                       pcg_max_iterations=3, **FIXED))
    assert s.pcg_unconverged_solves == 2
    # AUTO honours dense_max_size: a 40-camera problem (360 unknowns) with the limit at 100 runs the PCG
    engine.problem_set(PROBLEMS["bal"])
    s = engine.solve(capi.make_options(max_num_iterations=2, linear_solver=capi.DBA_LS_AUTO, dense_max_size=100, **FIXED))
    assert s.linear_solver_used == capi.DBA_LS_PCG
    engine.params_reset()
    s = engine.solve(capi.make_options(max_num_iterations=2, linear_solver=capi.DBA_LS_AUTO, **FIXED))
    assert s.linear_solver_used == capi.DBA_LS_DENSE


@pytest.mark.parametrize("kind", ["rig_long", "teabottle", "many_blocks"])
def test_dense_schur_shapes(engine, oracle, kind):
    """Long tracks (a point seen by 300 cameras of a 65-block rig: one point per batch), the
    teabottle stand-in (50 pose blocks = 300 unknowns, several pair groups) and 120 blocks of
    6-dof poses (720 unknowns, 7260 block pairs)."""
    if kind == "rig_long":
        p = synthetic.arc_rig(n_arc=6, n_ring=60, n_pts=40, obs_per_point=300, seed=17)
    elif kind == "teabottle":
        p = synthetic.teabottle_like(n_pts=1500, obs_per_point=8)
    else:
        p = synthetic.bal_like(n_cam=120, n_pts=3000, obs_per_point=5, window=25, seed=43, free_intrinsics=0)
    sg, so = _compare_solve(engine, oracle, p, n_iter=3, linear_solver=capi.DBA_LS_DENSE)
    assert sg.linear_solver_used == capi.DBA_LS_DENSE


# ------------------------------------------------------ device-resident outer loop (dba_problem_update)
def _filtered_copy(p, x, obs_remove, pt_remove, freeze):
    """What a caller that re-gathers the scene would upload after the filter: survivors in their
    original order, points re-indexed, parameters = the values the first solve left."""
    keep_pt = pt_remove == 0
    keep_ob = (obs_remove == 0) & keep_pt[p.obs_pt]
    new_pt = np.cumsum(keep_pt) - 1
    q = p.copy()
    for k in ("obs_xy", "obs_pose_a", "obs_pose_b", "obs_intr"):
        setattr(q, k, getattr(p, k)[keep_ob])
    q.obs_pt = new_pt[p.obs_pt[keep_ob]].astype(np.int32)
    q.pts = x["pts"][keep_pt]
    q.ext_rot, q.ext_trans = x["ext_rot"].copy(), x["ext_trans"].copy()
    q.intr_focal, q.intr_dist = x["intr_focal"].copy(), x["intr_dist"].copy()
    q.pts_rgb = None
    q.obs_col0 = q.obs_col1 = None
    q.freeze_camera = freeze
    return q


@pytest.mark.parametrize("name", ["rig", "plain", "bal"])
def test_problem_update_equals_reupload_bit_for_bit(engine, name):
    """The reference's outer loop (sfm.cc:111-127): points-only solve, filterPoint3d, full solve, filter, ...
    With dba_problem_update the engine drops the flagged observations / points itself and keeps the
    parameters on the device; the following solve is bit-identical to the one after gathering and
    uploading the filtered scene again — same structures, same summation order."""
    p = PROBLEMS[name].copy()
    rng = np.random.default_rng(21)
    order = rng.permutation(p.n_obs)  # a scene graph in file order, not point-sorted
    for k in ("obs_xy", "obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr"):
        setattr(p, k, getattr(p, k)[order].copy())
    p.freeze_camera = 1
    opts = capi.make_options(max_num_iterations=3, linear_solver=capi.DBA_LS_AUTO, **FIXED)
    engine.problem_set(p)
    engine.solve(opts)
    x1 = engine.params_get()
    mse = engine.filter_mse()
    boundary = float(np.quantile(mse, 0.15))
    centre = x1["pts"].mean(axis=0)
    rho = 2.0 * float(np.quantile(((x1["pts"] - centre) ** 2).sum(axis=1), 0.9))
    obs_rm, pt_rm = engine.filter(boundary, centre, rho)
    assert 0 < obs_rm.sum() < p.n_obs and 0 < pt_rm.sum() < p.n_pts
    # path A: engine-side compaction, cameras released for the second solve
    n_obs, n_pts = engine.problem_update(obs_rm, pt_rm, freeze_camera=0)
    ca = engine.eval(residuals=True)
    sa = engine.solve(opts)
    xa = engine.params_get()
    fa = engine.filter(boundary, centre, rho)  # caller-order outputs of the compacted problem
    # path B: the caller gathers the filtered scene and uploads it
    q = _filtered_copy(p, x1, obs_rm, pt_rm, 0)
    assert (q.n_obs, q.n_pts) == (n_obs, n_pts)
    engine.problem_set(q)
    cb = engine.eval(residuals=True)
    sb = engine.solve(opts)
    xb = engine.params_get()
    fb = engine.filter(boundary, centre, rho)
    assert ca["cost"] == cb["cost"] and np.array_equal(ca["residuals"], cb["residuals"])
    assert np.array_equal(sa.trace("cost"), sb.trace("cost")) and sa.num_iterations == sb.num_iterations
    for k in xa:
        assert np.array_equal(xa[k], xb[k]), k
    assert np.array_equal(fa[0], fb[0]) and np.array_equal(fa[1], fb[1])
    assert sa.final_cost < sa.initial_cost


# ------------------------------------------------------------ device-side problem construction
@pytest.mark.parametrize("name,freeze,ls", [("bal", 0, capi.DBA_LS_PCG), ("bal", 0, capi.DBA_LS_DENSE), ("plain", 0, capi.DBA_LS_PCG),
                                            ("plain", 1, capi.DBA_LS_AUTO), ("nd1", 0, capi.DBA_LS_PCG), ("small_angle", 0, capi.DBA_LS_DENSE)])
def test_device_build_equals_host_build(engine, monkeypatch, name, freeze, ls):
    """DBA_BUILD=device builds the per-observation records, point offsets, tile incidence, partial grouping and
    camera-sorted chunks on the GPU (ba_build.cu) instead of the host cores.  Same structures up to the order
    of a tile's camera blocks, which no sum depends on: residuals, traces, parameters and filter decisions are
    bit-identical; the engine-side compaction (dba_problem_update) works from the device-built image too."""
    p = PROBLEMS[name].copy()
    p.freeze_camera = freeze
    opts = capi.make_options(max_num_iterations=4, linear_solver=ls, pcg_rel_tolerance=0.0, pcg_max_iterations=15, **FIXED)
    runs = {}
    for mode in ("host", "device"):
        monkeypatch.setenv("DBA_BUILD", mode)
        engine.problem_set(p)
        e = engine.eval(residuals=True, jacobians=True)
        s = engine.solve(opts)
        x = engine.params_get()
        mse = engine.filter_mse()
        fl = engine.filter(float(np.quantile(mse, 0.2)))
        n2 = engine.problem_update(fl[0], fl[1], freeze_camera=0)
        s2 = engine.solve(opts)
        runs[mode] = (e, s, x, mse, fl, n2, s2, engine.params_get())
    monkeypatch.delenv("DBA_BUILD", raising=False)
    (ea, sa, xa, ma, fa, na, sa2, xa2), (eb, sb, xb, mb, fb, nb2, sb2, xb2) = runs["host"], runs["device"]
    for k in ("residuals", "jac_pt", "jac_pose_a", "jac_intr"):
        assert np.array_equal(ea[k], eb[k]), k
    assert ea["cost"] == eb["cost"]
    assert np.array_equal(sa.trace("cost"), sb.trace("cost")) and sa.kernel_launches == sb.kernel_launches
    for k in xa:
        assert np.array_equal(xa[k], xb[k]), k
    assert np.array_equal(ma, mb) and np.array_equal(fa[0], fb[0]) and np.array_equal(fa[1], fb[1]) and na == nb2
    assert np.array_equal(sa2.trace("cost"), sb2.trace("cost"))
    for k in xa2:
        assert np.array_equal(xa2[k], xb2[k]), k


def test_device_build_falls_back_for_unsorted_and_composed_input(engine, oracle, monkeypatch):
    """Anything the device build does not take (observations not sorted by point, composed arc o ring poses,
    an index out of range) goes through the host build, with the same results / the same error."""
    monkeypatch.setenv("DBA_BUILD", "device")
    p = PROBLEMS["plain"].copy()
    perm = np.random.default_rng(4).permutation(p.n_obs)
    for k in ("obs_xy", "obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr"):
        setattr(p, k, getattr(p, k)[perm].copy())
    _compare_solve(engine, oracle, p, n_iter=3)
    _compare_solve(engine, oracle, PROBLEMS["rig"], n_iter=3, linear_solver=capi.DBA_LS_DENSE)
    q = PROBLEMS["plain"].copy()
    q.obs_pose_a = q.obs_pose_a.copy()
    q.obs_pose_a[5] = q.n_ext + 3
    with pytest.raises(capi.EngineError) as e:
        engine.problem_set(q)
    assert e.value.status == capi.DBA_ERR_INVALID_ARGUMENT
    monkeypatch.delenv("DBA_BUILD", raising=False)


# ------------------------------------------------------------ plane-less front half of matrix-free solves
@pytest.mark.parametrize("name,ls", [("bal", capi.DBA_LS_PCG), ("plain", capi.DBA_LS_PCG), ("small_angle", capi.DBA_LS_PCG),
                                     ("nd1", capi.DBA_LS_PCG), ("bal", capi.DBA_LS_DENSE)])
def test_recomputed_front_half_equals_plane_reader(engine, monkeypatch, name, ls):
    """Single-pose PCG solves keep no camera planes: the camera-side gather (k_camera_gather_mf) and the
    back-substitution (k_back_substitute_mf) recompute F from the camera row, the Jacobian kernel writes r and E
    only.  DBA_MF_FRONT=0 restores the plane readers: same algorithm, so the traces agree to rounding
    (the recomputed F differs from the stored one by an ulp or two: fused multiply-adds contract differently,
    and F x goes through the geometric form of the product)."""
    p = PROBLEMS[name]
    opts = capi.make_options(max_num_iterations=5, linear_solver=ls, pcg_rel_tolerance=0.0, pcg_max_iterations=25, **FIXED)
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DBA_MF_FRONT", mode)
        engine.problem_set(p)
        s = engine.solve(opts)
        runs[mode] = (s, engine.params_get())
    monkeypatch.delenv("DBA_MF_FRONT", raising=False)
    engine.problem_set(p)  # leave the engine in its default mode
    (sa, xa), (sb, xb) = runs["0"], runs["1"]
    assert np.array_equal(sa.trace("step_is_successful"), sb.trace("step_is_successful"))
    np.testing.assert_allclose(sb.trace("cost"), sa.trace("cost"), rtol=1e-11)
    np.testing.assert_allclose(sb.trace("gradient_max_norm"), sa.trace("gradient_max_norm"), rtol=1e-9)
    for k in xa:
        np.testing.assert_allclose(xb[k], xa[k], rtol=1e-8, atol=1e-10, err_msg=k)


@pytest.mark.parametrize("n_cam", [2, 3, 9, 25, 26, 33])
def test_dense_factorisation_sizes(engine, oracle, n_cam):
    """Reduced systems of 12, 18, 54, 150, 156 and 198 unknowns (6-dof poses, first pose constant by the
    gauge rule of solve() where the generator says so): the blocked LDL^T of k_dense_ldlt_small with a full
    panel + a short one, a whole number of panels + the right-hand-side row, its largest size
    class (n <= 152), and the first sizes of k_dense_cholesky."""
    p = synthetic.bal_like(n_cam=n_cam, n_pts=max(60, 12 * n_cam), obs_per_point=min(4, n_cam), window=min(8, n_cam),
                           seed=100 + n_cam, free_intrinsics=0)
    sg, so = _compare_solve(engine, oracle, p, n_iter=3, linear_solver=capi.DBA_LS_DENSE)
    assert sg.linear_solver_used == capi.DBA_LS_DENSE and sg.linear_solver_failures == 0
    np.testing.assert_allclose(sg.trace("cost"), so.trace("cost"), rtol=1e-9)
