"""CPU tests (no GPU): pin the oracle — and the engine's analytic math compiled for the host —
against the 50-digit golden vectors, the reference's own compiled functor (oracle/_ref, when
built), scipy, and closed-form answers.  The reference ships no tests; these are
what anchors the oracle that the GPU parity tests then compare the CUDA path with."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.optimize
from scipy.spatial.transform import Rotation

from deeparc_sfm_b200 import capi, synthetic
from tests import golden_util, oracle_lib

EPS = np.finfo(np.float64).eps
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_residuals(r, truth, obs_xy):
    pred = np.abs(truth + obs_xy)
    err = np.abs(r - truth)
    assert np.all(err <= 1e-10 * np.abs(truth) + 64 * EPS * pred), float(np.max(err))


def _check_jac(j, truth, rtol, mask=None):
    scale = np.max(np.abs(truth), axis=(1, 2), keepdims=True)
    ok = np.abs(j - truth) <= rtol * scale + 1e-300
    if mask is not None:
        ok = ok | ~mask[:, None, None]
    assert np.all(ok), float(np.max(np.abs(j - truth) / np.maximum(scale, 1e-300)))


@pytest.fixture(scope="module")
def golden():
    return golden_util.load_golden()


@pytest.fixture(scope="module")
def harness():
    """The engine's device math (csrc/ba_math.cuh) compiled for the host by nvcc."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcpu_math_harness.so")
    src = os.path.join(ROOT, "tests", "cpu_math_harness.cu")
    dep = os.path.join(ROOT, "deeparc-sfm_b200", "csrc", "ba_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler",
                               "-Wno-unknown-pragmas", "-o", so, src])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.harness_eval.argtypes = [C.POINTER(capi.DbaProblem), dp, dp, dp, dp, dp]

    def run(p):
        m = capi.ProblemMarshal(p)
        n = m.p.n_obs
        out = {"residuals": np.zeros((n, 2)), "jac_pt": np.zeros((n, 2, 3)), "jac_pose_a": np.zeros((n, 2, 6)),
               "jac_pose_b": np.zeros((n, 2, 6)), "jac_intr": np.zeros((n, 2, 3))}
        ptr = lambda a: a.ctypes.data_as(dp)
        assert lib.harness_eval(C.byref(m.struct), ptr(out["residuals"]), ptr(out["jac_pt"]), ptr(out["jac_pose_a"]),
                                ptr(out["jac_pose_b"]), ptr(out["jac_intr"])) == 0
        return out
    return run


# ------------------------------------------------------------------------- golden vectors
def test_oracle_residuals_match_golden(oracle, golden):
    p, truth = golden
    o = oracle.eval(p, residuals=True, jacobians=False)
    _check_residuals(o["residuals"], truth["residuals"], p.obs_xy)


def test_oracle_autodiff_jacobians_match_golden(oracle, golden):
    """Dual-number autodiff (what Ceres does) is exact to rounding except for a NON-ZERO rotation
    of ~1e-8 rad, where dividing by theta costs ~7 digits (kind tiny_rot_above_eps)."""
    p, truth = golden
    o = oracle.eval(p, residuals=True, jacobians=True)
    regular = np.array([k != "tiny_rot_above_eps" for k in truth["kinds"]])
    for key in ("jac_pt", "jac_pose_a", "jac_pose_b"):
        _check_jac(o[key], truth[key], 1e-11, regular)
        _check_jac(o[key], truth[key], 5e-8, ~regular)
    nf1 = truth["nf"] == 1
    _check_jac(o["jac_intr"], truth["jac_intr"], 1e-11, nf1)
    _check_jac(o["jac_intr"][:, :, 1:], truth["jac_intr"][:, :, 1:], 1e-11)


def test_reference_functor_matches_golden(reference, golden):
    p, truth = golden
    r = reference.eval(p, residuals=True, jacobians=True)
    _check_residuals(r["residuals"], truth["residuals"], p.obs_xy)
    regular = np.array([k != "tiny_rot_above_eps" for k in truth["kinds"]])
    for key in ("jac_pt", "jac_pose_a", "jac_pose_b"):
        _check_jac(r[key], truth[key], 1e-11, regular)


def test_oracle_equals_reference_functor(oracle, reference):
    """Restated functor vs the reference's own sources compiled against the shim: same bits up
    to the compilers' freedom, on every pose mode / nf / nd combination."""
    for p in (synthetic.arc_rig(n_arc=4, n_ring=5, n_pts=300, obs_per_point=7, seed=21),
              synthetic.bal_like(n_cam=25, n_pts=400, window=8, seed=22, free_intrinsics=0)):
        for nd in (0, 1, 2):
            q = p.copy()
            q.intr_nd[:] = nd
            q.intr_dist[:] = [1e-2, -1e-3]
            a, b = oracle.eval(q, jacobians=True), reference.eval(q, jacobians=True)
            for k in ("residuals", "jac_pt", "jac_pose_a", "jac_pose_b", "jac_intr"):
                assert np.max(np.abs(a[k] - b[k])) <= 1e-9 * max(np.max(np.abs(b[k])), 1e-300), k
            assert abs(a["cost"] - b["cost"]) <= 1e-13 * b["cost"]


def test_engine_math_on_host_matches_golden(harness, golden):
    """The CUDA engine's closed-form Jacobian (pose rows + series coefficients), compiled for the
    host: accurate on ALL cases, including the tiny non-zero rotation where autodiff is not."""
    p, truth = golden
    h = harness(p)
    _check_residuals(h["residuals"], truth["residuals"], p.obs_xy)
    for key in ("jac_pt", "jac_pose_a", "jac_pose_b"):
        _check_jac(h[key], truth[key], 1e-11)
    nf1 = truth["nf"] == 1
    _check_jac(h["jac_intr"], truth["jac_intr"], 1e-11, nf1)
    _check_jac(h["jac_intr"][:, :, 1:], truth["jac_intr"][:, :, 1:], 1e-11)


def test_engine_math_on_host_matches_oracle(harness, oracle):
    for p in (synthetic.arc_rig(n_arc=5, n_ring=6, n_pts=2000, obs_per_point=9, seed=31),
              synthetic.bal_like(n_cam=60, n_pts=3000, window=12, seed=32)):
        h, o = harness(p), oracle.eval(p, jacobians=True)
        _check_residuals(h["residuals"], o["residuals"], p.obs_xy)
        for key in ("jac_pt", "jac_pose_a", "jac_pose_b", "jac_intr"):
            _check_jac(h[key], o[key], 1e-10)


# ------------------------------------------------------------------- rotation conversions
def test_rotation_helpers_against_scipy(oracle):
    rng = np.random.default_rng(7)
    for _ in range(200):
        aa = rng.standard_normal(3) * rng.choice([1e-9, 1e-3, 0.3, 1.5, 3.0])
        R = Rotation.from_rotvec(aa).as_matrix()
        np.testing.assert_allclose(oracle.angle_axis_to_matrix(aa), R, atol=1e-14)
        pt = rng.standard_normal(3)
        np.testing.assert_allclose(oracle.rotate_point(aa, pt), R @ pt, atol=1e-14)
        if np.linalg.norm(aa) < np.pi - 1e-3:
            np.testing.assert_allclose(oracle.matrix_to_angle_axis(R), aa, atol=1e-12)
            q = Rotation.from_rotvec(aa).as_quat()  # x, y, z, w
            np.testing.assert_allclose(oracle.quaternion_to_angle_axis([q[3], q[0], q[1], q[2]]), aa, atol=1e-12)
            np.testing.assert_allclose(oracle.quaternion_to_angle_axis([-q[3], -q[0], -q[1], -q[2]]), aa, atol=1e-12)
    assert np.array_equal(oracle.angle_axis_to_matrix(np.zeros(3)), np.eye(3))


# ------------------------------------------------------------------------------ mini-Ceres
def test_minimizer_reaches_scipy_optimum(oracle):
    """Independent check of the LM restatement: same minimum as scipy.optimize.least_squares."""
    p = synthetic.bal_like(n_cam=6, n_pts=60, obs_per_point=4, window=6, seed=41, free_intrinsics=0)
    s, x = oracle.solve(p, capi.make_options(max_num_iterations=200, function_tolerance=1e-14, gradient_tolerance=1e-14,
                                             parameter_tolerance=1e-14, linear_solver=capi.DBA_LS_DENSE))
    free_ext = np.flatnonzero(p.ext_const == 0)

    def fun(z):
        q = p.copy()
        q.pts = z[:3 * p.n_pts].reshape(-1, 3)
        rest = z[3 * p.n_pts:].reshape(-1, 6)
        q.ext_rot = p.ext_rot.copy()
        q.ext_trans = p.ext_trans.copy()
        q.ext_rot[free_ext] = rest[:, :3]
        q.ext_trans[free_ext] = rest[:, 3:]
        return oracle.eval(q)["residuals"].reshape(-1)

    z0 = np.concatenate([p.pts.reshape(-1), np.concatenate([p.ext_rot[free_ext], p.ext_trans[free_ext]], axis=1).reshape(-1)])
    ref = scipy.optimize.least_squares(fun, z0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=4000)
    assert abs(s.final_cost - ref.cost) <= 1e-8 * ref.cost
    assert s.termination == capi.DBA_CONVERGENCE


def test_cauchy_loss_reaches_scipy_optimum(oracle):
    """The robust loss the reference keeps in a comment (`new ceres::CauchyLoss(0.5)`, sfm.cc:49): the
    oracle's corrector restatement minimises 1/2 sum a^2 log(1 + |r|^2 / a^2) — the same optimum and cost
    as scipy's loss='cauchy' applied to the per-observation residual norm."""
    p = synthetic.bal_like(n_cam=6, n_pts=60, obs_per_point=4, window=6, seed=43, free_intrinsics=0)
    rng = np.random.default_rng(9)
    bad = rng.choice(p.n_obs, size=12, replace=False)
    p.obs_xy[bad] += rng.normal(0.0, 40.0, size=(12, 2))  # gross outliers
    a = 2.0
    kw = dict(max_num_iterations=400, function_tolerance=1e-15, gradient_tolerance=1e-14, parameter_tolerance=1e-14,
              linear_solver=capi.DBA_LS_DENSE)
    s, x = oracle.solve(p, capi.make_options(loss_type=capi.DBA_LOSS_CAUCHY, loss_scale=a, **kw))
    s_plain, _ = oracle.solve(p, capi.make_options(**kw))
    free_ext = np.flatnonzero(p.ext_const == 0)

    def fun(z):
        q = p.copy()
        q.pts = z[:3 * p.n_pts].reshape(-1, 3)
        rest = z[3 * p.n_pts:].reshape(-1, 6)
        q.ext_rot = p.ext_rot.copy()
        q.ext_trans = p.ext_trans.copy()
        q.ext_rot[free_ext] = rest[:, :3]
        q.ext_trans[free_ext] = rest[:, 3:]
        return np.linalg.norm(oracle.eval(q)["residuals"], axis=1)

    z0 = np.concatenate([x["pts"].reshape(-1), np.concatenate([x["ext_rot"][free_ext], x["ext_trans"][free_ext]], axis=1).reshape(-1)])
    r = fun(z0)
    cost_at_x = 0.5 * np.sum(a * a * np.log1p(r * r / (a * a)))
    assert abs(s.final_cost - cost_at_x) <= 1e-10 * cost_at_x            # the reported cost is 1/2 sum rho
    ref = scipy.optimize.least_squares(fun, z0, method="trf", loss="cauchy", f_scale=a, xtol=1e-15, ftol=1e-15, gtol=1e-12,
                                       max_nfev=300)
    assert abs(ref.cost - s.final_cost) <= 1e-7 * s.final_cost            # scipy cannot improve on the oracle's optimum
    assert s.final_cost < 0.5 * s_plain.final_cost                        # and the outliers no longer dominate


def test_dense_and_implicit_schur_agree(oracle):
    p = synthetic.bal_like(n_cam=20, n_pts=400, window=8, seed=42)
    kw = dict(max_num_iterations=5, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    s1, x1 = oracle.solve(p, capi.make_options(linear_solver=capi.DBA_LS_DENSE, **kw))
    s2, x2 = oracle.solve(p, capi.make_options(linear_solver=capi.DBA_LS_PCG, pcg_rel_tolerance=1e-14, pcg_max_iterations=5000, **kw))
    np.testing.assert_allclose(s1.trace("cost"), s2.trace("cost"), rtol=1e-9)
    for k in x1:
        assert np.max(np.abs(x1[k] - x2[k])) <= 1e-8 * max(np.max(np.abs(x1[k])), 1e-300)


def test_lm_trace_semantics(oracle):
    """Ceres trust-region bookkeeping: radius x3 after a near-perfect step, halved (then /4 ...)
    after rejected ones; iteration 0 carries the initial cost; costs never increase."""
    p = synthetic.arc_rig(n_arc=3, n_ring=3, n_pts=150, obs_per_point=5, seed=43)
    s, _ = oracle.solve(p, capi.make_options(max_num_iterations=12, function_tolerance=0.0, gradient_tolerance=0.0,
                                             parameter_tolerance=0.0, linear_solver=capi.DBA_LS_DENSE))
    it = s.iterations
    assert it[0].iteration == 0 and it[0].cost == s.initial_cost and it[0].trust_region_radius == 1e4
    radius, dec = 1e4, 2.0
    cost = s.initial_cost
    for k in it[1:]:
        if k.step_is_successful:
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * k.relative_decrease - 1.0) ** 3))
            dec = 2.0
            assert k.relative_decrease > 1e-3 and k.cost <= cost * (1 + 1e-12)
            cost = k.cost
        else:
            radius, dec = radius / dec, dec * 2.0
        assert abs(k.trust_region_radius - radius) <= 1e-12 * radius


def test_reference_solve_equals_oracle_solve(oracle, reference, tmp_path):
    """The reference's unmodified solve() (through its own DeepArcManager) and oracle_solve on the
    exported flat problem walk the same LM trajectory (both on the mini-Ceres shim)."""
    p = synthetic.arc_rig(n_arc=3, n_ring=4, n_pts=250, obs_per_point=6, seed=44)
    f = str(tmp_path / "rig.deeparc")
    synthetic.write_deeparc(p, f)
    h = reference.read(f)
    flat = reference.export(h)
    reference.set_overrides(quiet=1, max_num_iterations=6, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    try:
        sr = reference.solve(h, 100, 3600, False)
    finally:
        reference.set_overrides(quiet=1)
    after = reference.export(h)
    so, xo = oracle.solve(flat, capi.make_options(max_num_iterations=6, function_tolerance=0.0, gradient_tolerance=0.0,
                                                  parameter_tolerance=0.0, linear_solver=capi.DBA_LS_DENSE))
    # (two OpenMP runs of the same code differ by the order of their thread-private partial sums)
    np.testing.assert_allclose(sr.trace("cost"), so.trace("cost"), rtol=1e-10)
    for k in ("pts", "ext_rot", "ext_trans"):
        assert np.max(np.abs(getattr(after, k) - xo[k])) <= 1e-10 * np.max(np.abs(xo[k]))
    reference.free(h)


def test_real_ceres_pins_the_restatement(oracle, tmp_path):
    """SURVEY 8c-iv / BASELINE.md 2: when a real ceres-solver install exists, `make -C oracle WITH_CERES=1
    ref_ceres` builds the reference's UNMODIFIED solve() against it; this test then pins the mini-Ceres
    restatement: the scene the reference's solve() leaves (Ceres defaults, 100 iterations, as sfm.cc:121
    calls it) equals the oracle's to 1e-6.  Skipped in this image (no Ceres / Eigen / glog, no network)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libdeeparc_ref_ceres.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libdeeparc_ref_ceres.so not built: needs a ceres-solver install (make -C oracle WITH_CERES=1 ref_ceres)")
    real = oracle_lib.Reference(path)
    assert real.lib.ref_is_real_ceres() == 1
    p = synthetic.arc_rig(n_arc=3, n_ring=4, n_pts=250, obs_per_point=6, seed=44)
    f = str(tmp_path / "rig.deeparc")
    synthetic.write_deeparc(p, f)
    h = real.read(f)
    flat = real.export(h)
    real.solve(h, 100, 3600, False)
    after = real.export(h)
    so, xo = oracle.solve(flat, capi.make_options(max_num_iterations=100, linear_solver=capi.DBA_LS_DENSE))
    for k in ("pts", "ext_rot", "ext_trans"):
        assert np.max(np.abs(getattr(after, k) - xo[k])) <= 1e-6 * np.max(np.abs(xo[k])), k
    real.free(h)


# ------------------------------------------------------------------------------ hemisphere
def test_hemisphere_fit_closed_form(oracle):
    """|c - p|^2 - rho is linear in (c, rho - |c|^2): the LM fit must land on the lstsq answer."""
    rng = np.random.default_rng(9)
    c_true = np.array([0.02, -0.01, 0.5])
    d = rng.standard_normal((120, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = c_true + 0.5 * d + 1e-3 * rng.standard_normal((120, 3))
    c, rho, s = oracle.fit_hemisphere(pts, options=capi.make_options(max_num_iterations=1000, function_tolerance=1e-15,
                                                                     gradient_tolerance=1e-15, parameter_tolerance=1e-15))
    A = np.concatenate([-2 * pts, np.ones((120, 1))], axis=1)
    sol = np.linalg.lstsq(A, -np.sum(pts * pts, axis=1), rcond=None)[0]
    c_ls, rho_ls = sol[:3], sol[:3] @ sol[:3] - sol[3]
    np.testing.assert_allclose(c, c_ls, atol=1e-7)
    assert abs(rho - rho_ls) <= 1e-7
    assert abs(rho - 0.25) < 1e-2  # rho is the SQUARED radius (hemisphere_radius.hh:26)


def test_hemisphere_oracle_equals_reference(oracle, reference):
    rng = np.random.default_rng(10)
    pts = np.array([0.0, 0.0, 0.5]) + 0.5 * rng.standard_normal((60, 3)) / 3
    c1, r1, s1 = oracle.fit_hemisphere(pts)
    c2, r2, s2 = reference.fit_hemisphere(pts)
    np.testing.assert_allclose(c1, c2, rtol=1e-12, atol=1e-15)
    assert abs(r1 - r2) <= 1e-12 * abs(r2)
    assert s1.num_iterations == s2.num_iterations
