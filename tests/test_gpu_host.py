"""GPU tests of the drop-in layer: the C++ mirror of the reference's host surface (DeepArcManager,
solve(), filterPoint3d, the sfm driver) running on the CUDA engine, against the reference's own
sources (oracle/_ref: unmodified sfm.cc / DeepArcManager.cc on the mini-Ceres shim) on the same
files.  Also the multi-GPU path when more than one device is visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

from deeparc_sfm_b200 import capi, synthetic
from tests import oracle_lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    h = oracle_lib.HostMirror()
    yield h
    h.lib.dam_engine_release()


def _rig_file(tmp_path, name="rig", sigma=0.5, seed=81):
    p = synthetic.arc_rig(n_arc=4, n_ring=5, n_pts=500, obs_per_point=8, seed=seed, pixel_sigma=sigma)
    f = str(tmp_path / f"{name}.deeparc")
    synthetic.write_deeparc(p, f)
    return p, f


def _assert_same_scene(a, b, tol=1e-6):
    for k in ("obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    for k in ("pts", "ext_rot", "ext_trans"):
        x, y = getattr(a, k), getattr(b, k)
        assert x.shape == y.shape
        assert np.max(np.abs(x - y)) <= tol * max(np.max(np.abs(y)), 1e-300), k


@pytest.mark.parametrize("freeze", [True, False])
def test_manager_solve_matches_reference_solve(host, reference, tmp_path, freeze):
    """solve(manager, 100, 3600, freeze) exactly as the driver calls it (sfm.cc:111,121): Ceres
    default tolerances, results written in place into the scene graph."""
    _, f = _rig_file(tmp_path)
    hr, hh = reference.read(f), host.read(f)
    reference.set_overrides(quiet=1)
    sr = reference.solve(hr, 100, 3600, freeze)
    sh = host.solve(hh, 100, 3600, freeze)
    assert sh.termination == sr.termination == capi.DBA_CONVERGENCE
    assert sh.num_iterations == sr.num_iterations
    assert abs(sh.final_cost - sr.final_cost) <= 1e-6 * sr.final_cost
    np.testing.assert_allclose(sh.trace("cost"), sr.trace("cost"), rtol=1e-6)
    _assert_same_scene(host.export(hh), reference.export(hr))
    reference.free(hr), host.free(hh)


def test_filter_point3d_matches_reference(host, reference, tmp_path):
    """filterPoint3d (DeepArcManager.cc:331-424) with the GPU residuals: same surviving
    observations and points, including the `mse < boundary` direction and the rho/2 test."""
    _, f = _rig_file(tmp_path, sigma=3.0, seed=82)
    hr, hh = reference.read(f), host.read(f)
    centre, rho = np.array([0.0, 0.0, 0.5]), 0.02  # rho/2 = 0.01 = (0.1)^2: cuts through the point ball
    reference.filter(hr, 5.0, centre, rho)
    host.filter(hh, 5.0, centre, rho)
    cr, ch = reference.counts(hr), host.counts(hh)
    assert cr == ch
    assert 0 < ch["n_obs"] < 500 * 8 and 0 < ch["n_pts"] < 500
    _assert_same_scene(host.export(hh), reference.export(hr), tol=0.0)
    reference.free(hr), host.free(hh)


def test_sfm_driver_end_to_end_matches_reference_pipeline(reference, tmp_path):
    """bin/sfm on a noisy rig vs the reference pipeline (sfm.cc:77-131) replayed call by call
    through the bridge: hemisphere fit, frozen-camera pass, filter, outer loop, final write."""
    _, f = _rig_file(tmp_path, name="drv", sigma=3.0, seed=83)
    exe = os.path.join(ROOT, "deeparc-sfm_b200", "bin", "sfm")
    out = str(tmp_path / "drv_out.deeparc")
    r = subprocess.run([exe, "--input", f, "--output", out, "--ply-init", str(tmp_path / "init.ply"), "--ply-adjust",
                        str(tmp_path / "adj_"), "--ply-clear", str(tmp_path / "clear.ply")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    # reference pipeline
    h = reference.read(f)
    centres = reference.camera_centers(h)
    c, rho, _ = reference.fit_hemisphere(centres)
    reference.set_overrides(quiet=1)
    reference.solve(h, 100, 3600, True)
    reference.filter(h, 5.0, c, rho)
    old, cur, steps = 1, 10000000, 0
    while cur != old:
        steps += 1
        old = cur
        reference.solve(h, 100, 3600, False)
        reference.filter(h, 5.0, c, rho)
        cur = reference.counts(h)["n_pts"]
    ref_out = str(tmp_path / "ref_out.deeparc")
    reference.write(h, ref_out)
    assert f"TOTAL REPEAT: {steps}" in r.stdout
    a, b = synthetic.read_deeparc_text(out), synthetic.read_deeparc_text(ref_out)
    assert a["obs"].shape == b["obs"].shape and a["pts"].shape == b["pts"].shape
    assert np.array_equal(a["obs"][:, :3], b["obs"][:, :3])
    np.testing.assert_allclose(a["pts"][:, :3], b["pts"][:, :3], atol=2.1e-6)
    for (ta, ra), (tb, rb) in zip(a["ext"], b["ext"]):
        np.testing.assert_allclose(ta, tb, atol=2.1e-6)
        np.testing.assert_allclose(ra, rb, atol=2.1e-6)
    assert os.path.exists(tmp_path / "clear.ply") and os.path.exists(tmp_path / f"adj_{steps}.ply")
    reference.free(h)


def test_sfm_driver_device_resident_loop_equals_reupload(tmp_path):
    """The outer loop of main() (sfm.cc:118-127) with the scene resident on the device between solve()
    and filterPoint3d() (dba_filter on the engine's copy, dba_problem_update instead of a second gather +
    upload) writes byte-identical results to the same run with DEEPARC_RESIDENT=0, which flattens and
    uploads the scene before every solve and every filter as round 1 did."""
    _, f = _rig_file(tmp_path, name="res", sigma=3.0, seed=84)
    exe = os.path.join(ROOT, "deeparc-sfm_b200", "bin", "sfm")
    outs = {}
    for mode in ("1", "0"):
        out = str(tmp_path / f"out_{mode}.deeparc")
        r = subprocess.run([exe, "--input", f, "--output", out, "--ply-init", str(tmp_path / f"i{mode}.ply"), "--ply-adjust",
                            str(tmp_path / f"a{mode}_"), "--ply-clear", str(tmp_path / f"c{mode}.ply"), "--output-binary",
                            out + "b"], capture_output=True, text=True, timeout=600, env=dict(os.environ, DEEPARC_RESIDENT=mode))
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        outs[mode] = (out, r.stdout)
    import filecmp
    assert filecmp.cmp(outs["1"][0], outs["0"][0], shallow=False)
    assert filecmp.cmp(outs["1"][0] + "b", outs["0"][0] + "b", shallow=False)  # bit-identical parameters
    assert filecmp.cmp(str(tmp_path / "c1.ply"), str(tmp_path / "c0.ply"), shallow=False)
    rep = [l for l in outs["1"][1].splitlines() if l.startswith("TOTAL REPEAT")]
    assert rep and int(rep[0].split(":")[1]) >= 2  # the loop really went around


def test_empty_and_degenerate_problems(engine):
    p = synthetic.bal_like(n_cam=5, n_pts=20, obs_per_point=3, window=5, seed=84, free_intrinsics=0)
    empty = p.copy()
    for k in ("obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr"):
        setattr(empty, k, getattr(p, k)[:0])
    empty.obs_xy = p.obs_xy[:0]
    engine.problem_set(empty)
    s = engine.solve(capi.make_options(max_num_iterations=5))
    assert s.termination == capi.DBA_CONVERGENCE and s.final_cost == 0.0
    # points that nobody observes keep their coordinates
    q = p.copy()
    q.pts = np.concatenate([p.pts, [[1.0, 2.0, 3.0]]])
    engine.problem_set(q)
    engine.solve(capi.make_options(max_num_iterations=3))
    assert np.array_equal(engine.params_get()["pts"][-1], [1.0, 2.0, 3.0])


def test_two_gpu_sharded_solve_matches_single_gpu():
    if capi.load_library().dba_device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, REPO_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29741", os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU OK" in r.stdout
