"""TEST INFRASTRUCTURE: ctypes bindings of the CPU oracle (oracle/liboracle.so) and of the
reference's own sources compiled against the mini-Ceres shim (oracle/_ref/libdeeparc_ref.so).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from deeparc_sfm_b200 import capi
from deeparc_sfm_b200.synthetic import Problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libdeeparc_ref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _eval_common(fn, p: Problem, residuals=True, jacobians=False):
    m = capi.ProblemMarshal(p)
    n = m.p.n_obs
    cost = C.c_double()
    res = np.zeros((n, 2)) if residuals else None
    jp = np.zeros((n, 2, 3)) if jacobians else None
    ja = np.zeros((n, 2, 6)) if jacobians else None
    jb = np.zeros((n, 2, 6)) if jacobians else None
    ji = np.zeros((n, 2, 3)) if jacobians else None
    st = fn(C.byref(m.struct), C.byref(cost), _ptr(res), _ptr(jp), _ptr(ja), _ptr(jb), _ptr(ji))
    assert st == 0, st
    return {"cost": cost.value, "residuals": res, "jac_pt": jp, "jac_pose_a": ja, "jac_pose_b": jb, "jac_intr": ji}


class Oracle:
    def __init__(self, path: str = ORACLE_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle oracle`")
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_eval.argtypes = [C.POINTER(capi.DbaProblem), _dp, _dp, _dp, _dp, _dp, _dp]
        L.oracle_solve.argtypes = [C.POINTER(capi.DbaProblem), C.POINTER(capi.DbaSolveOptions),
                                   C.POINTER(capi.DbaSummary), C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.oracle_fit_hemisphere.argtypes = [_dp, C.c_int, _dp, _dp, C.POINTER(capi.DbaSolveOptions),
                                            C.POINTER(capi.DbaSummary), C.c_int]
        L.oracle_filter_mse.argtypes = [C.POINTER(capi.DbaProblem), _dp]
        for name in ("oracle_angle_axis_rotate_point", "oracle_angle_axis_to_rotation_matrix",
                     "oracle_rotation_matrix_to_angle_axis", "oracle_quaternion_to_angle_axis"):
            getattr(L, name).restype = None
        L.oracle_angle_axis_rotate_point.argtypes = [_dp, _dp, _dp]
        L.oracle_angle_axis_to_rotation_matrix.argtypes = [_dp, _dp]
        L.oracle_rotation_matrix_to_angle_axis.argtypes = [_dp, _dp]
        L.oracle_quaternion_to_angle_axis.argtypes = [_dp, _dp]

    def num_procs(self) -> int:
        return int(self.lib.oracle_num_procs())

    def eval(self, p: Problem, residuals=True, jacobians=False):
        return _eval_common(self.lib.oracle_eval, p, residuals, jacobians)

    def solve(self, p: Problem, options=None, num_threads: int = 0, capacity: int = 1024):
        m = capi.ProblemMarshal(p)
        q = m.p
        o = options or capi.default_options_struct()
        s = capi.Summary(capacity)
        out = {"pts": np.zeros_like(q.pts), "ext_rot": np.zeros_like(q.ext_rot), "ext_trans": np.zeros_like(q.ext_trans),
               "intr_focal": np.zeros_like(q.intr_focal), "intr_dist": np.zeros_like(q.intr_dist)}
        st = self.lib.oracle_solve(C.byref(m.struct), C.byref(o), C.byref(s.struct), num_threads, _ptr(out["pts"]),
                                   _ptr(out["ext_rot"]), _ptr(out["ext_trans"]), _ptr(out["intr_focal"]),
                                   _ptr(out["intr_dist"]))
        assert st == 0, st
        return s, out

    def fit_hemisphere(self, centres, centre0=(0.0, 0.0, 0.0), rho0=1.0, options=None, num_threads: int = 1):
        centres = np.ascontiguousarray(centres, dtype=np.float64)
        c = np.array(centre0, dtype=np.float64)
        rho = C.c_double(rho0)
        o = options or capi.make_options(max_num_iterations=1000)
        s = capi.Summary(1024)
        st = self.lib.oracle_fit_hemisphere(_ptr(centres), centres.shape[0], _ptr(c), C.byref(rho), C.byref(o),
                                            C.byref(s.struct), num_threads)
        assert st == 0, st
        return c, rho.value, s

    def filter_mse(self, p: Problem) -> np.ndarray:
        m = capi.ProblemMarshal(p)
        out = np.zeros(m.p.n_obs)
        st = self.lib.oracle_filter_mse(C.byref(m.struct), _ptr(out))
        assert st == 0, st
        return out

    # rotation helpers -------------------------------------------------------------
    def rotate_point(self, aa, pt):
        aa = np.ascontiguousarray(aa, dtype=np.float64)
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        out = np.zeros(3)
        self.lib.oracle_angle_axis_rotate_point(_ptr(aa), _ptr(pt), _ptr(out))
        return out

    def angle_axis_to_matrix(self, aa):
        aa = np.ascontiguousarray(aa, dtype=np.float64)
        out = np.zeros(9)
        self.lib.oracle_angle_axis_to_rotation_matrix(_ptr(aa), _ptr(out))
        return out.reshape(3, 3).T  # column-major -> numpy [row, col]

    def matrix_to_angle_axis(self, R):
        cm = np.ascontiguousarray(np.asarray(R, dtype=np.float64).T).reshape(-1)
        out = np.zeros(3)
        self.lib.oracle_rotation_matrix_to_angle_axis(_ptr(cm), _ptr(out))
        return out

    def quaternion_to_angle_axis(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = np.zeros(3)
        self.lib.oracle_quaternion_to_angle_axis(_ptr(q), _ptr(out))
        return out


class _ManagerApi:
    """Manager-level C entry points shared by the reference bridge (prefix ``ref_``) and the
    product's C++ host mirror (prefix ``dam_``)."""

    def _bind_manager(self, L, prefix):
        self._pfx = prefix
        f = lambda name: getattr(L, prefix + name)
        f("manager_read").argtypes = [C.c_char_p]
        f("manager_read").restype = C.c_void_p
        f("manager_free").argtypes = [C.c_void_p]
        f("manager_free").restype = None
        f("manager_is_shared").argtypes = [C.c_void_p]
        f("manager_counts").argtypes = [C.c_void_p, C.POINTER(C.c_int64)] + [C.POINTER(C.c_int)] * 5
        f("manager_counts").restype = None
        f("manager_export").argtypes = [C.c_void_p, _dp, _ip, _ip, _ip, _ip, _dp, _ip, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _bp]
        f("manager_solve").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        f("manager_filter").argtypes = [C.c_void_p, C.c_double, _dp, C.c_double]
        f("manager_write").argtypes = [C.c_void_p, C.c_char_p]
        f("manager_write").restype = None
        f("manager_write_ply").argtypes = [C.c_void_p, C.c_char_p]
        f("manager_write_ply").restype = None
        f("manager_camera_centers").argtypes = [C.c_void_p, _dp, C.c_int]
        f("last_summary").argtypes = [C.POINTER(capi.DbaSummary)]
        f("last_summary").restype = None

    def _f(self, name):
        return getattr(self.lib, self._pfx + name)

    def last_summary(self, capacity=1024):
        s = capi.Summary(capacity)
        self._f("last_summary")(C.byref(s.struct))
        return s

    def read(self, path: str):
        h = self._f("manager_read")(path.encode())
        if not h:
            raise IOError(f"DeepArcManager::read failed for {path}")
        return h

    def free(self, h):
        self._f("manager_free")(h)

    def is_shared(self, h) -> bool:
        return bool(self._f("manager_is_shared")(h))

    def counts(self, h):
        n_obs = C.c_int64()
        v = [C.c_int() for _ in range(5)]
        self._f("manager_counts")(h, C.byref(n_obs), *[C.byref(x) for x in v])
        return {"n_obs": n_obs.value, "n_pts": v[0].value, "n_ext": v[1].value, "n_intr": v[2].value,
                "n_arc": v[3].value, "n_ring": v[4].value}

    def export(self, h) -> Problem:
        c = self.counts(h)
        n, npt, ne, ni = c["n_obs"], c["n_pts"], c["n_ext"], c["n_intr"]
        p = Problem(obs_xy=np.zeros((n, 2)), obs_pt=np.zeros(n, np.int32), obs_pose_a=np.zeros(n, np.int32),
                    obs_pose_b=np.zeros(n, np.int32), obs_intr=np.zeros(n, np.int32), pts=np.zeros((npt, 3)),
                    ext_rot=np.zeros((ne, 3)), ext_trans=np.zeros((ne, 3)), intr_center=np.zeros((ni, 2)),
                    intr_focal=np.zeros((ni, 2)), intr_dist=np.zeros((ni, 2)), intr_nf=np.zeros(ni, np.int32),
                    intr_nd=np.zeros(ni, np.int32), ext_const=np.zeros(ne, np.uint8), n_arc=c["n_arc"], n_ring=c["n_ring"])
        p.pts_rgb = np.zeros((npt, 3), np.int32)
        ip = lambda a: a.ctypes.data_as(_ip)
        self._f("manager_export")(h, _ptr(p.obs_xy), ip(p.obs_pt), ip(p.obs_pose_a), ip(p.obs_pose_b), ip(p.obs_intr),
                                  _ptr(p.pts), ip(p.pts_rgb), _ptr(p.ext_rot), _ptr(p.ext_trans), _ptr(p.intr_center),
                                  _ptr(p.intr_focal), _ptr(p.intr_dist), ip(p.intr_nf), ip(p.intr_nd),
                                  p.ext_const.ctypes.data_as(_bp))
        return p

    def solve(self, h, max_iteration=100, max_second=3600, freeze_camera=False):
        self._f("manager_solve")(h, max_iteration, max_second, int(freeze_camera))
        return self.last_summary()

    def filter(self, h, error_boundary, centre, radius):
        centre = np.ascontiguousarray(centre, dtype=np.float64)
        self._f("manager_filter")(h, error_boundary, _ptr(centre), radius)

    def write(self, h, path):
        self._f("manager_write")(h, path.encode())

    def write_ply(self, h, path):
        self._f("manager_write_ply")(h, path.encode())

    def camera_centers(self, h):
        n = self._f("manager_camera_centers")(h, None, 0)
        out = np.zeros((max(n, 1), 3))
        self._f("manager_camera_centers")(h, _ptr(out), n)
        return out[:n]


class Reference(_ManagerApi):
    """The reference's own DeepArcManager / solve() / functors behind a C bridge (oracle/_ref)."""

    def __init__(self, path: str = REF_PATH):
        self.lib = C.CDLL(path)
        L = self.lib
        self._bind_manager(L, "ref_")
        L.ref_eval.argtypes = [C.POINTER(capi.DbaProblem), _dp, _dp, _dp, _dp, _dp, _dp]
        L.ref_set_overrides.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.ref_set_overrides.restype = None
        L.ref_fit_hemisphere.argtypes = [_dp, C.c_int, _dp, _dp, C.POINTER(capi.DbaSummary)]
        self.set_overrides(quiet=1)

    def set_overrides(self, quiet=-1, num_threads=-1, max_num_iterations=-1, function_tolerance=-1.0,
                      gradient_tolerance=-1.0, parameter_tolerance=-1.0):
        self.lib.ref_set_overrides(quiet, num_threads, max_num_iterations, function_tolerance, gradient_tolerance,
                                   parameter_tolerance)

    def eval(self, p: Problem, residuals=True, jacobians=False):
        return _eval_common(self.lib.ref_eval, p, residuals, jacobians)

    def fit_hemisphere(self, centres, centre0=(0.0, 0.0, 0.0), rho0=1.0):
        centres = np.ascontiguousarray(centres, dtype=np.float64)
        c = np.array(centre0, dtype=np.float64)
        rho = C.c_double(rho0)
        s = capi.Summary(1024)
        self.lib.ref_fit_hemisphere(_ptr(centres), centres.shape[0], _ptr(c), C.byref(rho), C.byref(s.struct))
        return c, rho.value, s


HOST_PATH = os.path.join(ROOT, "deeparc-sfm_b200", "lib", "libdeeparc_host.so")


class HostMirror(_ManagerApi):
    """The product's C++ mirror of the reference host surface (deeparc-sfm_b200/host)."""

    def __init__(self, path: str = HOST_PATH):
        capi.load_library()  # libdeeparc_ba.so first (RTLD_GLOBAL)
        self.lib = C.CDLL(path)
        self._bind_manager(self.lib, "dam_")
        self.lib.dam_last_error.restype = C.c_char_p
        self.lib.dam_manager_write_binary.argtypes = [C.c_void_p, C.c_char_p]
        self.lib.dam_manager_write_binary.restype = None

    def write_binary(self, h, path):
        self.lib.dam_manager_write_binary(h, path.encode())

    def last_error(self) -> str:
        return self.lib.dam_last_error().decode()
