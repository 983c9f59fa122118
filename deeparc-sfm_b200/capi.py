"""ctypes binding of the C ABI in include/deeparc_ba.h (libdeeparc_ba.so).

This is the reference-side binding a Python caller would use; the C++ binding used by the
``sfm`` driver is deeparc-sfm_b200/host/solve.cc.  There is no CPU fallback anywhere in this
module: if the shared library or a CUDA device is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .synthetic import Problem

_HERE = os.path.dirname(os.path.abspath(__file__))
# DBA_LIB=checked selects the build with device-side index assertions (make CHECKED=1)
LIB_PATH = os.path.join(_HERE, "lib", "libdeeparc_ba_checked.so" if os.environ.get("DBA_LIB") == "checked" else "libdeeparc_ba.so")

DBA_OK = 0
DBA_ERR_INVALID_ARGUMENT = -1
DBA_ERR_NO_DEVICE = -2
DBA_ERR_CUDA = -3
DBA_ERR_UNSUPPORTED = -4
DBA_ERR_NO_PROBLEM = -5
DBA_ERR_NCCL = -6
DBA_ERR_NUMERIC = -7

DBA_LS_AUTO, DBA_LS_PCG, DBA_LS_DENSE = 0, 1, 2
DBA_CONVERGENCE, DBA_NO_CONVERGENCE, DBA_FAILURE = 0, 1, 2
DBA_LOSS_NONE, DBA_LOSS_CAUCHY = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class DbaConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("verbose", C.c_int32)]


class DbaProblem(C.Structure):
    _fields_ = [("n_obs", C.c_int64), ("n_pts", C.c_int32), ("n_ext", C.c_int32), ("n_intr", C.c_int32),
                ("obs_xy", _dp), ("obs_pt", _ip), ("obs_pose_a", _ip), ("obs_pose_b", _ip), ("obs_intr", _ip),
                ("pts", _dp), ("ext_rot", _dp), ("ext_trans", _dp), ("intr_center", _dp), ("intr_focal", _dp),
                ("intr_dist", _dp), ("intr_nf", _ip), ("intr_nd", _ip), ("ext_const", _bp),
                ("freeze_camera", C.c_int32), ("free_intrinsics", C.c_int32)]


class DbaSolveOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("max_solver_time_in_seconds", C.c_double),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("jacobi_scaling", C.c_int32),
                ("max_num_consecutive_invalid_steps", C.c_int32), ("linear_solver", C.c_int32),
                ("pcg_max_iterations", C.c_int32), ("pcg_min_iterations", C.c_int32),
                ("pcg_rel_tolerance", C.c_double), ("dense_max_size", C.c_int32),
                ("progress_to_stdout", C.c_int32), ("loss_type", C.c_int32), ("loss_scale", C.c_double)]


class DbaIteration(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
                ("linear_solver_iterations", C.c_int32), ("cost", C.c_double), ("cost_change", C.c_double),
                ("gradient_max_norm", C.c_double), ("gradient_norm", C.c_double), ("step_norm", C.c_double),
                ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double),
                ("model_cost_change", C.c_double), ("iteration_time_in_seconds", C.c_double)]


class DbaSummary(C.Structure):
    _fields_ = [("termination", C.c_int32), ("num_iterations", C.c_int32), ("num_successful_steps", C.c_int32),
                ("num_unsuccessful_steps", C.c_int32), ("linear_solver_used", C.c_int32),
                ("reduced_system_size", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("total_time_in_seconds", C.c_double), ("device_time_in_seconds", C.c_double),
                ("loop_device_time_in_seconds", C.c_double),
                ("kernel_launches", C.c_int64), ("jacobian_evaluations", C.c_int64),
                ("residual_evaluations", C.c_int64), ("pcg_iterations_total", C.c_int64),
                ("message", C.c_char * 192), ("iterations", C.POINTER(DbaIteration)),
                ("iterations_capacity", C.c_int32), ("linear_solver_failures", C.c_int32),
                ("pcg_unconverged_solves", C.c_int32), ("reserved_", C.c_int32)]


class DbaKernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("total_ms", C.c_double),
                ("algorithmic_bytes", C.c_double)]


def default_options_struct() -> DbaSolveOptions:
    """Ceres defaults as the reference leaves them (src/sfm.cc:66-71); filled in Python so the
    oracle can use it without loading the product library."""
    o = DbaSolveOptions()
    o.max_num_iterations = 50
    o.max_solver_time_in_seconds = 1e9
    o.initial_trust_region_radius = 1e4
    o.max_trust_region_radius = 1e16
    o.min_trust_region_radius = 1e-32
    o.min_relative_decrease = 1e-3
    o.min_lm_diagonal = 1e-6
    o.max_lm_diagonal = 1e32
    o.function_tolerance = 1e-6
    o.gradient_tolerance = 1e-10
    o.parameter_tolerance = 1e-8
    o.jacobi_scaling = 1
    o.max_num_consecutive_invalid_steps = 5
    o.linear_solver = DBA_LS_AUTO
    o.pcg_max_iterations = 500
    o.pcg_min_iterations = 0
    o.pcg_rel_tolerance = 1e-12
    o.dense_max_size = 768
    o.progress_to_stdout = 0
    o.loss_type = DBA_LOSS_NONE
    o.loss_scale = 0.5
    return o


def make_options(**kw) -> DbaSolveOptions:
    o = default_options_struct()
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(f"dba_solve_options has no field {k!r}")
        setattr(o, k, v)
    return o


class ProblemMarshal:
    """Keeps the numpy buffers alive while a ``dba_problem`` struct points into them."""

    def __init__(self, p: Problem):
        self.p = p.normalised()
        q = self.p
        s = DbaProblem()
        s.n_obs, s.n_pts, s.n_ext, s.n_intr = q.n_obs, q.n_pts, q.n_ext, q.n_intr
        s.obs_xy = q.obs_xy.ctypes.data_as(_dp)
        s.obs_pt = q.obs_pt.ctypes.data_as(_ip)
        s.obs_pose_a = q.obs_pose_a.ctypes.data_as(_ip)
        s.obs_pose_b = q.obs_pose_b.ctypes.data_as(_ip)
        s.obs_intr = q.obs_intr.ctypes.data_as(_ip)
        s.pts = q.pts.ctypes.data_as(_dp)
        s.ext_rot = q.ext_rot.ctypes.data_as(_dp)
        s.ext_trans = q.ext_trans.ctypes.data_as(_dp)
        s.intr_center = q.intr_center.ctypes.data_as(_dp)
        s.intr_focal = q.intr_focal.ctypes.data_as(_dp)
        s.intr_dist = q.intr_dist.ctypes.data_as(_dp)
        s.intr_nf = q.intr_nf.ctypes.data_as(_ip)
        s.intr_nd = q.intr_nd.ctypes.data_as(_ip)
        s.ext_const = q.ext_const.ctypes.data_as(_bp)
        s.freeze_camera = int(q.freeze_camera)
        s.free_intrinsics = int(q.free_intrinsics)
        self.struct = s


class Summary:
    def __init__(self, capacity: int = 1024):
        self._iters = (DbaIteration * capacity)()
        self.struct = DbaSummary()
        self.struct.iterations = C.cast(self._iters, C.POINTER(DbaIteration))
        self.struct.iterations_capacity = capacity

    @property
    def iterations(self):
        return [self._iters[i] for i in range(self.struct.num_iterations)]

    def trace(self, field: str) -> np.ndarray:
        return np.array([getattr(it, field) for it in self.iterations])

    def __getattr__(self, name):
        if name in ("struct", "_iters"):
            raise AttributeError(name)
        v = getattr(self.struct, name)
        return v.decode() if isinstance(v, bytes) else v


class EngineError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"deeparc_ba status {status}: {text}")
        self.status = status


_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Loads libdeeparc_ba.so (built in-tree by ``__graft_entry__.build()``); raises if missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.dba_abi_version.restype = C.c_int
    lib.dba_device_count.restype = C.c_int
    lib.dba_nccl_unique_id.argtypes = [C.c_void_p]
    lib.dba_shard_plan.argtypes = [C.POINTER(DbaProblem), C.c_int32, _ip, C.POINTER(C.c_int64)]
    lib.dba_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(DbaConfig)]
    lib.dba_destroy.argtypes = [C.c_void_p]
    lib.dba_destroy.restype = None
    lib.dba_last_error.argtypes = [C.c_void_p]
    lib.dba_last_error.restype = C.c_char_p
    lib.dba_problem_set.argtypes = [C.c_void_p, C.POINTER(DbaProblem)]
    lib.dba_params_reset.argtypes = [C.c_void_p]
    lib.dba_problem_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    lib.dba_eval.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]
    lib.dba_solve_options_default.argtypes = [C.POINTER(DbaSolveOptions)]
    lib.dba_solve_options_default.restype = None
    lib.dba_solve.argtypes = [C.c_void_p, C.POINTER(DbaSolveOptions), C.POINTER(DbaSummary)]
    lib.dba_params_get.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
    lib.dba_fit_hemisphere.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, _dp, C.POINTER(DbaSolveOptions),
                                       C.POINTER(DbaSummary)]
    lib.dba_filter_mse.argtypes = [C.c_void_p, _dp]
    lib.dba_filter.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dba_filter.restype = C.c_int
    lib.dba_kernel_stats_enable.argtypes = [C.c_void_p, C.c_int32]
    lib.dba_kernel_stats_reset.argtypes = [C.c_void_p]
    lib.dba_kernel_stats.argtypes = [C.c_void_p, C.POINTER(DbaKernelStat), C.c_int32]
    if path == LIB_PATH:
        _lib = lib
    return lib


def shard_plan(p: Problem, world_size: int):
    """Host-only: (pt_begin[world+1], obs_count[world]) of the point sharding (no GPU needed)."""
    lib = load_library()
    m = ProblemMarshal(p)
    pt_begin = np.zeros(world_size + 1, np.int32)
    obs_count = np.zeros(world_size, np.int64)
    st = lib.dba_shard_plan(C.byref(m.struct), world_size, pt_begin.ctypes.data_as(_ip),
                            obs_count.ctypes.data_as(C.POINTER(C.c_int64)))
    if st != DBA_OK:
        raise EngineError(st, "dba_shard_plan")
    return pt_begin, obs_count


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


class Engine:
    """One handle == one GPU.  Method names follow the C ABI."""

    def __init__(self, device: int = 0, rank: int = 0, world_size: int = 1, nccl_unique_id: Optional[bytes] = None,
                 verbose: int = 0):
        self.lib = load_library()
        cfg = DbaConfig()
        cfg.device, cfg.rank, cfg.world_size, cfg.verbose = device, rank, world_size, verbose
        self._id_buf = None
        if nccl_unique_id is not None:
            self._id_buf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            cfg.nccl_unique_id = C.cast(self._id_buf, C.c_void_p)
        self.h = C.c_void_p()
        st = self.lib.dba_create(C.byref(self.h), C.byref(cfg))
        if st != DBA_OK:
            msg = self.lib.dba_last_error(None)
            self.h = None
            raise EngineError(st, msg.decode() if msg else "")
        self.problem: Optional[ProblemMarshal] = None

    def _check(self, st: int):
        if st != DBA_OK:
            msg = self.lib.dba_last_error(self.h)
            raise EngineError(st, msg.decode() if msg else "")

    def close(self):
        if self.h:
            self.lib.dba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def problem_set(self, p: Problem):
        self.problem = ProblemMarshal(p)
        self._check(self.lib.dba_problem_set(self.h, C.byref(self.problem.struct)))

    def problem_update(self, obs_remove=None, pt_remove=None, freeze_camera: int = 0):
        """Drops flagged observations / points on the engine side (no second upload); returns (n_obs, n_pts)."""
        o = None if obs_remove is None else np.ascontiguousarray(obs_remove, dtype=np.uint8)
        t = None if pt_remove is None else np.ascontiguousarray(pt_remove, dtype=np.uint8)
        n_obs, n_pts = C.c_int64(0), C.c_int32(0)
        self._check(self.lib.dba_problem_update(self.h, None if o is None else o.ctypes.data, None if t is None else t.ctypes.data,
                                                int(freeze_camera), C.byref(n_obs), C.byref(n_pts)))
        # keep the marshalled problem in step with the engine (shapes of later eval / filter / params_get calls)
        p = self.problem.p
        keep_pt = np.ones(p.n_pts, bool) if t is None else (t == 0)
        keep_ob = (np.ones(p.n_obs, bool) if o is None else (o == 0)) & keep_pt[p.obs_pt]
        new_pt = np.cumsum(keep_pt) - 1
        q = p.copy()
        for k in ("obs_xy", "obs_pose_a", "obs_pose_b", "obs_intr"):
            setattr(q, k, getattr(p, k)[keep_ob])
        q.obs_pt = new_pt[p.obs_pt[keep_ob]].astype(np.int32)
        q.pts = p.pts[keep_pt]
        if p.pts_rgb is not None:
            q.pts_rgb = p.pts_rgb[keep_pt]
        q.obs_col0 = q.obs_col1 = None
        q.freeze_camera = int(freeze_camera)
        self.problem = ProblemMarshal(q)
        assert (n_obs.value, n_pts.value) == (q.n_obs, q.n_pts)
        return n_obs.value, n_pts.value

    def params_reset(self):
        self._check(self.lib.dba_params_reset(self.h))

    def eval(self, residuals=True, jacobians=False):
        n = self.problem.p.n_obs
        cost = C.c_double()
        res = np.zeros((n, 2)) if residuals else None
        jp = np.zeros((n, 2, 3)) if jacobians else None
        ja = np.zeros((n, 2, 6)) if jacobians else None
        jb = np.zeros((n, 2, 6)) if jacobians else None
        ji = np.zeros((n, 2, 3)) if jacobians else None
        self._check(self.lib.dba_eval(self.h, C.byref(cost), _ptr(res), _ptr(jp), _ptr(ja), _ptr(jb), _ptr(ji)))
        return {"cost": cost.value, "residuals": res, "jac_pt": jp, "jac_pose_a": ja, "jac_pose_b": jb, "jac_intr": ji}

    def solve(self, options: Optional[DbaSolveOptions] = None, capacity: int = 1024) -> Summary:
        o = options or default_options_struct()
        s = Summary(capacity)
        self._check(self.lib.dba_solve(self.h, C.byref(o), C.byref(s.struct)))
        return s

    def params_get(self, out=None):
        """Current parameters in caller order.  `out`: a dict returned by an earlier call, written in place
        (a caller that keeps its result buffers, as the C++ host mirror does, pays no page faults here)."""
        q = self.problem.p
        if out is None:
            out = {"pts": np.zeros_like(q.pts), "ext_rot": np.zeros_like(q.ext_rot), "ext_trans": np.zeros_like(q.ext_trans),
                   "intr_focal": np.zeros_like(q.intr_focal), "intr_dist": np.zeros_like(q.intr_dist)}
        else:
            for k, ref in (("pts", q.pts), ("ext_rot", q.ext_rot), ("ext_trans", q.ext_trans), ("intr_focal", q.intr_focal),
                           ("intr_dist", q.intr_dist)):
                if out[k].shape != ref.shape or out[k].dtype != np.float64 or not out[k].flags.c_contiguous:
                    raise ValueError(f"params_get(out=): {k} does not match the problem")
        self._check(self.lib.dba_params_get(self.h, _ptr(out["pts"]), _ptr(out["ext_rot"]), _ptr(out["ext_trans"]),
                                            _ptr(out["intr_focal"]), _ptr(out["intr_dist"])))
        return out

    def fit_hemisphere(self, centres: np.ndarray, centre0=(0.0, 0.0, 0.0), rho0: float = 1.0,
                       options: Optional[DbaSolveOptions] = None):
        centres = np.ascontiguousarray(centres, dtype=np.float64)
        c = np.array(centre0, dtype=np.float64)
        rho = C.c_double(rho0)
        o = options or make_options(max_num_iterations=1000)
        s = Summary(1024)
        self._check(self.lib.dba_fit_hemisphere(self.h, _ptr(centres), centres.shape[0], _ptr(c), C.byref(rho),
                                                C.byref(o), C.byref(s.struct)))
        return c, rho.value, s

    def filter(self, error_boundary: float, centre=None, rho: float = 0.0):
        """filterPoint3d decisions on the device: (obs_remove, pt_remove) uint8 arrays, caller order."""
        p = self.problem.p
        obs = np.zeros(p.n_obs, dtype=np.uint8)
        pts = np.zeros(p.n_pts, dtype=np.uint8)
        c = None if centre is None else np.ascontiguousarray(centre, dtype=np.float64)
        n_obs, n_pts = C.c_int64(0), C.c_int32(0)
        self._check(self.lib.dba_filter(self.h, float(error_boundary), None if c is None else c.ctypes.data, float(rho),
                                        obs.ctypes.data, pts.ctypes.data, C.byref(n_obs), C.byref(n_pts)))
        assert n_obs.value == int(obs.sum()) and n_pts.value == int(pts.sum())
        return obs, pts

    def filter_mse(self) -> np.ndarray:
        out = np.zeros(self.problem.p.n_obs)
        self._check(self.lib.dba_filter_mse(self.h, _ptr(out)))
        return out

    def kernel_stats_enable(self, enable: bool = True):
        self._check(self.lib.dba_kernel_stats_enable(self.h, int(enable)))

    def kernel_stats_reset(self):
        self._check(self.lib.dba_kernel_stats_reset(self.h))

    def kernel_stats(self):
        buf = (DbaKernelStat * 64)()
        n = self.lib.dba_kernel_stats(self.h, buf, 64)
        if n < 0:
            self._check(n)
        return [{"name": buf[i].name.decode(), "launches": buf[i].launches, "total_ms": buf[i].total_ms,
                 "algorithmic_bytes": buf[i].algorithmic_bytes} for i in range(n)]
