"""Synthetic bundle-adjustment problems in the flat SoA image of ``dba_problem``
(include/deeparc_ba.h) and as ``.deeparc`` text (reference format,
src/DeepArcManager.cc:26-164 reader / :426-499 writer).

The three real datasets the reference was run on are absent from the reference mount
(.MISSING_LARGE_BLOBS), so every workload here is generated:

* ``arc_rig``      – shared-extrinsic DeepArc rig (BASELINE.json configs[2]):
                     A arc poses x R ring poses, camera (a, r) = arc_a o ring_r with the
                     aliasing rules of src/ParameterBlock.hh:68-94 and
                     src/DeepArcManager.cc:166-171 (ring 0 == ext[0], gauge camera (0,0)).
* ``bal_like``     – non-shared, one extrinsic + one intrinsic per camera, nf=1, nd=2
                     (BASELINE.json configs[3], configs[4]).
* ``teabottle_like`` – stand-in for the missing teabottle_green*.deeparc files, built from the
                     two format sample lines the reference carries in comments
                     (src/DeepArcManager.cc:456, :475).  It is NOT the real dataset.

Everything is seeded (numpy Philox, counter based) and fp64.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

BASE_SEED = 20261018


@dataclasses.dataclass
class Problem:
    """Host-side flat problem; field names follow ``dba_problem``."""

    obs_xy: np.ndarray  # [n_obs, 2] f64
    obs_pt: np.ndarray  # [n_obs] i32
    obs_pose_a: np.ndarray  # [n_obs] i32
    obs_pose_b: np.ndarray  # [n_obs] i32 (-1 = none)
    obs_intr: np.ndarray  # [n_obs] i32
    pts: np.ndarray  # [n_pts, 3]
    ext_rot: np.ndarray  # [n_ext, 3]
    ext_trans: np.ndarray  # [n_ext, 3]
    intr_center: np.ndarray  # [n_intr, 2]
    intr_focal: np.ndarray  # [n_intr, 2]
    intr_dist: np.ndarray  # [n_intr, 2]
    intr_nf: np.ndarray  # [n_intr] i32
    intr_nd: np.ndarray  # [n_intr] i32
    ext_const: np.ndarray  # [n_ext] u8
    freeze_camera: int = 0
    free_intrinsics: int = 0
    # bookkeeping for the .deeparc writer (reference observation columns)
    n_arc: int = 0
    n_ring: int = 0
    obs_col0: Optional[np.ndarray] = None  # pos_arc / intrinsic_id column
    obs_col1: Optional[np.ndarray] = None  # pos_ring / extrinsic_id column
    pts_rgb: Optional[np.ndarray] = None
    truth: Optional[dict] = None
    name: str = ""

    @property
    def n_obs(self) -> int:
        return int(self.obs_pt.shape[0])

    @property
    def n_pts(self) -> int:
        return int(self.pts.shape[0])

    @property
    def n_ext(self) -> int:
        return int(self.ext_rot.shape[0])

    @property
    def n_intr(self) -> int:
        return int(self.intr_center.shape[0])

    def copy(self) -> "Problem":
        kw = {}
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            kw[f.name] = v.copy() if isinstance(v, np.ndarray) else v
        return Problem(**kw)

    def normalised(self) -> "Problem":
        """Contiguous arrays of the exact dtypes the C ABI expects.  Arrays that already conform are
        passed through without a copy (the C ABI copies what it needs; callers' buffers are not
        retained past dba_problem_set)."""
        p = dataclasses.replace(self)
        for name in ("obs_xy", "pts", "ext_rot", "ext_trans", "intr_center", "intr_focal", "intr_dist"):
            setattr(p, name, np.ascontiguousarray(getattr(p, name), dtype=np.float64))
        for name in ("obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr", "intr_nf", "intr_nd"):
            setattr(p, name, np.ascontiguousarray(getattr(p, name), dtype=np.int32))
        p.ext_const = np.ascontiguousarray(p.ext_const, dtype=np.uint8)
        return p


# ----------------------------------------------------------------------------- math
def rodrigues(w: np.ndarray) -> np.ndarray:
    """Angle-axis -> rotation matrix, batched [..., 3] -> [..., 3, 3] (exact formula)."""
    w = np.asarray(w, dtype=np.float64)
    theta = np.linalg.norm(w, axis=-1, keepdims=True)
    safe = np.where(theta > 0, theta, 1.0)
    k = w / safe
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s = np.sin(theta)[..., None]
    c = np.cos(theta)[..., None]
    eye = np.broadcast_to(np.eye(3), K.shape)
    return eye + s * K + (1.0 - c) * (K @ K)


def matrix_to_angle_axis(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> angle-axis for angles < pi (generator use only)."""
    R = np.asarray(R, dtype=np.float64)
    tr = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1.0) / 2.0, -1.0, 1.0)
    theta = np.arccos(tr)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    s = 2.0 * np.sin(theta)
    scale = np.where(np.abs(s) > 1e-12, theta / np.where(np.abs(s) > 1e-12, s, 1.0), 0.5)
    return v * scale[..., None]


def project(p: Problem, pts=None, ext_rot=None, ext_trans=None, chunk: int = 2_000_000) -> np.ndarray:
    """Predicted pixels of every observation (numpy restatement of the forward model used
    ONLY to synthesise observations; the parity oracle lives in oracle/).  Chunked so that the
    50M-observation workload does not materialise per-observation rotation matrices at once."""
    pts = p.pts if pts is None else pts
    ext_rot = p.ext_rot if ext_rot is None else ext_rot
    ext_trans = p.ext_trans if ext_trans is None else ext_trans
    R = rodrigues(ext_rot)
    n = p.obs_pt.shape[0]
    out = np.empty((n, 2))
    any_b = bool((p.obs_pose_b >= 0).any())
    for lo in range(0, n, chunk):
        sl = slice(lo, min(lo + chunk, n))
        X = pts[p.obs_pt[sl]]
        if any_b:
            has_b = p.obs_pose_b[sl] >= 0
            b = np.where(has_b, p.obs_pose_b[sl], 0)
            Xb = np.einsum("nij,nj->ni", R[b], X) + ext_trans[b]
            X = np.where(has_b[:, None], Xb, X)
        a = p.obs_pose_a[sl]
        cam = np.einsum("nij,nj->ni", R[a], X) + ext_trans[a]
        u = cam[:, 0] / cam[:, 2]
        v = cam[:, 1] / cam[:, 2]
        it = p.obs_intr[sl]
        fx = p.intr_focal[it, 0]
        fy = np.where(p.intr_nf[it] == 2, p.intr_focal[it, 1], p.intr_focal[it, 0])
        rr = u * u + v * v
        k0 = np.where(p.intr_nd[it] >= 1, p.intr_dist[it, 0], 0.0)
        k1 = np.where(p.intr_nd[it] >= 2, p.intr_dist[it, 1], 0.0)
        d = 1.0 + rr * (k0 + k1 * rr)
        out[sl, 0] = fx * d * u + p.intr_center[it, 0]
        out[sl, 1] = fy * d * v + p.intr_center[it, 1]
    return out


def _rng(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, stream]))


def _rot_axis(axis: int, angle: np.ndarray) -> np.ndarray:
    w = np.zeros(np.shape(angle) + (3,))
    w[..., axis] = angle
    return w


# ----------------------------------------------------------------- shared-extrinsic rig
def arc_rig(n_arc: int = 10, n_ring: int = 10, n_pts: int = 100_000, obs_per_point: int = 10,
            seed: int = BASE_SEED, pixel_sigma: float = 0.5, point_sigma: float = 1e-3,
            pose_sigma: float = 1e-3, focal: float = 4949.234294, center=(923.0, 1223.0),
            ring_span_deg: float = 360.0, arc_step_deg: float = 8.0, name: str = "arc_rig") -> Problem:
    """DeepArc rig: world frame == frame of camera (arc 0, ring 0); ext[0] = identity and constant.

    ring r (turntable yaw about the vertical axis through the object centre c):
        T_r X = R_y(psi_r) (X - c) + c
    arc a (camera elevation about c):  T_a X = R_x(phi_a) (X - c) + c
    camera (a, r): X_cam = T_a T_r X  — the composition order of
    src/snavely_reprojection_error.hh:96-108 (ring first, then arc).
    Camera centres all lie at distance |c| from c, so the hemisphere fit
    (src/hemisphere_radius.hh) recovers centre c and rho = |c|^2.
    """
    A, R = n_arc, n_ring
    n_ext = A + R - 1
    c = np.array([0.0, 0.0, 0.5])
    phi = np.deg2rad(arc_step_deg) * np.arange(A)
    psi = np.deg2rad(ring_span_deg / R) * np.arange(R)
    rot = np.zeros((n_ext, 3))
    rot[:A] = _rot_axis(0, phi)
    rot[A:] = _rot_axis(1, psi[1:])
    Rm = rodrigues(rot)
    trans = c[None, :] - np.einsum("nij,j->ni", Rm, c)
    trans[0] = 0.0

    g = _rng(seed, 1)
    # points: uniform in a ball of radius 0.1 around c
    d = g.standard_normal((n_pts, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = c[None, :] + d * (0.1 * g.random((n_pts, 1)) ** (1.0 / 3.0))

    n_cam = A * R
    k = min(obs_per_point, n_cam)
    # each point seen by k distinct cameras
    cams = np.argsort(g.random((n_pts, n_cam)), axis=1)[:, :k].astype(np.int32)
    arc = (cams // R).reshape(-1)
    ring = (cams % R).reshape(-1)
    obs_pt = np.repeat(np.arange(n_pts, dtype=np.int32), k)

    # src/ParameterBlock.hh:68-94: which extrinsics enter the residual
    ring_ext = np.where(ring == 0, 0, ring + A - 1).astype(np.int32)
    arc_only = ring == 0
    ring_only = (arc == 0) & (ring != 0)
    pose_a = np.where(ring_only, ring_ext, arc).astype(np.int32)
    pose_b = np.where(arc_only | ring_only, -1, ring_ext).astype(np.int32)

    n_intr = A
    prob = Problem(
        obs_xy=np.zeros((obs_pt.size, 2)), obs_pt=obs_pt, obs_pose_a=pose_a, obs_pose_b=pose_b,
        obs_intr=arc.astype(np.int32), pts=pts, ext_rot=rot, ext_trans=trans,
        intr_center=np.tile(np.asarray(center, dtype=np.float64), (n_intr, 1)),
        intr_focal=np.full((n_intr, 2), focal), intr_dist=np.zeros((n_intr, 2)),
        intr_nf=np.full(n_intr, 2, np.int32), intr_nd=np.zeros(n_intr, np.int32),
        ext_const=np.zeros(n_ext, np.uint8), n_arc=A, n_ring=R,
        obs_col0=arc.astype(np.int32), obs_col1=ring.astype(np.int32), name=name)
    prob.ext_const[0] = 1 if np.any((arc == 0) & (ring == 0)) else 0
    prob.obs_xy = project(prob) + pixel_sigma * g.standard_normal((obs_pt.size, 2))
    prob.truth = {"pts": pts.copy(), "ext_rot": rot.copy(), "ext_trans": trans.copy(),
                  "hemisphere_centre": c.copy(), "hemisphere_rho": float(c @ c)}
    # perturbed initial guess (gauge block stays exact)
    prob.pts = pts + point_sigma * g.standard_normal(pts.shape)
    prob.ext_rot = rot + pose_sigma * g.standard_normal(rot.shape)
    prob.ext_trans = trans + pose_sigma * g.standard_normal(trans.shape)
    prob.ext_rot[0] = rot[0]
    prob.ext_trans[0] = trans[0]
    prob.pts_rgb = g.integers(0, 256, size=(n_pts, 3)).astype(np.int32)
    return prob.normalised()


def teabottle_like(n_pts: int = 20_000, obs_per_point: int = 8, seed: int = BASE_SEED + 7) -> Problem:
    """STAND-IN for data/teabottle_green*.deeparc (missing from the reference mount):
    A=10 arcs, R=41 rings (assumed), intrinsics from the sample line at
    src/DeepArcManager.cc:456 with the principal point truncated to integers
    (src/Camera/Intrinsic.hh:24-27)."""
    return arc_rig(n_arc=10, n_ring=41, n_pts=n_pts, obs_per_point=obs_per_point, seed=seed,
                   focal=4949.234294, center=(923.0, 1223.0), arc_step_deg=6.0, name="teabottle_like")


# ---------------------------------------------------------------------------- BAL-like
def bal_like(n_cam: int = 1700, n_pts: int = 1_000_000, obs_per_point: int = 5, window: int = 50,
             seed: int = BASE_SEED, pixel_sigma: float = 0.5, point_sigma: float = 1e-2,
             pose_sigma: float = 1e-3, free_intrinsics: int = 1, shuffle_points: bool = False,
             name: str = "bal_like") -> Problem:
    """Non-shared problem: camera i = (intrinsic i, extrinsic i), nf=1, nd=2; every point is
    seen by ``obs_per_point`` distinct cameras out of a window of ``window`` consecutive
    cameras along a trajectory (banded co-visibility).  Observations are point-sorted."""
    g = _rng(seed, 2)
    window = min(window, n_cam)
    k = min(obs_per_point, window)
    # cameras: translate along x, look down +z, small random rotations
    cam_pos = np.stack([0.1 * np.arange(n_cam), 0.05 * g.standard_normal(n_cam), 0.05 * g.standard_normal(n_cam)], axis=1)
    rot = 0.05 * g.standard_normal((n_cam, 3))
    Rm = rodrigues(rot)
    trans = -np.einsum("nij,nj->ni", Rm, cam_pos)

    start = np.floor(np.arange(n_pts) * ((n_cam - window + 1) / max(n_pts, 1))).astype(np.int64)
    if shuffle_points:
        start = g.permutation(start)
    centre_x = 0.1 * (start + window / 2.0)
    pts = np.stack([centre_x + g.uniform(-1.5, 1.5, n_pts), g.uniform(-1.5, 1.5, n_pts), g.uniform(6.0, 10.0, n_pts)], axis=1)
    if n_pts * window <= 60_000_000:
        offs = np.argsort(g.random((n_pts, window)), axis=1)[:, :k]
    else:
        # large problems: k distinct offsets per point from a random base and a stride coprime
        # to the window (O(n_pts * k) memory instead of O(n_pts * window))
        strides = np.array([s_ for s_ in range(1, window) if np.gcd(s_, window) == 1], dtype=np.int64)
        base = g.integers(0, window, n_pts)
        stride = strides[g.integers(0, len(strides), n_pts)]
        offs = (base[:, None] + stride[:, None] * np.arange(k)[None, :]) % window
    cams = np.sort(start[:, None] + offs, axis=1).astype(np.int32).reshape(-1)
    obs_pt = np.repeat(np.arange(n_pts, dtype=np.int32), k)

    f = g.uniform(800.0, 1200.0, n_cam)
    k0 = 1e-2 * g.standard_normal(n_cam)
    k1 = 1e-3 * g.standard_normal(n_cam)
    prob = Problem(
        obs_xy=np.zeros((obs_pt.size, 2)), obs_pt=obs_pt, obs_pose_a=cams, obs_pose_b=np.full(cams.size, -1, np.int32),
        obs_intr=cams.copy(), pts=pts, ext_rot=rot, ext_trans=trans,
        intr_center=np.zeros((n_cam, 2)), intr_focal=np.stack([f, np.zeros(n_cam)], axis=1),
        intr_dist=np.stack([k0, k1], axis=1), intr_nf=np.ones(n_cam, np.int32), intr_nd=np.full(n_cam, 2, np.int32),
        ext_const=np.zeros(n_cam, np.uint8), free_intrinsics=free_intrinsics, n_arc=n_cam, n_ring=0,
        obs_col0=cams.copy(), obs_col1=cams.copy(), name=name)
    # gauge rule of src/sfm.cc:50-53: blocks with column0 == 0 and column1 == 0 pin their pose
    prob.ext_const[0] = 1 if np.any(cams == 0) else 0
    prob.obs_xy = project(prob) + pixel_sigma * g.standard_normal((obs_pt.size, 2))
    prob.truth = {"pts": pts.copy(), "ext_rot": rot.copy(), "ext_trans": trans.copy(),
                  "intr_focal": prob.intr_focal.copy(), "intr_dist": prob.intr_dist.copy()}
    prob.pts = pts + point_sigma * g.standard_normal(pts.shape)
    prob.ext_rot = rot + pose_sigma * g.standard_normal(rot.shape)
    prob.ext_trans = trans + pose_sigma * g.standard_normal(trans.shape)
    prob.ext_rot[0] = rot[0]
    prob.ext_trans[0] = trans[0]
    if free_intrinsics:
        prob.intr_focal[:, 0] = f * (1.0 + 1e-3 * g.standard_normal(n_cam))
    prob.pts_rgb = np.full((n_pts, 3), 255, np.int32)
    return prob.normalised()


# ------------------------------------------------------------------- .deeparc text I/O
def write_deeparc(p: Problem, path: str, version: float = 0.01, rotation_format: int = 3,
                  center_override=None) -> None:
    """Text format read by DeepArcManager::read (src/DeepArcManager.cc:26-164):
    version; n_obs n_intr n_arc n_ring n_pts; observations ``c0 c1 point x y``; intrinsics
    ``cx cy nf f.. nd k..``; extrinsics ``tx ty tz nrot rot..`` (nrot 3 angle-axis, 4 quaternion
    wxyz, 9 column-major matrix); points ``x y z r g b``.  Full precision (repr) so a
    read-back reproduces the arrays bit for bit."""
    assert p.obs_col0 is not None and p.obs_col1 is not None
    with open(path, "w") as fh:
        fh.write(f"{version:.6f}\n")
        fh.write(f"{p.n_obs} {p.n_intr} {p.n_arc} {p.n_ring} {p.n_pts}\n")
        xy = p.obs_xy
        lines = [f"{a} {b} {c} {x!r} {y!r}\n" for a, b, c, x, y in
                 zip(p.obs_col0.tolist(), p.obs_col1.tolist(), p.obs_pt.tolist(), xy[:, 0].tolist(), xy[:, 1].tolist())]
        fh.writelines(lines)
        for i in range(p.n_intr):
            cx, cy = (p.intr_center[i] if center_override is None else center_override)
            nf, nd = int(p.intr_nf[i]), int(p.intr_nd[i])
            toks = [repr(float(cx)), repr(float(cy)), str(nf)] + [repr(float(v)) for v in p.intr_focal[i, :nf]]
            toks += [str(nd)] + [repr(float(v)) for v in p.intr_dist[i, :nd]]
            fh.write(" ".join(toks) + "\n")
        Rm = rodrigues(p.ext_rot)
        for i in range(p.n_ext):
            t = p.ext_trans[i]
            toks = [repr(float(v)) for v in t]
            if rotation_format == 3:
                toks += ["3"] + [repr(float(v)) for v in p.ext_rot[i]]
            elif rotation_format == 9:
                toks += ["9"] + [repr(float(v)) for v in Rm[i].T.reshape(-1)]  # column-major
            elif rotation_format == 4:
                w = p.ext_rot[i]
                th = float(np.linalg.norm(w))
                q = [1.0, 0.0, 0.0, 0.0] if th == 0 else [np.cos(th / 2)] + list(np.sin(th / 2) * w / th)
                toks += ["4"] + [repr(float(v)) for v in q]
            else:
                raise ValueError("rotation_format must be 3, 4 or 9")
            fh.write(" ".join(toks) + "\n")
        rgb = p.pts_rgb if p.pts_rgb is not None else np.full((p.n_pts, 3), 255, np.int32)
        lines = [f"{x!r} {y!r} {z!r} {r} {g} {b}\n" for (x, y, z), (r, g, b) in zip(p.pts.tolist(), rgb.tolist())]
        fh.writelines(lines)


def read_deeparc_text(path: str) -> dict:
    """Tokenises a ``.deeparc`` file into its sections (test helper; the product reader is the
    C++ DeepArcManager)."""
    toks = open(path).read().split()
    pos = 0

    def take(n):
        nonlocal pos
        out = toks[pos:pos + n]
        pos += n
        return out

    version = float(take(1)[0])
    n_obs, n_intr, n_arc, n_ring, n_pts = (int(t) for t in take(5))
    n_ext = n_arc + n_ring - 1 if n_ring else n_arc
    obs = np.array(take(5 * n_obs), dtype=np.float64).reshape(n_obs, 5)
    intr = []
    for _ in range(n_intr):
        cx, cy = (float(t) for t in take(2))
        nf = int(take(1)[0])
        f = [float(t) for t in take(nf)]
        nd = int(take(1)[0])
        k = [float(t) for t in take(nd)]
        intr.append((cx, cy, f, k))
    ext = []
    for _ in range(n_ext):
        t = [float(v) for v in take(3)]
        nr = int(take(1)[0])
        r = [float(v) for v in take(nr)]
        ext.append((t, r))
    pts = np.array(take(6 * n_pts), dtype=np.float64).reshape(n_pts, 6)
    return {"version": version, "n_arc": n_arc, "n_ring": n_ring, "obs": obs, "intr": intr, "ext": ext, "pts": pts}
