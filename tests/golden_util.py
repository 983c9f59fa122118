"""Loads tests/golden/snavely_kat.json (50-digit mpmath vectors, see tests/golden/make_golden.py)
as one flat Problem with one observation per case (own point, intrinsic and two extrinsics each)."""
import json
import os

import numpy as np

from deeparc_sfm_b200.synthetic import Problem

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "snavely_kat.json")


def load_golden():
    data = json.load(open(PATH))
    cases = data["cases"]
    n = len(cases)
    P = np.array([[float.fromhex(h) for h in c["params"]] for c in cases])
    obs = np.array([[float.fromhex(h) for h in c["obs"]] for c in cases])
    compose = np.array([c["compose"] for c in cases], dtype=bool)
    p = Problem(
        obs_xy=obs, obs_pt=np.arange(n, dtype=np.int32), obs_pose_a=(2 * np.arange(n)).astype(np.int32),
        obs_pose_b=np.where(compose, 2 * np.arange(n) + 1, -1).astype(np.int32), obs_intr=np.arange(n, dtype=np.int32),
        pts=P[:, 0:3].copy(), ext_rot=np.stack([P[:, 9:12], P[:, 15:18]], axis=1).reshape(2 * n, 3),
        ext_trans=np.stack([P[:, 12:15], P[:, 18:21]], axis=1).reshape(2 * n, 3), intr_center=P[:, 3:5].copy(),
        intr_focal=P[:, 5:7].copy(), intr_dist=P[:, 7:9].copy(),
        intr_nf=np.array([c["nf"] for c in cases], np.int32), intr_nd=np.array([c["nd"] for c in cases], np.int32),
        ext_const=np.zeros(2 * n, np.uint8), name="golden")
    res = np.array([[float(v) for v in c["residual"]] for c in cases])
    J = np.array([[[float(v) for v in row] for row in c["jacobian"]] for c in cases])  # [n, 2, 21]
    truth = {"residuals": res, "jac_pt": J[:, :, 0:3], "jac_pose_a": J[:, :, 9:15],
             "jac_pose_b": np.where(compose[:, None, None], J[:, :, 15:21], 0.0),
             # d/d(f, k0, k1): f column = d/d focal[0]; distortion columns that are not parameters are 0
             "jac_intr": np.stack([J[:, :, 5], J[:, :, 7], J[:, :, 8]], axis=2),
             "kinds": [c["kind"] for c in cases], "nf": p.intr_nf.copy(), "nd": p.intr_nd.copy()}
    return p.normalised(), truth
