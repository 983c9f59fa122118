#!/usr/bin/env python
"""Generates tests/golden/snavely_kat.json: known-answer vectors for the Snavely reprojection
residual and its Jacobian, evaluated with 50-digit mpmath arithmetic.

The reference ships no tests or golden vectors, so these are produced here:
  * the residual follows reference src/snavely_reprojection_error.hh:38-118 operation by
    operation (rotatePoint -> ceres::AngleAxisRotatePoint incl. its theta^2 <= DBL_EPSILON
    first-order branch [Ceres-upstream], compose order ring-then-arc, projectPoint with
    nf in {1,2}, nd in {0,1,2}, no sign flip), on the float64 inputs, in 50-digit arithmetic;
  * the Jacobian is the exact derivative of that same function (mpmath.diff at 50 digits);
  * when oracle/_ref/libdeeparc_ref.so is present (built from /root/reference, which exists
    only in the authoring container), every vector is also checked here against the
    reference's OWN compiled functor before it is written.
Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
mp.mp.dps = 50
DBL_EPS = mp.mpf(2) ** -52


def rotate(w, x):
    th2 = sum(a * a for a in w)
    if th2 > DBL_EPS:
        th = mp.sqrt(th2)
        c, s = mp.cos(th), mp.sin(th)
        k = [a / th for a in w]
        kx = [k[1] * x[2] - k[2] * x[1], k[2] * x[0] - k[0] * x[2], k[0] * x[1] - k[1] * x[0]]
        kd = (k[0] * x[0] + k[1] * x[1] + k[2] * x[2]) * (1 - c)
        return [x[i] * c + kx[i] * s + k[i] * kd for i in range(3)]
    wx = [w[1] * x[2] - w[2] * x[1], w[2] * x[0] - w[0] * x[2], w[0] * x[1] - w[1] * x[0]]
    return [x[i] + wx[i] for i in range(3)]


def residual(params, obs, nf, nd, compose):
    """params: flat list [X(3), c(2), f(2), k(2), wa(3), ta(3), wb(3), tb(3)] of mpf."""
    X, c, f, k = params[0:3], params[3:5], params[5:7], params[7:9]
    wa, ta, wb, tb = params[9:12], params[12:15], params[15:18], params[18:21]
    if compose:
        m = rotate(wb, X)
        m = [m[i] + tb[i] for i in range(3)]
    else:
        m = X
    p = rotate(wa, m)
    p = [p[i] + ta[i] for i in range(3)]
    u, v = p[0] / p[2], p[1] / p[2]
    fx = f[0]
    fy = f[1] if nf == 2 else f[0]
    rr = u * u + v * v
    d = mp.mpf(1)
    if nd == 2:
        d = 1 + rr * (k[0] + k[1] * rr)
    if nd == 1:
        d = 1 + rr * k[0]
    return [fx * d * u + c[0] - obs[0], fy * d * v + c[1] - obs[1]]


def jacobian(params, obs, nf, nd, compose):
    J = [[mp.mpf(0)] * 21 for _ in range(2)]
    for j in range(21):
        for r in range(2):
            def g(t, j=j, r=r):
                q = list(params)
                q[j] = t
                return residual(q, obs, nf, nd, compose)[r]
            J[r][j] = mp.diff(g, params[j])
    return J


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    specs = []
    for nf in (1, 2):
        for nd in (0, 1, 2):
            for compose in (0, 1):
                specs.append((nf, nd, compose, "random"))
    specs += [(1, 2, 0, "zero_rot"), (2, 0, 1, "zero_rot"), (1, 2, 0, "tiny_rot_below_eps"), (1, 2, 0, "tiny_rot_above_eps"),
              (2, 0, 1, "tiny_rot_below_eps"), (1, 0, 0, "near_pi"), (2, 1, 1, "near_pi"), (1, 2, 0, "bal_like"),
              (2, 0, 1, "deeparc_sample_line")]
    for (nf, nd, compose, kind) in specs:
        for rep in range(2):
            X = rng.uniform(-0.1, 0.1, 3) + np.array([0, 0, 0.5])
            c = np.array([923.0, 1223.0])
            f = np.array([4949.234294, 4949.234294 * 1.003])
            k = np.array([1e-2, -2e-3]) * rng.standard_normal(2)
            wa = 0.3 * rng.standard_normal(3)
            ta = 0.05 * rng.standard_normal(3)
            wb = 0.5 * rng.standard_normal(3)
            tb = 0.05 * rng.standard_normal(3)
            if kind == "zero_rot":
                wa[:] = 0.0
                wb[:] = 0.0
            elif kind == "tiny_rot_below_eps":
                wa = 1e-9 * rng.standard_normal(3)
                wb = 3e-9 * rng.standard_normal(3)
            elif kind == "tiny_rot_above_eps":
                wa = np.array([2e-8, -1e-8, 1.5e-8])
            elif kind == "near_pi":
                # roll about (almost) the optical axis keeps the point in front of the camera
                wa = np.array([0.01, -0.02, 1.0]) + 0.01 * rng.standard_normal(3)
                wa = wa / np.linalg.norm(wa) * (np.pi - 1e-3)
            elif kind == "bal_like":
                X = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(6, 10)])
                c = np.zeros(2)
                f = np.array([rng.uniform(800, 1200), 0.0])
                ta = np.array([0.1, -0.05, 0.02])
                wa = 0.05 * rng.standard_normal(3)
            elif kind == "deeparc_sample_line":  # src/DeepArcManager.cc:456, :475
                ta = np.array([-0.000454, 0.371719, -0.037265])
                wa = np.array([0.059579, -0.003424, 0.003330])
                f = np.array([4949.234294, 4949.234294])
            params64 = np.concatenate([X, c, f, k, wa, ta, wb, tb]).astype(np.float64)
            pm = [mp.mpf(float(v)) for v in params64]
            pred = residual(pm, [mp.mpf(0), mp.mpf(0)], nf, nd, compose)
            obs64 = np.array([float(pred[0]), float(pred[1])]) + rng.standard_normal(2) * 0.5
            obs = [mp.mpf(float(v)) for v in obs64]
            r = residual(pm, obs, nf, nd, compose)
            J = jacobian(pm, obs, nf, nd, compose)
            cases.append({"kind": kind, "nf": nf, "nd": nd, "compose": compose,
                          "params": [float(v).hex() for v in params64], "obs": [float(v).hex() for v in obs64],
                          "residual": [mp.nstr(v, 30) for v in r],
                          "jacobian": [[mp.nstr(v, 30) for v in row] for row in J]})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "snavely_kat.json")

    # cross-check against the reference's own compiled functor when available
    checked = 0
    try:
        from tests import oracle_lib
        if os.path.exists(oracle_lib.REF_PATH):
            import ctypes as C
            R = oracle_lib.Reference()
            dp = C.POINTER(C.c_double)
            R.lib.ref_functor.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(dp), dp, C.POINTER(dp)]
            for case in cases:
                p = np.array([float.fromhex(h) for h in case["params"]])
                o = [float.fromhex(h) for h in case["obs"]]
                blocks = [p[0:3], p[3:5], p[5:7], p[7:9], p[9:12], p[12:15], p[15:18], p[18:21]]
                blocks = [np.ascontiguousarray(b) for b in blocks]
                arr = (dp * 8)(*[b.ctypes.data_as(dp) for b in blocks])
                res = np.zeros(2)
                st = R.lib.ref_functor(o[0], o[1], case["nf"], case["nd"], case["compose"], arr, res.ctypes.data_as(dp), None)
                assert st == 0
                truth = np.array([float(mp.mpf(s)) for s in case["residual"]])
                pred = np.abs(truth + np.array(o))
                assert np.all(np.abs(res - truth) <= 1e-10 * np.abs(truth) + 64 * np.finfo(float).eps * pred), (case["kind"], res, truth)
                checked += 1
    except ImportError:
        pass
    json.dump({"generator": "tests/golden/make_golden.py", "precision_digits": 50,
               "layout": "params = [X(3), centre(2), focal(2), dist(2), rot_a(3), trans_a(3), rot_b(3), trans_b(3)] as float64 hex; "
                         "jacobian[row][col] in the same column order",
               "checked_against_reference_functor": checked, "cases": cases}, open(out, "w"), indent=0)
    print(f"wrote {len(cases)} cases to {out}; {checked} cross-checked against the reference functor")


if __name__ == "__main__":
    main()
