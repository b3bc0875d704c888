"""Generates the golden fixtures of tests/golden/ (run in the build container, where /root/reference exists).

  python tests/golden/make_golden.py

1. pyproto_derivs.json — outputs of the reference's own Python prototype
   (/root/reference/py_proto/suriko/bundle_adjustment_kanatani_impl.py, imported unmodified) on a small ragged scene with
   f0 = K[2,2] = 1, the regime in which the prototype and the C++ implementation use the same formulas:
   the normalised state it differentiates, gradE / E / G / F of the first derivative pass, the first two-phase correction
   vector at hessian_factor = 1e-4 and the error before / after the first accepted step.
2. oracle_traces.json — per-iteration traces of the CPU oracle in "exact" and "faithful" mode for fixed seeds, so that the
   oracle itself is pinned against silent edits.

The fixtures are committed; the GPU box never needs /root/reference.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def small_scene(seed=5, M=6, N=24, drop=0.25):
    rng = np.random.default_rng(seed)
    th = np.linspace(-0.9, 0.9, M)
    pos = np.stack([6.0 * np.sin(th), 0.4 * np.cos(3 * th) - 0.3 * np.arange(M), -6.0 * np.cos(th)], axis=1)
    X = rng.uniform(-1.0, 1.0, (N, 3)) * np.array([1.2, 1.0, 0.8])
    cams = []
    for m in range(M):
        f = -pos[m] / np.linalg.norm(pos[m])
        x = np.cross(np.array([0.0, 1.0, 0.0]), f); x /= np.linalg.norm(x)
        y = np.cross(f, x)
        R = np.stack([x, y, f])
        cams.append((R, -R @ pos[m]))
    K = np.array([[880.0, 0.0, 400.0], [0.0, 660.0, 300.0], [0.0, 0.0, 1.0]])
    pix = [[None] * M for _ in range(N)]
    for p in range(N):
        seen = rng.uniform(size=M) > drop
        seen[rng.integers(0, M, 3)] = True
        first = int(np.argmax(seen))
        for m in range(first, M):          # contiguous-from-first visibility with gaps marked None
            if not seen[m]:
                continue
            xc = cams[m][0] @ X[p] + cams[m][1]
            h = K @ (xc / xc[2])
            pix[p][m] = h[:2] + rng.normal(0.0, 0.5, 2)
    Xn = X + rng.normal(0.0, 0.01, X.shape)
    cams_n = []
    for (R, T) in cams:
        w = rng.normal(0.0, 0.003, 3)
        ang = np.linalg.norm(w); n = w / ang
        S = np.array([[0, -n[2], n[1]], [n[2], 0, -n[0]], [-n[1], n[0], 0]])
        dR = np.eye(3) + np.sin(ang) * S + (1 - np.cos(ang)) * S @ S
        cams_n.append((dR @ R, T + rng.normal(0.0, 0.01, 3)))
    return K, Xn, cams_n, pix


def run_prototype():
    sys.path.insert(0, "/root/reference/py_proto")
    # suriko.mvg (only needed for the 13-line PointLife record) imports matplotlib/pygame, absent here: provide the record
    import suriko.bundle_adjustment_kanatani_impl as impl

    class PointLife:
        def __init__(self):
            self.track_id = None
            self.points_list_pixel = []

    K, X, cams, pix = small_scene()
    N, M = X.shape[0], len(cams)
    lives = []
    for p in range(N):
        pl = PointLife(); pl.track_id = p; pl.points_list_pixel = [None if v is None else np.array(v) for v in pix[p]]
        lives.append(pl)
    world = [X[p].copy() for p in range(N)]
    poses = [(R.copy(), T.copy()) for (R, T) in cams]
    Ks = [K.copy() for _ in range(M)]

    ba = impl.BundleAdjustmentKanatani(min_err_change_abs=0.0, min_err_change_rel=None, max_iter=1, debug=0)
    cap = {}
    cls = impl.BundleAdjustmentKanatani
    orig_deriv = getattr(cls, "_BundleAdjustmentKanatani__ComputeDerivativesCloseForm")
    orig_corr = getattr(cls, "_BundleAdjustmentKanatani__EstimateCorrectionsDecomposedInTwoPhases")

    def deriv_wrap(self, points_count, frames_count, check, gradE, gradE2, dsp, dsf, dspf):
        if "state_points" not in cap:
            cap["state_points"] = np.array([self.world_pnts[i] for i in self.bundle_pnt_ids])
            cap["state_cams"] = [(R.copy(), T.copy()) for (R, T) in self.framei_from_world_RT_list]
            cap["err_initial"] = float(self.err_value)
        orig_deriv(self, points_count, frames_count, check, gradE, gradE2, dsp, dsf, dspf)
        if "gradE" not in cap:
            cap["gradE"] = gradE.copy(); cap["E"] = dsp.copy(); cap["G"] = dsf.copy(); cap["F"] = dspf.copy()

    def corr_wrap(self, points_count, frames_count, hessian_factor, gradE, dsp, dsf, dspf, matG, left, right, corrections, ref=None):
        orig_corr(self, points_count, frames_count, hessian_factor, gradE, dsp, dsf, dspf, matG, left, right, corrections, ref)
        if "corrections" not in cap:
            cap["corrections"] = corrections.copy(); cap["hessian_factor"] = float(hessian_factor)
            cap["S"] = (matG - left).copy(); cap["rhs"] = right.copy()   # the function rebinds its local: left holds sum(F^T E^-1 F)

    setattr(cls, "_BundleAdjustmentKanatani__ComputeDerivativesCloseForm", deriv_wrap)
    setattr(cls, "_BundleAdjustmentKanatani__EstimateCorrectionsDecomposedInTwoPhases", corr_wrap)
    ba.ComputeInplace(lives, world, poses, cam_mat_pixel_from_meter_list=Ks)
    cap["err_after_first_attempt_sequence"] = float(ba.err_value)

    # flat problem of the NORMALISED state the prototype differentiated
    obs_cam, obs_pt, obs_xy = [], [], []
    for p in range(N):
        for m in range(M):
            if pix[p][m] is not None:
                obs_cam.append(m); obs_pt.append(p); obs_xy.append([float(pix[p][m][0]), float(pix[p][m][1])])
    cams_flat = [list(map(float, T)) + [float(v) for v in np.asarray(R).T.reshape(9)] for (R, T) in cap["state_cams"]]
    out = {
        "source": "reference py_proto/suriko/bundle_adjustment_kanatani_impl.py (unmodified), f0 = K22 = 1",
        "n_cams": M, "n_points": N, "f0": 1.0,
        "obs_cam": obs_cam, "obs_point": obs_pt, "obs_xy": obs_xy,
        "points": cap["state_points"].tolist(), "cams": cams_flat, "K": [float(v) for v in K.T.reshape(9)],
        "err_initial": cap["err_initial"], "hessian_factor": cap["hessian_factor"],
        "gradE": cap["gradE"].tolist(), "E": cap["E"].tolist(), "G": cap["G"].tolist(), "F": cap["F"].tolist(),
        "S": cap["S"].tolist(), "rhs": cap["rhs"].tolist(), "corrections": cap["corrections"].tolist(),
    }
    with open(os.path.join(HERE, "pyproto_derivs.json"), "w") as f:
        json.dump(out, f)
    print("pyproto_derivs.json: %d cams, %d points, %d obs, err_initial %.6f" % (M, N, len(obs_cam), cap["err_initial"]))


def run_oracle_traces():
    import oracle_lib as ol
    from surikatoko_b200 import scenes
    ol.build()
    out = {}
    cases = {"circle_grid_cell0.5": (ol.circle_grid_scene(cell_x=0.5, cell_y=0.5), 1e-10, 6),
             "ring_30x600x6_seed3": (None, 1e-10, 5), "dino_shape_300pts_seed4": (None, 4.56e-8, 5)}
    pr = scenes.ring_scene(30, 600, 6, seed=3)
    cases["ring_30x600x6_seed3"] = (ol.Problem(pr.obs_cam, pr.obs_point, pr.obs_xy, pr.points, pr.cams, pr.K, False, pr.f0), 1e-10, 5)
    pr = scenes.dino_shaped_scene(n_points=300, n_obs=1000, seed=4)
    cases["dino_shape_300pts_seed4"] = (ol.Problem(pr.obs_cam, pr.obs_point, pr.obs_xy, pr.points, pr.cams, pr.K, False, pr.f0), 4.56e-8, 5)
    for name, (p, thr, iters) in cases.items():
        ex = ol.ba_solve(p, err_change=thr, max_outer_iters=iters, flow="sparse", solve="chol", acc="ld")
        fa = ol.ba_solve(p, err_change=thr, max_outer_iters=iters, flow="dense", solve="qr", acc="double")
        out[name] = {"n_cams": p.n_cams, "n_points": p.n_points, "n_obs": p.n_obs, "err_change": thr, "max_outer_iters": iters,
                     "seen_points": ex.seen_points, "err_initial": ex.err_initial,
                     "exact": {"err_trace": ex.err_trace.tolist(), "attempts": ex.attempts.tolist(), "stop_reason": ex.stop_reason, "converged": ex.converged,
                               "points_head": ex.points[:4].tolist(), "cams_head": ex.cams[:2].tolist()},
                     "faithful": {"err_trace": fa.err_trace.tolist(), "attempts": fa.attempts.tolist(), "stop_reason": fa.stop_reason, "converged": fa.converged}}
        print(name, ex.stop_reason, len(ex.err_trace), ex.err_initial, "->", ex.err_final)
    with open(os.path.join(HERE, "oracle_traces.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    if os.path.isdir("/root/reference/py_proto"):
        run_prototype()
    else:
        print("no /root/reference: skipping the prototype fixtures")
    run_oracle_traces()
