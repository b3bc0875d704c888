"""TEST INFRASTRUCTURE ONLY — ctypes binding of the CPU oracle (oracle/_build/libsrk_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (surikatoko_b200) never does.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsrk_oracle.so")

STOP_REASONS = {0: "", 1: "abs err threshold", 2: "small relative err change", 3: "hessian overflow",
                4: "err converged to limit value", 5: "normalization failed", 6: "max iterations"}


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Report(C.Structure):
    _fields_ = [("converged", C.c_int32), ("stop_reason", C.c_int32), ("outer_iters", C.c_int32), ("n_attempts", C.c_int32),
                ("err_initial", C.c_double), ("err_final", C.c_double), ("hessian_factor_final", C.c_double),
                ("seen_points", C.c_int64), ("world_scale", C.c_double), ("seconds", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


@dataclass
class Problem:
    """Flat BA problem (layout of include/srk/ba_c_api.h)."""
    obs_cam: np.ndarray
    obs_point: np.ndarray
    obs_xy: np.ndarray
    points: np.ndarray
    cams: np.ndarray
    K: np.ndarray
    shared_K: bool
    f0: float
    gt_points: np.ndarray = None
    gt_cams: np.ndarray = None

    @property
    def n_cams(self): return self.cams.shape[0]
    @property
    def n_points(self): return self.points.shape[0]
    @property
    def n_obs(self): return self.obs_cam.shape[0]

    def copy(self):
        return Problem(self.obs_cam.copy(), self.obs_point.copy(), self.obs_xy.copy(), self.points.copy(), self.cams.copy(),
                       self.K.copy(), self.shared_K, self.f0, self.gt_points, self.gt_cams)


@dataclass
class Result:
    converged: bool
    stop_reason: str
    outer_iters: int
    err_initial: float
    err_final: float
    seen_points: int
    err_trace: np.ndarray
    attempts: np.ndarray  # [n,4]: hessian_factor, err_new, accepted, skipped_points
    seconds: float
    points: np.ndarray = None
    cams: np.ndarray = None
    world_scale: float = 0.0


def read_bundle(path):
    """Independent (numpy) reader of the bundle format of include/srk/bundle_c_api.h: the oracle's side of a shared scene file."""
    raw = open(path, "rb").read()
    if raw[:8] != b"SRKBNDL1":
        raise ValueError("not a bundle file")
    n_cams, n_points, n_obs = (int(v) for v in np.frombuffer(raw, dtype="<i8", count=3, offset=8))
    shared_K = int(np.frombuffer(raw, dtype="<i4", count=1, offset=32)[0]) != 0
    f0 = float(np.frombuffer(raw, dtype="<f8", count=1, offset=40)[0])
    off = 48

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a
    obs_cam = take("<i4", n_obs); obs_point = take("<i4", n_obs); obs_xy = take("<f8", 2 * n_obs).reshape(-1, 2)
    points = take("<f8", 3 * n_points).reshape(-1, 3); cams = take("<f8", 12 * n_cams).reshape(-1, 12)
    K = take("<f8", 9 * (1 if shared_K else n_cams)).reshape(-1, 9)
    h = 1469598103934665603
    # FNV-1a 64 over the bytes before the checksum (pure-Python loop: small files only; large ones skip the check)
    if off <= (1 << 20):
        for b in raw[:off]:
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        if h != int(np.frombuffer(raw, dtype="<u8", count=1, offset=off)[0]):
            raise ValueError("bundle checksum mismatch")
    if len(raw) != off + 8:
        raise ValueError("bundle length mismatch")
    return Problem(obs_cam, obs_point, obs_xy, points, cams, K, shared_K, f0)


_FLOW = {"dense": 0, "sparse": 1, "threaded": 2}   # "threaded": sparse arithmetic over host threads + skyline Cholesky (the timed CPU arm at full size)
_SOLVE = {"qr": 0, "chol": 1, "none": 2}   # "none": assemble S / rhs only (large systems; the corrections are then meaningless)


def _prob_args(pr):
    return (C.c_int64(pr.n_cams), C.c_int64(pr.n_points), C.c_int64(pr.n_obs), _p(pr.obs_cam, C.c_int32), _p(pr.obs_point, C.c_int32),
            _p(pr.obs_xy, C.c_double))


def ba_solve(pr, err_change=None, max_hessian_factor=None, unity_ind=1, unity_val=1.0, max_outer_iters=0,
             flow="dense", solve="qr", acc="double", cap=4096):
    """Oracle ComputeInplace.  Returns Result; pr is not modified."""
    pts = np.ascontiguousarray(pr.points, dtype=np.float64).copy()
    cams = np.ascontiguousarray(pr.cams, dtype=np.float64).copy()
    rep = _Report()
    tr = np.zeros(cap)
    att = np.zeros((cap, 4))
    rc = lib().srk_oracle_ba_solve(*_prob_args(pr), _p(pts, C.c_double), _p(cams, C.c_double), _p(pr.K, C.c_double),
                                   C.c_int(1 if pr.shared_K else 0), C.c_double(pr.f0),
                                   C.c_int(err_change is not None), C.c_double(err_change or 0.0),
                                   C.c_int(max_hessian_factor is not None), C.c_double(max_hessian_factor or 0.0),
                                   C.c_int(unity_ind), C.c_double(unity_val), C.c_int(max_outer_iters),
                                   C.c_int(_FLOW[flow]), C.c_int(_SOLVE[solve]),
                                   C.c_int(0 if acc == "double" else 1), C.byref(rep), _p(tr, C.c_double), C.c_int(cap),
                                   _p(att, C.c_double), C.c_int(cap))
    if rc != 0:
        raise RuntimeError("oracle ba_solve failed rc=%d" % rc)
    n_acc = int(np.sum(att[:rep.n_attempts, 2] > 0))
    return Result(bool(rep.converged), STOP_REASONS[rep.stop_reason], rep.outer_iters, rep.err_initial, rep.err_final, rep.seen_points,
                  tr[:n_acc].copy(), att[:rep.n_attempts].copy(), rep.seconds, pts, cams, rep.world_scale)


def reproj_error(pr):
    err = C.c_double()
    seen = C.c_int64()
    rc = lib().srk_oracle_reproj_error(*_prob_args(pr), _p(pr.points, C.c_double), _p(pr.cams, C.c_double), _p(pr.K, C.c_double),
                                       C.c_int(1 if pr.shared_K else 0), C.c_double(pr.f0), C.byref(err), C.byref(seen))
    if rc != 0:
        raise RuntimeError("oracle reproj_error failed")
    return err.value, seen.value


def normalize(points, cams, unity_ind=1, unity_val=1.0):
    """NormalizeSceneInplace on copies.  Returns (ok, points, cams, cam0_prenorm, world_scale)."""
    pts = np.ascontiguousarray(points, dtype=np.float64).copy()
    cs = np.ascontiguousarray(cams, dtype=np.float64).copy()
    cam0 = np.zeros(12)
    ws = C.c_double()
    rc = lib().srk_oracle_normalize(C.c_int64(cs.shape[0]), C.c_int64(pts.shape[0]), _p(pts, C.c_double), _p(cs, C.c_double),
                                    C.c_int(unity_ind), C.c_double(unity_val), C.c_int(0), _p(cam0, C.c_double), C.byref(ws))
    return rc == 0, pts, cs, cam0, ws.value


def revert_normalization(points, cams, cam0_prenorm, world_scale):
    pts = np.ascontiguousarray(points, dtype=np.float64).copy()
    cs = np.ascontiguousarray(cams, dtype=np.float64).copy()
    ws = C.c_double(world_scale)
    cam0 = np.ascontiguousarray(cam0_prenorm, dtype=np.float64).copy()
    lib().srk_oracle_normalize(C.c_int64(cs.shape[0]), C.c_int64(pts.shape[0]), _p(pts, C.c_double), _p(cs, C.c_double),
                               C.c_int(1), C.c_double(1.0), C.c_int(1), _p(cam0, C.c_double), C.byref(ws))
    return pts, cs


def derivs_and_solve(pr, c=None, unity_ind=1, flow="sparse", solve="qr", acc="double"):
    """One derivative pass (+ optional two-phase solve at damping c) on the scene as given (already normalised)."""
    N, M, O = pr.n_points, pr.n_cams, pr.n_obs
    nf = 10 * M - 7
    out = dict(gradE=np.zeros(3 * N + 10 * M), E=np.zeros((N, 3, 3)), G=np.zeros((M, 10, 10)), F=np.zeros((O, 3, 10)))
    if c is not None:
        out.update(S=np.zeros((nf, nf)), rhs=np.zeros(nf), skipped=np.zeros(N, dtype=np.uint8), corrections=np.zeros(3 * N + 10 * M))
    rc = lib().srk_oracle_derivs_and_solve(*_prob_args(pr), _p(pr.points, C.c_double), _p(pr.cams, C.c_double), _p(pr.K, C.c_double),
                                           C.c_int(1 if pr.shared_K else 0), C.c_double(pr.f0), C.c_int(unity_ind),
                                           C.c_double(-1.0 if c is None else c), C.c_int(_FLOW[flow]),
                                           C.c_int(_SOLVE[solve]), C.c_int(0 if acc == "double" else 1),
                                           _p(out["gradE"], C.c_double), _p(out["E"], C.c_double), _p(out["G"], C.c_double),
                                           _p(out["F"], C.c_double), _p(out.get("S"), C.c_double), _p(out.get("rhs"), C.c_double),
                                           _p(out.get("skipped"), C.c_ubyte), _p(out.get("corrections"), C.c_double))
    if rc < 0:
        raise RuntimeError("oracle derivs_and_solve failed")
    out["solve_ok"] = rc == 0
    if "S" in out:
        out["S"] = out["S"].T.copy()  # col-major -> [row, col]
    return out


def naive_solve(pr, c, unity_ind=1):
    """EstimateCorrectionsNaive (BA.cpp:1700-1769) on the scene as given (already normalised): corrections[3N+10M]."""
    out = np.zeros(3 * pr.n_points + 10 * pr.n_cams)
    rc = lib().srk_oracle_naive_solve(*_prob_args(pr), _p(pr.points, C.c_double), _p(pr.cams, C.c_double), _p(pr.K, C.c_double),
                                      C.c_int(1 if pr.shared_K else 0), C.c_double(pr.f0), C.c_int(unity_ind), C.c_double(c), _p(out, C.c_double))
    if rc != 0:
        raise RuntimeError("oracle naive_solve failed")
    return out


def apply_corrections(points, cams, corrections):
    pts = np.ascontiguousarray(points, dtype=np.float64).copy()
    cs = np.ascontiguousarray(cams, dtype=np.float64).copy()
    corr = np.ascontiguousarray(corrections, dtype=np.float64)
    lib().srk_oracle_apply_corrections(C.c_int64(cs.shape[0]), C.c_int64(pts.shape[0]), _p(pts, C.c_double), _p(cs, C.c_double),
                                       _p(corr, C.c_double))
    return pts, cs


CIRCLE_GRID_DEFAULTS = dict(f0=600.0, xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0, zmin=0.0, zmax=1.0, cell_x=0.5, cell_y=0.5,
                            ang_start=-np.pi / 2 + np.pi / 6, ang_end=2 * np.pi / 3, ang_step=np.pi / 180 * 5,
                            noise_R_hi=0.005, noise_x3D_hi=0.005, rot_radius=-1.0, ascentZ=-1.0)


def circle_grid_scene(seed=1234, **kw):
    """The circle-grid demo scene (gflags defaults unless overridden)."""
    p = dict(CIRCLE_GRID_DEFAULTS)
    p.update(kw)
    params = np.array([p[k] for k in ("f0", "xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "cell_x", "cell_y", "ang_start", "ang_end",
                                      "ang_step", "noise_R_hi", "noise_x3D_hi", "rot_radius", "ascentZ")], dtype=np.float64)
    nc, npnt, no = C.c_int64(), C.c_int64(), C.c_int64()
    L = lib()
    L.srk_oracle_circle_grid_scene(_p(params, C.c_double), C.c_uint(seed), C.byref(nc), C.byref(npnt), C.byref(no),
                                   None, None, None, None, None, None, None, None)
    M, N, O = nc.value, npnt.value, no.value
    obs_cam = np.zeros(O, dtype=np.int32); obs_point = np.zeros(O, dtype=np.int32); obs_xy = np.zeros((O, 2))
    points = np.zeros((N, 3)); cams = np.zeros((M, 12)); K = np.zeros((M, 9)); gtp = np.zeros((N, 3)); gtc = np.zeros((M, 12))
    L.srk_oracle_circle_grid_scene(_p(params, C.c_double), C.c_uint(seed), C.byref(nc), C.byref(npnt), C.byref(no),
                                   _p(obs_cam, C.c_int32), _p(obs_point, C.c_int32), _p(obs_xy, C.c_double), _p(points, C.c_double),
                                   _p(cams, C.c_double), _p(K, C.c_double), _p(gtp, C.c_double), _p(gtc, C.c_double))
    return Problem(obs_cam, obs_point, obs_xy, points, cams, K, False, float(p["f0"]), gtp, gtc)


def circle_camera_shots(center, radius, ascentZ, angles):
    a = np.ascontiguousarray(angles, dtype=np.float64)
    c = np.ascontiguousarray(center, dtype=np.float64)
    out = np.zeros((len(a), 12))
    lib().srk_oracle_circle_camera_shots(_p(c, C.c_double), C.c_double(radius), C.c_double(ascentZ), _p(a, C.c_double), C.c_int(len(a)),
                                         _p(out, C.c_double))
    return out


def rotmat_from_axis_angle(w):
    w = np.ascontiguousarray(w, dtype=np.float64); R = np.zeros(9)
    ok = lib().srk_oracle_rotmat_from_axis_angle(_p(w, C.c_double), _p(R, C.c_double))
    return bool(ok), R.reshape(3, 3).T.copy()


def rotmat_from_unity_dir_and_angle(d, ang):
    d = np.ascontiguousarray(d, dtype=np.float64); R = np.zeros(9)
    ok = lib().srk_oracle_rotmat_from_unity_dir_and_angle(_p(d, C.c_double), C.c_double(ang), _p(R, C.c_double))
    return bool(ok), R.reshape(3, 3).T.copy()


def axis_angle_from_rotmat(R):
    r = np.ascontiguousarray(np.asarray(R, dtype=np.float64).T).reshape(9).copy(); w = np.zeros(3)
    ok = lib().srk_oracle_axis_angle_from_rotmat(_p(r, C.c_double), _p(w, C.c_double))
    return bool(ok), w


def track_pushback_probe(frames, xy, n_frames):
    f = np.ascontiguousarray(frames, dtype=np.int32); p = np.ascontiguousarray(xy, dtype=np.float64)
    has = np.zeros(n_frames, dtype=np.int32); out = np.zeros((n_frames, 2))
    lib().srk_oracle_track_pushback_probe(_p(f, C.c_int32), _p(p, C.c_double), C.c_int(len(f)), C.c_int(n_frames), _p(has, C.c_int32),
                                          _p(out, C.c_double))
    return has, out


# ---- MonoSLAM EKF dense chain (EKF.cpp:639-694, :977-1125) -------------------------------------------------------------
def ekf_update(P, x, Hcam, Hpt, pt_off, z, hpred, meas_var, fix_symmetry=True):
    """Reference-style dense stacked update on copies.  Returns (ok, P_new, x_new, seconds)."""
    Pn = np.asfortranarray(np.array(P, dtype=np.float64)); xn = np.array(x, dtype=np.float64)
    n = xn.shape[0]; m = len(pt_off); s = Hpt.shape[1]
    Hc = np.ascontiguousarray(Hcam, dtype=np.float64); Hp = np.ascontiguousarray(Hpt, dtype=np.float64)
    off = np.ascontiguousarray(pt_off, dtype=np.int64)
    zz = np.ascontiguousarray(z, dtype=np.float64); hh = np.ascontiguousarray(hpred, dtype=np.float64)
    sec = C.c_double()
    rc = lib().srk_oracle_ekf_update(C.c_int64(n), C.c_int64(m), Pn.ctypes.data_as(C.POINTER(C.c_double)), _p(xn, C.c_double), _p(Hc, C.c_double),
                                     _p(Hp, C.c_double), _p(off, C.c_int64), C.c_int(s), _p(zz, C.c_double), _p(hh, C.c_double), C.c_double(meas_var),
                                     C.c_int(1 if fix_symmetry else 0), C.byref(sec))
    return rc == 0, np.array(Pn), xn, sec.value


def ekf_update_exact(P, x, Hcam, Hpt, pt_off, z, hpred, meas_var, fix_symmetry=True):
    """The same update in long double through S = L L^T, Y = L^-1 H P, P - Y^T Y (srk_oracle_ekf_exact.hpp): the parity target at the
    sizes where the reference's explicit-inverse chain is itself ~cond(S) * eps off.  Returns (ok, P_new, x_new)."""
    Pn = np.asfortranarray(np.array(P, dtype=np.float64)); xn = np.array(x, dtype=np.float64)
    n = xn.shape[0]; m = len(pt_off); s = Hpt.shape[1]
    Hc = np.ascontiguousarray(Hcam, dtype=np.float64); Hp = np.ascontiguousarray(Hpt, dtype=np.float64)
    off = np.ascontiguousarray(pt_off, dtype=np.int64)
    zz = np.ascontiguousarray(z, dtype=np.float64); hh = np.ascontiguousarray(hpred, dtype=np.float64)
    rc = lib().srk_oracle_ekf_update_exact(C.c_int64(n), C.c_int64(m), Pn.ctypes.data_as(C.POINTER(C.c_double)), _p(xn, C.c_double), _p(Hc, C.c_double),
                                           _p(Hp, C.c_double), _p(off, C.c_int64), C.c_int(s), _p(zz, C.c_double), _p(hh, C.c_double), C.c_double(meas_var),
                                           C.c_int(1 if fix_symmetry else 0))
    return rc == 0, np.array(Pn), xn


def ekf_camera(fx_pix, fy_pix, cx, cy, dx_mm, dy_mm, k1, k2, enable_distortion=True):
    return np.array([fx_pix, fy_pix, cx, cy, dx_mm, dy_mm, k1, k2, 1.0 if enable_distortion else 0.0], dtype=np.float64)


def ekf_project(x, pt_off, s, cam9):
    """ProjectInternalSalientPoint (EKF.cpp:2947-2958) of every listed salient point at state x -> [m, 2] distorted pixels."""
    xs = np.ascontiguousarray(x, dtype=np.float64); off = np.ascontiguousarray(pt_off, dtype=np.int64)
    out = np.zeros((len(off), 2))
    lib().srk_oracle_ekf_ransac(C.c_int64(len(xs)), C.c_int64(len(off)), None, _p(xs, C.c_double), None, None, _p(off, C.c_int64), C.c_int(s), None, C.c_double(0.0),
                                _p(np.ascontiguousarray(cam9), C.c_double), C.c_double(0.0), None, None, _p(out, C.c_double))
    return out


def ekf_jacobians(x, pt_off, s, cam9):
    """Deriv_hd_by_cam_state_and_sal_pnt for every listed point (EKF.cpp:3067-3159): (Hcam [2m, 13], Hpt [2m, s], hd [2m])."""
    xs = np.ascontiguousarray(x, dtype=np.float64); off = np.ascontiguousarray(pt_off, dtype=np.int64)
    m = len(off)
    Hc = np.zeros((2 * m, 13)); Hp = np.zeros((2 * m, s)); hd = np.zeros(2 * m)
    lib().srk_oracle_ekf_jacobians(C.c_int64(len(xs)), C.c_int64(m), _p(xs, C.c_double), _p(off, C.c_int64), C.c_int(s), _p(np.ascontiguousarray(cam9), C.c_double),
                                   _p(Hc, C.c_double), _p(Hp, C.c_double), _p(hd, C.c_double))
    return Hc, Hp, hd


def ekf_ransac(P, x, Hcam, Hpt, pt_off, z, meas_var, cam9, max_divergence_pix):
    """OnePointRansac_GetConsensusMatches (EKF.cpp:1271-1391).  Returns (best, support[m], best_inliers[m])."""
    Pn = np.asfortranarray(np.array(P, dtype=np.float64)); xs = np.ascontiguousarray(x, dtype=np.float64)
    Hc = np.ascontiguousarray(Hcam, dtype=np.float64); Hp = np.ascontiguousarray(Hpt, dtype=np.float64)
    off = np.ascontiguousarray(pt_off, dtype=np.int64); zz = np.ascontiguousarray(z, dtype=np.float64)
    m = len(off); s = Hp.shape[1]
    support = np.zeros(m, dtype=np.int32); inl = np.zeros(m, dtype=np.uint8)
    cam = np.ascontiguousarray(cam9, dtype=np.float64)
    best = lib().srk_oracle_ekf_ransac(C.c_int64(len(xs)), C.c_int64(m), Pn.ctypes.data_as(C.POINTER(C.c_double)), _p(xs, C.c_double), _p(Hc, C.c_double),
                                       _p(Hp, C.c_double), _p(off, C.c_int64), C.c_int(s), _p(zz, C.c_double), C.c_double(meas_var), _p(cam, C.c_double),
                                       C.c_double(max_divergence_pix), _p(support, C.c_int32), inl.ctypes.data_as(C.POINTER(C.c_ubyte)), None)
    return int(best), support, inl


def ekf_sequential_update(P, x, pt_off, s, z, meas_var, cam9, per_component=False):
    """ProcessFrame_OneObservationPerUpdate (EKF.cpp:1153-1269) or, per_component, ...OneComponentOfOneObservationPerUpdate (:1525-1650)."""
    Pn = np.asfortranarray(np.array(P, dtype=np.float64)); xn = np.array(x, dtype=np.float64)
    off = np.ascontiguousarray(pt_off, dtype=np.int64); zz = np.ascontiguousarray(z, dtype=np.float64); c9 = np.ascontiguousarray(cam9, dtype=np.float64)
    lib().srk_oracle_ekf_sequential_update(C.c_int64(xn.shape[0]), C.c_int64(off.shape[0]), Pn.ctypes.data_as(C.POINTER(C.c_double)), _p(xn, C.c_double),
                                           _p(off, C.c_int64), C.c_int(s), _p(zz, C.c_double), C.c_double(meas_var), _p(c9, C.c_double), C.c_int(1 if per_component else 0))
    return np.array(Pn), xn


def ekf_ransac_update(P, x, pt_off, s, z, meas_var, cam9, max_divergence_pix, chi2_thr=float(np.float32(9.21034))):
    """ProcessFrame_OnePointRansacUpdateCore (EKF.cpp:1393-1513): (P_new, x_new, low mask, high mask)."""
    Pn = np.asfortranarray(np.array(P, dtype=np.float64)); xn = np.array(x, dtype=np.float64)
    off = np.ascontiguousarray(pt_off, dtype=np.int64); zz = np.ascontiguousarray(z, dtype=np.float64)
    c9 = np.ascontiguousarray(cam9, dtype=np.float64)
    m = off.shape[0]
    low = np.zeros(m, dtype=np.uint8); high = np.zeros(m, dtype=np.uint8); counts = np.zeros(2, dtype=np.int64)
    lib().srk_oracle_ekf_ransac_update(C.c_int64(xn.shape[0]), C.c_int64(m), Pn.ctypes.data_as(C.POINTER(C.c_double)), _p(xn, C.c_double), _p(off, C.c_int64),
                                       C.c_int(s), _p(zz, C.c_double), C.c_double(meas_var), _p(c9, C.c_double), C.c_double(max_divergence_pix), C.c_double(chi2_thr),
                                       _p(low, C.c_ubyte), _p(high, C.c_ubyte), _p(counts, C.c_int64))
    return np.array(Pn), xn, low.astype(bool), high.astype(bool)


def ekf_new_point(cam9, cam13, corner_pix, inv_dist, inv_dist_std, meas_std_pix):
    """GetNewSphericalSalientPointState + the small Jacobians of GetNewSphericalSalientPointCovar (EKF.cpp:2398-2527) and the XYZ conversion:
    dict(spher[6], Jy6[6,7], Q6[6,6], xyz[3], Jy3[3,7], Q3[3,3], xyz_ok)."""
    out = np.zeros(117)
    c9 = np.ascontiguousarray(cam9, dtype=np.float64); c13 = np.ascontiguousarray(cam13, dtype=np.float64)
    px = np.ascontiguousarray(corner_pix, dtype=np.float64)
    ok = lib().srk_oracle_ekf_new_point(_p(c9, C.c_double), _p(c13, C.c_double), _p(px, C.c_double), C.c_double(inv_dist), C.c_double(inv_dist_std),
                                        C.c_double(meas_std_pix), _p(out, C.c_double))
    return dict(spher=out[0:6].copy(), Jy6=out[6:48].reshape(6, 7).copy(), Q6=out[48:84].reshape(6, 6).copy(), xyz=out[84:87].copy(),
                Jy3=out[87:108].reshape(3, 7).copy(), Q3=out[108:117].reshape(3, 3).copy(), xyz_ok=bool(ok))


def ekf_add_points(P, x, x_new, Jy, Qnew, diag_only=False):
    """AllocateAndInitStateForNewSalientPoint (EKF.cpp:2322-2396), the points appended one after the other: (P_new, x_new_full)."""
    Pn = np.asfortranarray(np.array(P, dtype=np.float64)); xn = np.ascontiguousarray(x, dtype=np.float64)
    xa = np.ascontiguousarray(x_new, dtype=np.float64); k, s = xa.shape
    Ja = np.ascontiguousarray(Jy, dtype=np.float64); Qa = np.ascontiguousarray(Qnew, dtype=np.float64)
    n = xn.shape[0]; n2 = n + k * s
    Pout = np.zeros((n2, n2), order="F"); xout = np.zeros(n2)
    lib().srk_oracle_ekf_add_points(C.c_int64(n), Pn.ctypes.data_as(C.POINTER(C.c_double)), _p(xn, C.c_double), C.c_int64(k), C.c_int(s), _p(xa, C.c_double),
                                    _p(Ja, C.c_double), _p(Qa, C.c_double), C.c_int(1 if diag_only else 0), Pout.ctypes.data_as(C.POINTER(C.c_double)),
                                    _p(xout, C.c_double))
    return Pout, xout


def ekf_predict(P, F13, GQGt13, fix_symmetry=True):
    Pn = np.asfortranarray(np.array(P, dtype=np.float64))
    F = np.asfortranarray(np.array(F13, dtype=np.float64)); Q = np.asfortranarray(np.array(GQGt13, dtype=np.float64))
    lib().srk_oracle_ekf_predict(C.c_int64(Pn.shape[0]), Pn.ctypes.data_as(C.POINTER(C.c_double)), F.ctypes.data_as(C.POINTER(C.c_double)),
                                 Q.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(1 if fix_symmetry else 0))
    return np.array(Pn)


def triangulate(track_begin, obs_frame, obs_xy, proj, f0):
    """Triangulate3DPointByLeastSquares per track (obs-geom.cpp:679-727); proj: [n_frames, 12] 3x4 column-major."""
    tb = np.ascontiguousarray(track_begin, dtype=np.int64); fr = np.ascontiguousarray(obs_frame, dtype=np.int32)
    xy = np.ascontiguousarray(obs_xy, dtype=np.float64); pm = np.ascontiguousarray(proj, dtype=np.float64)
    out = np.zeros((len(tb) - 1, 3))
    L = lib()
    L.srk_oracle_triangulate.argtypes = [C.c_int64] + [C.c_void_p] * 4 + [C.c_double, C.c_void_p]
    L.srk_oracle_triangulate(len(tb) - 1, tb.ctypes.data, fr.ctypes.data, xy.ctypes.data, pm.ctypes.data, float(f0), out.ctypes.data)
    return out
