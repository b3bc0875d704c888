"""ctypes binding of the MonoSLAM EKF dense covariance chain (include/srk/ekf_c_api.h) and a synthetic scenario generator.

Mirrors the reference's DavisonMonoSlam members on this path: PredictEstimVars (covariance part, EKF.cpp:669-693) and
ProcessFrame_StackedObservationsPerUpdateCore (EKF.cpp:977-1125).  No CPU fallback.
"""
import ctypes as C

import numpy as np

from .capi import SrkError, load_library

CAM = 13
EKF_FAMILIES = ("pht", "innov", "chol", "trsm", "syrk", "state", "predict", "ransac", "chol_trsm")


def _lib():
    L = load_library()
    if not getattr(L, "_ekf_ready", False):
        L.srk_ekf_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.srk_ekf_destroy.argtypes = [C.c_void_p]; L.srk_ekf_destroy.restype = None
        L.srk_ekf_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.srk_ekf_set_state.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.srk_ekf_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.srk_ekf_predict_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.srk_ekf_update_resident.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double,
                                              C.POINTER(C.c_int32)]
        L.srk_ekf_ransac_consensus_resident.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_double,
                                                        C.c_void_p, C.c_double, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
        L.srk_ekf_measurement_jacobians_resident.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.srk_ekf_add_points_resident.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.srk_ekf_projected_covariances_resident.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.srk_ekf_sequential_update_resident.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_int32]
        L.srk_ekf_state_size.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.srk_ekf_update.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.c_double]
        L.srk_ekf_predict.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.srk_ekf_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.srk_ekf_get_timing.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.srk_ekf_launches.argtypes = [C.c_void_p]; L.srk_ekf_launches.restype = C.c_int64
        L._ekf_ready = True
    return L


def _chk(rc):
    if rc < 0:
        raise SrkError(rc, load_library().srk_last_error().decode())
    return rc


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class EkfCamera(C.Structure):
    """srk_ekf_camera: what ProjectCameraSalientPoint / DistortPixel read (EKF.cpp:3007-3033, :2960-3005)."""
    _fields_ = [("fx_pix", C.c_double), ("fy_pix", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("dx_mm", C.c_double), ("dy_mm", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double), ("enable_distortion", C.c_int32)]

    def as_array(self):
        return np.array([self.fx_pix, self.fy_pix, self.cx, self.cy, self.dx_mm, self.dy_mm, self.k1, self.k2, float(self.enable_distortion)])


CHI2_99_DOF2 = float(np.float32(9.21034))     # the reference's float literal (EKF.cpp:1495)


def one_point_ransac_update(engine, pt_off, s, z, camera, meas_var, max_divergence_pix, chi2_thr=CHI2_99_DOF2):
    """ProcessFrame_OnePointRansacUpdateCore (EKF.cpp:1393-1513), the update the shipped flagfile selects (--monoslam_update_impl=4), as host
    control flow over the resident entry points.  Stage 1: measurement Jacobians at the predicted state, 1-point RANSAC consensus
    (OnePointRansac_GetConsensusMatches :1271-1391), stacked update with the low-innovation inliers (:1445-1446).  Stage 2: every other
    matched point whose corner lies inside the chi^2 ellipse of its projection under the UPDATED state (:1466-1498) is rescued and enters a
    second stacked update (:1507-1508).  Returns (low mask, high mask)."""
    off = np.ascontiguousarray(pt_off, dtype=np.int64); z = np.ascontiguousarray(z, dtype=np.float64)
    m = off.shape[0]
    rows = lambda mask: np.repeat(mask, 2)
    Hc, Hp, hp = engine.measurement_jacobians(off, s, camera)
    best, _, inl = engine.ransac_consensus(Hc, Hp, off, z, meas_var, camera, max_divergence_pix)
    low = inl.astype(bool) if best >= 0 else np.zeros(m, dtype=bool)
    high = np.zeros(m, dtype=bool)
    if low.any():
        engine.update(Hc[rows(low)], Hp[rows(low)], off[low], z[rows(low)], hp[rows(low)], meas_var)
    rest = np.flatnonzero(~low)
    if rest.size == 0:
        return low, high
    Hc2, Hp2, hp2 = engine.measurement_jacobians(off[rest], s, camera)
    cov = engine.projected_covariances(Hc2, Hp2, off[rest])
    d = (z.reshape(-1, 2)[rest] - hp2.reshape(-1, 2))
    tr = cov[:, 0, 0] + cov[:, 1, 1]; df = cov[:, 0, 0] - cov[:, 1, 1]
    lo = tr / 2 - np.sqrt(df * df / 4 + cov[:, 0, 1] * cov[:, 0, 1])
    ellipse_ok = ~((lo < 0) & ~(np.abs(0.0 - lo) <= (1.0e-8 + 1.0e-5 * np.abs(np.maximum(0.0, lo)))))     # CheckEllipseIsExtractableFrom2DCovarMat
    det = cov[:, 0, 0] * cov[:, 1, 1] - cov[:, 0, 1] * cov[:, 1, 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        i00, i01, i10, i11 = cov[:, 1, 1] / det, -cov[:, 0, 1] / det, -cov[:, 1, 0] / det, cov[:, 0, 0] / det
        dist = d[:, 0] * (i00 * d[:, 0] + i01 * d[:, 1]) + d[:, 1] * (i10 * d[:, 0] + i11 * d[:, 1])
    take = ellipse_ok & (dist < chi2_thr)
    high[rest[take]] = True
    if take.any():
        t2 = rows(take)
        engine.update(Hc2[t2], Hp2[t2], off[rest[take]], z[rows(high)], hp2[t2], meas_var)
    return low, high


def new_salient_point(cam13, corner_pix, camera, inv_dist, inv_dist_std, meas_std_pix, s=3):
    """State and small Jacobians of a new salient point seen at `corner_pix` from camera state `cam13` -- the host side of
    AllocateAndInitStateForNewSalientPoint (EKF.cpp:2322-2396), 13 scalars of work per point like the kinematic model:
    GetNewSphericalSalientPointState (:2398-2455: undistortion, A.58 back-projection, azimuth / elevation, constant initial inverse
    distance), the Jacobians of GetNewSphericalSalientPointCovar (:2457-2527, A.67-A.79) and, for s = 3, ConvertXyzFromSphericalSalientPoint
    (:405-416) with DerivSalPnt_xyz_by_spher (:3790-3828).  Returns (x_new [s], Jy [s, 7], Qnew [s, s]) for EkfEngine.add_points:
    Jy = d(point) / d(camera position, quaternion), Qnew = the auto-covariance term that does not come from P (pixel noise through
    sal_pnt_by_h_rho, initial inverse-distance variance)."""
    cam13 = np.asarray(cam13, dtype=np.float64); q = cam13[3:7]
    hd = np.asarray(corner_pix, dtype=np.float64)
    cx, cy, dx, dy, k1, k2 = camera.cx, camera.cy, camera.dx_mm, camera.dy_mm, camera.k1, camera.k2
    hu = hd.copy(); hu_by_hd = np.eye(2)
    if camera.enable_distortion:
        rd = np.sqrt((dx * (hd[0] - cx)) ** 2 + (dy * (hd[1] - cy)) ** 2)                       # A.24
        stretch = 1 + k1 * rd ** 2 + k2 * rd ** 4
        hu = np.array([cx + (hd[0] - cx) * stretch, cy + (hd[1] - cy) * stretch])
        kk = k1 + 2 * k2 * rd ** 2
        side = 2 * kk * (hd[1] - cy) * (hd[0] - cx)
        hu_by_hd = np.array([[stretch + 2 * kk * (dx * (hd[0] - cx)) ** 2, side * dy ** 2],   # A.32, entries as the reference assigns them (:2674-2678)
                             [side * dx ** 2, stretch + 2 * kk * (dy * (hd[1] - cy)) ** 2]])
    hc = np.array([-(hu[0] - cx) / camera.fx_pix, -(hu[1] - cy) / camera.fy_pix, 1.0])          # A.58
    R = _quat_to_R(q)
    hw = R @ hc
    theta = np.arctan2(hw[0], hw[2]); phi = np.arctan2(-hw[1], np.sqrt(hw[0] ** 2 + hw[2] ** 2))
    dR = [np.array([[2 * q[0], -2 * q[3], 2 * q[2]], [2 * q[3], 2 * q[0], -2 * q[1]], [-2 * q[2], 2 * q[1], 2 * q[0]]]),     # A.46-A.49
          np.array([[2 * q[1], 2 * q[2], 2 * q[3]], [2 * q[2], -2 * q[1], -2 * q[0]], [2 * q[3], 2 * q[0], -2 * q[1]]]),
          np.array([[-2 * q[2], 2 * q[1], 2 * q[0]], [2 * q[1], 2 * q[2], 2 * q[3]], [-2 * q[0], 2 * q[3], -2 * q[2]]]),
          np.array([[-2 * q[3], -2 * q[0], 2 * q[1]], [2 * q[0], -2 * q[3], 2 * q[2]], [2 * q[1], 2 * q[2], 2 * q[3]]])]
    hw_by_q = np.stack([d @ hc for d in dR], axis=1)                                             # A.73
    dxz2 = hw[0] ** 2 + hw[2] ** 2; d2 = dxz2 + hw[1] ** 2; dxz = np.sqrt(dxz2); sf = hw[1] / (d2 * dxz)
    th_by_hw = np.array([hw[2] / dxz2, 0.0, -hw[0] / dxz2]); ph_by_hw = np.array([hw[0] * sf, -dxz / d2, hw[2] * sf])
    Jy6 = np.zeros((6, 7)); Jy6[0:3, 0:3] = np.eye(3)
    Jy6[3, 3:7] = th_by_hw @ hw_by_q; Jy6[4, 3:7] = ph_by_hw @ hw_by_q
    hw_by_hd = R @ (np.array([[-1 / camera.fx_pix, 0.0], [0.0, -1 / camera.fy_pix], [0.0, 0.0]]) @ hu_by_hd)
    A = np.zeros((6, 3)); A[3, 0:2] = th_by_hw @ hw_by_hd; A[4, 0:2] = ph_by_hw @ hw_by_hd; A[5, 2] = 1.0
    Q6 = A @ np.diag([meas_std_pix ** 2, meas_std_pix ** 2, inv_dist_std ** 2]) @ A.T
    spher = np.array([cam13[0], cam13[1], cam13[2], theta, phi, inv_dist])
    if s == 6:
        return spher, Jy6, Q6
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    dist, dist2 = 1 / inv_dist, 1 / inv_dist ** 2
    xyz = cam13[0:3] + dist * np.array([cp * st, -sp, cp * ct])
    D = np.array([[1, 0, 0, dist * cp * ct, -dist * sp * st, -dist2 * cp * st],
                  [0, 1, 0, 0.0, -dist * cp, dist2 * sp],
                  [0, 0, 1, -dist * cp * st, -dist * sp * ct, -dist2 * cp * ct]])
    return xyz, D @ Jy6, D @ Q6 @ D.T


def scenario01_camera(enable_distortion=True, k1=0.06, k2=0.01):
    """The camera of cpp_impl/demo-monoslam-scenario01.json: 320x240, f = 1.95 mm, 0.01 mm pixels, k1 k2 = (0.06, 0.01)."""
    return EkfCamera(1.95 / 0.01, 1.95 / 0.01, 160.0, 120.0, 0.01, 0.01, k1, k2, 1 if enable_distortion else 0)


class EkfEngine:
    def __init__(self, device=0):
        self._L = _lib()
        self._h = C.c_void_p()
        _chk(self._L.srk_ekf_create(C.byref(self._h), device))
        self.n = 0

    def close(self):
        if self._h:
            self._L.srk_ekf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        _chk(self._L.srk_ekf_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_state(self, P, x):
        P = np.asfortranarray(P, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
        self.n = x.shape[0]
        _chk(self._L.srk_ekf_set_state(self._h, self.n, C.c_void_p(P.ctypes.data), _p(x)))

    def get_state(self):
        P = np.zeros((self.n, self.n), order="F"); x = np.zeros(self.n)
        _chk(self._L.srk_ekf_get_state(self._h, C.c_void_p(P.ctypes.data), _p(x)))
        return P, x

    def predict(self, F13, GQGt13, cam_state_new=None):
        F = np.asfortranarray(F13, dtype=np.float64); Q = np.asfortranarray(GQGt13, dtype=np.float64)
        cs = None if cam_state_new is None else np.ascontiguousarray(cam_state_new, dtype=np.float64)
        _chk(self._L.srk_ekf_predict_resident(self._h, C.c_void_p(F.ctypes.data), C.c_void_p(Q.ctypes.data), _p(cs)))

    def update(self, Hcam, Hpt, pt_off, z, h_pred, meas_var):
        Hc = np.ascontiguousarray(Hcam, dtype=np.float64); Hp = np.ascontiguousarray(Hpt, dtype=np.float64)
        off = np.ascontiguousarray(pt_off, dtype=np.int64)
        zz = np.ascontiguousarray(z, dtype=np.float64); hh = np.ascontiguousarray(h_pred, dtype=np.float64)
        info = C.c_int32(0)
        _chk(self._L.srk_ekf_update_resident(self._h, off.shape[0], _p(Hc), _p(Hp), _p(off), Hp.shape[1], _p(zz), _p(hh), float(meas_var), C.byref(info)))
        return info.value

    def add_points(self, x_new, Jy, Qnew, diag_only=False):
        """Covariance growth for k new salient points on the resident state (srk_ekf_add_points_resident; AllocateAndInitStateForNewSalientPoint,
        EKF.cpp:2322-2396): x_new [k, s], Jy [k, s, 7], Qnew [k, s, s] as new_salient_point() forms them.  Returns the new state size."""
        xa = np.ascontiguousarray(x_new, dtype=np.float64); k, s = xa.shape
        Ja = np.ascontiguousarray(Jy, dtype=np.float64).reshape(k, s, 7); Qa = np.ascontiguousarray(Qnew, dtype=np.float64).reshape(k, s, s)
        _chk(self._L.srk_ekf_add_points_resident(self._h, k, s, _p(xa), _p(Ja), _p(Qa), 1 if diag_only else 0))
        n = C.c_int64(0)
        _chk(self._L.srk_ekf_state_size(self._h, C.byref(n)))
        self.n = n.value
        return self.n

    def sequential_update(self, pt_off, s, z, camera, meas_var, per_component=False):
        """ProcessFrame_OneObservationPerUpdate (EKF.cpp:1153-1269) / per_component: ...OneComponentOfOneObservationPerUpdate (:1525-1650)."""
        off = np.ascontiguousarray(pt_off, dtype=np.int64); zz = np.ascontiguousarray(z, dtype=np.float64)
        _chk(self._L.srk_ekf_sequential_update_resident(self._h, off.shape[0], _p(off), s, _p(zz), C.addressof(camera), float(meas_var), 1 if per_component else 0))

    def projected_covariances(self, Hcam, Hpt, pt_off):
        """J P_in J^T per listed point at the resident state (GetSalientPointProjected2DPosWithUncertainty, EKF.cpp:3901-4025): [m, 2, 2]."""
        Hc = np.ascontiguousarray(Hcam, dtype=np.float64); Hp = np.ascontiguousarray(Hpt, dtype=np.float64)
        off = np.ascontiguousarray(pt_off, dtype=np.int64)
        cov = np.zeros((off.shape[0], 2, 2))
        _chk(self._L.srk_ekf_projected_covariances_resident(self._h, off.shape[0], _p(Hc), _p(Hp), _p(off), Hp.shape[1], _p(cov)))
        return cov

    def measurement_jacobians(self, pt_off, s, camera):
        """Deriv_hd_by_cam_state_and_sal_pnt for every listed point at the resident state (EKF.cpp:3067-3159): (Hcam, Hpt, h_pred)."""
        off = np.ascontiguousarray(pt_off, dtype=np.int64)
        m = off.shape[0]
        Hc = np.zeros((2 * m, CAM)); Hp = np.zeros((2 * m, s)); hp = np.zeros(2 * m)
        _chk(self._L.srk_ekf_measurement_jacobians_resident(self._h, m, _p(off), s, C.addressof(camera), _p(Hc), _p(Hp), _p(hp)))
        return Hc, Hp, hp

    def ransac_consensus(self, Hcam, Hpt, pt_off, z, meas_var, camera, max_divergence_pix):
        """OnePointRansac_GetConsensusMatches (EKF.cpp:1271-1391) on the resident state: (best, support[m], best_inliers[m])."""
        Hc = np.ascontiguousarray(Hcam, dtype=np.float64); Hp = np.ascontiguousarray(Hpt, dtype=np.float64)
        off = np.ascontiguousarray(pt_off, dtype=np.int64); zz = np.ascontiguousarray(z, dtype=np.float64)
        m = off.shape[0]
        support = np.zeros(m, dtype=np.int32); inl = np.zeros(m, dtype=np.uint8); best = C.c_int32(-1)
        _chk(self._L.srk_ekf_ransac_consensus_resident(self._h, m, _p(Hc), _p(Hp), _p(off), Hp.shape[1], _p(zz), float(meas_var), C.addressof(camera),
                                                       float(max_divergence_pix), _p(support), C.byref(best), _p(inl)))
        return best.value, support, inl

    def update_host(self, P, x, Hcam, Hpt, pt_off, z, h_pred, meas_var):
        """One-shot srk_ekf_update with host buffers; P (Fortran order) and x are updated in place."""
        off = np.ascontiguousarray(pt_off, dtype=np.int64)
        _chk(self._L.srk_ekf_update(self._h, x.shape[0], off.shape[0], C.c_void_p(P.ctypes.data), _p(x), _p(Hcam), _p(Hpt), _p(off), Hpt.shape[1], _p(z),
                                    _p(h_pred), float(meas_var)))

    def set_timing(self, on):
        _chk(self._L.srk_ekf_set_timing(self._h, 1 if on else 0))

    def get_timing(self):
        out = {}
        for f in EKF_FAMILIES:
            t, c = C.c_double(), C.c_int64()
            _chk(self._L.srk_ekf_get_timing(self._h, f.encode(), C.byref(t), C.byref(c)))
            out[f] = dict(ms_total=t.value, count=c.value)
        return out

    def launches(self):
        return int(self._L.srk_ekf_launches(self._h))


def _quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def synthetic_ekf_frame(n_points=2000, s=3, seed=1234, cov_rank=24, pix_sigma=1.0):
    """One MonoSLAM frame of the configs[3] shape: a camera (pos, quaternion, velocities) in front of a wall of
    `n_points` XYZ salient points (s = 3) or 6-component inverse-depth points (s = 6, the extra components enter H with the same
    sparsity), all observed; H = d(pinhole projection)/d(state) with the velocity columns structurally zero (EKF.cpp:2755-2816).
    Returns dict(P, x, Hcam, Hpt, pt_off, z, h, meas_var, F, GQGt)."""
    rng = np.random.default_rng(seed)
    n = CAM + s * n_points
    x = np.zeros(n)
    x[0:3] = [0.1, -0.05, 0.0]
    q = np.array([0.995, 0.02, -0.03, 0.01]); x[3:7] = q / np.linalg.norm(q) * 1.0004   # slightly off unit: the normalisation branch runs
    x[7:13] = rng.normal(0, 0.01, 6)
    g = int(np.ceil(np.sqrt(n_points)))
    gx, gy = np.meshgrid(np.linspace(-1.5, 1.5, g), np.linspace(-1.5, -0.4 + 1.1, g))
    pts = np.stack([gx.ravel()[:n_points], gy.ravel()[:n_points], 7.0 + 0.3 * rng.normal(size=n_points)], axis=1)
    off = CAM + s * np.arange(n_points, dtype=np.int64)
    for c in range(3):
        x[off + c] = pts[:, c]
    if s == 6:
        x[off[:, None] + np.arange(3, 6)] = rng.normal(0, 0.1, (n_points, 3))
    fpx = np.array([195.0, 195.0]); c0 = np.array([160.0, 120.0])

    def project(xs):
        R = _quat_to_R(xs[3:7] / np.linalg.norm(xs[3:7]))
        X = np.stack([xs[off], xs[off + 1], xs[off + 2]], axis=1)
        if s == 6:
            X = X + 0.05 * np.stack([xs[off + 3], xs[off + 4], xs[off + 5]], axis=1)
        Xc = (X - xs[0:3]) @ R       # R^T (X - pos)
        return (fpx * Xc[:, :2] / Xc[:, 2:3] + c0)

    h = project(x)
    eps = 1e-6
    Hcam = np.zeros((2 * n_points, CAM)); Hpt = np.zeros((2 * n_points, s))
    for c in range(7):                # velocity / angular-velocity columns stay zero
        d = np.zeros(n); d[c] = eps
        Hcam[:, c] = ((project(x + d) - project(x - d)) / (2 * eps)).reshape(-1)
    for c in range(s):
        d = np.zeros(n); d[off + c] = eps
        Hpt[:, c] = ((project(x + d) - project(x - d)) / (2 * eps)).reshape(-1)
    z = (h + rng.normal(0, pix_sigma, h.shape)).reshape(-1)
    # covariance: point-wise variances + a low-rank camera/point coupling, symmetric positive definite
    U = rng.normal(0, 0.02, (n, cov_rank)); U[:CAM] *= 3.0
    P = np.asfortranarray(U @ U.T)
    P[np.diag_indices(n)] += np.concatenate([np.full(CAM, 1e-3), np.full(n - CAM, 2.5e-3)])
    dt = 1.0 / 30
    F = np.eye(CAM); F[0:3, 7:10] = dt * np.eye(3)
    F[3:7, 10:13] = 0.5 * dt * rng.normal(0, 1, (4, 3)); F[3:7, 3:7] += 0.01 * rng.normal(0, 1, (4, 4))
    Gm = np.zeros((CAM, 6)); Gm[0:3, 0:3] = dt * np.eye(3); Gm[7:10, 0:3] = np.eye(3); Gm[10:13, 3:6] = np.eye(3); Gm[3:7, 3:6] = 0.5 * dt * rng.normal(0, 1, (4, 3))
    Q = np.diag([0.15 ** 2] * 3 + [0.01 ** 2] * 3)
    return dict(P=P, x=x, Hcam=Hcam, Hpt=Hpt, pt_off=off, z=z, h=h.reshape(-1), meas_var=pix_sigma ** 2, F=np.asfortranarray(F),
                GQGt=np.asfortranarray(Gm @ Q @ Gm.T), n=n, m=n_points, s=s)


def project_salient_points(x, pt_off, s, cam):
    """numpy statement of ProjectInternalSalientPoint (EKF.cpp:2947-2958; A.22 / A.21 scaled by the inverse distance, pinhole :3021-3022,
    radial distortion :2960-3005) used to SYNTHESISE measurements and Jacobians for the scenario below (the device has its own)."""
    x = np.asarray(x, dtype=np.float64); off = np.asarray(pt_off, dtype=np.int64)
    q = x[3:7]
    R = np.array([[q[0] ** 2 + q[1] ** 2 - q[2] ** 2 - q[3] ** 2, 2 * (q[1] * q[2] - q[0] * q[3]), 2 * (q[1] * q[3] + q[0] * q[2])],
                  [2 * (q[1] * q[2] + q[0] * q[3]), q[0] ** 2 - q[1] ** 2 + q[2] ** 2 - q[3] ** 2, 2 * (q[2] * q[3] - q[0] * q[1])],
                  [2 * (q[1] * q[3] - q[0] * q[2]), 2 * (q[2] * q[3] + q[0] * q[1]), q[0] ** 2 - q[1] ** 2 - q[2] ** 2 + q[3] ** 2]])
    sp = x[off[:, None] + np.arange(s)]
    if s == 3:
        v = sp[:, :3] - x[0:3]
    else:
        th, ph, rho = sp[:, 3], sp[:, 4], sp[:, 5]
        mdir = np.stack([np.cos(ph) * np.sin(th), -np.sin(ph), np.cos(ph) * np.cos(th)], axis=1)
        v = rho[:, None] * (sp[:, :3] - x[0:3]) + mdir
    pc = v @ R                                  # R^T v per row
    hu = np.stack([cam.cx - cam.fx_pix * pc[:, 0] / pc[:, 2], cam.cy - cam.fy_pix * pc[:, 1] / pc[:, 2]], axis=1)
    if not cam.enable_distortion:
        return hu
    ru = np.sqrt((cam.dx_mm * (hu[:, 0] - cam.cx)) ** 2 + (cam.dy_mm * (hu[:, 1] - cam.cy)) ** 2)
    if cam.k2 != 0:
        rd = ru.copy()
        for _ in range(60):
            rd = rd - (rd + cam.k1 * rd ** 3 + cam.k2 * rd ** 5 - ru) / (1 + 3 * cam.k1 * rd ** 2 + 5 * cam.k2 * rd ** 4)
    elif cam.k1 == 0:
        rd = ru
    else:
        third = float(np.float32(1.0) / np.float32(3.0))
        e = (9 * cam.k1 ** 2 * ru + np.sqrt(3 * cam.k1 ** 3 * (4 + 27 * cam.k1 * ru ** 2))) ** third
        rd = (-2 * 3.0 ** third * cam.k1 + 2.0 ** third * e * e) / (6.0 ** (2.0 / 3) * cam.k1 * e)
    stretch = 1 + cam.k1 * rd ** 2 + cam.k2 * rd ** 4
    return np.stack([cam.cx + (hu[:, 0] - cam.cx) / stretch, cam.cy + (hu[:, 1] - cam.cy) / stretch], axis=1)


def synthetic_ransac_frame(n_points=200, s=3, seed=7, camera=None, outlier_frac=0.2, pix_sigma=0.1, cov_rank=16):
    """A frame for the 1-point RANSAC scoring (EKF.cpp:1271-1391) in the reference's own camera model: a camera in front of a wall of
    salient points (XYZ, s = 3, or first-camera / azimuth / elevation / inverse-distance, s = 6); every point is matched, a fraction of
    the corners are gross outliers (tens of pixels off).  H by central differences of project_salient_points.
    Returns dict(P, x, Hcam, Hpt, pt_off, z, meas_var, camera, outliers)."""
    rng = np.random.default_rng(seed)
    cam = camera or scenario01_camera()
    n = CAM + s * n_points
    x = np.zeros(n)
    x[0:3] = [0.05, -0.02, 0.1]
    q = np.array([0.998, 0.02, -0.03, 0.015]); x[3:7] = q / np.linalg.norm(q)
    x[7:13] = rng.normal(0, 0.01, 6)
    g = int(np.ceil(np.sqrt(n_points)))
    gx, gy = np.meshgrid(np.linspace(-1.2, 1.2, g), np.linspace(-0.9, 0.9, g))
    pts = np.stack([gx.ravel()[:n_points], gy.ravel()[:n_points], 4.0 + 0.5 * rng.normal(size=n_points)], axis=1)
    off = CAM + s * np.arange(n_points, dtype=np.int64)
    if s == 3:
        x[off[:, None] + np.arange(3)] = pts
    else:
        first = np.array([0.3, 0.1, -0.2]) + 0.05 * rng.normal(size=(n_points, 3))
        d = pts - first
        dist = np.linalg.norm(d, axis=1)
        x[off[:, None] + np.arange(3)] = first
        x[off + 3] = np.arctan2(d[:, 0], d[:, 2])            # azimuth:  m = (cos(phi) sin(theta), -sin(phi), cos(phi) cos(theta))
        x[off + 4] = -np.arcsin(d[:, 1] / dist)              # elevation
        x[off + 5] = 1.0 / dist
    h = project_salient_points(x, off, s, cam)
    Hcam = np.zeros((2 * n_points, CAM)); Hpt = np.zeros((2 * n_points, s))
    eps = 1e-6
    for c in range(7):
        d = np.zeros(n); d[c] = eps
        Hcam[:, c] = ((project_salient_points(x + d, off, s, cam) - project_salient_points(x - d, off, s, cam)) / (2 * eps)).reshape(-1)
    for c in range(s):
        d = np.zeros(n); d[off + c] = eps
        Hpt[:, c] = ((project_salient_points(x + d, off, s, cam) - project_salient_points(x - d, off, s, cam)) / (2 * eps)).reshape(-1)
    z = h + rng.normal(0, pix_sigma, h.shape)
    outliers = rng.random(n_points) < outlier_frac
    z[outliers] += rng.choice([-1.0, 1.0], size=(int(outliers.sum()), 2)) * rng.uniform(15.0, 40.0, size=(int(outliers.sum()), 2))
    U = rng.normal(0, 0.01, (n, cov_rank)); U[:CAM] *= 2.0
    P = np.asfortranarray(U @ U.T)
    P[np.diag_indices(n)] += np.concatenate([np.full(CAM, 2e-4), np.full(n - CAM, 1e-3)])
    if s == 6:      # inverse distances are known more tightly; scaled as a congruence D P D so that P stays positive definite
        d = np.ones(n); d[off + 5] = 0.3
        P = np.asfortranarray(P * d[:, None] * d[None, :])
    return dict(P=P, x=x, Hcam=Hcam, Hpt=Hpt, pt_off=off, z=z.reshape(-1), h=h.reshape(-1), meas_var=1.0, camera=cam, outliers=outliers, n=n, m=n_points, s=s)
