"""Seeded synthetic BA scenes of the shapes BASELINE.json names, as flat SoA problems.

Host-side numpy only (input generation, not the hot path).  Conventions follow the reference demos: inverse poses
(camera-from-world) stored as T[3], R column-major[9]; K per frame with rows 0,1 divided by f0
(demo-bundle-adj-circle-grid.cpp:149-163); pixels = f0 * (K * X_cam / z) with no visibility clipping (:183-222).
"""
import numpy as np

from .capi import BAProblem

K_PIX_CIRCLE_GRID = np.array([[880.0, 0.0, 400.0], [0.0, 660.0, 300.0], [0.0, 0.0, 1.0]])
K_PIX_DINO = np.array([[3217.4, 0.0, 289.9], [0.0, 2292.5, -1070.5], [0.0, 0.0, 1.0]])  # test_bundle_adjustment_kanatani.py:268-271


def _look_at(pos, target):
    """Inverse pose rows: x right, y down-ish, z forward (camera looks along +z)."""
    f = target - pos
    f = f / np.linalg.norm(f, axis=-1, keepdims=True)
    up = np.zeros_like(f); up[..., 2] = 1.0
    x = np.cross(up, f)
    x = x / np.linalg.norm(x, axis=-1, keepdims=True)
    y = np.cross(f, x)
    R = np.stack([x, y, f], axis=-2)                       # [.., 3(row), 3]
    T = -np.einsum("...ij,...j->...i", R, pos)
    return R, T


def _flat_cams(R, T):
    return np.concatenate([T, np.swapaxes(R, -1, -2).reshape(R.shape[:-2] + (9,))], axis=-1)


def _rodrigues(w):
    ang = np.linalg.norm(w, axis=-1)
    out = np.tile(np.eye(3), w.shape[:-1] + (1, 1))
    ok = ang > 1e-12
    n = np.zeros_like(w); n[ok] = w[ok] / ang[ok, None]
    S = np.zeros(w.shape[:-1] + (3, 3))
    S[..., 0, 1] = -n[..., 2]; S[..., 0, 2] = n[..., 1]; S[..., 1, 0] = n[..., 2]
    S[..., 1, 2] = -n[..., 0]; S[..., 2, 0] = -n[..., 1]; S[..., 2, 1] = n[..., 0]
    s = np.sin(ang)[..., None, None]; c = np.cos(ang)[..., None, None]
    return out + s * S + (1.0 - c) * (S @ S)


def project(R, T, K_norm, f0, X, obs_cam, obs_pt):
    """Pixels of X[obs_pt] in cameras obs_cam (K_norm = diag(1/f0,1/f0,1) K_pix), chunked to bound memory."""
    O = obs_cam.shape[0]
    out = np.empty((O, 2))
    step = 4_000_000
    for a in range(0, O, step):
        b = min(O, a + step)
        c = obs_cam[a:b]
        Xc = np.einsum("oij,oj->oi", R[c], X[obs_pt[a:b]]) + T[c]
        h = Xc / Xc[:, 2:3]
        out[a:b, 0] = f0 * (K_norm[0, 0] * h[:, 0] + K_norm[0, 2])
        out[a:b, 1] = f0 * (K_norm[1, 1] * h[:, 1] + K_norm[1, 2])
    return out


def _assemble(R, T, K_pix, f0, X_gt, obs_cam, obs_pt, rng, pix_sigma, rot_sigma, trans_sigma, point_sigma):
    M = R.shape[0]
    K_norm = np.diag([1.0 / f0, 1.0 / f0, 1.0]) @ K_pix
    xy = project(R, T, K_norm, f0, X_gt, obs_cam, obs_pt)
    if pix_sigma > 0:
        xy = xy + rng.normal(0.0, pix_sigma, xy.shape)
    # noisy initial state handed to BA
    Rn, Tn = R, T
    if rot_sigma > 0 or trans_sigma > 0:
        pos = -np.einsum("mji,mj->mi", R, T)                # camera centres
        dR = _rodrigues(rng.normal(0.0, 1.0, (M, 3)) * (rot_sigma / np.sqrt(3.0)))
        Rn = dR @ R
        pos_n = pos + rng.normal(0.0, trans_sigma, (M, 3)) if trans_sigma > 0 else pos
        Tn = -np.einsum("mij,mj->mi", Rn, pos_n)
    Xn = X_gt + rng.normal(0.0, point_sigma, X_gt.shape) if point_sigma > 0 else X_gt.copy()
    K = np.tile(K_norm.T.reshape(1, 9), (M, 1))             # column-major per frame
    prob = BAProblem(obs_cam.astype(np.int32), obs_pt.astype(np.int32), xy, Xn, _flat_cams(Rn, Tn), K, False, f0)
    prob.gt_points = X_gt
    prob.gt_cams = _flat_cams(R, T)
    return prob


def ring_scene(n_cams=1000, n_points=1_000_000, obs_per_point=10, seed=1234, f0=600.0, pix_sigma=0.5, rot_sigma=0.005,
               trans_rel=0.005, point_rel=0.005, point_offset=0, level_step=0.6):
    """Config 3 shape: cameras on a ring at 4 height levels looking inward, points uniform in a slab,
    each point observed by exactly `obs_per_point` cameras = the ring-nearest ones to its azimuth.  Points are ordered by
    azimuth (the order a sequential capture would create the tracks in), observations by (pnt_ind, frame_ind).
    `point_offset` rotates the azimuth window so that ranks of a weak-scaling run get different points of the same world."""
    rng = np.random.default_rng(seed + 7919 * point_offset)
    cam_rng = np.random.default_rng(seed)                  # cameras identical on every rank
    M, N, k = n_cams, n_points, obs_per_point
    Rc, Rp = 10.0, 6.0
    th = 2.0 * np.pi * np.arange(M) / M
    levels = level_step * np.arange(4)
    pos = np.stack([Rc * np.cos(th), Rc * np.sin(th), 2.0 + levels[np.arange(M) % 4]], axis=1)
    target = np.tile(np.array([0.0, 0.0, 0.0]), (M, 1))
    R, T = _look_at(pos, target)
    phi = np.sort(rng.uniform(0.0, 2.0 * np.pi, N))
    rho = Rp * np.sqrt(rng.uniform(0.05, 1.0, N))
    X = np.stack([rho * np.cos(phi), rho * np.sin(phi), rng.uniform(-1.0, 1.0, N)], axis=1)
    # k ring-nearest cameras: centred window around the nearest camera index, ascending frame order per point
    centre = np.floor(phi / (2.0 * np.pi) * M + 0.5).astype(np.int64)
    offs = np.arange(k) - k // 2
    cams = np.sort((centre[:, None] + offs[None, :]) % M, axis=1)
    obs_cam = cams.reshape(-1)
    obs_pt = np.repeat(np.arange(N, dtype=np.int64), k)
    # pose noise comes from the camera stream so that every rank perturbs the shared cameras identically
    prob = _assemble(R, T, K_PIX_CIRCLE_GRID, f0, X, obs_cam, obs_pt, rng, pix_sigma, 0.0, 0.0, point_rel * Rp)
    if rot_sigma > 0 or trans_rel > 0:
        dR = _rodrigues(cam_rng.normal(0.0, 1.0, (M, 3)) * (rot_sigma / np.sqrt(3.0)))
        Rn = dR @ R
        pos_n = pos + cam_rng.normal(0.0, trans_rel * Rc, (M, 3))
        Tn = -np.einsum("mij,mj->mi", Rn, pos_n)
        prob.cams = np.ascontiguousarray(_flat_cams(Rn, Tn))
    return prob


def dino_shaped_scene(n_cams=36, n_points=4983, n_obs=16432, seed=1234, f0=600.0, pix_sigma=0.5, point_rel=0.002, rot_sigma=0.002):
    """Config 1 shape (the Oxford dinosaur files are absent): 36 turntable views in 10 degree steps, ragged
    contiguous tracks of length >= 2 whose total is exactly n_obs."""
    rng = np.random.default_rng(seed)
    M, N = n_cams, n_points
    th = np.deg2rad(10.0) * np.arange(M)
    Rc = 20.0
    pos = np.stack([Rc * np.cos(th), Rc * np.sin(th), np.full(M, 4.0) + 0.3 * np.sin(3 * th)], axis=1)
    R, T = _look_at(pos, np.zeros((M, 3)))
    X = rng.normal(0.0, 1.0, (N, 3)) * np.array([1.5, 1.0, 1.2])
    lens = np.minimum(2 + rng.geometric(0.45, N) - 1, M)
    diff = n_obs - int(lens.sum())
    while diff != 0:
        idx = rng.integers(0, N, abs(diff))
        for i in idx:
            if diff > 0 and lens[i] < M: lens[i] += 1; diff -= 1
            elif diff < 0 and lens[i] > 2: lens[i] -= 1; diff += 1
            if diff == 0: break
    start = (rng.uniform(0.0, 1.0, N) * (M - lens + 1)).astype(np.int64)
    obs_pt = np.repeat(np.arange(N, dtype=np.int64), lens)
    first = np.cumsum(lens) - lens
    obs_cam = np.arange(obs_pt.shape[0], dtype=np.int64) - np.repeat(first, lens) + np.repeat(start, lens)
    # every camera must see something and cameras 0/1 anchor the gauge: force the first tracks to start at frame 0
    return _assemble(R, T, K_PIX_DINO, f0, X, obs_cam, obs_pt, rng, pix_sigma, rot_sigma, 0.0, point_rel * 3.0)


# ---- the circle-grid demo scene (demo-bundle-adj-circle-grid.cpp:64-291, scene-generator.cpp:9-55) ---------------------------------
# Scalar steps run as Python floats (IEEE double, libm through `math`, no contraction) in the order the demo writes them, the per-
# observation projection as numpy element-wise operations in the same order, and the random draws replay std::mt19937 (seed 1234) through
# libstdc++'s uniform_real_distribution<double> -- tests/test_cpu_host.py checks the result bit for bit against the oracle's C++ restatement.
import math as _m


def _mt19937_raw(seed, n):
    """n raw 32-bit outputs of std::mt19937 seeded with `seed` (numpy's legacy RandomState seeds MT19937 by the same init_genrand)."""
    return np.random.RandomState(seed)._bit_generator.random_raw(n).astype(np.float64)


class _StdUniform:
    """std::uniform_real_distribution<double>(a, b) over a std::mt19937: generate_canonical<double, 53> takes two draws, low word first."""

    def __init__(self, raw):
        self.raw, self.i = raw, 0

    def draw(self, a, b, count=None):
        n = 1 if count is None else count
        r = self.raw[self.i:self.i + 2 * n]; self.i += 2 * n
        canon = (r[0::2] + r[1::2] * 4294967296.0) / 18446744073709551616.0
        canon = np.where(canon >= 1.0, np.nextafter(1.0, 0.0), canon)
        v = canon * (b - a) + a
        return float(v[0]) if count is None else v


def _mm(A, B):
    return [[A[r][0] * B[0][c] + A[r][1] * B[1][c] + A[r][2] * B[2][c] for c in range(3)] for r in range(3)]


def _mv(A, x):
    return [A[r][0] * x[0] + A[r][1] * x[1] + A[r][2] * x[2] for r in range(3)]


def _is_close(a, b, rtol=1.0e-5, atol=1.0e-8):       # approx-alg.h:7-16
    return abs(a - b) <= (atol + rtol * abs(max(a, b)))


def _rot_unity_dir_angle(d, ang, check=True):        # obs-geom.cpp:520-551; None where the reference returns false
    if check:
        if not _is_close(1.0, _m.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])) or _is_close(0.0, ang):
            return None
    s, c = _m.sin(ang), _m.cos(ang)
    K = [[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]]
    K1 = [[(1.0 - c) * K[r][k] for k in range(3)] for r in range(3)]
    K2 = _mm(K1, K)
    return [[((1.0 if r == k else 0.0) + s * K[r][k]) + K2[r][k] for k in range(3)] for r in range(3)]


def _log_so3(R):                                     # obs-geom.cpp:563-596
    cos_ang = 0.5 * ((R[0][0] + R[1][1] + R[2][2]) - 1.0)
    cos_ang = min(max(cos_ang, -1.0), 1.0)
    sin_ang = _m.sqrt(1.0 - cos_ang * cos_ang)
    if _is_close(0.0, sin_ang, 0.0, float(np.float32(1e-3))):
        return None
    u = [R[2][1] - R[1][2], R[0][2] - R[2][0], R[1][0] - R[0][1]]
    k = 0.5 / sin_ang
    u = [v * k for v in u]
    k = 1.0 / _m.sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2])
    return [v * k for v in u], _m.acos(cos_ang)


def generate_circle_camera_shots(center, radius, ascent_z, angles):
    """GenerateCircleCameraShots (scene-generator.cpp:9-55): inverse poses (R as nested lists, T) looking at the circle's axis."""
    I = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    out = []
    for ang in angles:
        c2c = [radius * _m.cos(ang), radius * _m.sin(ang), ascent_z]
        shift = [center[k] + c2c[k] for k in range(3)]
        R, T = I, [0.0, 0.0, 0.0]

        def compose(Ra, Ta, Rb, Tb):
            t = _mv(Ra, Tb)
            return _mm(Ra, Rb), [t[k] + Ta[k] for k in range(3)]
        R, T = compose(I, [-shift[0], -shift[1], -shift[2]], R, T)
        tc = [-shift[0], -shift[1], 0.0]
        nrm = _m.sqrt(tc[0] * tc[0] + tc[1] * tc[1] + tc[2] * tc[2])
        tc = [v / nrm for v in tc]
        yaw = _m.acos(0.0 * tc[0] + 1.0 * tc[1] + 0.0 * tc[2])
        cross_z = 0.0 * tc[1] - 1.0 * tc[0]            # (oy x to_center) . oz
        yaw *= 1.0 if (0.0 * 0.0 + 0.0 * 0.0 + cross_z * 1.0) >= 0 else -1.0
        Ry = _rot_unity_dir_angle([0.0, 0.0, 1.0], -yaw) or I
        R, T = compose(Ry, [0.0, 0.0, 0.0], R, T)
        down = _m.atan2(c2c[2], _m.sqrt(c2c[0] * c2c[0] + c2c[1] * c2c[1] + 0.0))
        Rp = _rot_unity_dir_angle([1.0, 0.0, 0.0], down + _m.pi / 2) or I
        R, T = compose(Rp, [0.0, 0.0, 0.0], R, T)
        out.append((R, T))
    return out


def circle_grid_scene(seed=1234, f0=600.0, xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0, zmin=0.0, zmax=1.0, cell_x=0.5, cell_y=0.5,
                      ang_start=-_m.pi / 2 + _m.pi / 6, ang_end=2 * _m.pi / 3, ang_step=_m.pi / 180 * 5, noise_R_hi=0.005, noise_x3D_hi=0.005,
                      rot_radius=-1.0, ascentZ=-1.0):
    """The scene of demo-bundle-adj-circle-grid (defaults = its gflags, :46-62): a grid of points on z = zmin + cos((x - xmid)/xlen * pi) * zlen,
    cameras on a circle around (1, 0.5, 0) looking at its axis, EVERY point observed in EVERY frame (no clipping, :183-222), points
    perturbed by U(hi/2, hi) per coordinate, rotations by U(0,1)*hi on angle and axis; std::mt19937 seed 1234, draw order :109-128 then
    :224-257.  rot_radius / ascentZ < 0 take the demo's 15 * cell_x / 10 * cell_x."""
    if rot_radius < 0:
        rot_radius = 15 * cell_x
    if ascentZ < 0:
        ascentZ = 10 * cell_x
    gap = 1e-8
    xs, x = [], xmin
    while x < xmax + gap:
        xs.append(x); x += cell_x
    ys, y = [], ymin
    while y < ymax + gap:
        ys.append(y); y += cell_y
    xmid, xlen, zlen = (xmin + xmax) / 2, xmax - xmin, zmax - zmin
    zs = [zmin + _m.cos((x - xmid) / xlen * _m.pi) * zlen for x in xs]
    gt = np.empty((len(xs) * len(ys), 3))
    gt[:, 0] = np.repeat(np.array(xs), len(ys)); gt[:, 1] = np.tile(np.array(ys), len(xs)); gt[:, 2] = np.repeat(np.array(zs), len(ys))
    N = gt.shape[0]
    angles, ang = [], ang_start
    while not ((ang_start < ang_end and ang >= ang_end) or (ang_start > ang_end and ang <= ang_end)):
        angles.append(ang); ang += ang_step
    M = len(angles)
    dis = _StdUniform(_mt19937_raw(seed, 2 * (3 * N + 4 * M)))
    pts = gt.copy()
    if noise_x3D_hi > 0:
        pts = pts + dis.draw(noise_x3D_hi / 2, noise_x3D_hi, 3 * N).reshape(N, 3)
    Kn = [[(1 / f0) * 880, 0.0, (1 / f0) * (800 / 2.0)], [0.0, (1 / f0) * 660, (1 / f0) * (600 / 2.0)], [0.0, 0.0, 1.0]]
    shots = generate_circle_camera_shots([1.0, 0.5, 0.0], rot_radius, ascentZ, angles)
    X0, X1, X2 = gt[:, 0], gt[:, 1], gt[:, 2]
    xy = np.empty((N, M, 2))
    for i, (R, T) in enumerate(shots):
        pc = [((R[r][0] * X0 + R[r][1] * X1) + R[r][2] * X2) + T[r] for r in range(3)]
        img = [pc[0] / pc[2], pc[1] / pc[2], pc[2] / pc[2]]
        h = [(Kn[r][0] * img[0] + Kn[r][1] * img[1]) + Kn[r][2] * img[2] for r in range(3)]
        xy[:, i, 0] = (h[0] / h[2]) * f0; xy[:, i, 1] = (h[1] / h[2]) * f0
    noisy = []
    for R, T in shots:
        if noise_R_hi > 0:
            lg = _log_so3(R)
            if lg is not None:
                d, a = lg
                a += dis.draw(0.0, 1.0) * noise_R_hi
                d = [d[k] + dis.draw(0.0, 1.0) * noise_R_hi for k in range(3)]
                nrm = _m.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
                Rn = _rot_unity_dir_angle([v / nrm for v in d], a)
                if Rn is not None:
                    R = Rn
        noisy.append((R, T))

    def flat(sh):
        return np.array([list(T) + [R[r][c] for c in range(3) for r in range(3)] for R, T in sh])
    obs_cam = np.tile(np.arange(M, dtype=np.int32), N)
    obs_pt = np.repeat(np.arange(N, dtype=np.int32), M)
    K = np.tile(np.array([Kn[r][c] for c in range(3) for r in range(3)]).reshape(1, 9), (M, 1))
    prob = BAProblem(obs_cam, obs_pt, xy.reshape(-1, 2), pts, flat(noisy), K, False, f0)
    prob.gt_points = gt
    prob.gt_cams = flat(shots)
    return prob


def circle_grid_config(seed=1234):
    """BASELINE.json configs[1]: the demo's scene at 50 cameras x 100 x 100 points (every point in every frame: 500 000 observations).
    The demo's own flags with the grid refined to 100 points per axis and the angle step set for 50 shots; its default camera circle
    (15 x and 10 x the default cell: radius 7.5, ascent 5) is kept -- derived from the refined cell the cameras would sit inside the grid."""
    return circle_grid_scene(seed=seed, cell_x=2.0 / 99.5, cell_y=2.0 / 99.5, ang_step=_m.pi / 49.5, rot_radius=7.5, ascentZ=5.0)


CIRCLE_GRID_CONFIG = dict(cell_x=2.0 / 99.5, cell_y=2.0 / 99.5, ang_step=_m.pi / 49.5, rot_radius=7.5, ascentZ=5.0)


def shard_points(prob, rank, world):
    """Contiguous pnt_ind ranges balanced by observation count; cameras replicated."""
    O = prob.n_obs
    counts = np.bincount(prob.obs_point, minlength=prob.n_points)
    cum = np.concatenate([[0], np.cumsum(counts)])
    bounds = [int(np.searchsorted(cum, O * r / world, side="left")) for r in range(world)] + [prob.n_points]
    bounds[0] = 0
    p0, p1 = bounds[rank], bounds[rank + 1]
    o0, o1 = int(cum[p0]), int(cum[p1])
    return BAProblem(prob.obs_cam[o0:o1].copy(), prob.obs_point[o0:o1] - p0, prob.obs_xy[o0:o1].copy(), prob.points[p0:p1].copy(),
                     prob.cams.copy(), prob.K.copy(), prob.shared_K, prob.f0), (p0, p1)
