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
