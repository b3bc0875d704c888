"""CPU tests of the input front end: DecomposeProjMat and ReadMatrixFromFile are host code in the
C-ABI library (no GPU needed); the oracle restatement of Triangulate3DPointByLeastSquares is pinned by known answers."""
import numpy as np
import pytest


def _rot(w):
    th = np.linalg.norm(w)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def test_decompose_proj_mat_round_trip_and_sign():
    from surikatoko_b200 import frontend
    rng = np.random.default_rng(5)
    for trial in range(20):
        K = np.array([[3217.4 + 10 * trial, 12.0, 289.9], [0.0, 2292.5, -1070.5 + trial], [0.0, 0.0, 1.0]])
        R = _rot(rng.normal(size=3))                   # direct camera orientation
        t = rng.normal(size=3) * 5.0
        scale = (1.0 + 0.3 * trial) * (-1.0 if trial % 3 == 0 else 1.0)      # det(Q) < 0 for every third matrix (obs-geom.cpp:617-624)
        P = scale * K @ R.T @ np.concatenate([np.eye(3), -t[:, None]], axis=1)
        ok, s, K2, (R2, t2) = frontend.DecomposeProjMat(P)
        assert ok
        assert abs(s - scale) < 1e-9 * abs(scale)
        assert np.allclose(K2, K, rtol=1e-10, atol=1e-8)
        assert np.allclose(R2, R, atol=1e-12) and abs(np.linalg.det(R2) - 1.0) < 1e-12
        assert np.allclose(t2, t, atol=1e-10)
        assert np.allclose(K2[2], [0, 0, 1], atol=1e-14) and K2[1, 0] == 0.0
        P_back = s * K2 @ R2.T @ np.concatenate([np.eye(3), -t2[:, None]], axis=1)       # the reference's post-condition (:665-672)
        assert np.linalg.norm(P - P_back) < 1e-8 * np.linalg.norm(P)


def test_decompose_proj_mat_rejects_singular():
    import surikatoko_b200 as sb
    from surikatoko_b200 import frontend
    P = np.zeros((3, 4)); P[0, 0] = 1.0
    try:
        ok, *_ = frontend.DecomposeProjMat(P)
        assert not ok
    except sb.SrkError:
        pass


def test_read_matrix_from_file(tmp_path):
    import surikatoko_b200 as sb
    from surikatoko_b200 import frontend
    a = np.arange(12, dtype=np.float64).reshape(3, 4) * 1.5 - 2.25
    f = tmp_path / "m.txt"
    f.write_text("\n".join("\t".join(repr(float(v)) for v in r) for r in a) + "\n")
    assert np.array_equal(frontend.ReadMatrixFromFile(f, "\t"), a)
    g = tmp_path / "viff.xy"
    g.write_text("1.5 -1 3e2\n-1 -1 7\n")                                     # the '-1 = no corner' convention of viff.xy
    assert np.array_equal(frontend.ReadMatrixFromFile(g, " "), np.array([[1.5, -1, 300.0], [-1, -1, 7.0]]))
    h = tmp_path / "ragged.txt"
    h.write_text("1 2 3\n4 5\n")
    with pytest.raises(sb.SrkError) as ei:
        frontend.ReadMatrixFromFile(h, " ")
    assert "inconsistent number of columns" in str(ei.value)
    k = tmp_path / "bad.txt"
    k.write_text("1 2x 3\n")
    with pytest.raises(sb.SrkError) as ei:
        frontend.ReadMatrixFromFile(k, " ")
    assert "Can't parse number (2x) on line 0" in str(ei.value)
    with pytest.raises(sb.SrkError) as ei:
        frontend.ReadMatrixFromFile(tmp_path / "missing.txt", " ")
    assert "Can't open file" in str(ei.value)
    e = tmp_path / "empty.txt"
    e.write_text("")
    assert frontend.ReadMatrixFromFile(e, " ").shape == (0, 0)


def _tracks(seed, n_frames=12, n_points=200, noise=0.0):
    from surikatoko_b200 import scenes
    prob = scenes.dino_shaped_scene(n_cams=n_frames, n_points=n_points, n_obs=int(3.5 * n_points), seed=seed, pix_sigma=noise, point_rel=0.0, rot_sigma=0.0)
    f0 = prob.f0
    P = []
    for i in range(n_frames):
        T = prob.gt_cams[i, :3]; R = prob.gt_cams[i, 3:].reshape(3, 3).T
        Kn = prob.K[i].reshape(3, 3).T
        P.append(Kn @ np.concatenate([R, T[:, None]], axis=1))
    counts = np.bincount(prob.obs_point, minlength=n_points)
    tb = np.concatenate([[0], np.cumsum(counts)])
    return prob, np.array(P), tb, f0


def test_oracle_triangulation_known_answers(oracle):
    prob, P, tb, f0 = _tracks(21)
    pm = np.ascontiguousarray(P.transpose(0, 2, 1)).reshape(-1, 12)
    X = oracle.triangulate(tb, prob.obs_cam, prob.obs_xy, pm, f0)
    assert np.max(np.abs(X - prob.gt_points)) < 1e-9                        # exact pixels: the generating points come back
    # noisy pixels: the minimiser of the same linear system, independently by numpy
    prob, P, tb, f0 = _tracks(22, noise=0.7)
    pm = np.ascontiguousarray(P.transpose(0, 2, 1)).reshape(-1, 12)
    X = oracle.triangulate(tb, prob.obs_cam, prob.obs_xy, pm, f0)
    for t in range(0, len(tb) - 1, 17):
        A, B = [], []
        for o in range(tb[t], tb[t + 1]):
            x, y = prob.obs_xy[o]; Pf = P[prob.obs_cam[o]]
            A += [x * Pf[2, :3] - f0 * Pf[0, :3], y * Pf[2, :3] - f0 * Pf[1, :3]]
            B += [-(x * Pf[2, 3] - f0 * Pf[0, 3]), -(y * Pf[2, 3] - f0 * Pf[1, 3])]
        ref = np.linalg.lstsq(np.array(A), np.array(B), rcond=None)[0]
        assert np.max(np.abs(X[t] - ref)) < 1e-9 * max(1.0, np.max(np.abs(ref)))
