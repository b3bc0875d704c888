"""New salient point (AllocateAndInitStateForNewSalientPoint, EKF.cpp:2322-2396): the host-side state / small Jacobians of the product
(surikatoko_b200.ekf.new_salient_point) against the oracle's restatement, the analytic Jacobians against central differences of the
state function, and the oracle's sequential covariance growth against the closed batch form the device kernel evaluates."""
import numpy as np
import pytest

from conftest import relerr


def _cam13(seed=0):
    rng = np.random.default_rng(seed)
    q = np.array([0.98, 0.05, -0.12, 0.08]); q /= np.linalg.norm(q)
    return np.concatenate([[0.3, -0.2, 0.1], q, rng.normal(0, 0.01, 6)])


@pytest.mark.parametrize("distort", [True, False])
@pytest.mark.parametrize("s", [3, 6])
def test_new_point_state_and_jacobians_match_oracle(oracle, distort, s):
    from surikatoko_b200.ekf import new_salient_point, scenario01_camera
    cam = scenario01_camera(enable_distortion=distort)
    c13 = _cam13()
    for px in ([100.5, 80.25], [200.0, 150.0], [161.0, 119.0]):
        x_new, Jy, Q = new_salient_point(c13, px, cam, 0.4, 0.3, 1.0, s=s)
        ref = oracle.ekf_new_point(cam.as_array(), c13, px, 0.4, 0.3, 1.0)
        assert ref["xyz_ok"]
        r_x, r_J, r_Q = (ref["xyz"], ref["Jy3"], ref["Q3"]) if s == 3 else (ref["spher"], ref["Jy6"], ref["Q6"])
        assert relerr(x_new, r_x) < 1e-14 and relerr(Jy, r_J) < 1e-13 and relerr(Q, r_Q) < 1e-12


def test_new_point_jacobian_matches_central_differences(oracle):
    """d(point) / d(camera position, quaternion) of the oracle's restatement against central differences of its own state function."""
    from surikatoko_b200.ekf import scenario01_camera
    cam = scenario01_camera(enable_distortion=True).as_array()
    c13 = _cam13(1); px = [120.0, 90.0]
    ref = oracle.ekf_new_point(cam, c13, px, 0.5, 0.3, 1.0)
    eps = 1e-6
    J6 = np.zeros((6, 7)); J3 = np.zeros((3, 7))
    for k in range(7):
        d = np.zeros(13); d[k] = eps
        a = oracle.ekf_new_point(cam, c13 + d, px, 0.5, 0.3, 1.0); b = oracle.ekf_new_point(cam, c13 - d, px, 0.5, 0.3, 1.0)
        J6[:, k] = (a["spher"] - b["spher"]) / (2 * eps); J3[:, k] = (a["xyz"] - b["xyz"]) / (2 * eps)
    assert np.max(np.abs(J6 - ref["Jy6"])) < 1e-7 and np.max(np.abs(J3 - ref["Jy3"])) < 1e-7
    # the pixel / inverse-distance part: Q = A diag(var) A^T with A = d(point) / d(pixel, rho) by central differences
    A = np.zeros((6, 3))
    for k in range(2):
        d = np.zeros(2); d[k] = 1e-4
        A[:, k] = (oracle.ekf_new_point(cam, c13, np.array(px) + d, 0.5, 0.3, 1.0)["spher"] - oracle.ekf_new_point(cam, c13, np.array(px) - d, 0.5, 0.3, 1.0)["spher"]) / 2e-4
    A[:, 2] = (oracle.ekf_new_point(cam, c13, px, 0.5 + eps, 0.3, 1.0)["spher"] - oracle.ekf_new_point(cam, c13, px, 0.5 - eps, 0.3, 1.0)["spher"]) / (2 * eps)
    Q = A @ np.diag([1.0, 1.0, 0.09]) @ A.T
    # the reference's A.32 assigns the two off-diagonal entries of hu_by_hd with dx^2 / dy^2 swapped relative to the true derivative; with
    # square pixels (dx == dy, the scenario's camera) both agree
    assert np.max(np.abs(Q - ref["Q6"])) < 1e-6 * np.max(np.abs(ref["Q6"]))


@pytest.mark.parametrize("s", [3, 6])
def test_sequential_growth_equals_the_batch_form(oracle, s):
    """Adding k points one after the other (the reference: conservativeResize per point) == the closed form the device evaluates in one pass:
    P[new_i, old] = Jy_i P[0:7, old], P[new_i, new_l] = Jy_i P[0:7, 0:7] Jy_l^T + [i == l] Q_i."""
    from surikatoko_b200.ekf import synthetic_ekf_frame
    fr = synthetic_ekf_frame(12, s, seed=3)
    rng = np.random.default_rng(5)
    k = 4
    xn = rng.normal(0, 1, (k, s)); Jy = rng.normal(0, 1, (k, s, 7)); Q = np.stack([(lambda a: a @ a.T)(rng.normal(0, 0.1, (s, s))) for _ in range(k)])
    P1, x1 = oracle.ekf_add_points(fr["P"], fr["x"], xn, Jy, Q)
    n = fr["n"]; P = fr["P"]
    Pb = np.zeros((n + k * s, n + k * s)); Pb[:n, :n] = P
    Jall = Jy.reshape(k * s, 7)
    Pb[n:, :n] = Jall @ P[0:7, :]; Pb[:n, n:] = Pb[n:, :n].T
    Pb[n:, n:] = Jall @ P[0:7, 0:7] @ Jall.T
    for i in range(k):
        Pb[n + i * s:n + (i + 1) * s, n + i * s:n + (i + 1) * s] += Q[i]
    assert relerr(P1, Pb) < 1e-14 and np.array_equal(x1, np.concatenate([fr["x"], xn.reshape(-1)]))
    Pd, _ = oracle.ekf_add_points(fr["P"], fr["x"], xn, Jy, Q, diag_only=True)
    assert np.all(Pd[n:, :n] == 0) and np.all(Pd[:n, n:] == 0) and relerr(Pd[n:n + s, n:n + s], Q[0]) < 1e-16
