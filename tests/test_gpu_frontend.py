"""GPU parity of the input front end, called through the C ABI: batched triangulation against the
oracle restatement of Triangulate3DPointByLeastSquares; the headless dino-style demo (file formats -> DecomposeProjMat ->
triangulation -> BA) end to end; full-size property check."""
import os
import sys

import numpy as np
import pytest

from test_cpu_frontend import _tracks

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("noise", [0.0, 0.7])
def test_triangulation_matches_oracle(oracle, noise):
    from surikatoko_b200 import frontend
    prob, P, tb, f0 = _tracks(31, n_frames=36, n_points=3000, noise=noise)
    pm = np.ascontiguousarray(P.transpose(0, 2, 1)).reshape(-1, 12)
    ref = oracle.triangulate(tb, prob.obs_cam, prob.obs_xy, pm, f0)
    got = frontend.Triangulate3DPointByLeastSquares(tb, prob.obs_cam, prob.obs_xy, P, f0)
    # Householder QR (oracle, like the reference's colPivHouseholderQr) vs row-wise Givens QR (GPU): same minimiser, rounding apart
    assert np.max(np.abs(got - ref)) < 1e-9 * np.max(np.abs(ref))
    if noise == 0.0:
        assert np.max(np.abs(got - prob.gt_points)) < 1e-9


def test_triangulation_rejects_single_corner_tracks():
    import surikatoko_b200 as sb
    from surikatoko_b200 import frontend
    P = np.tile(np.concatenate([np.eye(3), np.ones((3, 1))], axis=1), (2, 1, 1))
    with pytest.raises(sb.SrkError) as ei:
        frontend.Triangulate3DPointByLeastSquares([0, 1], [0], [[1.0, 2.0]], P, 600.0)
    assert "2 or more projections" in str(ei.value)


def test_headless_dino_demo(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import demo_dino
    demo_dino.write_synthetic_dinosaur(str(tmp_path), n_frames=36, n_points=800, seed=4, pix_sigma=0.3)
    out = demo_dino.run(str(tmp_path), f0=600.0, allowed_repr_err=1e-10, max_outer_iters=20, verbose=False)
    assert out["frames"] == 36 and out["tracks"] == 800
    # linear triangulation from exact cameras and 0.3 px noise starts near the noise level; BA must not make it worse and ends below it
    assert out["rms_px_final"] <= out["rms_px_initial"] and out["rms_px_final"] < 0.35


def test_triangulation_full_size_property():
    """1M tracks x 10 corners (the shape of configs[2]): exact pixels give the generating points back."""
    from surikatoko_b200 import frontend, scenes
    prob = scenes.ring_scene(1000, 1_000_000, 10, seed=1234, pix_sigma=0.0, rot_sigma=0.0, trans_rel=0.0, point_rel=0.0)
    f0 = prob.f0
    cams = prob.gt_cams
    R = cams[:, 3:].reshape(-1, 3, 3).transpose(0, 2, 1); T = cams[:, :3]
    Kn = prob.K.reshape(-1, 3, 3).transpose(0, 2, 1)
    P = np.einsum("mij,mjk->mik", Kn, np.concatenate([R, T[:, :, None]], axis=2))
    tb = np.arange(0, prob.n_obs + 1, 10)
    X = frontend.Triangulate3DPointByLeastSquares(tb, prob.obs_cam, prob.obs_xy, P, f0)
    assert np.max(np.abs(X - prob.gt_points)) < 1e-7


def test_bundle_file_is_a_shared_scene(tmp_path, oracle, engine):
    """one scene file, read by the engine's C reader and by the oracle's own numpy reader; both refine it alike."""
    import surikatoko_b200 as sb
    from surikatoko_b200 import bundle, scenes
    path = tmp_path / "ring.srkb"
    bundle.write_bundle(path, scenes.ring_scene(40, 3000, 8, seed=7))
    prob = bundle.read_bundle(path)
    pr = oracle.read_bundle(str(path))
    ref = oracle.ba_solve(pr, err_change=1e-10, max_outer_iters=3, flow="sparse", solve="chol", acc="ld")
    rep = engine.solve(prob, sb.BAOptions(err_change=1e-10, max_outer_iters=3))
    assert rep.seen_points == ref.seen_points and abs(rep.err_initial - ref.err_initial) <= 1e-12 * ref.err_initial
    assert np.array_equal(rep.attempts[:, 2], ref.attempts[:, 2])
    dev = np.abs(np.sqrt(rep.err_trace) - np.sqrt(ref.err_trace)) / np.sqrt(ref.err_trace)
    assert dev[0] < 1e-9 and np.all(dev < 1e-7), dev
