"""Bundle files (include/srk/bundle_c_api.h): host code, no GPU needed.  The C writer / reader of the library is
checked against an independent numpy reader on the oracle side (oracle_lib.read_bundle) and against damaged files."""
import os

import numpy as np
import pytest


def small_problem(shared_K=True):
    import surikatoko_b200 as sb
    from surikatoko_b200 import scenes
    pr = scenes.ring_scene(12, 150, 4, seed=3)
    K0 = pr.K.reshape(-1, 9)[:1]
    K = K0 if shared_K else np.repeat(K0, pr.n_cams, axis=0) * (1.0 + 0.01 * np.arange(pr.n_cams)[:, None])
    return sb.BAProblem(pr.obs_cam, pr.obs_point, pr.obs_xy, pr.points, pr.cams, K, shared_K, pr.f0)


@pytest.mark.parametrize("shared_K", [True, False])
def test_bundle_round_trip_and_oracle_reader(tmp_path, oracle, shared_K):
    from surikatoko_b200 import bundle
    pr = small_problem(shared_K)
    path = tmp_path / "scene.srkb"
    bundle.write_bundle(path, pr)
    h = bundle.read_bundle_header(path)
    assert h == dict(n_cams=pr.n_cams, n_points=pr.n_points, n_obs=pr.n_obs, shared_K=shared_K, f0=pr.f0)
    assert os.path.getsize(path) == 48 + 24 * pr.n_obs + 24 * pr.n_points + 96 * pr.n_cams + 72 * (1 if shared_K else pr.n_cams) + 8
    back = bundle.read_bundle(path)
    ref = oracle.read_bundle(str(path))
    for got in (back, ref):
        for name in ("obs_cam", "obs_point", "obs_xy", "points", "cams", "K"):
            assert np.array_equal(np.asarray(getattr(got, name)).reshape(-1), np.asarray(getattr(pr, name)).reshape(-1)), name
        assert got.shared_K == shared_K and got.f0 == pr.f0
    # the oracle evaluates the loaded scene: same reprojection error as on the original arrays
    op = oracle.Problem(pr.obs_cam.copy(), pr.obs_point.copy(), pr.obs_xy.copy(), pr.points.copy(), pr.cams.copy(), pr.K.copy(), shared_K, pr.f0)
    assert oracle.reproj_error(ref) == oracle.reproj_error(op)


def test_empty_scene_round_trip(tmp_path):
    import surikatoko_b200 as sb
    from surikatoko_b200 import bundle
    pr = sb.BAProblem(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 2)), np.zeros((0, 3)), np.zeros((2, 12)), np.eye(3).reshape(1, 9), True, 600.0)
    path = tmp_path / "empty.srkb"
    bundle.write_bundle(path, pr)
    back = bundle.read_bundle(path)
    assert back.n_obs == 0 and back.n_points == 0 and back.n_cams == 2


def test_damaged_files_fail_loudly(tmp_path):
    import surikatoko_b200 as sb
    from surikatoko_b200 import bundle
    pr = small_problem()
    path = tmp_path / "scene.srkb"
    bundle.write_bundle(path, pr)
    raw = bytearray(open(path, "rb").read())

    def expect(data, text):
        q = tmp_path / "bad.srkb"
        open(q, "wb").write(bytes(data))
        with pytest.raises(sb.SrkError) as ei:
            bundle.read_bundle(q)
        assert text in str(ei.value), str(ei.value)
    expect(b"NOTABNDL" + raw[8:], "bad magic")
    expect(raw[:20], "shorter than its header")
    expect(raw[:-100], "truncated")
    flipped = bytearray(raw); flipped[200] ^= 0x40
    expect(flipped, "checksum mismatch")
    expect(raw + b"x", "trailing bytes")
    neg = bytearray(raw); neg[8:16] = (-5).to_bytes(8, "little", signed=True)
    expect(neg, "impossible sizes")
    with pytest.raises(sb.SrkError) as ei:
        bundle.read_bundle(tmp_path / "missing.srkb")
    assert "Can't open file" in str(ei.value)
