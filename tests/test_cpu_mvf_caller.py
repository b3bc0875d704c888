"""Host logic of the second caller of the BA path (surikatoko_b200/mvf.py, multi-view-factorization.cpp:255-397) without a GPU: the
driver runs with the CPU oracle plugged in as the adjuster (checker only) and must grow the model the way the reference does."""
import numpy as np

from test_gpu_mvf_caller import OracleAdjuster, run_walk


def test_growing_model_and_trigger_rule(oracle):
    OracleAdjuster.oracle = oracle
    fac, log = run_walk(OracleAdjuster)
    assert [e["frames"] for e in log] == list(range(1, len(log) + 1))
    assert log[0]["err"] is None and log[0]["points"] == 0                  # a single frame reconstructs nothing (:329)
    assert all(b["points"] >= a["points"] and b["tracks"] >= a["tracks"] for a, b in zip(log, log[1:]))
    assert log[-1]["tracks"] > log[-1]["points"] > 0                          # tracks without a SalientPointId stay out of the problem (Q10)
    for e in log[1:]:
        assert e["ba"] == (e["err"] > 1e-3)                                   # :379
        if e["ba"]:
            assert e["err_after"] < e["err"]
    # the adjuster saw problems of growing size, all with the shared K and f0 = 1, and stopped on the 1e-3 change rule
    reps = fac.bundle_adjuster_.reports
    assert len(reps) == sum(e["ba"] for e in log)
    assert all(r.stop_reason in ("small relative err change", "err converged to limit value", "abs err threshold") for r in reps)
    seen = [r.seen_points for r in reps]
    assert seen == sorted(seen) and seen[-1] > seen[0]


def test_flattening_skips_unreconstructed_tracks(oracle):
    from surikatoko_b200.ba import flatten_scene
    OracleAdjuster.oracle = oracle
    fac, _ = run_walk(OracleAdjuster)
    prob, ids = flatten_scene(1.0, fac.map_, fac.cam_orient_cfw_, fac.track_rep_, fac.K_, None)
    owned = [t for t in fac.track_rep_.CornerTracks if t.SalientPointId is not None]
    assert prob.n_points == len(owned) == len(ids) and prob.shared_K and prob.K.shape == (1, 9)
    assert prob.n_obs == sum(t.CornersCount() for t in owned)
    assert np.all(np.diff(prob.obs_point) >= 0) and prob.obs_point.max() == prob.n_points - 1
