"""The second caller of the BA path: MultiViewIterativeFactorizer::IntegrateNewFrameCorners
(multi-view-factorization.cpp:255-397) -- ONE adjuster object re-used over a growing scene (new frame + newly reconstructed points at
every call, tracks without a SalientPointId skipped), shared K, kF0 = 1 (where quirk Q3 vanishes), ReprojError trigger 1e-3,
AllowedReprojErrRelativeChange(1e-3).  The same driver runs twice on the same scenario: with the GPU engine and with the CPU oracle
plugged in as the adjuster."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class OracleAdjuster:
    """The oracle behind the BundleAdjustmentKanatani surface mvf.py uses (checker side of the comparison)."""
    oracle = None

    def __init__(self):
        self.reports = []

    @staticmethod
    def _problem(f0, map_, cams, tracks, K):
        from surikatoko_b200.ba import flatten_scene
        prob, ids = flatten_scene(f0, map_, cams, tracks, K, None)
        ol = OracleAdjuster.oracle
        return prob, ids, ol.Problem(prob.obs_cam.copy(), prob.obs_point.copy(), prob.obs_xy.copy(), prob.points.copy(), prob.cams.copy(), prob.K.copy(), True, f0)

    @staticmethod
    def ReprojError(f0, map_, cams, tracks, shared_K=None, Ks=None, engine=None):
        _, _, op = OracleAdjuster._problem(f0, map_, cams, tracks, shared_K)
        return OracleAdjuster.oracle.reproj_error(op)[0]

    def ComputeInplace(self, f0, map_, cams, tracks, shared_K=None, Ks=None, term_crit=None):
        from surikatoko_b200.ba import scatter_scene
        prob, ids, op = self._problem(f0, map_, cams, tracks, shared_K)
        res = OracleAdjuster.oracle.ba_solve(op, err_change=term_crit.AllowedReprojErrRelativeChange(), flow="sparse", solve="chol", acc="ld")
        self.reports.append(res)
        prob.points[:] = res.points; prob.cams[:] = res.cams
        scatter_scene(prob, ids, map_, cams)
        return res.converged


def run_walk(adjuster_factory):
    from surikatoko_b200 import mvf
    K, X, X_init, frames = mvf.synthetic_walk()
    fac = mvf.MultiViewIterativeFactorizer(K, bundle_adjuster=adjuster_factory())
    track_of_point = {}
    log = []
    for est, cor in frames:
        corners, new_points = {None: []}, {}
        fresh = []
        for j, xy in cor.items():
            if j in track_of_point:
                corners[track_of_point[j]] = xy
                tr = fac.track_rep_.GetPointTrackById(track_of_point[j])
                if tr.SalientPointId is None and j % 7 != 3:          # some tracks are never reconstructed (quirk Q10)
                    new_points[track_of_point[j]] = X_init[j]
            else:
                fresh.append(j); corners[None].append(xy)
        n0 = fac.track_rep_.CornerTracksCount()
        for i, j in enumerate(fresh):
            track_of_point[j] = n0 + i
        assert fac.IntegrateNewFrameCorners(est, corners, new_points)
        log.append(dict(err=fac.last_reproj_err, ba=fac.last_ba_ran, ok=fac.last_ba_result, frames=fac.FramesCount(),
                        points=fac.map_.SalientPointsCount(), tracks=fac.track_rep_.CornerTracksCount(),
                        err_after=fac.ReprojError() if fac.last_reproj_err is not None else None))
    return fac, log


def test_incremental_caller_matches_oracle(oracle, engine):
    import surikatoko_b200 as sb
    OracleAdjuster.oracle = oracle
    gpu_fac, gpu_log = run_walk(lambda: sb.BundleAdjustmentKanatani(engine=engine))
    ref_fac, ref_log = run_walk(OracleAdjuster)
    assert any(e["ba"] for e in ref_log), "scenario never triggers the bundle adjustment"
    assert ref_log[-1]["tracks"] > ref_log[-1]["points"], "scenario has no unreconstructed track"
    launches = 0
    for g, r in zip(gpu_log, ref_log):
        assert (g["frames"], g["points"], g["tracks"], g["ba"], g["ok"]) == (r["frames"], r["points"], r["tracks"], r["ba"], r["ok"])
        if r["err"] is None:
            assert g["err"] is None
            continue
        # both sides refine their own copy of the model, so the deviation accumulates over the calls: first call 1e-9, later ones 1e-6
        tol = 1e-9 if launches == 0 else 1e-6
        assert abs(g["err"] - r["err"]) <= tol * r["err"], (g, r)
        assert abs(np.sqrt(g["err_after"]) - np.sqrt(r["err_after"])) <= tol * np.sqrt(r["err_after"]), (g, r)
        launches += 1 if r["ba"] else 0
    rep = gpu_fac.bundle_adjuster_.last_report
    assert rep.gpu_launches > 0 and rep.stop_reason == ref_fac.bundle_adjuster_.reports[-1].stop_reason
    # final model: same poses and points
    for a, b in zip(gpu_fac.cam_orient_cfw_, ref_fac.cam_orient_cfw_):
        assert np.max(np.abs(a.R - b.R)) < 1e-6 and np.max(np.abs(a.T - b.T)) < 1e-6
    pa = np.array([gpu_fac.map_.GetSalientPoint(i) for i in gpu_fac.map_.GetSalientPointsIds()])
    pb = np.array([ref_fac.map_.GetSalientPoint(i) for i in ref_fac.map_.GetSalientPointsIds()])
    assert np.max(np.abs(pa - pb)) < 1e-6
    # the intrinsics are never touched (quirk Q2)
    assert np.array_equal(gpu_fac.K_, ref_fac.K_)
