"""The second caller of the BA path: the bundle-adjustment step of
MultiViewIterativeFactorizer::IntegrateNewFrameCorners (multi-view-factorization.cpp:255-397).

What is mirrored here is the part of that function that touches the path: the growing model (poses pushed back one frame at a time,
multi-view-factorization.cpp:296-304; salient points created in one batch for the tracks that just became reconstructible, :313-369,
so that some tracks never own a SalientPointId -- quirk Q10), the static ReprojError with the shared K and kF0 = 1 (:372,
multi-view-factorization.h:21), the `err > 1e-3` trigger and the ComputeInplace call with AllowedReprojErrRelativeChange(1e-3) on ONE
long-lived BundleAdjustmentKanatani object (:379-391), i.e. one engine handle re-bound to a larger problem at every frame.
Corner matching, anchor-frame localisation and depth estimation (:259-293, :331-336) feed the path but are not on it: the caller
supplies their results (the reference itself has the same switch: fake_localization_ / fake_mapping_, multi-view-factorization.h:43-44).
"""
import numpy as np

from .ba import (BundleAdjustmentKanatani, BundleAdjustmentKanataniTermCriteria, CornerTrackRepository, FragmentMap, SE3Transform)


class MultiViewIterativeFactorizer:
    kF0 = 1.0               # multi-view-factorization.h:21
    kMaxReprojErr = 1e-3    # multi-view-factorization.cpp:378
    kBAErrChange = 1e-3     # multi-view-factorization.cpp:387

    def __init__(self, K, device=0, engine=None, bundle_adjuster=None):
        self.map_ = FragmentMap()
        self.cam_orient_cfw_ = []
        self.track_rep_ = CornerTrackRepository()
        self.K_ = np.asarray(K, dtype=np.float64).reshape(3, 3).copy()
        # one long-lived adjuster (multi-view-factorization.h:46); tests plug the CPU oracle in here as the checker
        self.bundle_adjuster_ = bundle_adjuster if bundle_adjuster is not None else BundleAdjustmentKanatani(device=device, engine=engine)
        self.last_reproj_err = None
        self.last_ba_ran = False
        self.last_ba_result = None

    def FramesCount(self):
        return len(self.cam_orient_cfw_)

    def ReprojError(self):
        """multi-view-factorization.cpp:372: static BundleAdjustmentKanatani::ReprojError with the shared K."""
        return type(self.bundle_adjuster_).ReprojError(self.kF0, self.map_, self.cam_orient_cfw_, self.track_rep_, self.K_, None,
                                                       engine=getattr(self.bundle_adjuster_, "_engine", None))

    def IntegrateNewFrameCorners(self, cam_new_from_world, corners, new_points):
        """One frame of multi-view-factorization.cpp:255-397 with the off-path estimates supplied by the caller.

        cam_new_from_world: SE3Transform estimate of the new frame (what :296-304 pushes back);
        corners: {track_id or None: (x, y)} corners matched in the new frame; key None entries (a list under None) start new tracks;
        new_points: {track_id: xyz} world estimates for tracks that become reconstructed in this frame (:340-366).
        Returns True like the reference (False only when nothing could be integrated)."""
        new_frame_ind = self.FramesCount()
        for track_id, xy in corners.items():
            if track_id is None:
                continue
            self.track_rep_.GetPointTrackById(track_id).AddCorner(new_frame_ind, np.asarray(xy, dtype=np.float64))
        for xy in corners.get(None, []):
            self.track_rep_.AddCornerTrackObj().AddCorner(new_frame_ind, np.asarray(xy, dtype=np.float64))
        self.cam_orient_cfw_.append(cam_new_from_world)
        for track_id, xyz in new_points.items():
            track = self.track_rep_.GetPointTrackById(track_id)
            if track.SalientPointId is not None:      # already reconstructed (:323)
                continue
            if track.CornersCount() <= 1:             # one projection only: cannot be reconstructed (:329)
                continue
            _, track.SalientPointId = self.map_.AddSalientPointTempl(np.asarray(xyz, dtype=np.float64))
        self.last_ba_ran = False
        self.last_ba_result = None
        if self.track_rep_.ReconstructedCornerTracksCount() == 0 or self.FramesCount() < 2:
            self.last_reproj_err = None
            return True
        err = self.ReprojError()
        self.last_reproj_err = err
        if err > self.kMaxReprojErr:                  # :379
            term_crit = BundleAdjustmentKanataniTermCriteria()
            term_crit.AllowedReprojErrRelativeChange(self.kBAErrChange)
            self.last_ba_result = self.bundle_adjuster_.ComputeInplace(self.kF0, self.map_, self.cam_orient_cfw_, self.track_rep_, self.K_, None, term_crit)
            self.last_ba_ran = True
        return True


def synthetic_walk(n_frames=8, n_points=120, seed=5, pix_sigma=2e-4, pose_sigma=0.01, point_sigma=0.02, window=5):
    """Deterministic scenario for the caller above: a camera walks on an arc around a point cloud; every frame sees the points of a
    sliding window plus a common core, in normalised image coordinates (K = diag(1, 1, 1) with a principal-point offset, kF0 = 1).
    Yields per frame (pose_estimate, corners_by_point{point: xy}); points carry ground truth + a noisy initial estimate."""
    rng = np.random.default_rng(seed)
    K = np.array([[1.2, 0.0, 0.05], [0.0, 1.1, -0.03], [0.0, 0.0, 1.0]])
    X = np.column_stack([rng.uniform(-1, 1, n_points), rng.uniform(-0.6, 0.6, n_points), rng.uniform(-0.5, 0.5, n_points)])
    X_init = X + rng.normal(0.0, point_sigma, X.shape)
    frames = []
    for f in range(n_frames):
        ang = -0.5 + 1.0 * f / max(1, n_frames - 1)
        eye = np.array([4.0 * np.sin(ang), 0.3 * np.cos(3 * ang), -4.0 * np.cos(ang)])
        z = -eye / np.linalg.norm(eye); up = np.array([0.0, 1.0, 0.0])
        x = np.cross(up, z); x /= np.linalg.norm(x); y = np.cross(z, x)
        R = np.stack([x, y, z])                      # camera-from-world rotation
        T = -R @ eye
        vis = [j for j in range(n_points) if j % window == f % window or j < n_points // 3 or (j + f) % 3 == 0]
        cor = {}
        for j in vis:
            pc = K @ (R @ X[j] + T)
            cor[j] = pc[:2] / pc[2] + rng.normal(0.0, pix_sigma, 2)
        # noisy pose estimate (what localisation would deliver): small rotation about a random axis + translation noise
        w = rng.normal(0.0, pose_sigma, 3); th = np.linalg.norm(w); k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        dR = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        est = SE3Transform(dR @ R, T + rng.normal(0.0, pose_sigma, 3)) if f >= 2 else SE3Transform(R.copy(), T.copy())
        frames.append((est, cor))
    return K, X, X_init, frames
