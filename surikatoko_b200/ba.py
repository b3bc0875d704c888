"""Host-side mirror of the reference's operator interface for the BA path, above the C ABI.

Names, argument meaning and error behaviour follow suriko-engine:
  FragmentMap / SalientPointFragment        include/suriko/obs-geom.h:199-243, src/obs-geom.cpp:152-256
  CornerTrack / CornerTrackRepository       include/suriko/obs-geom.h:245-304, src/obs-geom.cpp:258-416
  SE3Transform                              include/suriko/obs-geom.h:177-190
  BundleAdjustmentKanatani(+TermCriteria)   include/suriko/bundle-adj-kanatani.h:68-261
The C++ form of the same adapter (for the reference's own build) is include/suriko_compat/bundle-adj-kanatani.h.
Every numerical step happens in the CUDA library; this module only flattens the containers into the SoA problem
(quirks Q10/Q11 of SURVEY.md section 8) and scatters the refined values back.
"""
import numpy as np

from .capi import BAOptions, BAProblem, Engine


class SE3Transform:
    """T (3) and R (3x3); `as_flat()` is the 12-double record T[3], R column-major[9] of the reference struct."""

    def __init__(self, R=None, T=None):
        self.R = np.eye(3) if R is None else np.array(R, dtype=np.float64).reshape(3, 3)
        self.T = np.zeros(3) if T is None else np.array(T, dtype=np.float64).reshape(3)

    @staticmethod
    def NoTransform():
        return SE3Transform()

    def as_flat(self):
        return np.concatenate([self.T, self.R.T.reshape(9)])

    @staticmethod
    def from_flat(f):
        f = np.asarray(f, dtype=np.float64)
        return SE3Transform(f[3:12].reshape(3, 3).T.copy(), f[0:3].copy())


class SalientPointFragment:
    __slots__ = ("synthetic_virtual_point_id", "coord", "user_obj")

    def __init__(self):
        self.synthetic_virtual_point_id = None
        self.coord = None
        self.user_obj = None


class FragmentMap:
    def __init__(self, fragment_id_offset=1000000):
        self._salient_points = []
        self._fragment_id_offset = fragment_id_offset
        self._next_salient_point_id = fragment_id_offset + 1

    def AddSalientPointTempl(self, coord):
        """Returns (fragment, salient_point_id)  (obs-geom.cpp:158-174)."""
        new_id = self._next_salient_point_id
        self._next_salient_point_id += 1
        frag = SalientPointFragment()
        frag.coord = None if coord is None else np.array(coord, dtype=np.float64).reshape(3)
        self._salient_points.append(frag)
        return frag, new_id

    def SalientPointIdToInd(self, salient_point_id):
        return salient_point_id - self._fragment_id_offset - 1

    def SalientPointIndToId(self, ind):
        return ind + self._fragment_id_offset + 1

    def GetSalientPointNew(self, salient_point_id):
        ind = self.SalientPointIdToInd(salient_point_id)
        if not (0 <= ind < len(self._salient_points)):
            raise IndexError("CHECK(ind < salient_points_.size())")
        return self._salient_points[ind]

    def GetSalientPoint(self, salient_point_id):
        return self.GetSalientPointNew(salient_point_id).coord

    def SetSalientPoint(self, point_track_id, coord):
        self._salient_points[point_track_id].coord = np.array(coord, dtype=np.float64).reshape(3)

    def SalientPointsCount(self):
        return len(self._salient_points)

    def SalientPoints(self):
        return self._salient_points

    def GetSalientPointsIds(self):
        return [self.SalientPointIndToId(i) for i in range(len(self._salient_points))]


class CornerData:
    __slots__ = ("pixel_coord", "image_coord")

    def __init__(self, pixel_coord=None):
        self.pixel_coord = np.zeros(2) if pixel_coord is None else np.array(pixel_coord, dtype=np.float64).reshape(2)
        self.image_coord = np.zeros(3)


class CornerTrack:
    def __init__(self):
        self.TrackId = 0
        self._StartFrameInd = -1
        self._CoordPerFramePixels = []
        self.SalientPointId = None
        self.SyntheticVirtualPointId = None

    def HasCorners(self):
        return self._StartFrameInd != -1

    def CornersCount(self):
        return len(self._CoordPerFramePixels)

    def _check_start(self, frame_ind):
        if self._StartFrameInd == -1:
            self._StartFrameInd = frame_ind
        elif not (self._StartFrameInd <= frame_ind):
            raise ValueError("Can insert points later than the initial (start) frame StartFrameInd=%d frame_ind=%d" %
                             (self._StartFrameInd, frame_ind))

    def AddCorner(self, frame_ind, value=None):
        """With a value: push_back — the k-th added corner is reported at frame Start+k, gaps collapse
        (obs-geom.cpp:277-292, quirk Q11).  Without: resize to frame_ind, gaps stay empty (obs-geom.cpp:294-314)."""
        self._check_start(frame_ind)
        if value is not None:
            self._CoordPerFramePixels.append(CornerData(value))
            return self._CoordPerFramePixels[-1]
        local_ind = frame_ind - self._StartFrameInd
        cur = len(self._CoordPerFramePixels)
        if local_ind + 1 > cur:
            self._CoordPerFramePixels.extend([None] * (local_ind + 1 - cur))
        else:
            del self._CoordPerFramePixels[local_ind + 1:]
        self._CoordPerFramePixels[-1] = CornerData()
        return self._CoordPerFramePixels[-1]

    def GetCornerData(self, frame_ind):
        if self._StartFrameInd == -1:
            raise RuntimeError("CHECK(StartFrameInd != -1)")
        local_ind = frame_ind - self._StartFrameInd
        if local_ind < 0 or local_ind >= len(self._CoordPerFramePixels):
            return None
        return self._CoordPerFramePixels[local_ind]

    def GetCorner(self, frame_ind):
        cd = self.GetCornerData(frame_ind)
        return None if cd is None else cd.pixel_coord

    def EachCorner(self, on_item):
        for i, cd in enumerate(self._CoordPerFramePixels):
            on_item(self._StartFrameInd + i, cd)


class CornerTrackRepository:
    def __init__(self):
        self.CornerTracks = []

    def AddCornerTrackObj(self):
        t = CornerTrack()
        t.TrackId = len(self.CornerTracks)
        self.CornerTracks.append(t)
        return t

    def CornerTracksCount(self):
        return len(self.CornerTracks)

    def ReconstructedCornerTracksCount(self):
        return sum(1 for t in self.CornerTracks if t.SalientPointId is not None)

    def FramesCount(self):
        return self.CornerTracks[0].CornersCount() if self.CornerTracks else 0

    def GetPointTrackById(self, point_track_id):
        return self.CornerTracks[point_track_id]


class BundleAdjustmentKanataniTermCriteria:
    """bundle-adj-kanatani.h:68-92."""

    def __init__(self):
        self._allowed_reproj_err_rel_change = None
        self._max_hessian_factor = None

    def AllowedReprojErrRelativeChange(self, allowed_reproj_err_rel_change=None):
        if allowed_reproj_err_rel_change is not None:
            self._allowed_reproj_err_rel_change = float(allowed_reproj_err_rel_change)
        return self._allowed_reproj_err_rel_change

    def MaxHessianFactor(self, max_hessian_factor=None):
        if max_hessian_factor is not None:
            self._max_hessian_factor = float(max_hessian_factor)
        return self._max_hessian_factor


def flatten_scene(f0, map_, inverse_orient_cams, track_rep, shared_K=None, Ks=None):
    """Containers -> flat SoA problem, through GetCorner/EachCorner semantics.

    pnt_ind is the running index over tracks that own a SalientPointId (BA.cpp:1161-1171, quirk Q10); frames beyond the
    pose vector are never probed by the reference loops and are dropped.  Returns (BAProblem, salient_point_ids[pnt_ind])."""
    if (shared_K is None) == (Ks is None):
        raise ValueError("Provide either shared K or separate K for each camera frame")  # BA.cpp:421
    M = len(inverse_orient_cams)
    ids, pts, oc, op, oxy = [], [], [], [], []
    for track in track_rep.CornerTracks:
        if track.SalientPointId is None:
            continue
        pnt_ind = len(ids)
        ids.append(track.SalientPointId)
        pts.append(map_.GetSalientPoint(track.SalientPointId))

        def on_item(frame_ind, cd, pnt_ind=pnt_ind):
            if cd is None or frame_ind >= M:
                return
            oc.append(frame_ind); op.append(pnt_ind); oxy.append(cd.pixel_coord)
        if track.HasCorners():
            track.EachCorner(on_item)
    cams = np.stack([c.as_flat() for c in inverse_orient_cams]) if M else np.zeros((0, 12))
    if shared_K is not None:
        K = np.asarray(shared_K, dtype=np.float64).reshape(3, 3).T.reshape(1, 9)
    else:
        K = np.stack([np.asarray(k, dtype=np.float64).reshape(3, 3).T.reshape(9) for k in Ks])
    prob = BAProblem(np.array(oc, dtype=np.int32), np.array(op, dtype=np.int32), np.array(oxy, dtype=np.float64).reshape(-1, 2),
                     np.array(pts, dtype=np.float64).reshape(-1, 3), cams, K, shared_K is not None, f0)
    return prob, ids


def scatter_scene(prob, ids, map_, inverse_orient_cams):
    for pnt_ind, sp_id in enumerate(ids):
        map_.GetSalientPoint(sp_id)[:] = prob.points[pnt_ind]
    for i, cam in enumerate(inverse_orient_cams):
        t = SE3Transform.from_flat(prob.cams[i])
        cam.R[:, :] = t.R
        cam.T[:] = t.T


class BundleAdjustmentKanatani:
    """Drop-in for suriko::BundleAdjustmentKanatani on the BA path (bundle-adj-kanatani.h:98-261)."""

    kPointVarsCount = 3
    kIntrinsicVarsCount = 4
    kTVarsCount = 3
    kWVarsCount = 3

    def __init__(self, device=0, engine=None):
        self._engine = engine or Engine(device)
        self.unity_t1_comp_value_ = 1.0
        self.unity_t1_comp_ind_ = 1
        self.vars_count_per_frame_ = 10
        self._points_count = 0
        self._frames_count = 0
        self._stop_reason = ""
        self.last_report = None
        self.solver = 0
        self.max_outer_iters = 0  # 0 = unlimited, the reference's behaviour (quirk Q9)

    # --- static helpers of the reference class
    @staticmethod
    def ReprojError(f0, map_, inverse_orient_cams, track_rep, shared_intrinsic_cam_mat=None, intrinsic_cam_mats=None, engine=None,
                    return_seen_points=False):
        """bundle-adj-kanatani.h:167-172 / BA.cpp:589-600: sum of squared residuals in (pix/f0)^2, no normalisation."""
        prob, _ = flatten_scene(f0, map_, inverse_orient_cams, track_rep, shared_intrinsic_cam_mat, intrinsic_cam_mats)
        eng = engine or Engine(0)
        err, seen = eng.reproj_error(prob)
        return (err, seen) if return_seen_points else err

    def ReprojErrorPixPerPoint(self, reproj_err, seen_points_count):
        """BA.cpp:602-615 (quirk Q15: no -7 dof correction)."""
        return self._f0 * float(np.sqrt(reproj_err / float(seen_points_count)))

    def ComputeInplace(self, f0, map_, inverse_orient_cams, track_rep, shared_intrinsic_cam_mat=None, intrinsic_cam_mats=None,
                       term_crit=None):
        """BA.cpp:617-718.  Refines map_ and inverse_orient_cams in place; K is left untouched (quirk Q2).  Returns bool."""
        if not (0 <= self.unity_t1_comp_ind_ < self.kTVarsCount):
            raise ValueError("Can normalize only one of [T1x, T1y, Tz] components")
        term_crit = term_crit or BundleAdjustmentKanataniTermCriteria()
        self._f0 = float(f0)
        prob, ids = flatten_scene(f0, map_, inverse_orient_cams, track_rep, shared_intrinsic_cam_mat, intrinsic_cam_mats)
        self._points_count, self._frames_count = prob.n_points, prob.n_cams
        opt = BAOptions(err_change=term_crit.AllowedReprojErrRelativeChange(), max_hessian_factor=term_crit.MaxHessianFactor(),
                        unity_comp_ind=self.unity_t1_comp_ind_, unity_comp_value=self.unity_t1_comp_value_,
                        max_outer_iters=self.max_outer_iters, solver=self.solver)
        rep = self._engine.solve(prob, opt)
        self.last_report = rep
        self._stop_reason = rep.stop_reason
        scatter_scene(prob, ids, map_, inverse_orient_cams)
        return rep.converged

    def PointsCount(self): return self._points_count
    def FramesCount(self): return self._frames_count
    def VarsCount(self): return self.kPointVarsCount * self._points_count + self.vars_count_per_frame_ * self._frames_count
    def NormalizedVarsCount(self): return self.VarsCount() - 7
    def OptimizationStatusString(self): return self._stop_reason
