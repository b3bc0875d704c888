#!/usr/bin/env python
"""Headless equivalent of the reference's demo-dino (cpp_impl/demos/demo-bundle-adj-dinosaur.cpp:79-260) on the new engine:

  ReadMatrixFromFile(dinoPs_as_mat108x4.txt, '\\t'), ReadMatrixFromFile(viff.xy, ' ')  ->  tracks (PopulateCornersPerFrame)
  DecomposeProjMat per frame  ->  K / f0, inverse poses, f0-scaled projection matrices
  Triangulate3DPointByLeastSquares per track (GPU, batched)  ->  FragmentMap
  BundleAdjustmentKanatani.ComputeInplace  ->  refined poses / points, RMS reprojection error before and after

    python examples/demo_dino.py --testdata <dir containing oxfvisgeom/dinosaur/>      # the Oxford files, when available
    python examples/demo_dino.py --synthetic                                           # writes files of the same format first
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_synthetic_dinosaur(dirpath, n_frames=36, n_points=600, seed=3, pix_sigma=0.3):
    """Turntable cameras and ragged tracks written in the two file formats of testdata/oxfvisgeom/dinosaur."""
    from surikatoko_b200 import scenes
    rng = np.random.default_rng(seed)
    prob = scenes.dino_shaped_scene(n_cams=n_frames, n_points=n_points, n_obs=int(3.3 * n_points), seed=seed, pix_sigma=0.0, point_rel=0.0, rot_sigma=0.0)
    K_pix = scenes.K_PIX_DINO
    rows = []
    for i in range(n_frames):
        T = prob.gt_cams[i, :3]; R = prob.gt_cams[i, 3:].reshape(3, 3).T
        P = K_pix @ np.concatenate([R, T[:, None]], axis=1)
        rows.append(P * (1.0 + 0.1 * i))          # projective scale is arbitrary, as in the Oxford file
    os.makedirs(dirpath, exist_ok=True)
    np.savetxt(os.path.join(dirpath, "dinoPs_as_mat108x4.txt"), np.concatenate(rows, axis=0), delimiter="\t", fmt="%.17g")
    viff = -np.ones((n_points, 2 * n_frames))
    xy = prob.obs_xy + rng.normal(0.0, pix_sigma, prob.obs_xy.shape)
    viff[prob.obs_point, 2 * prob.obs_cam] = xy[:, 0]
    viff[prob.obs_point, 2 * prob.obs_cam + 1] = xy[:, 1]
    with open(os.path.join(dirpath, "viff.xy"), "w") as f:
        for r in viff:
            f.write(" ".join("-1" if v == -1 else "%.17g" % v for v in r) + "\n")
    return prob


def run(dino_dir, f0=600.0, allowed_repr_err=4.56e-8, max_outer_iters=0, verbose=True):
    import surikatoko_b200 as sb
    from surikatoko_b200 import ba, frontend
    P_rows = frontend.ReadMatrixFromFile(os.path.join(dino_dir, "dinoPs_as_mat108x4.txt"), "\t")
    viff = frontend.ReadMatrixFromFile(os.path.join(dino_dir, "viff.xy"), " ")
    n_frames = P_rows.shape[0] // 3
    assert viff.shape[1] == 2 * n_frames, "Inconsistent frames_count"
    # PopulateCornersPerFrame (:24-57): AddCorner push_backs, so a track's corners sit at consecutive frames from its first one
    track_rep = ba.CornerTrackRepository()
    for pnt_ind in range(viff.shape[0]):
        track = ba.CornerTrack()
        for frame_ind in range(n_frames):
            x, y = viff[pnt_ind, 2 * frame_ind], viff[pnt_ind, 2 * frame_ind + 1]
            if x == -1 or y == -1:
                continue
            track.AddCorner(frame_ind, (x, y))
        if not track.HasCorners():
            continue
        track.SyntheticVirtualPointId = 10000 + pnt_ind
        track.TrackId = len(track_rep.CornerTracks)
        track_rep.CornerTracks.append(track)
    num_stab = np.diag([1.0 / f0, 1.0 / f0, 1.0])
    Ks, inv_cams, P_f0 = [], [], []
    for i in range(n_frames):                       # :118-150
        ok, scale, K, (R, T) = frontend.DecomposeProjMat(P_rows[3 * i:3 * i + 3])
        assert ok, "Can't decompose projection matrix for frame_ind=%d" % i
        Knew = num_stab @ K
        Knew[0, 1] = 0.0                            # zero_cam_intrinsic_mat_01
        Ri, Ti = R.T, -R.T @ T                      # SE3Inv(direct)
        Ks.append(Knew); inv_cams.append(ba.SE3Transform(Ri, Ti))
        P_f0.append(np.concatenate([Knew @ Ri, (Knew @ Ti)[:, None]], axis=1))
    # triangulation (:153-186): every track through GetCorner, batched on the GPU
    tb, fr, xy = [0], [], []
    for track in track_rep.CornerTracks:
        for frame_ind in range(n_frames):
            c = track.GetCorner(frame_ind)
            if c is None:
                continue
            fr.append(frame_ind); xy.append(c)
        tb.append(len(fr))
    X = frontend.Triangulate3DPointByLeastSquares(tb, fr, np.array(xy), np.array(P_f0), f0)
    fmap = ba.FragmentMap()
    for track, x3 in zip(track_rep.CornerTracks, X):
        frag, sid = fmap.AddSalientPointTempl(x3)
        frag.synthetic_virtual_point_id = track.SyntheticVirtualPointId
        track.SalientPointId = sid
    badj = sb.BundleAdjustmentKanatani()
    err0 = badj.ReprojError(f0, fmap, inv_cams, track_rep, None, Ks)
    n_obs = len(fr)
    crit = sb.BundleAdjustmentKanataniTermCriteria()
    crit.AllowedReprojErrRelativeChange(allowed_repr_err)
    if max_outer_iters:
        badj.max_outer_iters = max_outer_iters
    ok = badj.ComputeInplace(f0, fmap, inv_cams, track_rep, None, Ks, crit)
    err1 = badj.ReprojError(f0, fmap, inv_cams, track_rep, None, Ks)
    out = dict(converged=ok, reason=badj.OptimizationStatusString(), frames=n_frames, tracks=len(track_rep.CornerTracks), observations=n_obs,
               rms_px_initial=f0 * np.sqrt(err0 / n_obs), rms_px_final=f0 * np.sqrt(err1 / n_obs))
    if verbose:
        print(out)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--testdata", default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--f0", type=float, default=600.0)
    ap.add_argument("--allowed_repr_err", type=float, default=4.56e-8)   # flagfile-demo-dino.txt:10
    a = ap.parse_args()
    if a.testdata and not a.synthetic:
        d = os.path.join(a.testdata, "oxfvisgeom", "dinosaur")
        run(d, a.f0, a.allowed_repr_err)
    else:
        with tempfile.TemporaryDirectory() as td:
            write_synthetic_dinosaur(td)
            run(td, a.f0, a.allowed_repr_err)


if __name__ == "__main__":
    main()
