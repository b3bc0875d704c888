"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances (FP64): per-observation / per-block quantities 1e-11 relative to the largest entry; sparsity structure, skip
mask, accept/reject flags and stop reasons exact; accepted per-iteration errors 1e-9 relative (north_star); final RMS
reprojection error 1e-6 px.
"""
import os

import numpy as np
import pytest

from conftest import relerr, to_problem

pytestmark = pytest.mark.gpu


def small_scene(oracle, cell=0.25):
    return oracle.circle_grid_scene(cell_x=cell, cell_y=cell)


def test_reproj_error_matches_oracle(oracle, engine):
    pr = small_scene(oracle)
    e_ref, seen_ref = oracle.reproj_error(pr)
    e_gpu, seen_gpu = engine.reproj_error(to_problem(pr))
    assert seen_gpu == seen_ref == pr.n_obs
    assert abs(e_gpu - e_ref) <= 1e-13 * abs(e_ref)


def test_normalization_matches_oracle(oracle, engine):
    pr = small_scene(oracle)
    ok, pts, cams, cam0, ws = oracle.normalize(pr.points, pr.cams)
    assert ok
    assert engine.bind(to_problem(pr))
    gp, gc = engine.debug_get_state()
    assert relerr(gp, pts) < 1e-14
    assert relerr(gc, cams) < 1e-14
    # round trip through fetch = RevertNormalization
    out = engine.fetch(to_problem(pr))
    assert relerr(out.points, pr.points) < 1e-13
    assert relerr(out.cams, pr.cams) < 1e-13


def normalized_problem(oracle, pr):
    ok, pts, cams, _, _ = oracle.normalize(pr.points, pr.cams)
    assert ok
    q = pr.copy(); q.points = pts; q.cams = cams
    return q


def test_derivative_blocks_match_oracle(oracle, engine):
    pr = small_scene(oracle)
    ref = oracle.derivs_and_solve(normalized_problem(oracle, pr), c=None, flow="dense")
    assert engine.bind(to_problem(pr))
    got = engine.debug_derivs_and_solve(c=None)
    for k in ("gradE", "E", "G", "F"):
        assert relerr(got[k], ref[k]) < 1e-11, k
    # sparsity structure of F: identical non-zero 3x10 blocks (one per observation) -- every block is populated
    assert np.all(np.any(got["F"].reshape(len(got["F"]), -1) != 0, axis=1))


@pytest.mark.parametrize("c", [1e-4, 1e-2, 1.0])
def test_schur_system_and_corrections_match_oracle(oracle, engine, c):
    pr = small_scene(oracle)
    ref = oracle.derivs_and_solve(normalized_problem(oracle, pr), c=c, flow="sparse", solve="chol", acc="ld")
    assert engine.bind(to_problem(pr))
    got = engine.debug_derivs_and_solve(c=c)
    assert np.array_equal(got["skipped"], ref["skipped"])
    assert relerr(got["S"], ref["S"]) < 1e-11
    assert relerr(got["rhs"], ref["rhs"]) < 1e-10
    assert np.array_equal(got["S"] != 0, ref["S"] != 0)
    # the 7 gauge variables get exact zeros (BA.cpp:1600-1679)
    N = pr.n_points
    fixed = [3 * N + i for i in (4, 5, 6, 7, 8, 9, 15)]
    assert np.all(got["corrections"][fixed] == 0.0)
    # the solve is ill-conditioned (cond(S) ~ 1e13 at c = 1e-4): compare through the error after the step
    _, p2, c2, = (None,) + oracle.apply_corrections(normalized_problem(oracle, pr).points, normalized_problem(oracle, pr).cams, ref["corrections"])
    q = normalized_problem(oracle, pr); q.points = p2; q.cams = c2
    e_ref, _ = oracle.reproj_error(q)
    e_gpu = engine.debug_apply(got["corrections"])
    assert abs(e_gpu - e_ref) <= 1e-7 * abs(e_ref)


def run_pair(oracle, engine, pr, err_change, max_outer_iters, **kw):
    """Returns (exact oracle run, GPU report, refined problem).  The parity target is the oracle in "exact" mode (Schur
    complement and solve in long double); ref.noise holds the deviation of the "faithful" oracle (the reference's own plain
    double arithmetic with Householder QR, BA.cpp:1911) from it -- the reference's intrinsic floating-point noise on this
    scene (cond(S) ~ 1e13), which grows along the trajectory."""
    import surikatoko_b200 as sb
    ref = oracle.ba_solve(pr, err_change=err_change, max_outer_iters=max_outer_iters, flow="sparse", solve="chol", acc="ld")
    faithful = oracle.ba_solve(pr, err_change=err_change, max_outer_iters=max_outer_iters, flow="sparse", solve="qr", acc="double")
    n = min(len(ref.err_trace), len(faithful.err_trace))
    ref.noise = np.zeros(len(ref.err_trace))
    ref.noise[:n] = np.abs(np.sqrt(faithful.err_trace[:n]) - np.sqrt(ref.err_trace[:n])) / np.sqrt(ref.err_trace[:n])
    if n < len(ref.err_trace):
        ref.noise[n:] = np.inf
    prob = to_problem(pr)
    rep = engine.solve(prob, sb.BAOptions(err_change=err_change, max_outer_iters=max_outer_iters, **kw))
    return ref, rep, prob


def report_deviation(label, dev, noise):
    """Makes the tolerance visible: per accepted iteration the achieved deviation of the GPU residual norm from the exact oracle next to
    the faithful oracle's own deviation (the reference's FP64 noise on that scene); printed (pytest -s / -rP) and appended to
    gpurun_out/parity_deviation.jsonl so that DESIGN.md can quote measured numbers.  Iterations whose bound had to be relaxed beyond the
    north star's 1e-9 are counted explicitly."""
    import json
    relaxed = int(np.sum((dev > 1e-9)))
    line = {"scene": label, "iterations": int(len(dev)), "max_dev_vs_exact": float(np.max(dev)) if len(dev) else 0.0,
            "iters_within_1e-9": int(np.sum(dev <= 1e-9)), "iters_relaxed_to_reference_noise": relaxed,
            "max_reference_noise": float(np.max(noise[np.isfinite(noise)])) if np.any(np.isfinite(noise)) else None,
            "dev_first8": [float(v) for v in dev[:8]], "noise_first8": [float(v) for v in noise[:8]]}
    print("PARITY " + json.dumps(line))
    try:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_deviation.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    except OSError:
        pass


def check_trajectory(ref, rep, pr, prob, f0, label="scene"):
    assert rep.seen_points == ref.seen_points
    assert abs(rep.err_initial - ref.err_initial) <= 1e-12 * ref.err_initial
    n = min(len(ref.attempts), len(rep.attempts))
    assert len(ref.attempts) == len(rep.attempts)
    assert np.array_equal(rep.attempts[:n, 2], ref.attempts[:n, 2]), "accept/reject flags differ"
    assert np.array_equal(rep.attempts[:n, 3], ref.attempts[:n, 3]), "skipped-point counts differ"
    assert np.allclose(rep.attempts[:n, 0], ref.attempts[:n, 0], rtol=1e-15, atol=0)
    assert len(rep.err_trace) == len(ref.err_trace)
    # per-iteration residual norms sqrt(err): 1e-9 relative to the exact oracle; the first iteration starts from the identical
    # state and must meet it outright, later ones get max(1e-9, the reference's own double-precision noise at that iteration)
    dev = np.abs(np.sqrt(rep.err_trace) - np.sqrt(ref.err_trace)) / np.sqrt(ref.err_trace)
    report_deviation(label, dev, ref.noise)
    assert dev[0] < 1e-9
    assert np.all(dev <= np.maximum(1e-9, ref.noise)), (dev, ref.noise)
    assert rep.stop_reason == ref.stop_reason
    assert rep.converged == ref.converged
    rms_ref = f0 * np.sqrt(ref.err_final / ref.seen_points)
    rms_gpu = f0 * np.sqrt(rep.err_final / rep.seen_points)
    assert abs(rms_ref - rms_gpu) < 1e-6
    assert relerr(prob.points, ref.points) < 1e-6
    assert relerr(prob.cams, ref.cams) < 1e-6


def test_lm_trajectory_matches_oracle_small(oracle, engine):
    pr = small_scene(oracle)
    ref, rep, prob = run_pair(oracle, engine, pr, 1e-10, 8)
    check_trajectory(ref, rep, pr, prob, pr.f0)
    assert rep.gpu_launches > 0


def test_lm_converges_with_demo_threshold(oracle, engine):
    # the demo's default stop rule (flagfile: allowed relative change 1e-8 in (pix/f0)^2 units)
    pr = small_scene(oracle, cell=0.5)
    ref, rep, prob = run_pair(oracle, engine, pr, 1e-8, 0)
    assert ref.stop_reason in ("small relative err change", "err converged to limit value")
    check_trajectory(ref, rep, pr, prob, pr.f0)


def test_reset_and_rerun_is_repeatable(oracle, engine):
    import surikatoko_b200 as sb
    pr = small_scene(oracle)
    assert engine.bind(to_problem(pr))
    opt = sb.BAOptions(max_outer_iters=2)
    a = engine.run(opt)
    engine.reset()
    b = engine.run(opt)
    assert a.err_initial == b.err_initial
    assert np.max(np.abs(a.err_trace - b.err_trace) / a.err_trace) < 1e-9


def test_normalization_failure_returns_false_and_leaves_scene(oracle, engine):
    import surikatoko_b200 as sb
    pr = small_scene(oracle)
    pr.cams[1] = pr.cams[0]       # cam1 == cam0: T01 = 0 -> NormalizeSceneInplace fails (BA.cpp:215-217)
    prob = to_problem(pr)
    before_p, before_c = prob.points.copy(), prob.cams.copy()
    rep = engine.solve(prob, sb.BAOptions(max_outer_iters=2))
    assert not rep.converged and rep.stop_reason == ""
    assert np.array_equal(prob.points, before_p) and np.array_equal(prob.cams, before_c)


def test_invalid_arguments_fail_loudly(oracle, engine):
    import surikatoko_b200 as sb
    pr = small_scene(oracle)
    bad = to_problem(pr); bad.f0 = 0.0
    with pytest.raises(sb.SrkError):
        engine.solve(bad)
    bad = to_problem(pr); bad.obs_point[:] = bad.obs_point[::-1].copy()
    with pytest.raises(sb.SrkError):
        engine.solve(bad)
    with pytest.raises(sb.SrkError):
        engine.solve(to_problem(pr), sb.BAOptions(unity_comp_ind=3))


# ---------------------------------------------------------------------------------------------------------------------
# Sparse-visibility scenes: these go through the tiled Schur kernel (short tracks, localized cameras); the dinosaur-shaped
# scene mixes it with the per-point fallback (tracks longer than 16 frames) and ragged track lengths.

def as_oracle_problem(oracle, prob):
    return oracle.Problem(prob.obs_cam.copy(), prob.obs_point.copy(), prob.obs_xy.copy(), prob.points.copy(), prob.cams.copy(), prob.K.copy(),
                          prob.shared_K, prob.f0)


def scene_by_name(name):
    from surikatoko_b200 import scenes
    if name == "ring":
        return scenes.ring_scene(40, 3000, 8, seed=7)
    if name == "ring_wide":     # 14 cameras per point > the 12-camera tile table: every point is deferred to the per-point kernel
        return scenes.ring_scene(60, 1500, 14, seed=8)
    if name == "dino":
        return scenes.dino_shaped_scene(n_points=1200, n_obs=4000, seed=9)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["ring", "ring_wide", "dino"])
@pytest.mark.parametrize("c", [1e-4, 1.0])
def test_schur_system_sparse_scenes(oracle, engine, name, c):
    prob = scene_by_name(name)
    pr = as_oracle_problem(oracle, prob)
    ref = oracle.derivs_and_solve(normalized_problem(oracle, pr), c=c, flow="sparse", solve="chol", acc="ld")
    assert engine.bind(to_problem(pr))
    got = engine.debug_derivs_and_solve(c=c)
    for k in ("gradE", "E", "G", "F"):
        assert relerr(got[k], ref[k]) < 1e-11, k
    assert np.array_equal(got["skipped"], ref["skipped"])
    assert np.array_equal(got["S"] != 0, ref["S"] != 0), "sparsity structure of the reduced camera system differs"
    assert relerr(got["S"], ref["S"]) < 1e-11
    assert relerr(got["rhs"], ref["rhs"]) < 1e-10


@pytest.mark.parametrize("name", ["ring", "dino"])
def test_lm_trajectory_sparse_scenes(oracle, engine, name):
    prob = scene_by_name(name)
    pr = as_oracle_problem(oracle, prob)
    ref, rep, out = run_pair(oracle, engine, pr, 1e-10, 5)
    check_trajectory(ref, rep, pr, out, pr.f0)


def test_ordered_partitioned_solve_matches_oracle(oracle, engine):
    """170 cameras on a ring: large enough (n_f = 1693, 27 block columns) for the nested-dissection order of csrc/solve_order.cu --
    parts factored concurrently, separator last.  The corrections must still be the reference's (BA.cpp:1911 solves in capture order;
    a symmetric permutation does not change the solution): compared with the exact oracle through the error after the step, and
    with the same engine forced to capture order (SRK_SOLVE_ORDER=0)."""
    import surikatoko_b200 as sb
    from surikatoko_b200 import scenes
    prob = scenes.ring_scene(170, 4000, 6, seed=9)
    pr = as_oracle_problem(oracle, prob)
    c = 1e-4
    ref = oracle.derivs_and_solve(normalized_problem(oracle, pr), c=c, flow="sparse", solve="chol", acc="ld")
    assert engine.bind(to_problem(pr))
    got = engine.debug_derivs_and_solve(c=c)
    st = engine.solve_stats()
    assert st["parts"] >= 2 and st["ordered_n"] >= st["n_f"] and st["separator_blocks"] >= 1
    assert np.array_equal(got["skipped"], ref["skipped"])
    assert np.array_equal(got["S"] != 0, ref["S"] != 0)
    assert relerr(got["S"], ref["S"]) < 1e-11
    q = normalized_problem(oracle, pr)
    _, p2, c2 = (None,) + oracle.apply_corrections(q.points, q.cams, ref["corrections"])
    q.points = p2; q.cams = c2
    e_ref, _ = oracle.reproj_error(q)
    e_gpu = engine.debug_apply(got["corrections"])
    assert abs(e_gpu - e_ref) <= 1e-7 * abs(e_ref)
    os.environ["SRK_SOLVE_ORDER"] = "0"
    try:
        eng_nat = sb.Engine(0)
    finally:
        del os.environ["SRK_SOLVE_ORDER"]
    try:
        assert eng_nat.bind(to_problem(pr))
        nat = eng_nat.debug_derivs_and_solve(c=c)
        assert eng_nat.solve_stats()["parts"] == 0
        e_nat = eng_nat.debug_apply(nat["corrections"])
        assert abs(e_gpu - e_nat) <= 1e-9 * abs(e_nat)
        # whole LM trajectories, ordered vs capture order: same decisions, same residual norms
        opt = sb.BAOptions(err_change=1e-10, max_outer_iters=4)
        r1 = engine.solve(to_problem(pr), opt); r0 = eng_nat.solve(to_problem(pr), opt)
        assert np.array_equal(r1.attempts[:, 2], r0.attempts[:, 2]) and r1.stop_reason == r0.stop_reason
        assert np.all(np.abs(np.sqrt(r1.err_trace) - np.sqrt(r0.err_trace)) <= 1e-9 * np.sqrt(r0.err_trace))
    finally:
        eng_nat.close()


def test_skipped_points_mask_is_reproduced(oracle, engine):
    # two-frame tracks with a narrow baseline have det(E_damped) <= 1e-12 in the demos' f0 = 600 units (quirk Q5)
    from surikatoko_b200 import scenes
    prob = scenes.ring_scene(200, 2000, 2, seed=11, level_step=0.05)
    pr = as_oracle_problem(oracle, prob)
    ref = oracle.derivs_and_solve(normalized_problem(oracle, pr), c=1e-4, flow="sparse", solve="chol", acc="ld")
    assert engine.bind(to_problem(pr))
    got = engine.debug_derivs_and_solve(c=1e-4)
    assert ref["skipped"].sum() > 0, "scene does not exercise the skip rule"
    assert np.array_equal(got["skipped"], ref["skipped"])
    assert np.array_equal(got["S"] != 0, ref["S"] != 0)
    N = pr.n_points
    sk = np.nonzero(ref["skipped"])[0]
    assert np.all(got["corrections"][:3 * N].reshape(N, 3)[sk] == 0.0)


# ---------------------------------------------------------------------------------------------------------------------
def test_gpu_matches_reference_python_prototype_fixture(engine):
    """The committed outputs of the reference's own Python prototype (tests/golden/pyproto_derivs.json, f0 = K22 = 1)."""
    import json, os
    import surikatoko_b200 as sb
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyproto_derivs.json")))
    M, N = g["n_cams"], g["n_points"]
    prob = sb.BAProblem(g["obs_cam"], g["obs_point"], g["obs_xy"], g["points"], g["cams"], np.tile(np.array(g["K"]), (M, 1)), False, g["f0"])
    err, seen = engine.reproj_error(prob.copy())
    assert seen == len(g["obs_cam"]) and abs(err - g["err_initial"]) <= 1e-12 * g["err_initial"]
    assert engine.bind(prob)       # the fixture state is already normalised: normalising again is the identity up to rounding
    got = engine.debug_derivs_and_solve(c=g["hessian_factor"])
    assert relerr(got["gradE"], np.array(g["gradE"])) < 1e-11
    assert relerr(got["E"].reshape(3 * N, 3), np.array(g["E"])) < 1e-11
    assert relerr(got["G"].reshape(10 * M, 10), np.array(g["G"])) < 1e-11
    Fd = np.array(g["F"]); mine = np.zeros_like(Fd)
    for o in range(len(g["obs_cam"])):
        p, f = g["obs_point"][o], g["obs_cam"][o]
        mine[3 * p:3 * p + 3, 10 * f:10 * f + 10] = got["F"][o]
    assert relerr(mine, Fd) < 1e-11 and np.array_equal(mine != 0, Fd != 0)
    assert relerr(got["S"], np.array(g["S"])) < 1e-10
    assert relerr(got["rhs"], np.array(g["rhs"])) < 1e-9
    assert relerr(got["corrections"], np.array(g["corrections"])) < 1e-4    # cond(S) ~ 1e11: LA.solve vs refined Cholesky


def test_cpp_drop_in_adapter_builds_and_runs(tmp_path):
    """include/suriko_compat (C++17, the reference's container API) -> C ABI -> GPU."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "surikatoko_b200", "_lib")
    exe = str(tmp_path / "compat_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(root, "tests", "cpp", "compat_demo.cpp"), "-L" + lib, "-lsrk_ba",
                           "-Wl,-rpath," + lib, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    f = out.stdout.split()
    e0, e1 = float(f[0]), float(f[1])
    assert e1 < 0.5 * e0 and int(f[-3]) > 0 and int(f[-2]) == 12 * 64


def test_python_mirror_of_reference_interface(oracle, engine):
    """BundleAdjustmentKanatani.ComputeInplace on FragmentMap / CornerTrackRepository containers == flat path."""
    import surikatoko_b200 as sb
    from surikatoko_b200 import ba
    pr = small_scene(oracle, cell=0.5)
    m = ba.FragmentMap(); rep = ba.CornerTrackRepository()
    for p in range(pr.n_points):
        _, sp_id = m.AddSalientPointTempl(pr.points[p])
        rep.AddCornerTrackObj().SalientPointId = sp_id
    for o in range(pr.n_obs):
        rep.GetPointTrackById(int(pr.obs_point[o])).AddCorner(int(pr.obs_cam[o]), pr.obs_xy[o])
    cams = [ba.SE3Transform.from_flat(c) for c in pr.cams]
    Ks = [k.reshape(3, 3).T for k in pr.K]
    bak = sb.BundleAdjustmentKanatani(engine=engine)
    bak.max_outer_iters = 4
    tc = sb.BundleAdjustmentKanataniTermCriteria(); tc.AllowedReprojErrRelativeChange(1e-10)
    e0 = sb.BundleAdjustmentKanatani.ReprojError(pr.f0, m, cams, rep, None, Ks, engine=engine)
    ok = bak.ComputeInplace(pr.f0, m, cams, rep, None, Ks, tc)
    flat = engine.solve(to_problem(pr), sb.BAOptions(err_change=1e-10, max_outer_iters=4))
    assert abs(e0 - flat.err_initial) <= 1e-14 * e0
    assert ok == flat.converged and bak.OptimizationStatusString() == flat.stop_reason
    assert np.allclose(bak.last_report.err_trace, flat.err_trace, rtol=1e-9)
    assert bak.PointsCount() == pr.n_points and bak.FramesCount() == pr.n_cams and bak.NormalizedVarsCount() == 3 * pr.n_points + 10 * pr.n_cams - 7
    rms = bak.ReprojErrorPixPerPoint(bak.last_report.err_final, bak.last_report.seen_points)
    assert abs(rms - pr.f0 * np.sqrt(flat.err_final / flat.seen_points)) < 1e-9


# ---------------------------------------------------------------------------------------------------------------------
# K3b: block-sparse reduced camera system + block-Jacobi PCG

@pytest.mark.parametrize("name", ["ring", "dino", "ring_wide"])
def test_pcg_system_and_solution_match_dense_path(oracle, engine, name):
    import surikatoko_b200 as sb
    prob = scene_by_name(name)
    pr = as_oracle_problem(oracle, prob)
    ref = oracle.derivs_and_solve(normalized_problem(oracle, pr), c=1e-2, flow="sparse", solve="chol", acc="ld")
    assert engine.bind(to_problem(pr))
    got = engine.debug_derivs_and_solve(c=1e-2, solver=sb.SOLVER_BLOCK_PCG)
    assert got["pcg_iters"] > 0
    assert np.array_equal(got["skipped"], ref["skipped"])
    assert np.array_equal(got["S"] != 0, ref["S"] != 0), "sparsity structure of the block-sparse system differs"
    assert relerr(got["S"], ref["S"]) < 1e-11
    assert relerr(got["rhs"], ref["rhs"]) < 1e-10
    N = pr.n_points
    fixed = [3 * N + i for i in (4, 5, 6, 7, 8, 9, 15)]
    assert np.all(got["corrections"][fixed] == 0.0)
    q = normalized_problem(oracle, pr)
    p2, c2 = oracle.apply_corrections(q.points, q.cams, ref["corrections"])
    z = q.copy(); z.points = p2; z.cams = c2
    e_ref, _ = oracle.reproj_error(z)
    e_gpu = engine.debug_apply(got["corrections"])
    assert abs(e_gpu - e_ref) <= 1e-7 * abs(e_ref)


def test_pcg_lm_trajectory(oracle, engine):
    import surikatoko_b200 as sb
    prob = scene_by_name("ring")
    pr = as_oracle_problem(oracle, prob)
    ref, rep, out = run_pair(oracle, engine, pr, 1e-10, 4, solver=sb.SOLVER_BLOCK_PCG)
    assert rep.solver_used == sb.SOLVER_BLOCK_PCG and rep.pcg_iters_last > 0
    check_trajectory(ref, rep, pr, out, pr.f0)


def test_ordered_solve_with_loop_closures(engine):
    """A ring scene with a few long-range co-visibilities (loop closures): the camera graph is no longer a clean band, the nested-
    dissection order has to route those couplings through its separator.  Ordered / partitioned solve vs capture order, same engine."""
    import surikatoko_b200 as sb
    from surikatoko_b200 import scenes
    base = scenes.ring_scene(170, 4000, 6, seed=12)
    rng = np.random.default_rng(5)
    obs_cam = base.obs_cam.copy(); obs_xy = base.obs_xy.copy()
    K = base.K.reshape(-1, 9)
    ends = np.searchsorted(base.obs_point, np.arange(base.n_points), side="right") - 1
    cand = np.nonzero((obs_cam[ends] >= 15) & (obs_cam[ends] < 25))[0]          # one revisited place: cameras 15..24 see cameras 85..97 again
    pts = rng.choice(cand, size=min(40, len(cand)), replace=False)
    first = np.searchsorted(base.obs_point, pts, side="left"); last = np.searchsorted(base.obs_point, pts, side="right") - 1
    for j, a, b in zip(pts, first, last):
        far = int(obs_cam[b] + 70 + rng.integers(0, 4))
        if far >= base.n_cams or far <= obs_cam[b]:
            continue                                  # keeps (pnt_ind, frame_ind) sorted and unique
        cam = base.cams[far]; T = cam[:3]; R = cam[3:].reshape(3, 3).T       # column-major R
        Kc = (K[0] if base.shared_K else K[far]).reshape(3, 3).T
        pc = Kc @ (R @ base.points[j] + T)
        if pc[2] <= 0:
            continue
        obs_cam[b] = far
        obs_xy[b] = base.f0 * pc[:2] / pc[2] + rng.normal(0, 0.3, 2)
    assert np.any(obs_cam != base.obs_cam)

    def problem():
        return sb.BAProblem(obs_cam, base.obs_point, obs_xy, base.points.copy(), base.cams.copy(), base.K, base.shared_K, base.f0)
    opt = sb.BAOptions(err_change=1e-10, max_outer_iters=3)
    r1 = engine.solve(problem(), opt)
    st = engine.solve_stats()
    assert st["parts"] >= 2, st
    os.environ["SRK_SOLVE_ORDER"] = "0"
    try:
        eng_nat = sb.Engine(0)
    finally:
        del os.environ["SRK_SOLVE_ORDER"]
    try:
        r0 = eng_nat.solve(problem(), opt)
        assert eng_nat.solve_stats()["parts"] == 0
    finally:
        eng_nat.close()
    assert np.array_equal(r1.attempts[:, 2], r0.attempts[:, 2]) and r1.stop_reason == r0.stop_reason
    assert np.all(np.abs(np.sqrt(r1.err_trace) - np.sqrt(r0.err_trace)) <= 1e-9 * np.sqrt(r0.err_trace))
