"""Full-size checks at BASELINE.json configs[2] (1,000 cameras x 1M points x 10M observations), where the dense-flow oracle
cannot run: size-independent properties instead of element-wise comparison.

  * the engine's initial error equals a float64 numpy evaluation of the reprojection error (BA.cpp:410-490) of the same
    scene, computed chunk by chunk on the host -- an independent restatement, not the engine's code;
  * the two independent implementations of the Schur accumulation (DMMA contraction, schur_mma.cu; vector-FMA tile kernel,
    ba_kernels.cu) give the same LM step;
  * reset + rerun is repeatable; one LM iteration decreases the error; at the exact (noise-free) solution the error is ~0
    and stays there;
  * the structure the dense factorisation reports is the block band + wrap-around the ring scene must produce.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M, N, K_OBS = 1000, 1_000_000, 10


def numpy_reproj_error(prob):
    """sum |K (R X + T) / z - x/f0|^2 over all observations, float64, chunked (BA.cpp:461-482)."""
    cams = prob.cams.reshape(-1, 12)
    T = cams[:, :3]; R = cams[:, 3:].reshape(-1, 3, 3).transpose(0, 2, 1)      # stored column-major
    Km = prob.K.reshape(-1, 9).reshape(-1, 3, 3).transpose(0, 2, 1)
    tot = 0.0
    step = 2_000_000
    for a in range(0, prob.n_obs, step):
        b = min(prob.n_obs, a + step)
        c = prob.obs_cam[a:b]; p = prob.obs_point[a:b]
        Xc = np.einsum("oij,oj->oi", R[c], prob.points[p]) + T[c]
        pqr = np.einsum("oij,oj->oi", Km[c] if Km.shape[0] > 1 else np.broadcast_to(Km[0], (b - a, 3, 3)), Xc)
        rx = pqr[:, 0] / pqr[:, 2] - prob.obs_xy[a:b, 0] / prob.f0
        ry = pqr[:, 1] / pqr[:, 2] - prob.obs_xy[a:b, 1] / prob.f0
        tot += float(np.sum(rx * rx + ry * ry))
    return tot


@pytest.fixture(scope="module")
def c3_scene():
    from surikatoko_b200 import scenes
    return scenes.ring_scene(M, N, K_OBS, seed=1234)


def test_full_size_lm_step_properties(c3_scene, engine):
    import surikatoko_b200 as sb
    prob = c3_scene
    assert prob.n_obs == 10_000_000
    e_np = numpy_reproj_error(prob)
    e_gpu, seen = engine.reproj_error(prob)
    assert seen == prob.n_obs
    assert abs(e_gpu - e_np) <= 1e-11 * e_np, (e_gpu, e_np)

    opt = sb.BAOptions(max_outer_iters=1)
    assert engine.bind(prob, opt)
    rep1 = engine.run(opt)
    assert rep1.solver_used == sb.SOLVER_DENSE_CHOLESKY
    # the gauge normalisation rescales the world but the error is invariant under it (pixels over f0)
    assert abs(rep1.err_initial - e_np) <= 1e-10 * e_np
    assert rep1.err_final < rep1.err_initial and rep1.outer_iters == 1
    st = engine.solve_stats()
    # nested-dissection order (csrc/solve_order.cu): the ring splits into independent arcs + a separator, every part padded to a 64-column boundary
    assert st["n_f"] == 10 * M - 7 and st["block_rows"] == (st["ordered_n"] + 63) // 64 and st["ordered_n"] >= st["n_f"]
    assert st["parts"] >= 4 and st["max_part_blocks"] + st["separator_blocks"] < st["block_rows"] // 3
    # 10 ring-nearest cameras per point: block band of 2-3 tiles per block column + the wrap-around rows, far from dense
    assert st["block_rows"] < st["nonzero_tiles"] < 8 * st["block_rows"]
    engine.reset()
    rep2 = engine.run(opt)
    assert abs(rep2.err_final - rep1.err_final) <= 1e-12 * rep1.err_final      # atomics reorder sums: not bit-exact, but far below 1e-9

    # capture-order factorisation (one chain of 157 block columns) as a cross-check of the ordered, partitioned one
    os.environ["SRK_SOLVE_ORDER"] = "0"
    try:
        eng_nat = sb.Engine(0)
    finally:
        del os.environ["SRK_SOLVE_ORDER"]
    try:
        assert eng_nat.bind(prob, opt)
        rep_nat = eng_nat.run(opt)
        st_nat = eng_nat.solve_stats()
        assert st_nat["parts"] == 0 and st_nat["block_rows"] == (10 * M - 7 + 63) // 64
        assert abs(rep_nat.err_final - rep1.err_final) <= 1e-11 * rep1.err_final, (rep_nat.err_final, rep1.err_final)
    finally:
        eng_nat.close()

    # second, independent implementation of K2
    os.environ["SRK_SCHUR_IMPL"] = "1"
    try:
        eng2 = sb.Engine(0)
    finally:
        del os.environ["SRK_SCHUR_IMPL"]
    try:
        assert eng2.bind(prob, opt)
        rep3 = eng2.run(opt)
        n1, n3 = np.sqrt(rep1.err_final), np.sqrt(rep3.err_final)
        assert abs(n1 - n3) <= 1e-9 * n1, (rep1.err_final, rep3.err_final)
        assert rep3.attempts_count == rep1.attempts_count
    finally:
        eng2.close()


def test_full_size_exact_scene_is_a_fixed_point(engine):
    """No pixel / pose / point noise: the error at the generating state is rounding noise and an LM iteration keeps it there."""
    import surikatoko_b200 as sb
    from surikatoko_b200 import scenes
    prob = scenes.ring_scene(M, N, K_OBS, seed=99, pix_sigma=0.0, rot_sigma=0.0, trans_rel=0.0, point_rel=0.0)
    e0, _ = engine.reproj_error(prob)
    assert e0 < 1e-18 * prob.n_obs
    opt = sb.BAOptions(max_outer_iters=1, max_hessian_factor=1e3)   # bounds the damping retries when nothing can decrease
    assert engine.bind(prob, opt)
    rep = engine.run(opt)
    assert rep.err_final <= max(rep.err_initial, 1e-18 * prob.n_obs)
