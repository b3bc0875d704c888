"""The multi-rank data path under the driver's eyes on ONE GPU: `world` engine handles in one process, every handle bound to its shard of
the points (scenes.shard_points, cameras replicated), coupled by a host all-reduce callback that sums the handles' device buffers in rank
order (what NCCL does between GPUs; tests/multi_gpu_parity.py and bench.py's multi_gpu_parity field exercise NCCL itself at N > 1).
Checked against ONE handle on the whole problem: every exchange of SURVEY.md 8e is on the path -- G / g_f after the derivative pass, the
tiles of S + rhs per attempt (or the PCG mat-vec results), the error slots added in rank order, the camera co-visibility graph as a union
over ranks for the elimination order -- and all ranks must take bit-identical decisions."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def run_sharded(full, world, opt):
    import torch
    import surikatoko_b200 as sb
    from surikatoko_b200 import scenes
    barrier = threading.Barrier(world)
    turn = threading.Lock()      # one handle's host code at a time (the library's threading contract: calls are serialised by the caller)
    pending, shared, results, errors = [None] * world, {}, [None] * world, []

    def make_cb(r):
        def cb(ptr, count, _stream):
            torch.cuda.synchronize()
            pending[r] = torch.as_tensor(_DevView(ptr, count), device="cuda:0")
            turn.release()
            barrier.wait()
            if r == 0:
                tot = pending[0].clone()
                for q in range(1, world):
                    tot += pending[q]            # rank order, like a deterministic ring
                shared["tot"] = tot
                torch.cuda.synchronize()
            barrier.wait()
            pending[r].copy_(shared["tot"])
            torch.cuda.synchronize()
            barrier.wait()
            turn.acquire()
        return cb

    def worker(r):
        try:
            shard, (p0, p1) = scenes.shard_points(full, r, world)
            with turn:
                eng = sb.Engine(0)
                try:
                    eng.set_allreduce(make_cb(r), r, world)
                    rep = eng.solve(shard, opt)
                    st = eng.solve_stats() if rep.solver_used == 1 else None
                finally:
                    eng.close()
            results[r] = (rep, shard, (p0, p1), st)
        except Exception as ex:  # pragma: no cover
            errors.append(ex)
            barrier.abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errors:
        raise errors[0]
    return results


@pytest.mark.parametrize("scene,world,solver", [("ring40", 2, "auto"), ("ring170", 2, "auto"), ("ring170", 3, "auto"), ("ring40", 2, "pcg")])
def test_sharded_job_matches_the_whole_problem(scene, world, solver):
    import surikatoko_b200 as sb
    from surikatoko_b200 import scenes
    full = scenes.ring_scene(170, 4000, 6, seed=9) if scene == "ring170" else scenes.ring_scene(40, 3000, 8, seed=7)
    opt = sb.BAOptions(err_change=1e-10, max_outer_iters=4, solver=sb.SOLVER_BLOCK_PCG if solver == "pcg" else sb.SOLVER_AUTO)
    one = sb.Engine(0)
    try:
        whole = full.copy()
        ref = one.solve(whole, opt)
        ref_stats = one.solve_stats() if ref.solver_used == 1 else None
    finally:
        one.close()
    res = run_sharded(full, world, opt)
    rep0, shard0, _, st0 = res[0]
    for r in range(1, world):   # every rank: the same decisions, bit-identical errors and cameras
        rep, shard, _, _ = res[r]
        assert np.array_equal(rep.err_trace, rep0.err_trace) and np.array_equal(rep.attempts, rep0.attempts)
        assert rep.stop_reason == rep0.stop_reason and rep.seen_points == rep0.seen_points == full.n_obs
        assert np.array_equal(shard.cams, shard0.cams), "ranks diverged"
    assert rep0.err_initial == pytest.approx(ref.err_initial, rel=1e-13)
    assert np.array_equal(rep0.attempts[:, 2], ref.attempts[:, 2]) and np.array_equal(rep0.attempts[:, 3], ref.attempts[:, 3])
    assert rep0.stop_reason == ref.stop_reason and len(rep0.err_trace) == len(ref.err_trace)
    dev = np.abs(np.sqrt(rep0.err_trace) - np.sqrt(ref.err_trace)) / np.sqrt(ref.err_trace)
    print("PARITY sharded x%d (%s, %s): per-iteration residual-norm deviation vs the whole problem %s" % (world, scene, solver, dev))
    # identical start state: the first iteration differs only by summation order (1e-13); afterwards the differences are amplified by the
    # conditioning of the reduced camera system (cond(S) ~ 1e13 in the demos' units: two plain-double runs of the REFERENCE drift apart by
    # 1e-9 .. 1e-6 over five iterations on this scene)
    assert dev[0] < (1e-9 if solver == "auto" else 1e-7), dev
    assert np.all(dev < 1e-6), dev
    pts = np.concatenate([res[r][1].points for r in range(world)], axis=0)
    assert pts.shape == whole.points.shape
    assert np.max(np.abs(pts - whole.points)) / np.max(np.abs(whole.points)) < 1e-6
    assert np.max(np.abs(shard0.cams - whole.cams)) / np.max(np.abs(whole.cams)) < 1e-6
    if scene == "ring170" and ref_stats is not None:   # the elimination order comes from the union of the ranks' camera graphs
        assert st0["parts"] == ref_stats["parts"] >= 2 and st0["ordered_n"] == ref_stats["ordered_n"]


def test_cpp_host_drives_two_devices_in_one_process(tmp_path):
    """tests/cpp/two_device_demo.cpp: C++17 host, one process, one thread + one engine handle per device, the library's own NCCL exchange
    (srk_nccl_unique_id / srk_ba_nccl_init) -- no Python in the loop.  Needs two sm_100 devices (exit code 77 = skipped)."""
    import os
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "surikatoko_b200", "_lib")
    exe = str(tmp_path / "two_device_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I/usr/local/cuda/include", "-o", exe, os.path.join(root, "tests", "cpp", "two_device_demo.cpp"),
                           "-L" + lib, "-lsrk_ba", "-Wl,-rpath," + lib, "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64"])
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    env = dict(os.environ)
    try:   # the NCCL that torch ships, when no system libnccl.so.2 is on the loader path
        import nvidia.nccl
        cand = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            env.setdefault("SRK_NCCL_LIB", cand)
    except Exception:
        pass
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    print(out.stdout)
    assert out.returncode == 0, out.stdout
