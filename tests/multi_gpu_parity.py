"""Multi-GPU parity check, launched by torchrun (one process per GPU, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_parity.py

Every rank binds its shard of points (cameras replicated), the engine all-reduces G/g_f, S/rhs and the error partials through
the callback; the trajectory must match the single-rank oracle on the whole scene, and every rank must hold the same poses.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import surikatoko_b200 as sb                      # noqa: E402
from surikatoko_b200 import scenes                # noqa: E402
from surikatoko_b200.dist import attach_allreduce  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # SRK_TEST_SCENE=ring170: large enough (n_f = 1693) for the nested-dissection order of the dense solve (csrc/solve_order.cu), whose
    # camera graph is the union over ranks; checked against a single-process engine on the whole scene instead of the (slow) oracle
    big = os.environ.get("SRK_TEST_SCENE", "") == "ring170"
    full = scenes.ring_scene(170, 4000, 6, seed=9) if big else scenes.ring_scene(40, 3000, 8, seed=7)
    shard, (p0, p1) = scenes.shard_points(full, rank, world)
    stream = torch.cuda.Stream(device=dev)
    eng = sb.Engine(local)
    eng.set_stream(stream.cuda_stream)
    attach_allreduce(eng, stream, dev)
    solver = sb.SOLVER_BLOCK_PCG if os.environ.get("SRK_TEST_SOLVER", "") == "pcg" else sb.SOLVER_AUTO
    opt = sb.BAOptions(err_change=1e-10, max_outer_iters=4, solver=solver)
    with torch.cuda.stream(stream):
        rep = eng.solve(shard, opt)
    torch.cuda.synchronize()
    # every rank holds identical cameras afterwards
    c = torch.from_numpy(shard.cams.copy()).to(dev)
    lst = [torch.zeros_like(c) for _ in range(world)]
    dist.all_gather(lst, c)
    assert all(torch.equal(x, lst[0]) for x in lst), "ranks diverged"
    pts = torch.zeros(full.n_points, 3, dtype=torch.float64, device=dev)
    pts[p0:p1] = torch.from_numpy(shard.points).to(dev)
    dist.all_reduce(pts)
    if big:
        st = eng.solve_stats()
        assert st["parts"] >= 2, st
        if rank == 0:
            one = sb.Engine(local)
            whole = sb.BAProblem(full.obs_cam.copy(), full.obs_point.copy(), full.obs_xy.copy(), full.points.copy(), full.cams.copy(), full.K.copy(), False, full.f0)
            ref = one.solve(whole, opt)
            assert one.solve_stats()["parts"] == st["parts"] and one.solve_stats()["ordered_n"] == st["ordered_n"]
            assert np.array_equal(rep.attempts[:, 2], ref.attempts[:, 2]) and rep.stop_reason == ref.stop_reason
            dev_ = np.abs(np.sqrt(rep.err_trace) - np.sqrt(ref.err_trace)) / np.sqrt(ref.err_trace)
            assert np.all(dev_ < 1e-9), dev_
            assert np.max(np.abs(pts.cpu().numpy() - whole.points)) / np.max(np.abs(whole.points)) < 1e-6
            assert np.max(np.abs(shard.cams - whole.cams)) < 1e-6
            print("multi-gpu parity ok (ring170, ordered solve, %d parts): world=%d, deviation from the single-GPU run %s" % (st["parts"], world, dev_))
            one.close()
    elif rank == 0:
        import oracle_lib as ol
        ol.build()
        op = ol.Problem(full.obs_cam, full.obs_point, full.obs_xy, full.points, full.cams, full.K, False, full.f0)
        ref = ol.ba_solve(op, err_change=1e-10, max_outer_iters=4, flow="sparse", solve="chol", acc="ld")
        fa = ol.ba_solve(op, err_change=1e-10, max_outer_iters=4, flow="sparse", solve="qr", acc="double")
        assert rep.seen_points == ref.seen_points == full.n_obs
        assert abs(rep.err_initial - ref.err_initial) <= 1e-12 * ref.err_initial
        assert np.array_equal(rep.attempts[:, 2], ref.attempts[:, 2])
        dev_ = np.abs(np.sqrt(rep.err_trace) - np.sqrt(ref.err_trace)) / np.sqrt(ref.err_trace)
        noise = np.abs(np.sqrt(fa.err_trace) - np.sqrt(ref.err_trace)) / np.sqrt(ref.err_trace)
        assert dev_[0] < 1e-9 and np.all(dev_ <= np.maximum(1e-9, noise)), (dev_, noise)
        assert np.max(np.abs(pts.cpu().numpy() - ref.points)) / np.max(np.abs(ref.points)) < 1e-6
        assert np.max(np.abs(shard.cams - ref.cams)) < 1e-6
        print("multi-gpu parity ok: world=%d, solver=%s, per-iteration deviation %s" % (world, "pcg" if rep.solver_used == sb.SOLVER_BLOCK_PCG else "cholesky", dev_))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
