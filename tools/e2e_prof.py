import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import surikatoko_b200 as sb
from surikatoko_b200 import scenes
prob = scenes.ring_scene(1000, 1_000_000, 10, seed=1234)
for attr in ("obs_cam", "obs_point", "obs_xy", "points", "cams", "K"):
    t = torch.from_numpy(np.ascontiguousarray(getattr(prob, attr))).pin_memory(); setattr(prob, "_keep_" + attr, t); setattr(prob, attr, t.numpy())
pts0, cams0 = prob.points.copy(), prob.cams.copy()
eng = sb.Engine(0)
opt = sb.BAOptions(max_outer_iters=1)
for i in range(4):
    prob.points[:] = pts0; prob.cams[:] = cams0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ok = eng.bind(prob, opt); torch.cuda.synchronize(); t1 = time.perf_counter()
    rep = eng.run(opt); torch.cuda.synchronize(); t2 = time.perf_counter()
    eng.fetch(prob); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("bind %.2f run %.2f fetch %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
for i in range(3):
    prob.points[:] = pts0; prob.cams[:] = cams0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rep = eng.solve(prob, opt); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("solve %.2f ms" % ((t1 - t0) * 1e3))
