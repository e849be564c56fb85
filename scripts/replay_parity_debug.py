#!/usr/bin/env python
"""For every window the it3f replay assembles (tests/test_replay_gpu.py): GPU trace against oracle trace -- where do the
cost sequences first differ, do the accept / reject decisions agree, is the window gauge-free?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from slslam_b200 import capi, replay
from oracle import oracle

traj = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "traj_it3f_wolc.npy"))
kw = dict(max_keyframes=36, sigma_px=0.2, seed=3, odo_noise=(5e-3, 5e-2), lines_per_kf=24, max_iters=10)
windows = []
est, st = replay.run(traj, lambda w, it: capi.lba_solve(w, max_iters=it), record=windows, **kw)
for i, w in enumerate(windows):
    b = capi.LbaBatch([w], max_iters=10)
    b.solve()
    (pg,), (sg,) = b.download(trace=True)
    b.close()
    po, so = oracle.lba_solve(w, max_iters=10, solver=1)
    tg, to = sg["trace"], so["trace"]
    n = min(sg["iterations"], so["iterations"])
    first = next((k for k in range(n) if abs(tg[k, 0] - to[k, 0]) > 1e-9 * abs(to[k, 0])), -1)
    dec = next((k for k in range(n) if tg[k, 5] != to[k, 5]), -1)
    fixed = int(np.any(w.fixed_index.reshape(-1, 2)[:, 0] != 0))
    rel = abs(sg["final_cost"] - so["final_cost"]) / so["final_cost"]
    relk = [abs(tg[k, 0] - to[k, 0]) / abs(to[k, 0]) for k in range(n)]
    print(f"win {i:2d} C={w.num_cameras:2d} N={w.num_observations:4d} anchored={fixed} it={sg['iterations']}/{so['iterations']} "
          f"rel_final={rel:.2e} first_cost_diff_iter={first} first_decision_diff={dec} radius_last={tg[n-1,3]:.2e} "
          f"rel_by_iter={' '.join(f'{r:.0e}' for r in relk)}")
