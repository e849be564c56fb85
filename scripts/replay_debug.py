import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from slslam_b200 import capi, replay
from oracle import oracle
traj = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "traj_it3f_wolc.npy"))
kw = dict(max_keyframes=36, sigma_px=0.2, seed=3, odo_noise=(5e-3, 5e-2), lines_per_kf=24)
rec = []
est_g, st_g = replay.run(traj, lambda w, it: capi.lba_solve(w, max_iters=it), record=rec, **kw)
for w, sg in zip(rec, st_g):
    pc, sc = oracle.lba_solve(w, max_iters=10, solver=1)
    pg, s2 = capi.lba_solve(w, max_iters=10)
    rel = abs(s2["final_cost"] - sc["final_cost"]) / sc["final_cost"]
    dp = np.abs(pg[:6 * w.num_cameras] - pc[:6 * w.num_cameras]).max()
    print(f"kf {w.meta['keyframe']:3d} C={w.num_cameras:2d} L={w.num_lines:4d} N={w.num_observations:5d} gpu it={s2['iterations']} {s2['termination'][:8]} "
          f"cpu it={sc['iterations']} {sc['termination'][:8]} init {s2['initial_cost']:.6e}/{sc['initial_cost']:.6e} final {s2['final_cost']:.9e}/{sc['final_cost']:.9e} rel {rel:.1e} dpose {dp:.1e}")
