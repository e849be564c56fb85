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
    ri = abs(s2["initial_cost"] - sc["initial_cost"]) / sc["initial_cost"]
    rs = abs(sg["initial_cost"] - sc["initial_cost"]) / sc["initial_cost"]
    if ri > 1e-12 or rs > 1e-12 or s2["initial_cost"] != sg["initial_cost"]:
        print(f"   INIT kf {w.meta['keyframe']}: re-solve vs oracle {ri:.2e}; replay-solve vs oracle {rs:.2e}; replay==resolve {s2['initial_cost'] == sg['initial_cost']}")
    dp = np.abs(pg[:6 * w.num_cameras] - pc[:6 * w.num_cameras]).max()
    print(f"kf {w.meta['keyframe']:3d} C={w.num_cameras:2d} L={w.num_lines:4d} N={w.num_observations:5d} gpu it={s2['iterations']} {s2['termination'][:8]} "
          f"cpu it={sc['iterations']} {sc['termination'][:8]} init {s2['initial_cost']:.6e}/{sc['initial_cost']:.6e} final {s2['final_cost']:.9e}/{sc['final_cost']:.9e} rel {rel:.1e} dpose {dp:.1e}")

# determinism stress: the same window many times, bitwise
w = rec[18]
ref, sref = capi.lba_solve(w, max_iters=10)
bad = 0
for k in range(200):
    p, s = capi.lba_solve(w, max_iters=10)
    if not np.array_equal(p, ref) or s["final_cost"] != sref["final_cost"] or s["initial_cost"] != sref["initial_cost"]:
        bad += 1
print("determinism stress: mismatches", bad, "of 200")
