import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
from slslam_b200 import capi, synth, replay
S = synth.house_segments()
P, Q = np.stack([a for a, _ in S]), np.stack([b for _, b in S])
traj = synth.house_trajectory()
for W in (20, 40):
    windows = []
    replay.run(traj, lambda w_, it_: capi.lba_solve(w_, max_iters=it_), window_size=W, max_iters=10, sigma_px=0.2, seed=1,
               max_keyframes=2 * W + 6, scene=(P, Q), odo_noise=(5e-4, 2e-3), record=windows)
    w = windows[-1]
    print("=== W", W, w.num_cameras, w.num_observations, flush=True)
    for _ in range(2):
        t0 = time.perf_counter(); p, s = capi.lba_solve(w, max_iters=10); print("ms", 1e3 * (time.perf_counter() - t0), s["iterations"], flush=True)
