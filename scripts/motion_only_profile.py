#!/usr/bin/env python
"""Motion-only BA (reference src/slam.cpp:578-675: one free camera, every line constant), the per-frame call:
end-to-end latency through the C ABI against the single-thread oracle."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from slslam_b200 import capi, synth
from oracle import oracle

for nl in (60, 200):
    w = synth.motion_only_window(11, num_lines=nl)
    for _ in range(5):
        capi.lba_solve(w, max_iters=10)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); p, s = capi.lba_solve(w, max_iters=10); ts.append(time.perf_counter() - t0)
    split = capi.last_timings()
    tc = []
    for _ in range(20):
        t0 = time.perf_counter(); po, so = oracle.lba_solve(w, max_iters=10, solver=1); tc.append(time.perf_counter() - t0)
    print(f"motion-only BA, {w.num_lines} lines / {w.num_observations} obs: GPU e2e {np.median(ts)*1e3:.3f} ms ({s['iterations']} it, "
          f"kernel {split['dev_kernel_ms']:.3f} ms, plan {split['plan_ms']:.3f} ms)   oracle 1 thread {np.median(tc)*1e3:.3f} ms ({so['iterations']} it)   "
          f"final cost rel diff {abs(s['final_cost']-so['final_cost'])/so['final_cost']:.1e}")
