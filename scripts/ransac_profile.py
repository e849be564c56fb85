#!/usr/bin/env python
"""RANSAC hypothesis scoring (reference src/slam.cpp:398-412): all hypotheses x all lines in one launch through the C ABI
(host buffers in, scores + inlier masks out) against the single-thread oracle."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from slslam_b200 import capi
from oracle import oracle
from test_ransac import make_case, BASELINE, THR

for n_lines, n_hyp in ((150, 150), (300, 1000), (1000, 5000)):
    poses, lines, obs, _ = make_case(0, n_lines, n_hyp)
    for _ in range(3):
        capi.ransac_score(poses, lines, obs, BASELINE, THR, want_errors=False)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); sg, ig, _ = capi.ransac_score(poses, lines, obs, BASELINE, THR, want_errors=False); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); so, io, _ = oracle.ransac_score(poses, lines, obs, BASELINE, THR); tc = time.perf_counter() - t0
    assert np.array_equal(sg, so) and np.array_equal(ig, io)
    print(f"RANSAC scoring, {n_hyp} hypotheses x {n_lines} lines: GPU end to end {np.median(ts)*1e3:.3f} ms "
          f"({n_hyp*n_lines/np.median(ts)/1e6:.0f} M line tests/s) | oracle 1 thread {tc*1e3:.3f} ms ({n_hyp*n_lines/tc/1e6:.1f} M/s) | "
          f"scores and inlier masks identical", flush=True)
