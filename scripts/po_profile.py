#!/usr/bin/env python
"""Pose-graph solve on the myungdong-scale synthetic graph: device time of the LM loop, end-to-end time, oracle parity."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from slslam_b200 import capi, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = synth.make_pose_graph(0)
for _ in range(reps):
    t0 = time.perf_counter()
    p, s = capi.po_solve(g, max_iters=10)
    wall = time.perf_counter() - t0
    ms = capi.lib().slslam_po_last_solve_ms()
    print(f"PO K={g.num_poses} E={g.num_edges}: iterations {s['iterations']} final {s['final_cost']:.9e} term {s['termination']} "
          f"device {ms:.3f} ms ({ms / max(1, s['iterations']):.3f} ms/iter) wall {wall * 1e3:.1f} ms", flush=True)
    print("   ", capi.po_last_stats(), flush=True)
