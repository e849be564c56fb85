#!/usr/bin/env python
"""Small invocations of every kernel, for compute-sanitizer (memcheck / racecheck / synccheck) under gpurun."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from slslam_b200 import capi, synth

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "lba"):
    w = synth.make_window(0, 6, 60, 260, sigma_px=0.5)
    p, s = capi.lba_solve(w, max_iters=3)                       # plan kernel + gather + solve kernel (group of CTAs)
    print("lba", s["iterations"], s["final_cost"])
    ws = [synth.make_window(i, 5, 40, 150) for i in range(3)]
    ps, ss = capi.lba_solve_batch(ws, max_iters=2)
    print("lba batch", [x["final_cost"] for x in ss])
if which in ("all", "moba"):
    p, s = capi.lba_solve(synth.motion_only_window(1, num_lines=40), max_iters=4)
    print("moba", s["iterations"], s["final_cost"])
if which in ("all", "ransac"):
    from test_ransac import make_case
    poses, lines, obs, _ = make_case(0, 70, 9)
    print("ransac", capi.ransac_score(poses, lines, obs)[0])
if which in ("all", "po"):
    g = synth.make_pose_graph(0, num_poses=16, neighbours=2, num_loops=2)
    p, s = capi.po_solve(g, max_iters=3)
    print("po", s["final_cost"])
