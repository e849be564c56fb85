#!/usr/bin/env python
"""PO: level order + warp-per-column kernel against minimum-degree order + column-at-a-time kernel, on the bench's graph
(myungdong trajectory, 10 loop closures) and on a band graph."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from slslam_b200 import capi, synth
traj = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "traj_myungdong_wolc.npy"))
graphs = {"myungdong + 10 loops": synth.pose_graph_from_trajectory(traj, seed=0, num_loops=10),
          "261 poses, 3 neighbours, 3 loops": synth.make_pose_graph(0)}
for name, g in graphs.items():
    for mode in ("levels", "columns"):
        if mode == "columns":
            os.environ["SLSLAM_PO_COLUMNS"] = "1"
        for _ in range(3):
            p, s = capi.po_solve(g, max_iters=10)
        ms = []
        for _ in range(10):
            p, s = capi.po_solve(g, max_iters=10)
            ms.append(float(capi.lib().slslam_po_last_solve_ms()))
        st = capi.po_last_stats()
        os.environ.pop("SLSLAM_PO_COLUMNS", None)
        print(f"{name} [{mode}]: {s['iterations']} iterations, device {np.median(ms):.3f} ms ({np.median(ms)/s['iterations']:.3f} ms/iter), "
              f"blocks {st['factor_blocks']}, updates {st['block_updates']}, max rows {st['max_column_rows']}, cycles {st['factor_cycles']}, cost {s['final_cost']:.12e}")
