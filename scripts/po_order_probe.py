import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["SLSLAM_PO_DEBUG"] = "1"
import numpy as np
from slslam_b200 import capi, synth
traj = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "traj_myungdong_wolc.npy"))
for g in (synth.pose_graph_from_trajectory(traj, seed=0, num_loops=10), synth.make_pose_graph(0)):
    try:
        capi.po_solve(g, max_iters=1)
    except Exception as e:
        print("solve:", str(e)[:80])
