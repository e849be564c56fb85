"""BASELINE.json configs[2] and [4] substitutes: sliding-window LBA replay around the reference's it3f
output trajectory, and pose-graph optimisation on the myungdong output trajectory with synthetic loop closures.
The committed fixtures come from tests/golden/make_trajectories.py."""
import os

import numpy as np
import pytest

from slslam_b200 import replay, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_it3f_sliding_window_replay_matches_oracle(gpu):
    from oracle import oracle
    traj = np.load(os.path.join(GOLD, "traj_it3f_wolc.npy"))
    kw = dict(max_keyframes=36, sigma_px=0.2, seed=3, odo_noise=(5e-3, 5e-2), lines_per_kf=24, max_iters=10)

    def gpu_solve(w, it):
        return gpu.lba_solve(w, max_iters=it)

    def cpu_solve(w, it):
        return oracle.lba_solve(w, max_iters=it, solver=1)

    def no_solve(w, it):
        return w.parameters.copy(), dict(initial_cost=0.0, final_cost=0.0, iterations=0)

    windows = []
    est_g, st_g = replay.run(traj, gpu_solve, record=windows, **kw)
    est_c, st_c = replay.run(traj, cpu_solve, **kw)
    est_0, _ = replay.run(traj, no_solve, **kw)
    rmse_g, rmse_c, rmse_0 = (replay.trajectory_rmse(e, traj) for e in (est_g, est_c, est_0))
    # LBA must beat dead reckoning clearly
    assert rmse_g < 0.25 * rmse_0, (rmse_g, rmse_0)
    # (1) parity on IDENTICAL inputs: every window the GPU replay assembled, solved again by the oracle.  With the
    # reference's max_num_iterations = 10 nearly every window stops mid-descent, and for some windows (the first, 2-camera
    # one above all) the 10-iteration LM map amplifies rounding noise: no accept / reject decision differs, the gain
    # ratio's effect on the radius compounds a 1e-12 difference by a decade or two per iteration
    # (scripts/replay_parity_debug.py prints the traces).  So: 1e-6 wherever the window's result is defined that well,
    # and otherwise within 10x the width the ORACLE ITSELF shows under a one-ulp perturbation of its input.
    from test_lba_gpu import _path_sensitivity
    rel, loose = [], []
    for i, (w, sg) in enumerate(zip(windows, st_g)):
        _, so = oracle.lba_solve(w, max_iters=10, solver=1)
        assert abs(sg["initial_cost"] - so["initial_cost"]) <= 1e-11 * so["initial_cost"]
        assert sg["iterations"] == so["iterations"]
        d = abs(sg["final_cost"] - so["final_cost"]) / so["final_cost"]
        rel.append(d)
        if d > 1e-6:
            sens = _path_sensitivity(w, 10, reps=4)
            loose.append((i, w.num_cameras, d, sens["final"]))
            assert d <= 10.0 * sens["final"], loose[-1]
    print("windows above 1e-6 (index, cameras, |gpu - oracle| / cost, oracle one-ulp width):", loose)
    assert np.median(rel) < 1e-9 and len(loose) <= 2, (np.median(rel), loose)
    # (2) the two replays, each feeding its own write-back into the next window.  Because the windows stop unconverged,
    # the outcome of a window depends on its last accept / reject decisions and a 1e-8 difference is carried and
    # amplified through 35 dependent windows: these are two equally valid LBA runs, compared as such (both cut the
    # dead-reckoning error by > 20x and agree to centimetres), while (1) is the parity statement.
    assert max(rmse_g, rmse_c) < 0.05 * rmse_0 and abs(rmse_g - rmse_c) < 0.6 * rmse_c, (rmse_g, rmse_c, rmse_0)
    assert np.abs(est_g[:, 3:] - est_c[:, 3:]).max() < 5e-2          # metres
    assert np.abs(est_g[:, :3] - est_c[:, :3]).max() < 1e-2          # radians
    assert [a["observations"] for a in st_g] == [b["observations"] for b in st_c]
    # steady-state windows have the reference's shape: W free + W constant cameras
    assert st_g[-1]["cameras"] == 20


def test_myungdong_pose_graph(gpu):
    from oracle import oracle
    traj = np.load(os.path.join(GOLD, "traj_myungdong_wolc.npy"))
    g = synth.pose_graph_from_trajectory(traj, seed=0, num_loops=10)
    assert g.num_poses == 253
    pg, sg = gpu.po_solve(g, max_iters=10)
    po, so = oracle.po_solve(g, max_iters=10)
    assert abs(sg["final_cost"] - so["final_cost"]) < 1e-6 * so["final_cost"]
    assert sg["iterations"] == so["iterations"] and sg["termination"] == so["termination"]
    assert np.abs(pg - po).max() < 1e-5
    assert sg["final_cost"] < 0.5 * sg["initial_cost"]


def test_replay_with_motion_only_ba_per_keyframe(gpu):
    """The reference's per-frame step in front of the window solve (src/slam.cpp:578-675): every incoming keyframe is
    first refined by motion-only BA against the current map -- through the dedicated kernel -- then enters the LBA
    window.  Motion-only windows are checked against the oracle one by one.  In this synthetic replay the odometry
    stand-in is the TRUE relative motion plus noise, so replacing it by a map-based pose removes the only absolute scale
    reference besides the 12 cm stereo baseline: the trajectory drifts in scale (as the real pipeline does) and is only
    required to stay better than dead reckoning, while every single motion-only solve must cut its cost."""
    from oracle import oracle
    traj = np.load(os.path.join(GOLD, "traj_it3f_wolc.npy"))
    kw = dict(max_keyframes=16, sigma_px=0.2, seed=5, odo_noise=(5e-3, 5e-2), lines_per_kf=24, max_iters=10)
    seen = []

    def gpu_solve(w, it):
        return gpu.lba_solve(w, max_iters=it)

    def gpu_motion_only(w, it):
        p, s = gpu.lba_solve(w, max_iters=it)
        assert gpu.last_timings()["plan_ms"] == 0.0            # the motion-only fast path has no plan stage
        seen.append((w, p, s))
        return p, s

    def no_solve(w, it):
        return w.parameters.copy(), dict(initial_cost=0.0, final_cost=0.0, iterations=0)

    est, st = replay.run(traj, gpu_solve, motion_only=gpu_motion_only, **kw)
    est_0, _ = replay.run(traj, no_solve, **kw)
    assert len(seen) >= 10
    for w, p, s in seen:
        po, so = oracle.lba_solve(w, max_iters=10, solver=1)
        assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
        assert np.abs(p[:6] - po[:6]).max() < 1e-7
        assert s["final_cost"] < 0.9 * s["initial_cost"]
    assert replay.trajectory_rmse(est, traj) < 0.8 * replay.trajectory_rmse(est_0, traj)
