"""The reference's file formats around the LBA path: observation files as SLAM::grab_new_frame
reads them (reference src/slam.cpp:62-135), trajectory files as SLAM::save_trajectory writes them (:1473-1496)."""
import os

import numpy as np

from slslam_b200 import dataset_io, replay, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_observation_file_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    obs = {7: rng.uniform(0, 480, 8), 3: rng.uniform(0, 640, 8), 120: rng.uniform(0, 480, 8)}
    path = dataset_io.write_frame_observations(str(tmp_path), 12, obs)
    assert os.path.basename(path) == "0012.txt"
    first = open(path).readline().split()
    assert len(first) == 10 and first[0] == "3"                       # id, 8 pixel coordinates, trailing token
    back = dataset_io.read_frame_observations(str(tmp_path), 12)
    assert sorted(back) == [3, 7, 120]
    for k, v in obs.items():
        want = np.array([(v[0] - synth.CX) / synth.FOCAL, (v[1] - synth.CY) / synth.FOCAL] * 4)
        want[2:4] = [(v[2] - synth.CX) / synth.FOCAL, (v[3] - synth.CY) / synth.FOCAL]
        want[4:6] = [(v[4] - synth.CX) / synth.FOCAL, (v[5] - synth.CY) / synth.FOCAL]
        want[6:8] = [(v[6] - synth.CX) / synth.FOCAL, (v[7] - synth.CY) / synth.FOCAL]
        assert np.abs(back[k] - want).max() < 1e-15
        assert np.abs(dataset_io.to_pixels(back[k]) - v).max() < 1e-10
    assert dataset_io.read_frame_observations(str(tmp_path), 13) is None   # grab_new_frame: missing file ends the run


def test_trajectory_file_round_trip(tmp_path):
    traj = np.load(os.path.join(GOLD, "traj_it3f_wolc.npy"))
    path = str(tmp_path / "traj.txt")
    dataset_io.write_trajectory(path, traj)
    cols = open(path).readline().rstrip("\n").split("\t")
    assert len(cols) == 7 and cols[0] == "0"
    raw = np.loadtxt(path)
    assert np.allclose(raw[:, 1], traj[:, 5]) and np.allclose(raw[:, 2], -traj[:, 3]) and np.allclose(raw[:, 3], -traj[:, 4])
    assert np.array_equal(dataset_io.read_trajectory(path), traj)


def test_replay_from_exported_dataset_matches_in_memory(tmp_path):
    """Exported observation files read back through the reference's reader give the same windows as the in-memory
    generator (to the last bit of the normalisation round trip)."""
    traj = np.load(os.path.join(GOLD, "traj_it3f_wolc.npy"))
    kw = dict(max_keyframes=8, sigma_px=0.2, seed=3, lines_per_kf=12)
    n = replay.export_dataset(str(tmp_path), traj, **kw)
    assert n == 8 and os.path.exists(dataset_io.frame_path(str(tmp_path), 7))

    def no_solve(w, it):
        return w.parameters.copy(), dict(initial_cost=0.0, final_cost=0.0, iterations=0)

    wa, wb = [], []
    replay.run(traj, no_solve, record=wa, odo_noise=(5e-3, 5e-2), **kw)
    replay.run(traj, no_solve, record=wb, odo_noise=(5e-3, 5e-2), obs_dir=str(tmp_path), **kw)
    assert len(wa) == len(wb) > 0
    for a, b in zip(wa, wb):
        assert np.array_equal(a.camera_index, b.camera_index) and np.array_equal(a.line_index, b.line_index)
        assert np.abs(a.observations - b.observations).max() < 1e-14
        assert np.abs(a.parameters - b.parameters).max() < 1e-9


def test_replay_motion_only_windows_have_the_reference_shape():
    """The per-keyframe motion-only step builds what SLAM::motion_only_ba packs (reference src/slam.cpp:578-640): two
    cameras (camera 1 = identity, constant), every line constant, two observations per line.  CPU oracle as the solver."""
    from oracle import oracle
    traj = np.load(os.path.join(GOLD, "traj_it3f_wolc.npy"))
    seen = []

    def cpu(w, it):
        return oracle.lba_solve(w, max_iters=it, solver=1)

    def cpu_motion_only(w, it):
        seen.append(w)
        return oracle.lba_solve(w, max_iters=it, solver=1)

    est, stats = replay.run(traj, cpu, motion_only=cpu_motion_only, max_keyframes=6, sigma_px=0.2, seed=5,
                            odo_noise=(5e-3, 5e-2), lines_per_kf=16, max_iters=6)
    assert len(seen) >= 3 and len(stats) == 5
    for w in seen:
        assert w.num_cameras == 2 and w.num_observations == 2 * w.num_lines >= 12
        assert np.array_equal(w.parameters[6:12], np.zeros(6))                     # camera 1: identity
        fx = w.fixed_index.reshape(-1, 2)
        assert (fx[:, 1] == 1).all()                                               # every line constant
        assert (fx[w.camera_index == 1, 0] == 1).all() and (fx[w.camera_index == 0, 0] == 0).all()
        p, s = oracle.lba_solve(w, max_iters=6, solver=1)
        assert s["final_cost"] < s["initial_cost"] and s["fixed_cost"] > 0
        assert np.array_equal(p[6:], w.parameters[6:])                             # only camera 0 moves
