"""The reference's house simulation (SURVEY.md §8c anchor 4): the only numbers the reference ships for the LBA path are
the 80 end-of-run summaries under matlab_script/result_comp_ancdir_orthonorm/ (mean iterations, mean initial and final
LBA cost per frame for pixel noise 0.2 .. 1.0, windows 5 .. 40), measured on the 74-segment model of matlab_script/house.m.
The simulator and its ground-truth trajectory are not shipped, so the run is rebuilt here: the same line model
(synth.house_segments), a trajectory circling it (synth.house_trajectory), the sliding-window driver (replay.run).
What can be compared without the original trajectory is what does not depend on it:
  * the cost per window is  1/2 (4 N - p) (sigma / 406.05)^2  in the quadratic regime (N observation blocks, p free
    parameters): checked at sigma = 0.2 px;
  * how the mean final cost grows with sigma -- 1 : 8.17 : 12.80 : 17.60 for 0.2 : 0.6 : 0.8 : 1.0 px in the reference's
    table, NOT the 1 : 9 : 16 : 25 of a plain least-squares cost: that curve is the signature of HuberLoss(1 / 406.05)
    applied to the squared norm of the 4-vector of a residual block (reference src/lba_problem.cpp:78-80), and it pins
    exactly the loss semantics the oracle and the kernels restate;
  * the ratio between our cost level and the reference's is then one constant for every sigma (the reference's windows
    hold ~1.5 x more observation blocks than 2 W x 74; its simulator is not available to say why).
The fixture tests/golden/house_ba_results.json is produced from the reference's files by tests/golden/make_house_results.py."""
import json
import os

import numpy as np
import pytest

from slslam_b200 import replay, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _reference_table(window, max_iter=10):
    recs = json.load(open(os.path.join(GOLD, "house_ba_results.json")))
    return {r["sigma_px"]: r for r in recs if r["parameterisation"] == "orthonorm" and r["window"] == window and r["max_iter"] == max_iter}


def _house_run(solve, sigma_px, window=10, keyframes=90, steady_from=45, seed=1, max_iters=10, record=None):
    S = synth.house_segments()
    P, Q = np.stack([a for a, _ in S]), np.stack([b for _, b in S])
    traj = synth.house_trajectory()
    est, st = replay.run(traj, solve, window_size=window, max_iters=max_iters, sigma_px=sigma_px, seed=seed, max_keyframes=keyframes,
                         scene=(P, Q), odo_noise=(5e-4, 2e-3), record=record)
    ss = [s for s in st if s["keyframe"] >= steady_from]
    return dict(final=float(np.mean([s["final_cost"] for s in ss])), initial=float(np.mean([s["initial_cost"] for s in ss])),
                iterations=float(np.mean([s["iterations"] for s in ss])), observations=float(np.mean([s["observations"] for s in ss])),
                cameras=ss[-1]["cameras"], lines=ss[-1]["lines"], windows=len(ss))


def test_house_model_is_the_references():
    S = synth.house_segments()
    assert len(S) == 74
    pts = np.concatenate([np.stack([a for a, _ in S]), np.stack([b for _, b in S])])
    assert np.allclose(pts.min(0), [0, 0, 0]) and np.allclose(pts.max(0), [4.5, 4.5, 3.5])      # house.m:20-22
    assert np.allclose(S[12][0], [0, 2.25, 3.5]) and np.allclose(S[12][1], [4.5, 2.25, 3.5])   # the ridge, house.m:45
    assert np.allclose(S[0][1], [0, 0, 0.65 * 3.5])                                            # wall height r h, house.m:30
    assert all(np.linalg.norm(b - a) > 0.1 for a, b in S)


def test_house_simulation_reproduces_the_published_noise_curve():
    """CPU (oracle).  W = 10, max 10 iterations, sigma 0.2 .. 1.0 px, statistics over the steady-state windows."""
    from oracle import oracle
    ref = _reference_table(10)

    def solve(w, it):
        return oracle.lba_solve(w, max_iters=it, solver=1)

    runs = {s: _house_run(solve, s) for s in (0.2, 0.6, 0.8, 1.0)}
    r02 = runs[0.2]
    assert r02["cameras"] == 20 and r02["lines"] >= 70
    # quadratic regime: the cost is what the noise puts in, minus the degrees of freedom the fit absorbs
    theory = 0.5 * (4 * r02["observations"] - (6 * 10 + 4 * r02["lines"])) * (0.2 / 406.05) ** 2
    assert abs(r02["final"] - theory) < 0.05 * theory, (r02, theory)
    # the Huber signature: growth of the mean final cost with sigma, against the reference's own table
    for s in (0.6, 0.8, 1.0):
        ours = runs[s]["final"] / r02["final"]
        theirs = ref[s]["mean_final_cost"] / ref[0.2]["mean_final_cost"]
        assert abs(ours - theirs) < 0.08 * theirs, (s, ours, theirs)
    # ... clearly below the sigma^2 law of an un-robustified cost where the noise reaches the Huber threshold of 1 px
    assert runs[1.0]["final"] / r02["final"] < 0.8 * 25.0 and runs[0.8]["final"] / r02["final"] < 0.88 * 16.0
    # one constant between the two cost levels at every sigma (a difference in observation count, not in the noise model)
    k = [ref[s]["mean_final_cost"] / runs[s]["final"] for s in (0.2, 0.6, 0.8, 1.0)]
    assert max(k) / min(k) < 1.08 and 1.2 < np.mean(k) < 2.0, k
    # LM needs more iterations as the noise grows, as in the reference's table (2.2 -> 5.3 per frame there)
    assert runs[1.0]["iterations"] > runs[0.2]["iterations"]
    assert ref[1.0]["mean_iterations"] > ref[0.2]["mean_iterations"]
    # the un-robustified cost follows sigma^2 (the curve above is the loss, not the generator)
    def solve_l2(w, it):
        return oracle.lba_solve(w, max_iters=it, solver=1, robust=False)
    l2 = {s: _house_run(solve_l2, s, keyframes=70, steady_from=40) for s in (0.2, 1.0)}
    assert 20.0 < l2[1.0]["final"] / l2[0.2]["final"] < 28.0 and runs[1.0]["final"] / r02["final"] < 19.0


@pytest.mark.gpu
@pytest.mark.parametrize("window,keyframes", [(10, 60), (20, 70), (40, 100)])
def test_house_windows_on_the_gpu(gpu, window, keyframes):
    """The reference's --ba_window_size 10 / 20 / 40 runs (ba_result_*_basize{10,20,40}_*): 2 W cameras per window, W of
    them free, every line seen by nearly every camera.  W = 10 fits the tiled kernel; W = 20 and 40 (40 / 80 camera
    blocks, lines with up to 80 observations) go through the general kernel.  Every window the replay assembles is solved
    again by the oracle: same iteration count, termination and steps, final cost rel 1e-6, poses 1e-6; and the steady
    state sits at the cost level the noise dictates, growing with W as in the reference's table."""
    from oracle import oracle
    windows = []

    def solve(w, it):
        return gpu.lba_solve(w, max_iters=it)

    run = _house_run(solve, 0.2, window=window, keyframes=keyframes, steady_from=keyframes - 15, record=windows)
    assert run["cameras"] == 2 * window
    lim = gpu.Limits()
    gpu.lib().slslam_lba_get_limits(__import__("ctypes").byref(lim))
    assert (2 * window > lim.max_cameras) == (window > 10) and lim.max_free_cameras_general >= 40
    checked = 0
    for w in windows[-12:] + windows[:6]:
        p, s = gpu.lba_solve(w, max_iters=10)
        po, so = oracle.lba_solve(w, max_iters=10, solver=1)
        assert s["iterations"] == so["iterations"] and s["termination"] == so["termination"], (s, so)
        assert s["num_successful_steps"] == so["num_successful_steps"]
        assert abs(s["initial_cost"] - so["initial_cost"]) <= 1e-11 * so["initial_cost"]
        assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"], (w.num_cameras, s, so)
        C = w.num_cameras
        assert np.abs(p[:6 * C] - po[:6 * C]).max() < 1e-6
        checked += 1
    assert checked == 18
    theory = 0.5 * (4 * run["observations"] - (6 * window + 4 * run["lines"])) * (0.2 / 406.05) ** 2
    if window <= 20:
        assert abs(run["final"] - theory) < 0.08 * theory, (run, theory)
    else:   # 100 keyframes are barely more than one 80-camera window: the map has not settled yet, the level is only bracketed
        assert 0.9 * theory < run["final"] < 2.0 * theory, (run, theory)
    ref = _reference_table(window)[0.2]["mean_final_cost"] / _reference_table(10)[0.2]["mean_final_cost"]
    assert abs(ref - window / 10.0) < 0.06 * window / 10.0          # the reference's cost is linear in W, as N is
