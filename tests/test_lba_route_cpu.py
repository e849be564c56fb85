"""Which kernel a window goes to (slslam_lba_route: host logic of the product, no device needed): the default
--ba_window_size 10 windows to the tiled kernel, the per-frame motion-only problem to its own kernel, the reference's
--ba_window_size 20 / 40 shapes (reference src/slam.cpp:1376-1382, matlab_script/result_comp_ancdir_orthonorm/*basize{20,40}*)
to the general kernel, and what is beyond every kernel to an error."""
import numpy as np
import pytest

from slslam_b200 import capi, synth


def test_routes_of_the_reference_window_shapes():
    assert capi.lba_route(synth.window_S(0)) == "tiled"
    assert capi.lba_route(synth.window_M(1)) == "tiled"
    assert capi.lba_route(synth.make_window(2, 10, 80, 600, num_fixed_cameras=10)) == "tiled"           # steady state of W = 10: 20 cameras
    assert capi.lba_route(synth.motion_only_window(3, num_lines=60)) == "motion_only"
    assert capi.lba_route(synth.make_window(4, 20, 100, 1500, num_fixed_cameras=20)) == "general"       # W = 20: 40 camera blocks
    assert capi.lba_route(synth.make_window(5, 40, 100, 3000, num_fixed_cameras=40)) == "general"       # W = 40: 80 camera blocks
    assert capi.lba_route(synth.make_window(6, 28, 100, 1500)) == "general"                             # 28 free cameras > 24
    lim = capi.Limits()
    capi.lib().slslam_lba_get_limits(__import__("ctypes").byref(lim))
    assert lim.max_cameras == 32 and lim.max_free_cameras == 24 and lim.max_free_cameras_general == 64


def test_a_line_with_more_than_32_observations_needs_the_general_kernel():
    # the house simulation: every camera sees every line
    from slslam_b200 import replay
    S = synth.house_segments()
    P, Q = np.stack([a for a, _ in S]), np.stack([b for _, b in S])
    traj = synth.house_trajectory()
    windows = []
    replay.run(traj, lambda w_, it_: (w_.parameters.copy(), dict(iterations=0, initial_cost=0.0, final_cost=0.0, termination="NO_CONVERGENCE",
                                                                  num_successful_steps=0, num_unsuccessful_steps=0)),
               window_size=20, max_iters=1, sigma_px=0.2, seed=1, max_keyframes=44, scene=(P, Q), odo_noise=(5e-4, 2e-3), record=windows)
    w = windows[-1]
    assert w.num_cameras == 40 and np.bincount(w.line_index).max() > 32
    assert capi.lba_route(w) == "general"
    assert capi.lba_route(windows[5]) == "tiled"                                # the first windows of a run are small


def test_beyond_every_kernel_and_bad_input():
    w = synth.make_window(7, 70, 60, 1200)                                      # 70 free cameras
    with pytest.raises(capi.SlslamError) as e:
        capi.lba_route(w)
    assert e.value.code == -2
    w = synth.window_S(8)
    w.camera_index = w.camera_index.copy()
    w.camera_index[3] = 99
    with pytest.raises(capi.SlslamError) as e:
        capi.lba_route(w)
    assert e.value.code == -1
