#!/usr/bin/env python
"""One small invocation of every kernel family of the library (for an `ncu --set full` capture of the kernels beside
lba_solve_kernel): device planner + observation gather + tiled solve (S window), motion-only BA, general (wide) kernel,
RANSAC scoring, PO (linearise, block-sparse assemble / factor+solve, step, decide, ...), resident-map assembly and
write-back, geometry conversions, K1 evaluation."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from slslam_b200 import capi, synth, replay


def map_case():
    """resident map: a few keyframes seeing the same lines, one bundle adjustment"""
    dm = capi.DeviceMap(8, 256, 4096)
    ws = synth.make_window(5, 3, 60, 170, sigma_px=0.3)
    C = ws.num_cameras
    T = [np.concatenate([synth.rodrigues(ws.parameters[6 * c:6 * c + 3]).ravel(), ws.parameters[6 * c + 3:6 * c + 6]]) for c in range(C)]
    obs = ws.observations.reshape(-1, 8)
    for c in range(C):
        idx = np.flatnonzero(ws.camera_index == c)
        dm.add_keyframe(c, T[c], ws.line_index[idx], obs[idx])
    av = np.zeros((ws.num_lines, 6)); first = np.zeros(ws.num_lines, np.int32)
    for l in range(ws.num_lines):
        i0 = np.flatnonzero(ws.line_index == l)[0]; first[l] = ws.camera_index[i0]
        cp, dv = synth.orth_to_av(ws.parameters[6 * C + 4 * l:6 * C + 4 * l + 4])
        R, t = T[first[l]][:9].reshape(3, 3), T[first[l]][9:]
        av[l, :3] = R @ cp + t; av[l, 3:] = R @ dv
    dm.add_landmarks(np.arange(ws.num_lines), first, av)
    dm.bundle_adjust(list(range(C)), [C + 1] + list(range(1, C)), C, max_iters=3)
    dm.close()


if __name__ == "__main__":
    w = synth.window_S(0, sigma_px=0.5)
    capi.lba_solve(w, max_iters=4)
    capi.lba_evaluate(w)
    capi.lba_solve(synth.motion_only_window(1, num_lines=120), max_iters=6)
    capi.lba_solve(synth.make_window(77, 20, 120, 1500, num_fixed_cameras=16, sigma_px=0.5), max_iters=3)     # general kernel
    rng = np.random.default_rng(0)
    H, K = 256, 200
    poses = np.concatenate([np.tile(np.eye(3).ravel(), (H, 1)), rng.normal(0, 0.05, (H, 3))], axis=1)
    lines = np.concatenate([rng.normal(0, 1, (K, 3)) + [0, 0, 6], rng.normal(0, 1, (K, 3))], axis=1)
    capi.ransac_score(poses, lines, rng.normal(0, 0.2, (K, 8)))
    capi.po_solve(synth.make_pose_graph(0), max_iters=2)                       # level order: po_sp_factor_levels (a cluster)
    os.environ["SLSLAM_PO_COLUMNS"] = "1"
    capi.po_solve(synth.make_pose_graph(0), max_iters=1)                       # minimum-degree order: po_sp_factor_solve
    del os.environ["SLSLAM_PO_COLUMNS"]
    capi.geometry_convert(0, np.concatenate([rng.normal(0, 1, (64, 3)), rng.normal(0, 1, (64, 3))], axis=1))
    map_case()
    print("all kernel families launched")
