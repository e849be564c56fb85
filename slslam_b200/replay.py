"""Sliding-window LBA replay over a keyframe trajectory: the substitute SURVEY.md §8d names for BASELINE.json
configs[2] (the it3f dataset is not shipped with the reference; only its output trajectory is).

A synthetic line scene is laid around the given keyframe poses, stereo observations are generated per keyframe in the
reader's convention (normalised endpoints, reference src/slam.cpp:121-128), and for every new keyframe a window is
assembled the way SLAM::local_bundle_adjustment / SLAM::bundle_adjustment do (reference src/slam.cpp:1370-1386,
795-920): the newest W keyframes are free, up to W older ones are appended as constant cameras, only landmarks with at
least two observations in free keyframes enter (slam.cpp:839), the window frame is re-anchored so that the newest
keyframe is exactly identity (slam.cpp:1322), and the result is written back (slam.cpp:957-972).  New landmarks are
initialised by stereo triangulation of their first observation (the role of slam.cpp:190-219).

The solver is injected: `solve(window, max_iters) -> parameters`.  The product passes the C-ABI call; tests also run
the CPU oracle through the same driver to compare trajectories.  Pure numpy host logic; no solver lives here.
"""
from __future__ import annotations

import numpy as np

from . import synth
from .synth import BASELINE, CX, CY, FOCAL, HEIGHT, WIDTH, Window, av_to_orth, log_so3, orth_to_av, rodrigues


def pose_compose(T21, T10):
    R21, R10 = rodrigues(T21[:3]), rodrigues(T10[:3])
    return np.concatenate([log_so3(R21 @ R10), R21 @ T10[3:] + T21[3:]])


def pose_inverse(T):
    R = rodrigues(T[:3])
    return np.concatenate([-T[:3], -R.T @ T[3:]])


def line_transform(orth, T):
    """4-parameter line expressed in frame 0 -> the same line in frame 1, x1 = R x0 + t (gc_line_to_pose's role)."""
    cp, d = orth_to_av(orth)
    R = rodrigues(T[:3])
    p, dd = R @ cp + T[3:], R @ d
    p = p - dd * (p @ dd)
    return av_to_orth(p, dd)


def triangulate(ob, baseline=BASELINE):
    """Stereo observation (normalised x0 y0 x1 y1 | x2 y2 x3 y3) -> line in the camera frame as (closest point, dir):
    intersection of the two interpretation planes.  Returns None for a degenerate configuration."""
    nA = np.cross([ob[0], ob[1], 1.0], [ob[2], ob[3], 1.0])
    nB = np.cross([ob[4], ob[5], 1.0], [ob[6], ob[7], 1.0])
    d = np.cross(nA, nB)
    nd = np.linalg.norm(d)
    if nd < 1e-9 * np.linalg.norm(nA) * np.linalg.norm(nB):
        return None
    d = d / nd
    A = np.stack([nA, nB, d])
    try:
        X = np.linalg.solve(A, np.array([0.0, nB[0] * baseline, 0.0]))
    except np.linalg.LinAlgError:
        return None
    if not np.all(np.isfinite(X)) or X[2] < 0.5 or X[2] > 60.0:
        return None
    return X - d * (X @ d), d


def make_scene(poses_wc, seed=0, lines_per_kf=30):
    """Segments in front of every keyframe (depth 3-12 m), world frame.  poses_wc: camera->world (angle-axis, t) rows."""
    rng = np.random.default_rng(seed)
    P, Q = [], []
    for T in poses_wc:
        R = rodrigues(T[:3])
        for _ in range(lines_per_kf):
            z = rng.uniform(3.0, 12.0)
            mid = np.array([rng.uniform(-0.7, 0.7) * z, rng.uniform(-0.5, 0.5) * z, z])
            v = rng.normal(size=3)
            v = v / np.linalg.norm(v) * rng.uniform(0.5, 2.0) * 0.5
            P.append(R @ (mid - v) + T[3:]); Q.append(R @ (mid + v) + T[3:])
    return np.asarray(P), np.asarray(Q)


def observe(poses_cw, P, Q, sigma_px, seed, lines_per_kf, reach=14):
    """Per keyframe: {line id: 8 normalised coordinates}.  A line is observed when its four stereo endpoints fall in
    the image and lie 1-40 m ahead.  Only lines created within `reach` keyframes are tested (reach None: all lines)."""
    rng = np.random.default_rng(seed + 1)
    obs = []
    K = len(poses_cw)
    for k, T in enumerate(poses_cw):
        R = rodrigues(T[:3])
        lo, hi = (0, len(P)) if reach is None else (max(0, k - reach) * lines_per_kf, min(K, k + reach + 1) * lines_per_kf)
        cur = {}
        pa, qa = (R @ P[lo:hi].T).T + T[3:], (R @ Q[lo:hi].T).T + T[3:]
        for j in range(hi - lo):
            ends = [pa[j], qa[j], pa[j] - [BASELINE, 0, 0], qa[j] - [BASELINE, 0, 0]]
            if min(e[2] for e in ends) < 1.0 or max(e[2] for e in ends) > 40.0:
                continue
            px = np.array([[FOCAL * e[0] / e[2] + CX, FOCAL * e[1] / e[2] + CY] for e in ends])
            if px[:, 0].min() < 0 or px[:, 0].max() > WIDTH or px[:, 1].min() < 0 or px[:, 1].max() > HEIGHT:
                continue
            px = px + rng.normal(0, sigma_px, px.shape)
            cur[lo + j] = np.stack([(px[:, 0] - CX) / FOCAL, (px[:, 1] - CY) / FOCAL], 1).ravel()
        obs.append(cur)
    return obs


def export_dataset(obs_dir, poses_wc_true, sigma_px=0.5, seed=0, lines_per_kf=30, max_keyframes=None):
    """Writes the synthetic scene's stereo observations as the reference's per-frame files (`%04d.txt`, pixels;
    dataset_io.write_frame_observations), one file per keyframe.  Returns the number of files."""
    from . import dataset_io
    K = len(poses_wc_true) if max_keyframes is None else min(max_keyframes, len(poses_wc_true))
    truth_cw = np.stack([pose_inverse(T) for T in poses_wc_true[:K]])
    P, Q = make_scene(poses_wc_true[:K], seed, lines_per_kf)
    for k, cur in enumerate(observe(truth_cw, P, Q, sigma_px, seed, lines_per_kf)):
        dataset_io.write_frame_observations(obs_dir, k, {lid: dataset_io.to_pixels(ob) for lid, ob in cur.items()})
    return K


def run(poses_wc_true, solve, window_size=10, max_iters=10, sigma_px=0.5, seed=0, lines_per_kf=30,
        odo_noise=(2e-3, 2e-2), anchor_first=True, max_keyframes=None, record=None, obs_dir=None, motion_only=None, scene=None):
    """Replays the trajectory.  Returns the estimated camera->world poses [K][6] and per-window statistics.
    obs_dir: read the observations from the reference's per-frame files (export_dataset) instead of generating them.
    motion_only: optional solver for the per-frame motion-only BA (reference src/slam.cpp:578-675): before a keyframe
    enters the window its pose is refined against the current map with every line held constant."""
    K = len(poses_wc_true) if max_keyframes is None else min(max_keyframes, len(poses_wc_true))
    truth_cw = np.stack([pose_inverse(T) for T in poses_wc_true[:K]])
    if obs_dir is not None:
        from . import dataset_io
        obs = [dataset_io.read_frame_observations(obs_dir, k) for k in range(K)]
    elif scene is not None:
        # a given line model seen from every keyframe (the reference's house simulation: every segment tested in every view)
        P, Q = scene
        obs = observe(truth_cw, np.asarray(P), np.asarray(Q), sigma_px, seed, lines_per_kf, reach=None)
    else:
        P, Q = make_scene(poses_wc_true[:K], seed, lines_per_kf)
        obs = observe(truth_cw, P, Q, sigma_px, seed, lines_per_kf)
    rng = np.random.default_rng(seed + 2)
    est = np.zeros((K, 6))                  # world->camera estimates
    est[0] = truth_cw[0]
    lines = {}                              # line id -> orth parameters in the world frame
    refined = set()                         # lines written back by at least one LBA window
    stats = []
    W = window_size
    for k in range(K):
        if k > 0:                           # visual-odometry stand-in: true relative motion + noise
            rel = pose_compose(truth_cw[k], pose_inverse(truth_cw[k - 1]))
            rel = rel + np.concatenate([rng.normal(0, odo_noise[0], 3), rng.normal(0, odo_noise[1], 3)])
            est[k] = pose_compose(rel, est[k - 1])
        if k > 0 and motion_only is not None:
            # motion-only BA as SLAM::motion_only_ba packs it: camera 0 = this keyframe (free), camera 1 = identity
            # (constant), two observations per common line (this frame, previous keyframe), lines expressed in the
            # previous keyframe's frame and constant
            # (only map lines an LBA window has already refined: a line triangulated from one stereo pair has a depth
            # error of decimetres and, held constant, would drag the pose with it)
            common = [l for l in obs[k] if l in refined and l in obs[k - 1]]
            if len(common) >= 6:
                Tp = est[k - 1]
                ci, li, fi, ob_arr = [], [], [], []
                for j, l in enumerate(common):
                    ci += [0, 1]; li += [j, j]; fi += [0, 1, 1, 1]; ob_arr += [obs[k][l], obs[k - 1][l]]
                params = np.zeros(12 + 4 * len(common))
                params[:6] = pose_compose(est[k], pose_inverse(Tp))
                for j, l in enumerate(common):
                    params[12 + 4 * j:16 + 4 * j] = line_transform(lines[l], Tp)
                wm = Window(2, len(common), np.asarray(ci, np.int32), np.asarray(li, np.int32), np.asarray(fi, np.int32),
                            np.asarray(ob_arr, np.float64).ravel(), params, params.copy(), dict(keyframe=k, kind="motion_only"))
                out, _ = motion_only(wm, max_iters)
                est[k] = pose_compose(out[:6], Tp)
        for lid, ob in obs[k].items():      # landmark initialisation from the first stereo view
            if lid not in lines:
                tri = triangulate(ob)
                if tri is not None:
                    T = pose_inverse(est[k])
                    R = rodrigues(T[:3])
                    p, d = R @ tri[0] + T[3:], R @ tri[1]
                    lines[lid] = av_to_orth(p - d * (p @ d), d)
        if k == 0:
            continue
        free = list(range(k, max(-1, k - W), -1))
        fixed = list(range(k - W, max(-1, k - 2 * W), -1)) if k - W >= 0 else []
        cams = free + fixed
        cnt = {}
        for c in free:
            for lid in obs[c]:
                if lid in lines:
                    cnt[lid] = cnt.get(lid, 0) + 1
        lids = sorted(l for l, n in cnt.items() if n >= 2)
        if not lids:
            continue
        lslot = {l: i for i, l in enumerate(lids)}
        Tn = est[k].copy()                                   # window frame := newest keyframe
        Tn_inv = pose_inverse(Tn)
        ci, li, fi, ob_arr = [], [], [], []
        for l in lids:                                       # grouped by landmark, as slam.cpp:899-920 packs them
            for slot, c in enumerate(cams):
                if l in obs[c]:
                    is_const = (c not in free) or (anchor_first and c == 0)
                    ci.append(slot); li.append(lslot[l]); fi.extend([1 if is_const else 0, 0]); ob_arr.append(obs[c][l])
        params = np.zeros(6 * len(cams) + 4 * len(lids))
        for slot, c in enumerate(cams):
            params[6 * slot:6 * slot + 6] = 0.0 if c == k else pose_compose(est[c], Tn_inv)
        for l in lids:
            params[6 * len(cams) + 4 * lslot[l]:6 * len(cams) + 4 * lslot[l] + 4] = line_transform(lines[l], Tn)
        w = Window(len(cams), len(lids), np.asarray(ci, np.int32), np.asarray(li, np.int32), np.asarray(fi, np.int32),
                   np.asarray(ob_arr, np.float64).ravel(), params, params.copy(), dict(keyframe=k))
        if record is not None:
            record.append(w)
        out, summ = solve(w, max_iters)
        stats.append(dict(keyframe=k, cameras=len(cams), lines=len(lids), observations=len(ci), **{
            a: summ[a] for a in ("initial_cost", "final_cost", "iterations")}))
        for slot, c in enumerate(cams):
            if c in free and not (anchor_first and c == 0):
                est[c] = pose_compose(out[6 * slot:6 * slot + 6], Tn)
        for l in lids:
            lines[l] = line_transform(out[6 * len(cams) + 4 * lslot[l]:6 * len(cams) + 4 * lslot[l] + 4], Tn_inv)
            refined.add(l)
    return np.stack([pose_inverse(T) for T in est]), stats


def trajectory_rmse(est_wc, true_wc):
    n = min(len(est_wc), len(true_wc))
    return float(np.sqrt(np.mean(np.sum((est_wc[:n, 3:] - true_wc[:n, 3:]) ** 2, axis=1))))
