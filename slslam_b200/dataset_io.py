"""The reference's file formats on either side of the LBA path: the per-frame observation files
SLAM::grab_new_frame reads and the trajectory file SLAM::save_trajectory writes.  Host-side numpy / text only.

Observation file `<dir>/%04d.txt` (reference src/slam.cpp:62-106): one line segment per text line,
    feature_id x0 y0 x1 y1 x2 y2 x3 y3 <token>
pixels; (x0,y0)-(x1,y1) in the left image, (x2,y2)-(x3,y3) in the right one; the trailing token is read and dropped.
SLAM::insert_curr_obs (src/slam.cpp:110-135) normalises every coordinate as v / f - c / f with the left intrinsics.

Trajectory file (reference src/slam.cpp:1473-1496): per keyframe `idx  t_z  -t_x  -t_y  w_x  w_y  w_z`, tab-separated,
(w, t) the camera->world pose with keyframe 0 at the origin.
"""
from __future__ import annotations

import os

import numpy as np

from .synth import CX, CY, FOCAL


def frame_path(obs_dir: str, frame_id: int) -> str:
    return os.path.join(obs_dir, "%04d.txt" % frame_id)


def write_frame_observations(obs_dir: str, frame_id: int, obs_px: dict) -> str:
    """obs_px: {feature id: 8 pixel coordinates}.  Full double precision so that a read-back is exact."""
    os.makedirs(obs_dir, exist_ok=True)
    path = frame_path(obs_dir, frame_id)
    with open(path, "w") as f:
        for fid in sorted(obs_px):
            f.write("%d %s 0\n" % (fid, " ".join("%.17g" % v for v in np.asarray(obs_px[fid], np.float64))))
    return path


def read_frame_observations(obs_dir: str, frame_id: int, focal=(FOCAL, FOCAL), centre=(CX, CY)):
    """{feature id: 8 normalised coordinates} or None when the file does not exist (grab_new_frame returns false)."""
    path = frame_path(obs_dir, frame_id)
    if not os.path.exists(path):
        return None
    fx, fy = focal
    cx, cy = centre
    scale = np.array([1 / fx, 1 / fy] * 4)
    shift = np.array([cx / fx, cy / fy] * 4)
    out = {}
    with open(path) as f:
        for line in f:
            tok = line.split()
            if len(tok) < 9:
                continue
            out[int(tok[0])] = np.array([float(t) for t in tok[1:9]]) * scale - shift     # v / f - c / f, slam.cpp:121-128
    return out


def to_pixels(obs_norm, focal=(FOCAL, FOCAL), centre=(CX, CY)) -> np.ndarray:
    fx, fy = focal
    cx, cy = centre
    return np.asarray(obs_norm, np.float64) * np.array([fx, fy] * 4) + np.array([cx, cy] * 4)


def write_trajectory(path: str, poses_wc: np.ndarray) -> None:
    """poses_wc: camera->world (angle-axis[3], t[3]) rows."""
    with open(path, "w") as f:
        for i, T in enumerate(np.asarray(poses_wc, np.float64)):
            f.write("%d\t%.17g\t%.17g\t%.17g\t%.17g\t%.17g\t%.17g\n" % (i, T[5], -T[3], -T[4], T[0], T[1], T[2]))


def read_trajectory(path: str) -> np.ndarray:
    a = np.atleast_2d(np.loadtxt(path))
    out = np.zeros((a.shape[0], 6))
    out[:, :3] = a[:, 4:7]
    out[:, 3], out[:, 4], out[:, 5] = -a[:, 2], -a[:, 3], a[:, 1]
    return out
