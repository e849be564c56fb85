#!/usr/bin/env python
"""Derives keyframe-pose fixtures from the reference's published output trajectories (run in the build container only:
/root/reference does not exist on the GPU box).

  /root/reference/matlab_script/traj_slslam_itbt3f_basize10_wolc.txt     -> traj_it3f_wolc.npy      (102 keyframes)
  /root/reference/matlab_script/traj_slslam_myungdong_basize10_wolc.txt  -> traj_myungdong_wolc.npy (253 keyframes)

File format (reference src/slam.cpp:1489-1493): idx, t_z, -t_x, -t_y, angle-axis of the camera->world rotation.
The fixtures hold camera->world poses as (angle-axis[3], t[3]) rows.  They are the substitutes SURVEY.md §8d names
for BASELINE.json configs 3 and 5 (the datasets themselves are not shipped with the reference).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/matlab_script"


def convert(name):
    a = np.loadtxt(os.path.join(SRC, name))
    out = np.zeros((a.shape[0], 6))
    out[:, :3] = a[:, 4:7]
    out[:, 3] = -a[:, 2]
    out[:, 4] = -a[:, 3]
    out[:, 5] = a[:, 1]
    return out


if __name__ == "__main__":
    for src, dst in (("traj_slslam_itbt3f_basize10_wolc.txt", "traj_it3f_wolc.npy"),
                     ("traj_slslam_myungdong_basize10_wolc.txt", "traj_myungdong_wolc.npy")):
        t = convert(src)
        np.save(os.path.join(HERE, dst), t)
        steps = np.linalg.norm(np.diff(t[:, 3:], axis=0), axis=1)
        print(dst, t.shape, "path length %.1f m" % steps.sum(), "median step %.3f m" % np.median(steps))
