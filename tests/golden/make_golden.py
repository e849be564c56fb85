"""Generates the committed golden fixtures in this directory.

The reference ships no tests or golden vectors for the LBA / PO path and cannot be built in this image
(SURVEY.md §8c), so these fixtures come from an INDEPENDENT restatement, written differently from
oracle/slslam_oracle.cpp on purpose:
  * residuals through rotation MATRICES and the Pluecker form of SURVEY.md Appendix B
    (n = R n_w + [t - o_c]x R v_w), not through AngleAxisRotatePoint on (closest point, direction);
  * Jacobians from torch.autograd (fp64), not from dual numbers;
  * PO residual through matrix composition and a matrix log map, not through quaternions;
  * converged costs from scipy.optimize.least_squares, not from the restated Ceres LM loop.
Nothing here imports oracle/ or the product.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from slslam_b200 import synth  # noqa: E402

torch.set_default_dtype(torch.float64)
BASELINE = 0.12


def skew(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def rot(w):
    """exp([w]x) by the closed form with series coefficients near zero (smooth for autograd at w = 0)."""
    th2 = (w * w).sum()
    K = skew(w)
    if float(th2) < 1e-16:
        A = 1.0 - th2 / 6.0
        B = 0.5 - th2 / 24.0
    else:
        th = torch.sqrt(th2)
        A = torch.sin(th) / th
        B = (1.0 - torch.cos(th)) / th2
    return torch.eye(3) + A * K + B * (K @ K)


def log_rot(R):
    c = ((R[0, 0] + R[1, 1] + R[2, 2]) - 1.0) * 0.5
    v = torch.stack([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = torch.sqrt((v * v).sum()) * 0.5
    th = torch.atan2(s, c)
    if float(th) < 1e-8:
        return 0.5 * v
    return th / (2.0 * s) * v


def lba_residual(cam, line, ob):
    a, b, g, t = line[0], line[1], line[2], line[3]
    Rx = torch.stack([torch.stack([torch.ones(()), torch.zeros(()), torch.zeros(())]),
                      torch.stack([torch.zeros(()), torch.cos(a), -torch.sin(a)]),
                      torch.stack([torch.zeros(()), torch.sin(a), torch.cos(a)])])
    Ry = torch.stack([torch.stack([torch.cos(b), torch.zeros(()), torch.sin(b)]),
                      torch.stack([torch.zeros(()), torch.ones(()), torch.zeros(())]),
                      torch.stack([-torch.sin(b), torch.zeros(()), torch.cos(b)])])
    Rz = torch.stack([torch.stack([torch.cos(g), -torch.sin(g), torch.zeros(())]),
                      torch.stack([torch.sin(g), torch.cos(g), torch.zeros(())]),
                      torch.stack([torch.zeros(()), torch.zeros(()), torch.ones(())])])
    U = Rz @ Ry @ Rx                        # columns x^, y^, z^
    n_w = U[:, 0] / torch.tan(t)            # Pluecker normal for unit direction y^: cp x dv = cot(t) x^
    v_w = U[:, 1]
    R = rot(cam[:3])
    tt = cam[3:]
    out = []
    for k, off in enumerate((0.0, BASELINE)):
        o = torch.tensor([off, 0.0, 0.0])
        n = R @ n_w + torch.linalg.cross(tt - o, R @ v_w)
        s = torch.sqrt(n[0] ** 2 + n[1] ** 2)
        for e in range(2):
            x, y = ob[4 * k + 2 * e], ob[4 * k + 2 * e + 1]
            out.append(-(x * n[0] + y * n[1] + n[2]) / s)
    return torch.stack(out)


def po_residual(p1, p2, c):
    R1, R2, Rc = rot(p1[:3]), rot(p2[:3]), rot(c[:3])
    Rtc = Rc @ R1
    ttc = Rc @ p1[3:] + c[3:]
    Re = R2.T @ Rtc
    te = R2.T @ (ttc - p2[3:])
    return torch.cat([log_rot(Re), te])


def main():
    rng = np.random.default_rng(20261017)
    jac = torch.autograd.functional.jacobian
    cams, lines, obs, rs, Jcs, Jls = [], [], [], [], [], []
    for i in range(64):
        cam = np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 1.0, 3)])
        if i % 8 == 0:
            cam[:3] = 0.0                   # exact identity rotation: Taylor branch in the reference's helper
        if i % 8 == 1:
            cam[:3] = rng.normal(0, 1e-7, 3)
        # a line in front of the camera, converted with the reference's own parameterisation
        P = np.array([rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(4, 12)])
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        cp = P - d * np.dot(P, d)
        line = synth.av_to_orth(cp, d)
        ob = rng.normal(0, 0.3, 8)
        ct, lt, ot = torch.tensor(cam), torch.tensor(line), torch.tensor(ob)
        r = lba_residual(ct, lt, ot)
        Jc, Jl = jac(lambda c_, l_: lba_residual(c_, l_, ot), (ct, lt))
        cams.append(cam); lines.append(line); obs.append(ob)
        rs.append(r.numpy()); Jcs.append(Jc.numpy()); Jls.append(Jl.numpy())
    np.savez(os.path.join(HERE, "lba_residual_cases.npz"), cam=np.array(cams), line=np.array(lines), obs=np.array(obs),
             r=np.array(rs), Jc=np.array(Jcs), Jl=np.array(Jls))

    p1s, p2s, cs, rs, J1s, J2s = [], [], [], [], [], []
    for i in range(48):
        p1 = np.concatenate([rng.normal(0, 0.4, 3), rng.normal(0, 2.0, 3)])
        p2 = np.concatenate([rng.normal(0, 0.4, 3), rng.normal(0, 2.0, 3)])
        c = np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.5, 3)])
        if i % 6 == 0:
            p1[:] = 0.0                     # keyframe 0 is the identity in the reference's pose graph
        a, b, cc = torch.tensor(p1), torch.tensor(p2), torch.tensor(c)
        r = po_residual(a, b, cc)
        J1, J2 = jac(lambda x, y: po_residual(x, y, cc), (a, b))
        p1s.append(p1); p2s.append(p2); cs.append(c); rs.append(r.numpy()); J1s.append(J1.numpy()); J2s.append(J2.numpy())
    np.savez(os.path.join(HERE, "po_residual_cases.npz"), p1=np.array(p1s), p2=np.array(p2s), c=np.array(cs),
             r=np.array(rs), J1=np.array(J1s), J2=np.array(J2s))

    # converged minima of small non-robust problems from scipy (independent optimiser)
    from scipy.optimize import least_squares
    minima = []
    for seed in (1, 2, 3):
        w = synth.make_window(seed, num_cameras=4, num_lines=24, num_observations=80, sigma_px=0.5, start="near")
        C = w.num_cameras
        free_c = sorted(set(int(c) for c, f in zip(w.camera_index, w.fixed_index[0::2]) if not f))
        sel = np.concatenate([np.arange(6 * c, 6 * c + 6) for c in free_c] + [np.arange(6 * C, w.num_parameters)])
        ob_t = torch.tensor(w.observations.reshape(-1, 8))

        def fun(x):
            p = w.parameters.copy(); p[sel] = x
            pt = torch.tensor(p)
            out = [lba_residual(pt[6 * c:6 * c + 6], pt[6 * C + 4 * l:6 * C + 4 * l + 4], ob_t[i])
                   for i, (c, l) in enumerate(zip(w.camera_index, w.line_index))]
            return torch.cat(out).numpy()

        sol = least_squares(fun, w.parameters[sel], method="trf", x_scale="jac", xtol=1e-15, ftol=1e-15, gtol=1e-15,
                            max_nfev=400)
        minima.append(dict(seed=seed, num_cameras=4, num_lines=24, num_observations=80, sigma_px=0.5, start="near",
                           robust=False, cost=float(sol.cost), nfev=int(sol.nfev)))
        print("lba minimum", seed, sol.cost, sol.nfev, sol.status)
    po_min = []
    for seed in (1, 2):
        g = synth.make_pose_graph(seed, num_poses=12, neighbours=2, num_loops=2)
        konst = int(g.pose_index_1[0])
        sel = np.concatenate([np.arange(6 * k, 6 * k + 6) for k in range(g.num_poses) if k != konst])
        ct = torch.tensor(g.constraints.reshape(-1, 6))

        def fun(x):
            p = g.parameters.copy(); p[sel] = x
            pt = torch.tensor(p)
            out = [po_residual(pt[6 * a:6 * a + 6], pt[6 * b:6 * b + 6], ct[e])
                   for e, (a, b) in enumerate(zip(g.pose_index_1, g.pose_index_2))]
            return torch.cat(out).numpy()

        sol = least_squares(fun, g.parameters[sel], method="trf", x_scale="jac", xtol=1e-15, ftol=1e-15, gtol=1e-15,
                            max_nfev=400)
        po_min.append(dict(seed=seed, num_poses=12, neighbours=2, num_loops=2, cost=float(sol.cost)))
        print("po minimum", seed, sol.cost, sol.nfev, sol.status)
    with open(os.path.join(HERE, "minima.json"), "w") as f:
        json.dump(dict(lba=minima, po=po_min), f, indent=1)

    # the two restatement-derived vectors quoted in SURVEY.md §8c (NOT from the reference binary)
    kat = dict(
        lba=dict(cam=[0.01, -0.02, 0.03, 0.1, -0.2, 0.3], line=[0.3, -0.4, 0.5, 0.6],
                 obs=[0.1, 0.2, -0.1, 0.25, 0.05, 0.2, -0.15, 0.25],
                 r=[-0.7485569598465627, -0.5967185255271112, -0.6086360697783584, -0.46079325099172763]),
        po=dict(p1=[0.1, -0.2, 0.3, 1, 2, 3], p2=[0.15, -0.1, 0.25, 1.2, 1.9, 3.3], c=[0.02, 0.05, -0.03, 0.2, -0.1, 0.3],
                r=[-0.03136810784688467, -0.04967635148980306, 0.02079791289278603, 0.1766986832032107,
                   -0.14405007360055877, -0.01762362280937646]))
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)


if __name__ == "__main__":
    main()
