"""Seeded synthetic LBA windows and pose graphs.

The reference's datasets are not shipped (reference README:18-23, .gitignore:3), so every test and
benchmark input is generated here from the reference's own camera constants
(reference src/parameter.h:43-49) and packed exactly as SLAM::bundle_adjustment packs its arrays
(reference src/slam.cpp:899-920): observations grouped by landmark, camera blocks first in the
parameter vector, `fixed_index[2i]` = camera constant, `fixed_index[2i+1]` = line constant.

Pure numpy; no product or oracle code is used to *generate* inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FOCAL = 406.05          # reference src/parameter.h:47
CX, CY = 327.783, 237.172
WIDTH, HEIGHT = 640, 480
BASELINE = 0.12         # reference src/parameter.h:46 (and the literal at lba_problem.h:101)
HUBER_DELTA = 1.0 / FOCAL  # reference src/lba_problem.cpp:78-80


def rodrigues(w: np.ndarray) -> np.ndarray:
    """Angle-axis -> rotation matrix (what gc_Rodriguez returns, reference src/gc.cpp:24-35)."""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    if th < 1e-12:
        return np.eye(3) + K
    k = K / th
    return np.eye(3) + np.sin(th) * k + (1 - np.cos(th)) * (k @ k)


def log_so3(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> angle-axis (any log map within 1e-12 is adequate)."""
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-10:
        return 0.5 * v
    if np.pi - th < 1e-6:
        A = (R + np.eye(3)) * 0.5
        ax = np.sqrt(np.maximum(np.diag(A), 0.0))
        i = int(np.argmax(ax))
        ax = A[:, i] / ax[i]
        ax /= np.linalg.norm(ax)
        if np.dot(ax, v) < 0:
            ax = -ax
        return th * ax
    return th / (2.0 * np.sin(th)) * v


def av_to_orth(cp: np.ndarray, v: np.ndarray) -> np.ndarray:
    """(closest point, direction) -> 4-parameter orthonormal line (reference src/gc.cpp:361-417)."""
    n = np.cross(cp, v)
    x = n / np.linalg.norm(n)
    y = v / np.linalg.norm(v)
    z = np.cross(x, y)
    a = np.arctan2(y[2], z[2])
    b = np.arcsin(-x[2])
    g = np.arctan2(x[1], x[0])
    wv = np.array([np.linalg.norm(n), np.linalg.norm(v)])
    wv = wv / np.linalg.norm(wv)
    return np.array([a, b, g, np.arcsin(wv[1])])


def orth_to_av(o: np.ndarray):
    """4-parameter line -> (closest point, unit direction) (reference src/gc.cpp:419-460)."""
    a, b, g, t = o
    s1, c1, s2, c2, s3, c3 = np.sin(a), np.cos(a), np.sin(b), np.cos(b), np.sin(g), np.cos(g)
    R = np.array([[c2 * c3, s1 * s2 * c3 - c1 * s3, c1 * s2 * c3 + s1 * s3],
                  [c2 * s3, s1 * s2 * s3 + c1 * c3, c1 * s2 * s3 - s1 * c3],
                  [-s2, s1 * c2, c1 * c2]])
    d = np.cos(t) / np.sin(t)
    return -R[:, 2] * d, R[:, 1]


@dataclass
class Window:
    """One LBA problem in the reference's array layout (reference src/lba_problem.h:188-196)."""
    num_cameras: int
    num_lines: int
    camera_index: np.ndarray   # int32 [N]
    line_index: np.ndarray     # int32 [N]
    fixed_index: np.ndarray    # int32 [2N]
    observations: np.ndarray   # float64 [8N]
    parameters: np.ndarray     # float64 [6C+4L] initial guess
    truth: np.ndarray          # float64 [6C+4L] generating parameters
    meta: dict = field(default_factory=dict)

    @property
    def num_observations(self) -> int:
        return int(self.camera_index.shape[0])

    @property
    def num_parameters(self) -> int:
        return 6 * self.num_cameras + 4 * self.num_lines


def _project(Rcw, tcw, P):
    """World point -> normalised coords in stereo cam A and cam B (B sits at +baseline in A's x)."""
    pc = Rcw @ P + tcw
    pb = pc - np.array([BASELINE, 0.0, 0.0])
    return pc, pb


def make_window(seed: int, num_cameras: int = 10, num_lines: int = 200, num_observations: int = 1000,
                sigma_px: float = 0.2, start: str = "near", anchored: bool = True,
                num_fixed_cameras: int = 0, newest_identity: bool = True, shuffle: bool = False) -> Window:
    """SURVEY.md §8d generator.  `num_fixed_cameras` extra constant cameras are appended after the free
    ones, as SLAM::bundle_adjustment does for keyframes beyond the window (reference slam.cpp:855-863)."""
    rng = np.random.default_rng(seed)
    C = num_cameras + num_fixed_cameras
    # camera trajectory: centres advance ~0.77 m along +z, small rotations (kf thresholds parameter.h:59-60)
    centres = np.zeros((C, 3))
    rots = []
    z = 0.0
    order = list(range(num_cameras, C)) + list(range(num_cameras))  # fixed (older) cameras come first in space
    for k, c in enumerate(order):
        centres[c] = np.array([rng.normal(0, 0.05), rng.normal(0, 0.02), z])
        rots.append((c, rng.normal(0, 0.03, 3)))
        z += 0.77 + rng.normal(0, 0.02)
    Rwc = {c: rodrigues(w) for c, w in rots}            # camera -> world
    newest = num_cameras - 1
    # world frame := newest camera (reference slam.cpp:1322 makes the newest keyframe exactly identity)
    Rn, cn = Rwc[newest], centres[newest].copy()
    cams_true = np.zeros((C, 6))
    Rcw_all, tcw_all = [], []
    for c in range(C):
        if newest_identity:
            Rw = Rn.T @ Rwc[c]
            cw = Rn.T @ (centres[c] - cn)
        else:
            Rw, cw = Rwc[c], centres[c]
        Rcw = Rw.T
        tcw = -Rcw @ cw
        w = log_so3(Rcw)
        if newest_identity and c == newest:
            w = np.zeros(3); tcw = np.zeros(3); Rcw = np.eye(3)
        cams_true[c, :3] = w
        cams_true[c, 3:] = tcw
        Rcw_all.append(rodrigues(w)); tcw_all.append(tcw)
    zmax = z

    # candidate segments in the unrotated scene frame, then expressed in the world frame
    lines_P, lines_Q, vis = [], [], []
    want = num_lines
    attempts = 0
    while len(lines_P) < want and attempts < 60:
        attempts += 1
        m = max(4 * want, 256)
        mid = np.stack([rng.uniform(-7, 7, m), rng.uniform(-3.5, 3.5, m), rng.uniform(3.0, zmax + 14.0, m)], 1)
        d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        ln = rng.uniform(0.5, 2.0, m)[:, None]
        P = mid - 0.5 * ln * d; Q = mid + 0.5 * ln * d
        if newest_identity:
            P = (P - cn) @ Rn; Q = (Q - cn) @ Rn       # rows: Rn^T (p - cn)
        Rs = np.stack(Rcw_all); ts = np.stack(tcw_all)                       # [C,3,3], [C,3]
        ok = np.ones((m, C), bool)
        for X in (P, Q):
            pa = np.einsum('cij,mj->mci', Rs, X) + ts[None]                  # [m,C,3]
            for off in (0.0, BASELINE):
                x = pa[..., 0] - off; y = pa[..., 1]; zc = pa[..., 2]
                zs = np.where(zc > 1e-6, zc, 1.0)
                u = FOCAL * x / zs + CX; v = FOCAL * y / zs + CY
                ok &= (zc >= 3.0) & (zc <= 16.0) & (u >= 0) & (u < WIDTH) & (v >= 0) & (v < HEIGHT)
        nfree_all = ok[:, :num_cameras].sum(1)
        for j in np.nonzero(nfree_all >= 2)[0]:                              # reference slam.cpp:839
            lines_P.append(P[j]); lines_Q.append(Q[j]); vis.append([int(c) for c in np.nonzero(ok[j])[0]])
            if len(lines_P) >= want:
                break
    L = len(lines_P)
    # thin observations to the target count, never below two free-camera observations per line
    total = sum(len(v) for v in vis)
    if total > num_observations:
        cand = [(l, c) for l, v in enumerate(vis) for c in v]
        rng.shuffle(cand)
        for (l, c) in cand:
            if total <= num_observations:
                break
            nfree = sum(1 for cc in vis[l] if cc < num_cameras)
            if c < num_cameras and nfree <= 2:
                continue
            vis[l].remove(c); total -= 1

    cam_idx, line_idx, fixed, obs = [], [], [], []
    lines_true = np.zeros((L, 4))
    for l in range(L):
        P, Q = lines_P[l], lines_Q[l]
        v = Q - P
        cp = P - v * (np.dot(P, v) / np.dot(v, v))
        lines_true[l] = av_to_orth(cp, v)
        for c in vis[l]:
            o = np.zeros(8)
            for e, X in enumerate((P, Q)):
                pa, pb = _project(Rcw_all[c], tcw_all[c], X)
                ua = FOCAL * pa[:2] / pa[2] + np.array([CX, CY]) + rng.normal(0, sigma_px, 2)
                ub = FOCAL * pb[:2] / pb[2] + np.array([CX, CY]) + rng.normal(0, sigma_px, 2)
                # pixel -> normalised, reference slam.cpp:121-128
                o[2 * e:2 * e + 2] = ua / FOCAL - np.array([CX, CY]) / FOCAL
                o[4 + 2 * e:6 + 2 * e] = ub / FOCAL - np.array([CX, CY]) / FOCAL
            cam_idx.append(c); line_idx.append(l)
            is_fixed = (c >= num_cameras) or (anchored and c == 0)
            fixed += [1 if is_fixed else 0, 0]
            obs.append(o)

    truth = np.concatenate([cams_true.ravel(), lines_true.ravel()])
    k = {"near": 1.0, "far": 10.0, "exact": 0.0}[start]
    p0 = truth.copy()
    for c in range(C):
        if (c >= num_cameras) or (anchored and c == 0):
            continue
        if newest_identity and c == newest and not anchored:
            pass
        p0[6 * c:6 * c + 3] += k * rng.normal(0, 1e-3, 3)
        p0[6 * c + 3:6 * c + 6] += k * rng.normal(0, 5e-3, 3)
    if newest_identity:
        p0[6 * newest:6 * newest + 6] = 0.0              # exactly zeros(6): Taylor branch of the rotation
    p0[6 * C:] += k * rng.normal(0, 2e-3, 4 * L)

    cam_idx = np.asarray(cam_idx, np.int32); line_idx = np.asarray(line_idx, np.int32)
    fixed = np.asarray(fixed, np.int32); obs = np.asarray(obs, np.float64).reshape(-1, 8)
    if shuffle:
        perm = rng.permutation(len(cam_idx))
        cam_idx, line_idx, obs = cam_idx[perm], line_idx[perm], obs[perm]
        fixed = fixed.reshape(-1, 2)[perm].ravel()
    return Window(C, L, cam_idx, line_idx, np.ascontiguousarray(fixed), np.ascontiguousarray(obs.ravel()),
                  p0, truth, dict(seed=seed, sigma_px=sigma_px, start=start, anchored=anchored,
                                  num_free_cameras=num_cameras, num_fixed_cameras=num_fixed_cameras))


def window_S(seed: int = 0, **kw) -> Window:
    """S: 10 cameras / 200 lines / ~1 k observations (BASELINE.json configs[0])."""
    return make_window(seed, 10, 200, 1000, **kw)


def window_M(seed: int = 0, **kw) -> Window:
    """M: 10 cameras / 2 k lines / ~10 k observations (BASELINE.json configs[1])."""
    return make_window(seed, 10, 2000, 10000, **kw)


def motion_only_window(seed: int, num_lines: int = 60, sigma_px: float = 0.5) -> Window:
    """What SLAM::motion_only_ba packs (reference slam.cpp:578-640): camera 0 free, camera 1 = identity and
    constant, every line constant, two observations per line (cam 0 then cam 1)."""
    w = make_window(seed, 2, num_lines, 2 * num_lines, sigma_px=sigma_px, start="near", anchored=False)
    # re-label so camera 1 is the identity (newest) and is constant, lines constant
    fixed = w.fixed_index.reshape(-1, 2).copy()
    fixed[:, 1] = 1
    fixed[:, 0] = (w.camera_index == 1).astype(np.int32)
    w.fixed_index = np.ascontiguousarray(fixed.ravel())
    w.meta["kind"] = "motion_only"
    return w


@dataclass
class PoseGraph:
    """One PO problem in the reference's array layout (reference src/po_problem.h:139-143)."""
    num_poses: int
    pose_index_1: np.ndarray   # int32 [E]
    pose_index_2: np.ndarray   # int32 [E]
    constraints: np.ndarray    # float64 [6E]
    parameters: np.ndarray     # float64 [6K] initial guess (drifted odometry)
    truth: np.ndarray          # float64 [6K]
    meta: dict = field(default_factory=dict)

    @property
    def num_edges(self) -> int:
        return int(self.pose_index_1.shape[0])


def _compose(T21, T10):
    """T20 = T21 o T10 on (angle-axis, t) world->camera poses (reference src/po_problem.h:55-64)."""
    R21, R10 = rodrigues(T21[:3]), rodrigues(T10[:3])
    return np.concatenate([log_so3(R21 @ R10), R21 @ T10[3:] + T21[3:]])


def _inverse(T):
    R = rodrigues(T[:3])
    return np.concatenate([-T[:3], -(R.T @ T[3:])])


def make_pose_graph(seed: int, num_poses: int = 261, neighbours: int = 3, num_loops: int = 8,
                    noise_rot: float = 2e-3, noise_tr: float = 1e-2, drift: float = 1.0) -> PoseGraph:
    """Odometry chain on a closed loop (so that loop-closure edges are geometrically meaningful) plus edges to
    the next `neighbours` keyframes (the reference links keyframes sharing landmarks, slam.cpp:1398-1416) and
    `num_loops` long-range loop-closure edges.  Constraint C of edge (n1,n2) is the measured T_{n2<-n1}, so the
    residual T2^-1 o C o T1 vanishes on a consistent graph (reference po_problem.h:73-105).  myungdong-scale
    default: 261 poses."""
    rng = np.random.default_rng(seed)
    K = num_poses
    radius = 0.75 * K / (2 * np.pi)          # keyframes every ~0.75 m around a circuit
    truth = np.zeros((K, 6))
    for k in range(K):
        ang = 2 * np.pi * k / K * (1.0 - 1.0 / K)
        # camera -> world: heading tangent to the circle, small wobble
        Rwc = rodrigues(np.array([0.0, ang, 0.0])) @ rodrigues(rng.normal(0, 0.02, 3))
        cw = np.array([radius * np.sin(ang), rng.normal(0, 0.02), radius * (1 - np.cos(ang))])
        Rcw = Rwc.T
        truth[k, :3] = log_so3(Rcw); truth[k, 3:] = -Rcw @ cw
    # world frame = pose 0 exactly identity
    T0inv = _inverse(truth[0])
    truth = np.stack([_compose(truth[k], T0inv) for k in range(K)])
    truth[0] = 0.0
    e1, e2, cons = [], [], []

    def add(n1, n2):
        C = _compose(truth[n2], _inverse(truth[n1]))
        C = C + np.concatenate([rng.normal(0, noise_rot, 3), rng.normal(0, noise_tr, 3)])
        e1.append(n1); e2.append(n2); cons.append(C)

    for k in range(K):
        for d in range(1, neighbours + 1):
            if k + d < K:
                add(k, k + d)
    for _ in range(num_loops):
        a = int(rng.integers(0, max(1, K // 10)))
        b = int(rng.integers(K - max(1, K // 10), K))
        if a != b:
            add(a, b)
    # std::set<pii> iteration order in the reference (slam.cpp:1249) is lexicographic
    orderv = sorted(range(len(e1)), key=lambda i: (e1[i], e2[i]))
    e1 = np.asarray([e1[i] for i in orderv], np.int32); e2 = np.asarray([e2[i] for i in orderv], np.int32)
    cons = np.asarray([cons[i] for i in orderv], np.float64)
    # initial guess: integrate noisy odometry only (drift), as the front end would have before PO
    p0 = np.zeros((K, 6))
    for k in range(1, K):
        C = _compose(truth[k], _inverse(truth[k - 1]))
        C = C + drift * np.concatenate([rng.normal(0, noise_rot, 3), rng.normal(0, noise_tr, 3)])
        p0[k] = _compose(C, p0[k - 1])
    return PoseGraph(K, e1, e2, np.ascontiguousarray(cons.ravel()), np.ascontiguousarray(p0.ravel()),
                     np.ascontiguousarray(truth.ravel()), dict(seed=seed))


def pose_graph_from_trajectory(poses_wc: np.ndarray, seed: int = 0, neighbours: int = 2, num_loops: int = 10,
                               noise_rot: float = 2e-3, noise_tr: float = 1e-2, drift: float = 1.0) -> PoseGraph:
    """Pose graph around a given keyframe trajectory (camera->world rows of (angle-axis, t), e.g. the fixture derived
    from the reference's myungdong output): odometry + neighbour edges with measurement noise, `num_loops`
    loop-closure edges between keyframes that are far apart in time but closest in space, and a drifted
    dead-reckoning initial guess.  Substitute for BASELINE.json configs[4]."""
    rng = np.random.default_rng(seed)
    K = len(poses_wc)
    truth = np.stack([_inverse(T) for T in poses_wc])          # world -> camera
    T0inv = _inverse(truth[0])
    truth = np.stack([_compose(truth[k], T0inv) for k in range(K)])
    truth[0] = 0.0
    centres = np.stack([-rodrigues(T[:3]).T @ T[3:] for T in truth])
    e1, e2, cons = [], [], []

    def add(n1, n2):
        C = _compose(truth[n2], _inverse(truth[n1]))
        C = C + np.concatenate([rng.normal(0, noise_rot, 3), rng.normal(0, noise_tr, 3)])
        e1.append(n1); e2.append(n2); cons.append(C)

    for k in range(K):
        for d in range(1, neighbours + 1):
            if k + d < K:
                add(k, k + d)
    # loop closures: spatially closest pairs at least K/4 keyframes apart, one per anchor spread along the path
    anchors = np.linspace(0, K - 1, num_loops + 2).astype(int)[1:-1]
    for a in anchors:
        far = [j for j in range(K) if abs(j - a) >= K // 4]
        j = min(far, key=lambda j: np.linalg.norm(centres[j] - centres[a]))
        add(min(a, j), max(a, j))
    orderv = sorted(range(len(e1)), key=lambda i: (e1[i], e2[i]))
    e1a = np.asarray([e1[i] for i in orderv], np.int32); e2a = np.asarray([e2[i] for i in orderv], np.int32)
    consa = np.asarray([cons[i] for i in orderv], np.float64)
    p0 = np.zeros((K, 6))
    for k in range(1, K):
        C = _compose(truth[k], _inverse(truth[k - 1]))
        C = C + drift * np.concatenate([rng.normal(0, noise_rot, 3), rng.normal(0, noise_tr, 3)])
        p0[k] = _compose(C, p0[k - 1])
    return PoseGraph(K, e1a, e2a, np.ascontiguousarray(consa.ravel()), np.ascontiguousarray(p0.ravel()),
                     np.ascontiguousarray(truth.ravel()), dict(seed=seed, source="trajectory"))


# ---------------------------------------------------------------------------------------------------------------------
# The reference's house simulation (matlab_script/house.m): the 74-segment line model its published LBA table
# (matlab_script/result_comp_ancdir_orthonorm/ba_result_*) was measured on.  The helpers house.m calls (zSegment,
# XYRectangle, ...) are not in the repository; their names say what they build.  The ground-truth trajectory
# (gt_trajectory_wave.txt) is not shipped either: house_trajectory() circles the model with a vertical wave.
# ---------------------------------------------------------------------------------------------------------------------
def house_segments(x=0.0, y=0.0, z=0.0):
    """[(P, Q)] x 74, reference matlab_script/house.m:20-133 (l = w = 4.5, h = 3.5 as overwritten at :20-22)."""
    l, w, h = 4.5, 4.5, 3.5
    a, b, c, d = .2, .4, .6, .8
    p, q, r = .25, .5, .65
    S = []

    def seg(P, Q):
        S.append((np.asarray(P, float), np.asarray(Q, float)))

    def zseg(x_, y_, z1, z2): seg([x_, y_, z1], [x_, y_, z2])
    def xseg(x1, x2, y_, z_): seg([x1, y_, z_], [x2, y_, z_])
    def yseg(x_, y1, y2, z_): seg([x_, y1, z_], [x_, y2, z_])

    def xyrect(x1, x2, y1, y2, z_):
        xseg(x1, x2, y1, z_); yseg(x2, y1, y2, z_); xseg(x2, x1, y2, z_); yseg(x1, y2, y1, z_)

    def yzrect(x_, y1, y2, z1, z2):
        yseg(x_, y1, y2, z1); zseg(x_, y2, z1, z2); yseg(x_, y2, y1, z2); zseg(x_, y1, z2, z1)

    zseg(x, y, z, z + r * h); zseg(x + l, y, z, z + r * h); zseg(x + l, y + w, z, z + r * h); zseg(x, y + w, z, z + r * h)   # 1-4 walls
    xyrect(x, x + l, y, y + w, z)                                                                                          # 5-8 floor
    seg([x, y, z + r * h], [x, y + w / 2, z + h]); seg([x, y + w / 2, z + h], [x, y + w, z + r * h])                        # 9-12 roof slopes
    seg([x + l, y, z + r * h], [x + l, y + w / 2, z + h]); seg([x + l, y + w / 2, z + h], [x + l, y + w, z + r * h])
    xseg(x, x + l, y + .5 * w, z + h); xseg(x, x + l, y, z + r * h); xseg(x, x + l, y + w, z + r * h)                        # 13-15 roof
    yzrect(x, y + c * w, y + d * w, z, z + q * h)                                                                          # 16-19 door
    yzrect(x, y + a * w, y + b * w, z + p * h, z + q * h)                                                                  # 20-23 window
    yseg(x, y, y + w, z + r * h); yseg(x + l, y, y + w, z + r * h)                                                          # 24-25
    yseg(x, y + a * w, y + b * w, (z + p * h + z + q * h) / 2); zseg(x, (y + a * w + y + b * w) / 2, z + p * h, z + q * h)   # 26-27
    for fx in (.5, .25, .75):                                                                                              # 28-30
        seg([x + l * fx, y, z + r * h], [x + l * fx, y + w / 2, z + h])
    for fx in (.5, .25, .75):                                                                                              # 31-33
        seg([x + l * fx, y + w / 2, z + h], [x + l * fx, y + w, z + r * h])
    for k in (1, 2, 3):                                                                                                    # 34-36
        xseg(x, x + l, y + w * k / 8, z + r * h + (h - r * h) * k / 4)
    for k, m in ((5, 3), (6, 2), (7, 1)):                                                                                  # 37-39
        xseg(x, x + l, y + w * k / 8, z + r * h + (h - r * h) * m / 4)
    for k in (1, 2, 3): zseg(x + l * k / 4, y, z, z + r * h)                                                                # 40-42
    for k in (1, 2, 3): zseg(x + l * k / 4, y + w, z, z + r * h)                                                            # 43-45
    for k in (1, 2, 3): zseg(x + l, y + w * k / 4, z, z + r * h)                                                            # 46-48
    seg([x, y + c * w, z], [x, y + d * w, z + q * h]); seg([x, y + d * w, z], [x, y + c * w, z + q * h])                    # 49-50
    for k in range(4):                                                                                                     # 51-58 front braces
        seg([x + k / 4 * l, y, z], [x + (k + 1) / 4 * l, y, z + r * h]); seg([x + (k + 1) / 4 * l, y, z], [x + k / 4 * l, y, z + r * h])
    for k in range(4):                                                                                                     # 59-66 side braces
        seg([x + l, y + k / 4 * w, z], [x + l, y + (k + 1) / 4 * w, z + r * h]); seg([x + l, y + (k + 1) / 4 * w, z], [x + l, y + k / 4 * w, z + r * h])
    for k in range(4):                                                                                                     # 67-74 back braces
        seg([x + k / 4 * l, y + w, z], [x + (k + 1) / 4 * l, y + w, z + r * h]); seg([x + (k + 1) / 4 * l, y + w, z], [x + k / 4 * l, y + w, z + r * h])
    assert len(S) == 74
    return S


def house_trajectory(num_keyframes=402, radius=11.0, wave=0.6, turns=1.0):
    """Camera->world poses (angle-axis, t) [K][6] circling the house at `radius` m, optical axis towards its centre, height
    on a vertical wave.  Consecutive keyframes are ~0.17 m apart at the defaults (402 frames per turn)."""
    centre = np.array([2.25, 2.25, 1.6])
    out = np.zeros((num_keyframes, 6))
    for k in range(num_keyframes):
        ang = 2 * np.pi * turns * k / num_keyframes
        c = centre + np.array([radius * np.cos(ang), radius * np.sin(ang), wave * np.sin(6 * ang)])
        zc = centre - c; zc /= np.linalg.norm(zc)                  # optical axis
        xc = np.cross(zc, [0.0, 0.0, 1.0]); xc /= np.linalg.norm(xc)  # image x: horizontal
        yc = np.cross(zc, xc)                                      # image y: down
        Rwc = np.stack([xc, yc, zc], axis=1)
        out[k, :3] = log_so3(Rwc); out[k, 3:] = c
    return out


def ransac_case(seed, n_lines=120, n_hyp=64, sigma_px=0.3, baseline=0.12):
    """RANSAC scoring input (reference src/slam.cpp:363-425).  Lines in the previous keyframe's frame, a true motion T (previous -> current), stereo observations in the current
    frame, and hypotheses = T perturbed by growing amounts (a few with |t| > 1, which the reference skips)."""
    rng = np.random.default_rng(seed)
    w = rng.normal(0, 0.03, 3); t = np.array([rng.normal(0, 0.05), rng.normal(0, 0.05), -0.75 + rng.normal(0, 0.05)])
    R = rodrigues(w)
    lines, obs = np.zeros((n_lines, 6)), np.zeros((n_lines, 8))
    for k in range(n_lines):
        z = rng.uniform(3, 12)
        mid = np.array([rng.uniform(-0.6, 0.6) * z, rng.uniform(-0.4, 0.4) * z, z])
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        lines[k, :3], lines[k, 3:] = mid - v * (mid @ v), v
        P, Q = R @ (mid - 0.6 * v) + t, R @ (mid + 0.6 * v) + t
        ends = [P, Q, P - [baseline, 0, 0], Q - [baseline, 0, 0]]
        obs[k] = np.concatenate([[e[0] / e[2], e[1] / e[2]] for e in ends]) + rng.normal(0, sigma_px / 406.05, 8)
    poses = np.zeros((n_hyp, 12))
    for h in range(n_hyp):
        s = 0.0 if h == 0 else 10.0 ** rng.uniform(-4, -0.5)
        Rh = rodrigues(w + rng.normal(0, s, 3))
        th = t + rng.normal(0, 3 * s, 3) + (np.array([0, 0, 2.0]) if h % 17 == 5 else 0)
        poses[h, :9], poses[h, 9:] = Rh.ravel(), th
    return poses, lines, obs, (R, t)
