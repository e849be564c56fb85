"""Device-resident map: the geometry conversions, the window assembly and the write-back that
SLAM::bundle_adjustment performs on the host around ceres::Solve (reference src/slam.cpp:795-975, src/gc.cpp), run as
kernels on a map that stays on the device -- against a host restatement of those loops written here with numpy."""
import numpy as np
import pytest

from slslam_b200 import replay, synth
from slslam_b200.synth import av_to_orth, log_so3, orth_to_av, rodrigues

pytestmark = pytest.mark.gpu


def test_geometry_conversions(gpu):
    """gc_av_to_orth / gc_orth_to_av (src/gc.cpp:361-460) and the rotation <-> angle-axis pair behind gc_Rt_to_wt /
    gc_wt_to_Rt on the device against the numpy restatements; round trip closest point / direction to 1e-9
    (SURVEY.md §8c anchor 2)."""
    rng = np.random.default_rng(0)
    n = 500
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    p = rng.normal(size=(n, 3)) * rng.uniform(0.5, 12.0, (n, 1))
    cp = p - d * np.sum(p * d, axis=1, keepdims=True)
    av = np.concatenate([cp, d], axis=1)
    orth = gpu.geometry_convert(0, av)
    ref = np.stack([av_to_orth(a[:3], a[3:]) for a in av])
    assert np.abs(orth - ref).max() < 1e-13
    back = gpu.geometry_convert(1, orth)
    assert np.abs(back - av).max() < 1e-9
    refb = np.stack([np.concatenate(orth_to_av(o)) for o in ref])
    assert np.abs(back - refb).max() < 1e-11
    # a non-unit direction and a scaled closest point give the same line parameters
    av2 = av.copy(); av2[:, 3:] *= 3.7
    assert np.abs(gpu.geometry_convert(0, av2) - orth).max() < 1e-12
    w = rng.normal(size=(n, 3)); w = w / np.linalg.norm(w, axis=1, keepdims=True) * rng.uniform(0.0, 3.0, (n, 1))   # |w| < pi
    w[0] = 0.0; w[1] = [1e-9, 0, 0]; w[2] = [np.pi - 1e-7, 0, 0]
    R = gpu.geometry_convert(3, w)
    Rref = np.stack([rodrigues(x).ravel() for x in w])
    assert np.abs(R - Rref).max() < 1e-13
    wb = gpu.geometry_convert(2, R)
    assert np.abs(wb - np.stack([log_so3(x.reshape(3, 3)) for x in Rref])).max() < 1e-8
    assert np.abs(wb[3:] - w[3:]).max() < 1e-9 and np.abs(wb[:2] - w[:2]).max() < 1e-15


class HostMap:
    """The reference's data structures and loops, restated: keyframes kfs[id].T, landmarks lms[id] = (line in the frame of
    init_kfid, init_kfid, obs_vec [(kf id, obs)]), keyframe.member_lms; bundle_adjustment() packs the window exactly as
    src/slam.cpp:799-920 and writes back as :957-972."""

    def __init__(self):
        self.T, self.lms, self.member = {}, {}, {}

    def add_keyframe(self, kf, T12, lm_ids, obs):
        self.T[kf] = np.asarray(T12, float).copy()
        self.member[kf] = set(int(l) for l in lm_ids)
        for l, o in zip(lm_ids, obs):
            if int(l) in self.lms:
                self.lms[int(l)]["obs"].append((kf, np.asarray(o, float)))
            else:
                self.lms[int(l)] = dict(line=None, init=None, obs=[(kf, np.asarray(o, float))])

    def add_landmark(self, lm, init_kf, av6):
        self.lms[int(lm)]["line"] = np.asarray(av6, float).copy()
        self.lms[int(lm)]["init"] = init_kf

    @staticmethod
    def line_to_pose(l, T):
        R, t = T[:9].reshape(3, 3), T[9:]
        return np.concatenate([R @ l[:3] + t, R @ l[3:]])

    @staticmethod
    def line_from_pose(l, T):
        R, t = T[:9].reshape(3, 3), T[9:]
        return np.concatenate([R.T @ (l[:3] - t), R.T @ l[3:]])

    def pack(self, ba, W):
        """ba: {kf id: rank}.  Returns the window in the reference's layout + (camera keyframes, line landmarks)."""
        cam_of, cams, lm_cnt = {}, [], {}
        for kf in sorted(ba):
            if ba[kf] >= W:
                continue
            for l in self.member[kf]:
                lm_cnt[l] = lm_cnt.get(l, 0) + 1
            cam_of[kf] = len(cams); cams.append(kf)
        ci, li, fi, ob, lines = [], [], [], [], []
        for l in sorted(lm_cnt):
            if lm_cnt[l] < 2 or self.lms[l]["line"] is None:
                continue
            for kf, o in self.lms[l]["obs"]:
                if kf not in ba:
                    continue
                if kf not in cam_of:
                    cam_of[kf] = len(cams); cams.append(kf)
                    fi += [1, 0]
                else:
                    fi += [0 if ba[kf] < W else 1, 0]
                ci.append(cam_of[kf]); li.append(len(lines)); ob.append(o)
            lines.append(l)
        params = np.zeros(6 * len(cams) + 4 * len(lines))
        for i, kf in enumerate(cams):
            T = self.T[kf]
            params[6 * i:6 * i + 3] = log_so3(T[:9].reshape(3, 3)); params[6 * i + 3:6 * i + 6] = T[9:]
        for j, l in enumerate(lines):
            lw = self.line_from_pose(self.lms[l]["line"], self.T[self.lms[l]["init"]])
            params[6 * len(cams) + 4 * j:6 * len(cams) + 4 * j + 4] = av_to_orth(lw[:3], lw[3:])
        w = synth.Window(len(cams), len(lines), np.asarray(ci, np.int32), np.asarray(li, np.int32), np.asarray(fi, np.int32),
                         np.asarray(ob, float).ravel(), params, params.copy(), {})
        return w, cams, lines

    def write_back(self, params, cams, lines):
        for i, kf in enumerate(cams):
            self.T[kf] = np.concatenate([rodrigues(params[6 * i:6 * i + 3]).ravel(), params[6 * i + 3:6 * i + 6]])
        for j, l in enumerate(lines):
            cp, dv = orth_to_av(params[6 * len(cams) + 4 * j:6 * len(cams) + 4 * j + 4])
            self.lms[l]["line"] = self.line_to_pose(np.concatenate([cp, dv]), self.T[self.lms[l]["init"]])


def _T12(pose6):
    return np.concatenate([rodrigues(pose6[:3]).ravel(), pose6[3:]])


def _scene(K=16, lines_per_kf=24, sigma_px=0.3, seed=5):
    import os
    traj = np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_it3f_wolc.npy"))[:K]
    truth_cw = np.stack([replay.pose_inverse(T) for T in traj])
    P, Q = replay.make_scene(traj, seed, lines_per_kf)
    obs = replay.observe(truth_cw, P, Q, sigma_px, seed, lines_per_kf)
    return truth_cw, obs


def test_resident_map_window_assembly_solve_and_write_back(gpu):
    """A sliding-window run on the device-resident map beside the host restatement of SLAM::bundle_adjustment fed with the
    same keyframes: per window (1) the assembled arrays agree -- same cameras, same lines, per line the same observation
    sequence bit for bit, parameters to 1e-12; (2) the solve on the resident window gives the bits the host-buffer entry
    point gives for the same arrays; (3) the written-back poses and lines agree with the numpy write-back to 1e-12, and
    with the fully host-side run to the solver's tolerance."""
    truth_cw, obs = _scene()
    K, W = len(truth_cw), 5
    rng = np.random.default_rng(11)
    dm = gpu.DeviceMap(64, 4096, 1 << 16)
    hm = HostMap()
    est = np.zeros((K, 6)); est[0] = truth_cw[0]
    known = set()
    checked = 0
    for k in range(K):
        if k > 0:
            rel = replay.pose_compose(truth_cw[k], replay.pose_inverse(truth_cw[k - 1]))
            rel = rel + np.concatenate([rng.normal(0, 3e-3, 3), rng.normal(0, 3e-2, 3)])
            est[k] = replay.pose_compose(rel, est[k - 1])
        lids = sorted(obs[k])
        ob = np.stack([obs[k][l] for l in lids]) if lids else np.zeros((0, 8))
        T12 = _T12(est[k])
        dm.add_keyframe(k, T12, lids, ob)
        hm.add_keyframe(k, T12, lids, ob)
        new_ids, new_av = [], []
        for l in lids:                                   # landmark initialisation: line in the frame of its first keyframe
            if l not in known:
                tri = replay.triangulate(obs[k][l])
                if tri is not None:
                    known.add(l); new_ids.append(l); new_av.append(np.concatenate(tri))
        if new_ids:
            dm.add_landmarks(new_ids, [k] * len(new_ids), np.stack(new_av))
            for l, a in zip(new_ids, new_av):
                hm.add_landmark(l, k, a)
        if k == 0:
            continue
        # the caller's metric embedding: window keyframes re-anchored at the newest one (reference slam.cpp:1317-1366)
        ba = {c: r for r, c in enumerate(range(k, max(-1, k - 2 * W), -1))}
        Tn_inv = replay.pose_inverse(est[k])
        rel12 = {c: _T12(np.zeros(6) if c == k else replay.pose_compose(est[c], Tn_inv)) for c in ba}
        dm.set_poses(list(ba), np.stack([rel12[c] for c in ba]))
        for c in ba:
            hm.T[c] = rel12[c].copy()
        # ---- (1) assembly ----
        s_dev = dm.bundle_adjust(list(ba), [ba[c] for c in ba], W, max_iters=6)
        wd = dm.last_window()
        wh, cams_h, lines_h = hm.pack(ba, W)
        if wh.num_lines == 0:
            assert dm.sizes[1] == 0
            continue
        assert list(wd["line_landmark"]) == lines_h
        assert dm.sizes[1] == wh.num_lines and dm.sizes[2] == wh.num_observations
        cams_d = list(wd["camera_keyframe"])
        nfree = sum(1 for c in ba if ba[c] < W)
        assert cams_d[:nfree] == cams_h[:nfree]                     # free cameras: same indices
        assert set(cams_h) <= set(cams_d)                           # constant cameras: the device keeps all of the window's
        for j in range(wh.num_lines):
            ih, idv = np.flatnonzero(wh.line_index == j), np.flatnonzero(wd["line_index"] == j)
            assert [cams_h[c] for c in wh.camera_index[ih]] == [cams_d[c] for c in wd["camera_index"][idv]]
            assert np.array_equal(wh.observations.reshape(-1, 8)[ih], wd["observations"].reshape(-1, 8)[idv])
            assert np.array_equal(wh.fixed_index.reshape(-1, 2)[ih], wd["fixed_index"].reshape(-1, 2)[idv])
        Ch, Cd = wh.num_cameras, wd["num_cameras"]
        for i, c in enumerate(cams_h):
            assert np.abs(wh.parameters[6 * i:6 * i + 6] - wd["parameters"][6 * cams_d.index(c):6 * cams_d.index(c) + 6]).max() < 1e-12
        assert np.abs(wh.parameters[6 * Ch:] - wd["parameters"][6 * Cd:]).max() < 1e-12
        # ---- (2) the resident solve = the host-buffer solve of the same arrays, bit for bit ----
        wdev = synth.Window(Cd, wd["num_lines"], wd["camera_index"], wd["line_index"], wd["fixed_index"], wd["observations"],
                            wd["parameters"], wd["parameters"].copy(), {})
        p_ref, s_ref = gpu.lba_solve(wdev, max_iters=6)
        assert s_dev == s_ref, (s_dev, s_ref)
        # ---- (3) write-back ----
        hm2_T = {c: np.concatenate([rodrigues(p_ref[6 * i:6 * i + 3]).ravel(), p_ref[6 * i + 3:6 * i + 6]]) for i, c in enumerate(cams_d)}
        got = dm.get_poses(cams_d)
        for i, c in enumerate(cams_d):
            assert np.abs(got[i] - hm2_T[c]).max() < 1e-12
        gl = dm.get_landmarks(lines_h)
        for j, l in enumerate(lines_h):
            cp, dv = orth_to_av(p_ref[6 * Cd + 4 * j:6 * Cd + 4 * j + 4])
            init = hm.lms[l]["init"]
            Tinit = hm2_T.get(init, hm.T[init])
            assert np.abs(gl[j] - HostMap.line_to_pose(np.concatenate([cp, dv]), Tinit)).max() < 1e-11
        # the fully host-side run (host pack -> host-buffer solve -> numpy write-back) stays with the device map
        p_h, s_h = gpu.lba_solve(wh, max_iters=6)
        assert abs(s_h["final_cost"] - s_dev["final_cost"]) <= 1e-6 * s_h["final_cost"]
        hm.write_back(p_h, cams_h, lines_h)
        for c in cams_h:
            assert np.abs(hm.T[c] - got[cams_d.index(c)]).max() < 1e-6
        # keep both maps identical for the next keyframe (the comparison above is per window, not accumulated)
        for i, c in enumerate(cams_d):
            hm.T[c] = got[i].copy()
        for j, l in enumerate(lines_h):
            hm.lms[l]["line"] = gl[j].copy()
        for i, c in enumerate(cams_d):                               # back to the world frame for the next odometry step
            if ba[c] < W:
                est[c] = replay.pose_compose(np.concatenate([log_so3(got[i][:9].reshape(3, 3)), got[i][9:]]), est[k]) if c != k else est[k]
        checked += 1
        tm = dm.last_timings()
        assert tm["h2d_bytes"] < 4096                                 # the call itself uploads only the window's keyframe list
    assert checked >= 10
    dm.close()


def test_map_errors(gpu):
    dm = gpu.DeviceMap(8, 16, 64)
    dm.add_keyframe(0, _T12(np.zeros(6)), [0, 1], np.zeros((2, 8)))
    with pytest.raises(gpu.SlslamError):
        dm.add_keyframe(0, _T12(np.zeros(6)), [], np.zeros((0, 8)))      # duplicate id
    with pytest.raises(gpu.SlslamError):
        dm.add_keyframe(9, _T12(np.zeros(6)), [], np.zeros((0, 8)))      # beyond the capacity
    with pytest.raises(gpu.SlslamError):
        dm.add_keyframe(1, _T12(np.zeros(6)), [99], np.zeros((1, 8)))    # landmark id out of range
    with pytest.raises(gpu.SlslamError):
        dm.bundle_adjust([0, 3], [0, 1], 2)                               # keyframe 3 is not in the map
    s = dm.bundle_adjust([0], [0], 2)                                     # nothing is seen twice: an empty window, zero cost
    assert dm.sizes[1] == 0 and s["initial_cost"] == 0.0 and s["iterations"] == 0
    dm.close()
