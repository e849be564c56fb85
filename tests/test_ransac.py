"""RANSAC hypothesis scoring (reference src/slam.cpp:398-412, 691-726): oracle known answers on
CPU, bit-exact GPU parity (scores, inlier masks and the float errors themselves)."""
import numpy as np
import pytest

from slslam_b200 import synth

BASELINE, THR = 0.12, 5.0 / 406.05


def _oracle():
    from oracle import oracle
    oracle.lib()
    return oracle


def make_case(seed, n_lines=120, n_hyp=64, sigma_px=0.3):
    return synth.ransac_case(seed, n_lines, n_hyp, sigma_px, baseline=BASELINE)


def test_oracle_known_answers():
    oracle = _oracle()
    poses, lines, obs, (R, t) = make_case(0, sigma_px=0.0)
    scores, inl, err = oracle.ransac_score(poses, lines, obs, BASELINE, THR)
    assert scores[0] == lines.shape[0] and err[0].max() < 1e-6          # the true motion explains noise-free observations
    assert (scores[np.linalg.norm(poses[:, 9:], axis=1) > 1] == -1).all() and (scores == -1).sum() >= 3
    ok = scores >= 0
    assert (scores[ok] == inl[ok].sum(1)).all() and (scores[ok] <= scores[0]).all()
    # independent numpy restatement of SLAM::reprojection_error for one pair (double precision; the oracle's float steps
    # move the result by < 1e-6 relative)
    h, k = 7, 11
    Rh, th = poses[h, :9].reshape(3, 3), poses[h, 9:].copy()
    e = 0.0
    for i in range(2):
        if i == 1:
            th[0] -= BASELINE
        n = np.cross(Rh @ lines[k, :3] + th, Rh @ lines[k, 3:])
        n = n / np.hypot(n[0], n[1])
        e += abs(n @ [obs[k, 4 * i], obs[k, 4 * i + 1], 1]) + abs(n @ [obs[k, 4 * i + 2], obs[k, 4 * i + 3], 1])
    assert abs(err[h, k] - e / 4) <= 2e-6 * max(e / 4, 1e-3)
    # the threshold is 5 px: with 0.3 px noise the true motion keeps every line, a 0.1 rad error keeps almost none
    poses, lines, obs, _ = make_case(1, sigma_px=0.3)
    scores, _, _ = oracle.ransac_score(poses, lines, obs, BASELINE, THR)
    assert scores[0] == lines.shape[0]
    bad = poses[:1].copy(); bad[0, :9] = (synth.rodrigues(np.array([0.0, 0.1, 0.0])) @ bad[0, :9].reshape(3, 3)).ravel()
    assert oracle.ransac_score(bad, lines, obs, BASELINE, THR)[0][0] < 0.2 * lines.shape[0]


def _reproj_error_numpy(ft, R, t, line, baseline):
    """SLAM::reprojection_error (reference src/slam.cpp:691-726) in numpy scalars with the reference's types: `error`
    and `sql` are float32, everything else float64; sums in source order.  Written independently of the C++ oracle."""
    f64, f32 = np.float64, np.float32
    error = f32(0)
    t = [f64(t[0]), f64(t[1]), f64(t[2])]
    cp, dv = [f64(v) for v in line[:3]], [f64(v) for v in line[3:]]
    R = [[f64(R[3 * r + c]) for c in range(3)] for r in range(3)]
    for i in range(2):
        if i == 1:
            t[0] = t[0] - f64(baseline)
        p1, p2 = [f64(ft[4 * i]), f64(ft[4 * i + 1]), f64(1)], [f64(ft[4 * i + 2]), f64(ft[4 * i + 3]), f64(1)]
        cpc = [((R[r][0] * cp[0] + R[r][1] * cp[1]) + R[r][2] * cp[2]) + t[r] for r in range(3)]
        dvc = [(R[r][0] * dv[0] + R[r][1] * dv[1]) + R[r][2] * dv[2] for r in range(3)]
        nc = [cpc[1] * dvc[2] - cpc[2] * dvc[1], cpc[2] * dvc[0] - cpc[0] * dvc[2], cpc[0] * dvc[1] - cpc[1] * dvc[0]]
        sql = f32(np.sqrt(nc[0] * nc[0] + nc[1] * nc[1]))
        nc = [v / f64(sql) for v in nc]
        error = f32(f64(error) + abs((nc[0] * p1[0] + nc[1] * p1[1]) + nc[2] * p1[2]))
        error = f32(f64(error) + abs((nc[0] * p2[0] + nc[1] * p2[1]) + nc[2] * p2[2]))
    return f32(f64(error) / f64(4.0))


def test_oracle_is_bit_exact_against_numpy_restatement():
    oracle = _oracle()
    poses, lines, obs, _ = make_case(7, n_lines=40, n_hyp=12)
    scores, inl, err = oracle.ransac_score(poses, lines, obs, BASELINE, THR)
    for h in range(poses.shape[0]):
        if np.linalg.norm(poses[h, 9:]) > 1:
            continue
        want = np.array([_reproj_error_numpy(obs[k], poses[h, :9], poses[h, 9:], lines[k], BASELINE) for k in range(lines.shape[0])],
                        np.float32)
        assert np.array_equal(want.view(np.uint32), err[h].view(np.uint32)), h
        assert np.array_equal((want.astype(np.float64) < THR).astype(np.uint8), inl[h]) and scores[h] == inl[h].sum()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_lines,n_hyp", [(0, 120, 64), (1, 300, 500), (2, 1, 3), (3, 257, 1), (4, 1000, 2000)])
def test_gpu_scoring_is_bit_exact(gpu, seed, n_lines, n_hyp):
    oracle = _oracle()
    poses, lines, obs, _ = make_case(seed, n_lines, n_hyp)
    so, io, eo = oracle.ransac_score(poses, lines, obs, BASELINE, THR)
    sg, ig, eg = gpu.ransac_score(poses, lines, obs, BASELINE, THR)
    assert np.array_equal(sg, so)
    assert np.array_equal(ig, io)
    assert np.array_equal(eg.view(np.uint32), eo.view(np.uint32))            # the float errors, bit for bit
    s2, i2, e2 = gpu.ransac_score(poses, lines, obs, BASELINE, THR, want_inliers=False, want_errors=False)
    assert np.array_equal(s2, so) and i2 is None and e2 is None


@pytest.mark.gpu
def test_gpu_scoring_edge_cases(gpu):
    poses, lines, obs, _ = make_case(5, 10, 4)
    s, i, e = gpu.ransac_score(poses[:0], lines, obs)
    assert s.shape == (0,)
    s, i, e = gpu.ransac_score(poses, lines[:0], obs[:0])
    assert (s[np.linalg.norm(poses[:, 9:], axis=1) <= 1] == 0).all()
