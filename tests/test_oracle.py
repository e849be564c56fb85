"""CPU tests: the oracle against the committed golden fixtures, the known-answer tests of SURVEY.md §8c and
independent optimiser minima.  The oracle is 'parity unpinned' (no reference golden vectors exist, the reference
cannot be built here): these tests are what stands in for that pin."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from slslam_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def test_kat_vectors():
    k = json.load(open(os.path.join(G, "kat.json")))
    r, _, _ = oracle.lba_residual_jacobian(k["lba"]["cam"], k["lba"]["line"], k["lba"]["obs"])
    assert np.abs(r - np.array(k["lba"]["r"])).max() < 1e-15
    r, _, _ = oracle.po_residual_jacobian(k["po"]["p1"], k["po"]["p2"], k["po"]["c"])
    assert np.abs(r - np.array(k["po"]["r"])).max() < 1e-15


def test_lba_residual_jacobian_vs_golden():
    d = np.load(os.path.join(G, "lba_residual_cases.npz"))
    for i in range(len(d["cam"])):
        r, Jc, Jl = oracle.lba_residual_jacobian(d["cam"][i], d["line"][i], d["obs"][i])
        assert np.abs(r - d["r"][i]).max() < 1e-13
        assert np.abs(Jc - d["Jc"][i]).max() < 2e-9 * max(1.0, np.abs(Jc).max())
        assert np.abs(Jl - d["Jl"][i]).max() < 2e-9 * max(1.0, np.abs(Jl).max())


def test_po_residual_jacobian_vs_golden():
    d = np.load(os.path.join(G, "po_residual_cases.npz"))
    for i in range(len(d["p1"])):
        r, J1, J2 = oracle.po_residual_jacobian(d["p1"][i], d["p2"][i], d["c"][i])
        assert np.abs(r - d["r"][i]).max() < 1e-12
        assert np.abs(J1 - d["J1"][i]).max() < 1e-9 and np.abs(J2 - d["J2"][i]).max() < 1e-9


def test_geometric_kat_noise_free_projection():
    """A noise-free stereo projection of a 3-D segment has |r| <= 1e-14 (SURVEY.md §8c anchor 1)."""
    w = synth.make_window(0, 6, 60, 300, sigma_px=0.0, start="exact")
    C = w.num_cameras
    for i in range(w.num_observations):
        cam = w.truth[6 * w.camera_index[i]:6 * w.camera_index[i] + 6]
        ln = w.truth[6 * C + 4 * w.line_index[i]:6 * C + 4 * w.line_index[i] + 4]
        r, _, _ = oracle.lba_residual_jacobian(cam, ln, w.observations[8 * i:8 * i + 8])
        assert np.abs(r).max() < 1e-13


def test_orth_round_trip():
    rng = np.random.default_rng(1)
    for _ in range(200):
        P = rng.normal(size=3) * 3 + np.array([0, 0, 6]); d = rng.normal(size=3); d /= np.linalg.norm(d)
        cp = P - d * np.dot(P, d)
        cp2, d2 = synth.orth_to_av(synth.av_to_orth(cp, d))
        assert np.abs(cp2 - cp).max() < 1e-9 and min(np.abs(d2 - d).max(), np.abs(d2 + d).max()) < 1e-9


def test_po_consistent_edge_has_zero_residual():
    rng = np.random.default_rng(2)
    for _ in range(200):
        T1 = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 2, 3)])
        Cc = np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 1, 3)])
        T2 = synth._compose(Cc, T1)
        r, _, _ = oracle.po_residual_jacobian(T1, T2, Cc)
        assert np.abs(r).max() < 1e-13


def test_jacobian_vs_central_differences():
    rng = np.random.default_rng(3)
    w = synth.window_S(0)
    C = w.num_cameras
    for i in rng.integers(0, w.num_observations, 10):
        cam = w.parameters[6 * w.camera_index[i]:6 * w.camera_index[i] + 6].copy()
        ln = w.parameters[6 * C + 4 * w.line_index[i]:6 * C + 4 * w.line_index[i] + 4].copy()
        ob = w.observations[8 * i:8 * i + 8]
        _, Jc, Jl = oracle.lba_residual_jacobian(cam, ln, ob)
        h = 1e-6
        for j in range(6):
            a, b = cam.copy(), cam.copy(); a[j] += h; b[j] -= h
            fd = (oracle.lba_residual_jacobian(a, ln, ob)[0] - oracle.lba_residual_jacobian(b, ln, ob)[0]) / (2 * h)
            assert np.abs(fd - Jc[:, j]).max() < 1e-6 * max(1.0, np.abs(Jc).max())
        for j in range(4):
            a, b = ln.copy(), ln.copy(); a[j] += h; b[j] -= h
            fd = (oracle.lba_residual_jacobian(cam, a, ob)[0] - oracle.lba_residual_jacobian(cam, b, ob)[0]) / (2 * h)
            assert np.abs(fd - Jl[:, j]).max() < 1e-6 * max(1.0, np.abs(Jl).max())


def test_minima_vs_scipy():
    m = json.load(open(os.path.join(G, "minima.json")))
    for c in m["lba"]:
        w = synth.make_window(c["seed"], num_cameras=c["num_cameras"], num_lines=c["num_lines"],
                              num_observations=c["num_observations"], sigma_px=c["sigma_px"], start=c["start"])
        for solver in (0, 1):
            _, s = oracle.lba_solve(w, max_iters=300, robust=False, solver=solver, lm_opts=[1e-14, 1e-16, 1e-14, 1e4])
            assert abs(s["final_cost"] / c["cost"] - 1) < 1e-9
    for c in m["po"]:
        g = synth.make_pose_graph(c["seed"], num_poses=c["num_poses"], neighbours=c["neighbours"], num_loops=c["num_loops"])
        _, s = oracle.po_solve(g, max_iters=100, lm_opts=[1e-14, 1e-16, 1e-14, 1e4])
        assert abs(s["final_cost"] / c["cost"] - 1) < 1e-9


def test_schur_equals_full_normal_equations():
    """The reference always solves the full normal equations (lba_problem.cpp:96-101); Schur is the same step."""
    w = synth.window_S(2, sigma_px=0.5)
    pa, sa = oracle.lba_solve(w, max_iters=6, solver=0)
    pb, sb = oracle.lba_solve(w, max_iters=6, solver=1)
    assert sa["iterations"] == sb["iterations"]
    assert np.allclose(sa["trace"][:, 0], sb["trace"][:, 0], rtol=1e-6, atol=0)
    assert abs(sa["final_cost"] / sb["final_cost"] - 1) < 1e-6


def test_summary_semantics():
    w = synth.motion_only_window(5)
    p, s = oracle.lba_solve(w, max_iters=10)
    assert s["fixed_cost"] > 0                       # cam-1 residual blocks are all-constant but still counted (Q6)
    assert np.array_equal(p[6:], w.parameters[6:])   # constant camera and constant lines untouched
    assert s["final_cost"] <= s["initial_cost"]
    assert abs(s["initial_cost"] - oracle.lba_cost(w)) < 1e-15
    p0, s0 = oracle.lba_solve(w, max_iters=0)
    assert np.array_equal(p0, w.parameters) and s0["iterations"] == 0


def test_huber_semantics():
    """HuberLoss(1/406.05) acts on the squared norm of the 4-vector (lba_problem.cpp:78-80)."""
    w = synth.window_S(1, sigma_px=3.0)
    robust, plain = oracle.lba_cost(w, robust=True), oracle.lba_cost(w, robust=False)
    assert robust < plain
    a = 1 / 406.05
    C = w.num_cameras
    tot = 0.0
    for i in range(w.num_observations):
        r, _, _ = oracle.lba_residual_jacobian(w.parameters[6 * w.camera_index[i]:6 * w.camera_index[i] + 6],
                                               w.parameters[6 * C + 4 * w.line_index[i]:6 * C + 4 * w.line_index[i] + 4],
                                               w.observations[8 * i:8 * i + 8])
        s = float(r @ r)
        tot += 0.5 * (s if s <= a * a else 2 * a * np.sqrt(s) - a * a)
    assert abs(tot / robust - 1) < 1e-12


def test_po_sparse_cholesky_matches_dense():
    """The oracle's sparse Cholesky (what SPARSE_NORMAL_CHOLESKY ends in: ordering, elimination tree, up-looking
    factorisation) against its dense Cholesky on the same normal equations: same LM path to rounding."""
    from oracle import oracle
    for seed, K, nbr, loops in ((0, 24, 2, 2), (1, 60, 3, 5), (2, 33, 1, 0), (3, 120, 2, 9)):
        g = synth.make_pose_graph(seed, num_poses=K, neighbours=nbr, num_loops=loops)
        p1, s1 = oracle.po_solve(g, max_iters=10, solver=1, want_stats=True)
        p0, s0 = oracle.po_solve(g, max_iters=10, solver=0)
        assert s1["iterations"] == s0["iterations"] and s1["termination"] == s0["termination"]
        assert abs(s1["final_cost"] - s0["final_cost"]) <= 1e-11 * s0["final_cost"] + 1e-20
        assert np.abs(p1 - p0).max() < 1e-10
        n = 6 * (K - 1)
        assert s1["factor_nnz"] < 0.5 * n * (n + 1) / 2
