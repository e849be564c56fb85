"""GPU parity tests of the pose-graph path: CUDA (through the C ABI) against the CPU oracle on the same seeded inputs.
Reference behaviour under test: PoseConstraintError (src/po_problem.h:73-105), POProblem::build / set_options
(src/po_problem.cpp:40-77) and the ceres::Solve call at src/slam.cpp:1293."""
import json
import os

import numpy as np
import pytest

from slslam_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _oracle():
    from oracle import oracle
    return oracle


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def _graph_from_cases(p1, p2, c):
    n = len(p1)
    params = np.concatenate([p1, p2]).ravel().copy()
    return synth.PoseGraph(2 * n, np.arange(n, dtype=np.int32), np.arange(n, 2 * n, dtype=np.int32), c.ravel().copy(),
                           params, params.copy(), {})


def test_k5_golden_fixture(gpu):
    """K5 residual abs 1e-12, Jacobians rel 1e-9 against the committed torch-autograd vectors."""
    d = np.load(os.path.join(GOLD, "po_residual_cases.npz"))
    g = _graph_from_cases(d["p1"], d["p2"], d["c"])
    r, J1, J2, cost = gpu.po_evaluate(g)
    assert np.abs(r - d["r"]).max() < 1e-12
    assert np.abs(J1 - d["J1"]).max() < 1e-9 * max(1.0, np.abs(d["J1"]).max())
    assert np.abs(J2 - d["J2"]).max() < 1e-9 * max(1.0, np.abs(d["J2"]).max())
    assert _rel(cost, 0.5 * (d["r"] ** 2).sum()) < 1e-12


def test_k5_kat_vector(gpu):
    k = json.load(open(os.path.join(GOLD, "kat.json")))["po"]
    g = _graph_from_cases(np.array([k["p1"]]), np.array([k["p2"]]), np.array([k["c"]]))
    r, _, _, _ = gpu.po_evaluate(g)
    assert np.abs(r[0] - np.array(k["r"])).max() < 1e-12


def test_k5_vs_oracle_on_graph(gpu):
    oracle = _oracle()
    g = synth.make_pose_graph(3, num_poses=40, neighbours=2, num_loops=3)
    r, J1, J2, cost = gpu.po_evaluate(g)
    for e in range(g.num_edges):
        a, b = g.pose_index_1[e], g.pose_index_2[e]
        ro, J1o, J2o = oracle.po_residual_jacobian(g.parameters[6 * a:6 * a + 6], g.parameters[6 * b:6 * b + 6],
                                                   g.constraints[6 * e:6 * e + 6])
        assert np.abs(r[e] - ro).max() < 1e-12
        assert np.abs(J1[e] - J1o).max() < 1e-9 * max(1.0, np.abs(J1o).max())
        assert np.abs(J2[e] - J2o).max() < 1e-9 * max(1.0, np.abs(J2o).max())
    assert _rel(cost, oracle.po_cost(g)) < 1e-12


def test_consistent_edge_has_zero_residual(gpu):
    g = synth.make_pose_graph(5, num_poses=30, neighbours=1, num_loops=0, noise_rot=0.0, noise_tr=0.0)
    r, _, _, cost = gpu.po_evaluate(g, params=g.truth)
    assert np.abs(r).max() < 1e-12 and cost < 1e-24


def _compare(gpu, g, max_iters=10, lm_opts=None, tol_cost=1e-6):
    oracle = _oracle()
    po, so = oracle.po_solve(g, max_iters=max_iters, lm_opts=lm_opts)
    pg, sg = gpu.po_solve(g, max_iters=max_iters, lm_opts=lm_opts)
    assert _rel(sg["initial_cost"], so["initial_cost"]) < 1e-11
    n = min(sg["iterations"], so["iterations"])
    agree = True
    for k in range(n):
        if sg["trace"][k, 5] != so["trace"][k, 5]:
            agree = False
            break
        tol = 1e-9 * 10 ** min(k, 4)
        assert _rel(sg["trace"][k, 0], so["trace"][k, 0]) < tol, (k, sg["trace"][k], so["trace"][k])
        assert _rel(sg["trace"][k, 3], so["trace"][k, 3]) < 1e-6
    if agree:
        assert sg["iterations"] == so["iterations"]
        assert sg["num_successful_steps"] == so["num_successful_steps"]
        assert sg["termination"] == so["termination"]
    # (a consistent graph converges to cost ~ 1e-21, where a relative difference means nothing: absolute floor)
    assert abs(sg["final_cost"] - so["final_cost"]) < tol_cost * so["final_cost"] + 1e-18, (sg["final_cost"], so["final_cost"])
    return pg, sg, po, so


@pytest.mark.parametrize("seed,K,nbr,loops", [(0, 24, 2, 2), (1, 60, 3, 4), (2, 33, 1, 1)])
def test_solve_small_graphs(gpu, seed, K, nbr, loops):
    g = synth.make_pose_graph(seed, num_poses=K, neighbours=nbr, num_loops=loops)
    pg, sg, po, so = _compare(gpu, g)
    assert sg["final_cost"] < sg["initial_cost"]
    # pose parity with the oracle (rad / m); pose idx1[0] is the constant anchor (po_problem.cpp:62-63)
    assert np.abs(pg - po).max() < 1e-6
    a = g.pose_index_1[0]
    assert np.array_equal(pg[6 * a:6 * a + 6], g.parameters[6 * a:6 * a + 6])


def test_solve_myungdong_scale(gpu):
    """BASELINE.json configs[4] substitute: 261 poses, neighbour + loop-closure edges."""
    g = synth.make_pose_graph(0)
    pg, sg, po, so = _compare(gpu, g)
    assert np.abs(pg - po).max() < 1e-5


def test_panel_boundaries(gpu):
    """Reduced sizes straddling the 32-wide panels and the 64-wide trailing tiles of the dense factorisation."""
    for K in (6, 7, 12, 17, 23):       # n = 6 (K - 1): 30, 36, 66, 96, 132
        g = synth.make_pose_graph(10 + K, num_poses=K, neighbours=2, num_loops=1)
        _compare(gpu, g)


def test_unused_pose_untouched_and_zero_iterations(gpu):
    g = synth.make_pose_graph(4, num_poses=20, neighbours=1, num_loops=1)
    params = np.concatenate([g.parameters, [9.0, 8.0, 7.0, 6.0, 5.0, 4.0]])
    g2 = synth.PoseGraph(21, g.pose_index_1, g.pose_index_2, g.constraints, params, params.copy(), {})
    pg, sg = gpu.po_solve(g2, max_iters=5)
    assert np.array_equal(pg[-6:], params[-6:])
    p0, s0 = gpu.po_solve(g, max_iters=0)
    assert np.array_equal(p0, g.parameters) and s0["iterations"] == 0 and s0["initial_cost"] == s0["final_cost"]


def test_determinism_and_errors(gpu):
    g = synth.make_pose_graph(6, num_poses=30, neighbours=2, num_loops=2)
    a, sa = gpu.po_solve(g)
    b, sb = gpu.po_solve(g)
    assert np.array_equal(a, b) and sa["final_cost"] == sb["final_cost"]
    bad = synth.PoseGraph(g.num_poses, g.pose_index_1.copy(), g.pose_index_2, g.constraints, g.parameters, g.truth, {})
    bad.pose_index_1[2] = 999
    with pytest.raises(gpu.SlslamError) as e:
        gpu.po_solve(bad)
    assert e.value.code == -1
    nanp = g.parameters.copy(); nanp[7] = np.inf
    with pytest.raises(gpu.SlslamError) as e:
        gpu.po_solve(g, params=nanp)
    assert e.value.code == -4


def test_sparse_and_dense_paths_agree(gpu, monkeypatch):
    """The block-sparse factorisation (minimum-degree order, one CTA, one launch per LM iteration) is the default; the
    dense blocked factorisation remains for nearly full factors.  Same graph through both, both against the oracle
    (whose default is its own sparse Cholesky, cross-checked against its dense one)."""
    oracle = _oracle()
    for g in (synth.make_pose_graph(3, num_poses=90, neighbours=3, num_loops=6), synth.make_pose_graph(0)):
        pg, sg, po, so = _compare(gpu, g)
        st = gpu.po_last_stats()
        assert st["sparse"] == 2 and st["free_poses"] == g.num_poses - 1       # 2: level order, warp per column
        assert st["factor_blocks"] < 0.25 * st["free_poses"] * (st["free_poses"] + 1) / 2      # the factor is sparse
        assert st["iterations_enqueued"] <= 10
        # the minimum-degree order with the column-at-a-time kernel (what takes over when a column of the level order is
        # too long for the per-warp staging): same LM path
        monkeypatch.setenv("SLSLAM_PO_COLUMNS", "1")
        pc, sc = gpu.po_solve(g, max_iters=10)
        monkeypatch.delenv("SLSLAM_PO_COLUMNS")
        assert gpu.po_last_stats()["sparse"] == 1
        assert _rel(sc["final_cost"], sg["final_cost"]) < 1e-9 and sc["iterations"] == sg["iterations"]
        assert np.abs(pc - pg).max() < 1e-8
        # run to run the level kernel repeats itself bit for bit (no atomics: the stages order every update)
        pg2, sg2 = gpu.po_solve(g, max_iters=10)
        assert np.array_equal(pg2, pg) and sg2["final_cost"] == sg["final_cost"]
        # ... and whether the level kernel runs as a cluster of CTAs or as one CTA (what it falls back to when the cluster
        # cannot be launched) changes nothing: the stages fix the order of every update
        monkeypatch.setenv("SLSLAM_PO_CLUSTER", "1")
        p1, s1 = gpu.po_solve(g, max_iters=10)
        monkeypatch.delenv("SLSLAM_PO_CLUSTER")
        assert gpu.po_last_stats()["sparse"] == 2 and np.array_equal(p1, pg) and s1["final_cost"] == sg["final_cost"]
        monkeypatch.setenv("SLSLAM_PO_DENSE", "1")
        pd, sd = gpu.po_solve(g, max_iters=10)
        monkeypatch.delenv("SLSLAM_PO_DENSE")
        assert gpu.po_last_stats()["sparse"] == 0
        assert _rel(sd["final_cost"], sg["final_cost"]) < 1e-9 and sd["iterations"] == sg["iterations"]
        assert np.abs(pd - pg).max() < 1e-8
        pod, sod = oracle.po_solve(g, max_iters=10, solver=0)
        assert _rel(sod["final_cost"], so["final_cost"]) < 1e-10 and np.abs(pod - po).max() < 1e-9


def test_sparse_structures(gpu):
    """Graph shapes that stress the symbolic plan: a star (one pose linked to all: the centre is eliminated last), a
    complete graph (no sparsity at all, still through the sparse kernel below 64 poses), two components joined by one
    edge, a pure chain, duplicate edges and a self edge."""
    rng = np.random.default_rng(0)
    base = synth.make_pose_graph(7, num_poses=40, neighbours=1, num_loops=0)

    def graph(pairs):
        truth = base.truth.reshape(-1, 6)
        e1 = np.asarray([a for a, _ in pairs], np.int32); e2 = np.asarray([b for _, b in pairs], np.int32)
        cons = np.stack([synth._compose(truth[b], synth._inverse(truth[a])) + rng.normal(0, 2e-3, 6) for a, b in pairs])
        return synth.PoseGraph(base.num_poses, e1, e2, cons.ravel(), base.parameters.copy(), base.truth, {})

    K = base.num_poses
    chain = [(k, k + 1) for k in range(K - 1)]
    cases = {
        "star": chain + [(5, k) for k in range(K) if abs(k - 5) > 1],
        "complete": [(a, b) for a in range(14) for b in range(a + 1, 14)] + [(k, k + 1) for k in range(13, K - 1)],
        "two_components": [(k, k + 1) for k in range(19)] + [(k, k + 1) for k in range(20, K - 1)] + [(3, 30)],
        "chain": chain,
        "duplicates_and_self": chain + chain[:7] + [(9, 9), (4, 17), (4, 17)],
    }
    for name, pairs in cases.items():
        g = graph(pairs)
        pg, sg, po, so = _compare(gpu, g)
        assert gpu.po_last_stats()["sparse"] in (1, 2), name
        assert np.abs(pg - po).max() < 1e-6, name
