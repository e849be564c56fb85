"""The symbolic plan of the pose-graph factorisation (host code of the product, no device needed): slslam_po_plan_check
builds what slslam_po_solve would build -- level order (independent poses side by side, stages of columns with disjoint
blocks) or minimum-degree order -- and verifies its invariants in C++: rows later than their column, update destinations
present, columns of a stage pairwise non-adjacent with private destination blocks and right-hand-side rows."""
import numpy as np
import pytest

from slslam_b200 import capi, synth


def _graph(num_poses, pairs):
    e1 = np.asarray([a for a, _ in pairs], np.int32)
    e2 = np.asarray([b for _, b in pairs], np.int32)
    cons = np.zeros(6 * len(pairs))
    return synth.PoseGraph(num_poses, e1, e2, cons, np.zeros(6 * num_poses), np.zeros(6 * num_poses), {})


def _cases():
    rng = np.random.default_rng(5)
    K = 120
    chain = [(k, k + 1) for k in range(K - 1)]
    yield "chain", _graph(K, chain)
    yield "band3 + loops", _graph(K, [(k, k + d) for d in (1, 2, 3) for k in range(K - d)] + [(3, 90), (20, 110), (45, 70)])
    yield "star", _graph(K, chain + [(5, k) for k in range(K) if abs(k - 5) > 1])
    yield "complete 14 + tail", _graph(K, [(a, b) for a in range(14) for b in range(a + 1, 14)] + [(k, k + 1) for k in range(13, K - 1)])
    yield "two components", _graph(K, [(k, k + 1) for k in range(59)] + [(k, k + 1) for k in range(60, K - 1)] + [(3, 90)])
    yield "duplicates and self edges", _graph(K, chain + chain[:9] + [(9, 9), (4, 17), (4, 17)])
    yield "random sparse", _graph(K, chain + [tuple(sorted(rng.choice(K, 2, replace=False))) for _ in range(40)])
    yield "grid 10 x 12", _graph(K, [(12 * r + c, 12 * r + c + 1) for r in range(10) for c in range(11)] + [(12 * r + c, 12 * r + 12 + c) for r in range(9) for c in range(12)])
    g = synth.make_pose_graph(0)
    yield "bench-like band graph", g
    traj = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "traj_myungdong_wolc.npy"))
    yield "myungdong + 10 loops", synth.pose_graph_from_trajectory(traj, seed=0, num_loops=10)


@pytest.mark.parametrize("name,g", list(_cases()))
def test_plan_invariants(name, g):
    info = capi.po_plan_check(g)
    assert info["free_poses"] == g.num_poses - 1 or name == "two components" or info["free_poses"] <= g.num_poses
    cols = capi.po_plan_check(g, force_columns=True)
    assert cols["order"] in (0, 1)
    if info["order"] == 2:
        assert info["stages"] >= 1 and info["widest_stage"] >= 1 and info["max_column_rows"] <= 12
        # the level order trades some fill for parallelism, not an order of magnitude of it
        if cols["order"] == 1:
            assert info["factor_blocks"] <= 2.5 * cols["factor_blocks"], (info, cols)
    if name == "star":
        assert info["order"] == 1          # a hub adjacent to every pose leaves nothing to eliminate side by side
    if name in ("chain", "myungdong + 10 loops", "bench-like band graph", "band3 + loops"):
        assert info["order"] == 2 and info["widest_stage"] >= 8, info       # trajectory graphs are what the level order is for
        assert info["stages"] < 0.5 * info["free_poses"]


def test_plan_check_rejects_bad_graphs():
    g = _graph(5, [(0, 1), (1, 7)])
    with pytest.raises(capi.SlslamError) as e:
        capi.po_plan_check(g)
    assert e.value.code == -1
