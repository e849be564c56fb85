"""Multi-GPU host logic on CPU: world_size-2 gloo runs of the window scatter / gather.  The solver is
injected; here it is the CPU oracle (tests may use it), on the GPU box it is capi.lba_solve_batch."""
import os
import socket
import sys

import numpy as np
import pytest

from slslam_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_and_pack_round_trip():
    assert shard.local_indices(7, 1, 3) == [1, 4]
    assert sorted(sum((shard.local_indices(64, r, 8) for r in range(8)), [])) == list(range(64))
    w = synth.make_window(3, 4, 30, 100, num_fixed_cameras=2)
    b = shard.pack_window(w)
    assert b.dtype == np.uint8 and b.size == shard.packed_size(w.num_cameras, w.num_lines, w.num_observations)
    u = shard.unpack_window(b)
    for k in ("camera_index", "line_index", "fixed_index", "observations", "parameters"):
        assert np.array_equal(getattr(u, k), getattr(w, k))
    with pytest.raises(ValueError):
        shard.unpack_window(b[:-8])
    # device-path layout: aligned sections, result region contiguous, round trip through the packed buffer
    ws = [synth.make_window(3 + i, 4, 30 + i, 100 + 3 * i) for i in range(3)]
    buf, lay = shard.pack_rank_buffer(ws)
    assert buf.size == lay.total and lay.total % 256 == 0 and lay.result_begin % 256 == 0
    assert all(o["observations"] % 16 == 0 and o["parameters"] % 16 == 0 for o in lay.offsets())
    assert lay.param_off[0] == lay.result_begin and lay.summary_off + 48 * 3 == lay.result_end
    ps, ss = shard.unpack_results(buf[lay.result_begin:lay.result_end], lay)
    assert all(np.array_equal(p, w_.parameters) for p, w_ in zip(ps, ws)) and len(ss) == 3
    lay2 = shard.RankLayout([(w_.num_cameras, w_.num_lines, w_.num_observations) for w_ in ws])
    assert lay2.offsets() == lay.offsets() and lay2.total == lay.total
    assert shard.RankLayout([]).total == 256 and shard.RankLayout([]).result_bytes == 0
    m = synth.window_M(0)
    # ~0.87 MB per M window
    assert abs(shard.pack_window(m).size - 864_560) < 4000


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle
    from slslam_b200 import shard as sh, synth as sy
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def solve(ws, max_iters):
        res = [oracle.lba_solve(w, max_iters=max_iters, solver=1) for w in ws]
        return [p for p, _ in res], [s for _, s in res]

    windows = [sy.make_window(40 + i, 4, 30 + 3 * i, 110 + 7 * i, sigma_px=0.5) for i in range(5)] if rank == 0 else None
    local, idx = sh.scatter_windows(windows)
    np.save(os.path.join(out_dir, f"idx{rank}.npy"), np.array(idx))
    ps, ss = sh.solve_sharded(windows, solve, max_iters=6)
    # the packed form the bench uses: per-rank buffers, known result sizes (no size all-gather)
    bufs = sh.pack_for_ranks(windows, world) if rank == 0 else None
    mine = sh.scatter_packed(bufs)
    mine_w = [windows[w] for w in sh.local_indices(5, 0, world)] if rank == 0 else sh.unpack_many(mine)
    p2, s2 = solve(mine_w, 6)
    sizes = sh.result_sizes(windows, world) if rank == 0 else [0] * world
    gp, gs = sh.gather_results(p2, s2, sh.local_indices(5, rank, world), 5, sizes=sizes)
    if rank == 0:
        assert all(np.array_equal(a, b) for a, b in zip(gp, ps)) and [x["final_cost"] for x in gs] == [x["final_cost"] for x in ss]
        assert sizes == [sum(8 + 1 + 6 * windows[w].num_cameras + 4 * windows[w].num_lines for w in sh.local_indices(5, r, world)) for r in range(world)]
    # the device path's host logic (layout, in-place solve, one result slice per rank) on CPU tensors over gloo
    ds = sh.DeviceSharder(windows, "cpu")
    assert ds.num_windows == 5 and [len(l.shapes) for l in ds.layouts] == [3, 2]
    for origin in ("host", "device"):
        if origin == "device":
            ds.preload_device()
        ds.scatter(origin=origin)
        ds.wait_scatter()
        for i, w in enumerate(ds.views()):
            p_, s_ = oracle.lba_solve(w, max_iters=6, solver=1)
            w.parameters[:] = p_                       # in place, as slslam_lba_solve_batch_device does
            ds.store_summary(i, s_)
        dp_, dss_ = ds.gather()
        if rank == 0:
            assert all(np.array_equal(a, b) for a, b in zip(dp_, ps))
            for a, b in zip(dss_, ss):
                assert all(a[k] == b[k] for k in ("initial_cost", "final_cost", "iterations", "termination", "num_successful_steps"))
        else:
            assert dp_ is None and dss_ is None
    # every rank contributes its own windows (global index i * world + rank); rank 0 ends up holding all of them
    mine_l = [sy.make_window(70 + 2 * i + rank, 3, 20 + i, 60 + 5 * i, sigma_px=0.5) for i in range(2)]
    ds2 = sh.DeviceSharder(None, "cpu", local_windows=mine_l)
    assert ds2.num_windows == 4
    ds2.scatter(origin="host"); ds2.wait_scatter()
    for w_, m_ in zip(ds2.views(), mine_l):
        assert np.array_equal(w_.observations, m_.observations) and np.array_equal(w_.parameters, m_.parameters)
        assert np.array_equal(w_.camera_index, m_.camera_index) and np.array_equal(w_.fixed_index, m_.fixed_index)
        w_.parameters[:] = w_.parameters + 1.0
    for i in range(2):
        ds2.store_summary(i, dict(initial_cost=1.0, final_cost=0.5 + rank, num_successful_steps=1, num_unsuccessful_steps=0,
                                  termination="NO_CONVERGENCE", iterations=1))
    gp2, gs2 = ds2.gather()
    if rank == 0:
        assert [s_["final_cost"] for s_ in gs2] == [0.5, 1.5, 0.5, 1.5]
        assert np.array_equal(gp2[0], mine_l[0].parameters + 1.0) and np.array_equal(gp2[2], mine_l[1].parameters + 1.0)
    # the shared page-locked segment path (one node): every rank pulls its slice from / writes its results to ONE host
    # segment mapped by all processes; no send / recv at all
    ds2.enable_shared_host()
    ds2.scatter(origin="host_shared"); ds2.wait_scatter()
    for w_, m_ in zip(ds2.views(), mine_l):
        assert np.array_equal(w_.observations, m_.observations) and np.array_equal(w_.parameters, m_.parameters)
        w_.parameters[:] = w_.parameters + 2.0
    for i in range(2):
        ds2.store_summary(i, dict(initial_cost=1.0, final_cost=2.5 + rank, num_successful_steps=1, num_unsuccessful_steps=0,
                                  termination="NO_CONVERGENCE", iterations=1))
    if ds2.gather_raw_shared():
        gp3, gs3 = ds2.unpack_gathered(shared=True)
        assert [s_["final_cost"] for s_ in gs3] == [2.5, 3.5, 2.5, 3.5]
        assert np.array_equal(gp3[0], mine_l[0].parameters + 2.0) and np.array_equal(gp3[2], mine_l[1].parameters + 2.0)
        del gp3, gs3
    else:
        assert rank != 0
    ds2.close_shared_host()
    if rank == 0:
        np.save(os.path.join(out_dir, "cost.npy"), np.array([s["final_cost"] for s in ss]))
        np.save(os.path.join(out_dir, "iters.npy"), np.array([s["iterations"] for s in ss]))
        np.savez(os.path.join(out_dir, "params.npz"), *ps)
    else:
        assert ps is None and ss is None
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_solve_gather_world2(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert list(np.load(tmp_path / "idx0.npy")) == [0, 2, 4] and list(np.load(tmp_path / "idx1.npy")) == [1, 3]
    cost, iters = np.load(tmp_path / "cost.npy"), np.load(tmp_path / "iters.npy")
    params = np.load(tmp_path / "params.npz")
    for i in range(5):
        w = synth.make_window(40 + i, 4, 30 + 3 * i, 110 + 7 * i, sigma_px=0.5)
        p, s = oracle.lba_solve(w, max_iters=6, solver=1)
        assert cost[i] == s["final_cost"] and iters[i] == s["iterations"]
        assert np.array_equal(params[f"arr_{i}"], p)
