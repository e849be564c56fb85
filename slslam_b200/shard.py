"""Sharding independent LBA windows over the GPUs of one node (BASELINE.json configs[3]).

Only whole windows shard: inside the live pipeline consecutive windows depend on each other (reference
src/slam.cpp:957-972 -> :1317-1366), but a batch of independent windows partitions trivially.  Window w goes to rank
w mod G.  Rank 0 holds the packed windows; the exchange is one grouped point-to-point scatter of the packed byte
buffers (NCCL grouped ncclSend/ncclRecv under `torch.distributed.batch_isend_irecv`; gloo on CPU for the tests) and
one gather of parameters + summaries.  There is no collective inside the LM loop.

The solver is injected (`solve_fn(windows, max_iters) -> (params_list, summaries)`): the product passes
`capi.lba_solve_batch`; this module never imports a CPU solver.
"""
from __future__ import annotations

import numpy as np

from .synth import Window

_MAGIC = 0x534C4241  # 'SLBA'
SUMMARY_WIDTH = 8    # initial_cost final_cost fixed_cost gradient_max_norm successful unsuccessful termination iterations
_TERM = ["NO_CONVERGENCE", "GRADIENT_TOLERANCE", "FUNCTION_TOLERANCE", "PARAMETER_TOLERANCE", "NUMERICAL_FAILURE"]


def owner(window: int, world: int) -> int:
    return window % world


def local_indices(num_windows: int, rank: int, world: int) -> list[int]:
    return [w for w in range(num_windows) if owner(w, world) == rank]


def packed_size(C: int, L: int, N: int) -> int:
    """Bytes of one packed window: 4 int32 header, 4N int32 indices/flags, 8N observations, 6C+4L parameters."""
    return 16 + 4 * 4 * N + 8 * (8 * N + 6 * C + 4 * L)


def pack_window(w: Window) -> np.ndarray:
    """Window -> flat uint8 buffer in the reference's array layout (reference src/lba_problem.h:188-196)."""
    N = w.num_observations
    head = np.array([_MAGIC, w.num_cameras, w.num_lines, N], np.int32)
    parts = [head, np.ascontiguousarray(w.camera_index, np.int32), np.ascontiguousarray(w.line_index, np.int32),
             np.ascontiguousarray(w.fixed_index, np.int32), np.ascontiguousarray(w.observations, np.float64),
             np.ascontiguousarray(w.parameters, np.float64)]
    buf = np.concatenate([p.view(np.uint8).ravel() for p in parts])
    assert buf.size == packed_size(w.num_cameras, w.num_lines, N)
    return buf


def unpack_window(buf: np.ndarray) -> Window:
    buf = np.ascontiguousarray(buf, np.uint8)
    head = buf[:16].view(np.int32)
    if int(head[0]) != _MAGIC:
        raise ValueError("not a packed LBA window")
    C, L, N = int(head[1]), int(head[2]), int(head[3])
    if buf.size != packed_size(C, L, N):
        raise ValueError("packed LBA window has the wrong length")
    o = 16
    # views into the received buffer (no copies; the solver entry points never write into a window's arrays)
    ci = buf[o:o + 4 * N].view(np.int32); o += 4 * N
    li = buf[o:o + 4 * N].view(np.int32); o += 4 * N
    fi = buf[o:o + 8 * N].view(np.int32); o += 8 * N
    ob = buf[o:o + 64 * N].view(np.float64); o += 64 * N
    pr = buf[o:].view(np.float64)
    return Window(C, L, ci, li, fi, ob, pr, pr, {})


def summary_to_row(s: dict) -> np.ndarray:
    t = s["termination"]
    return np.array([s["initial_cost"], s["final_cost"], s.get("fixed_cost", 0.0), s.get("gradient_max_norm", 0.0),
                     s["num_successful_steps"], s["num_unsuccessful_steps"], _TERM.index(t) if t in _TERM else -1,
                     s["iterations"]], np.float64)


def row_to_summary(r: np.ndarray) -> dict:
    return dict(initial_cost=float(r[0]), final_cost=float(r[1]), fixed_cost=float(r[2]), gradient_max_norm=float(r[3]),
                num_successful_steps=int(r[4]), num_unsuccessful_steps=int(r[5]),
                termination=_TERM[int(r[6])] if 0 <= int(r[6]) < len(_TERM) else "?", iterations=int(r[7]))


_PINNED = {}


def _to_device(arr: np.ndarray, dev):
    """Host array -> tensor on `dev`.  For CUDA the bytes go through a cached pinned staging tensor (allocating pinned
    memory per call costs more than the copy)."""
    import torch
    t = torch.from_numpy(arr)
    if dev.type != "cuda":
        return t
    key = (t.dtype, dev.index)
    stage = _PINNED.get(key)
    if stage is None or stage.numel() < t.numel():
        stage = torch.empty(max(t.numel(), 1 << 16), dtype=t.dtype).pin_memory()
        _PINNED[key] = stage
    view = stage[:t.numel()]
    view.copy_(t)
    return view.to(dev)          # blocking: the staging tensor is reused for the next buffer


def _exchange(ops, dist):
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def pack_for_ranks(windows, world: int, pin: bool = False):
    """Rank-0 side assembly: one contiguous byte buffer per destination rank, holding that rank's windows back to back
    (int64 count, int64 byte sizes, then the packed windows).  Window w goes to rank w mod world.  With `pin` the buffers
    live in page-locked memory (numpy views of pinned torch tensors), so the scatter can DMA them without staging."""
    out = []
    for r in range(world):
        mine = [pack_window(windows[w]) for w in local_indices(len(windows), r, world)]
        head = np.array([len(mine)] + [m.size for m in mine], np.int64).view(np.uint8)
        parts = [head] + mine
        total = sum(p.size for p in parts)
        if pin:
            import torch
            buf = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory().numpy()[:total]
        else:
            buf = np.empty(total, np.uint8)
        o = 0
        for p in parts:
            buf[o:o + p.size] = p
            o += p.size
        out.append(buf)
    return out


def _is_pinned(arr) -> bool:
    try:
        import torch
        return torch.from_numpy(arr).is_pinned()
    except Exception:
        return False


def unpack_many(buf: np.ndarray):
    buf = np.ascontiguousarray(buf, np.uint8)
    cnt = int(buf[:8].view(np.int64)[0])
    sizes = buf[8:8 + 8 * cnt].view(np.int64)
    o, ws = 8 + 8 * cnt, []
    for sz in sizes:
        ws.append(unpack_window(buf[o:o + int(sz)]))
        o += int(sz)
    return ws


def scatter_packed(bufs, device=None, group=None):
    """Rank 0 passes the per-rank buffers of `pack_for_ranks` (other ranks None); every rank returns its own buffer as a
    host uint8 array.  One size broadcast, then ONE send per destination rank, grouped (`batch_isend_irecv` = grouped
    ncclSend / ncclRecv on NCCL; gloo on CPU)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cpu") if device is None else torch.device(device)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    if rank == 0:
        sizes[:] = torch.tensor([b.size for b in bufs], dtype=torch.int64)
    dist.broadcast(sizes, 0, group=group)
    sizes = sizes.cpu().numpy()
    ops, keep, recv = [], [], None
    if rank == 0:
        for r in range(1, world):
            if dev.type == "cuda" and _is_pinned(bufs[r]):
                t = torch.from_numpy(bufs[r]).to(dev, non_blocking=True)     # all copies in flight, stream-ordered before the sends
            else:
                t = _to_device(bufs[r], dev)
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, r, group=group))
    else:
        recv = torch.empty(int(sizes[rank]), dtype=torch.uint8, device=dev)
        ops.append(dist.P2POp(dist.irecv, recv, 0, group=group))
    _exchange(ops, dist)
    if rank == 0:
        return bufs[0]
    if dev.type != "cuda":
        return recv.numpy()
    # device -> page-locked host (cached): the solver DMAs the observations straight out of this buffer again
    key = ("recv", dev.index)
    stage = _PINNED.get(key)
    if stage is None or stage.numel() < recv.numel():
        stage = torch.empty(max(recv.numel(), 1 << 16), dtype=torch.uint8).pin_memory()
        _PINNED[key] = stage
    view = stage[:recv.numel()]
    view.copy_(recv, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return view.numpy()


def scatter_windows(windows, device=None, group=None):
    """Rank 0 passes the full list (other ranks pass None); every rank returns (its windows, their global indices)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cpu") if device is None else torch.device(device)
    meta = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        meta[0] = len(windows)
    dist.broadcast(meta, 0, group=group)
    n = int(meta.item())
    bufs = pack_for_ranks(windows, world) if rank == 0 else None
    mine = scatter_packed(bufs, device=device, group=group)
    idx = local_indices(n, rank, world)
    local = [windows[w] for w in idx] if rank == 0 else unpack_many(mine)
    return local, idx


def result_sizes(windows, world: int):
    """Float64 elements of every rank's result buffer, computable on rank 0 from the windows it scattered."""
    return [sum(SUMMARY_WIDTH + 1 + 6 * windows[w].num_cameras + 4 * windows[w].num_lines
                for w in local_indices(len(windows), r, world)) for r in range(world)]


def gather_results(params, summaries, indices, num_windows, device=None, group=None, sizes=None):
    """Every rank passes the parameters / summaries of its windows (global `indices`); rank 0 returns the full lists in
    window order, other ranks return (None, None).  One float64 buffer per rank: per window the summary row, the
    parameter count and the parameters; one size all-gather, then one grouped send/recv per rank."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cpu") if device is None else torch.device(device)
    parts = []
    for p, s in zip(params, summaries):
        parts += [summary_to_row(s), np.array([float(len(p))]), np.ascontiguousarray(p, np.float64)]
    mine = np.concatenate(parts) if parts else np.zeros(0)
    if sizes is None:
        # rank 0 does not know the other ranks' sizes: one small all-gather (skipped when the caller passes `sizes`,
        # which every rank can compute when it knows the window shapes; only rank 0 uses them)
        sz = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sz, torch.tensor([mine.size], dtype=torch.int64, device=dev), group=group)
        sizes = [int(t.item()) for t in sz]
    ops, keep, bufs = [], [], {}
    if rank == 0:
        for r in range(1, world):
            if sizes[r]:
                bufs[r] = torch.empty(sizes[r], dtype=torch.float64, device=dev)
                ops.append(dist.P2POp(dist.irecv, bufs[r], r, group=group))
    elif mine.size:
        t = _to_device(mine, dev)
        keep.append(t)
        ops.append(dist.P2POp(dist.isend, t, 0, group=group))
    _exchange(ops, dist)
    if rank != 0:
        return None, None
    out_p, out_s = [None] * num_windows, [None] * num_windows
    for p, s, w in zip(params, summaries, indices):
        out_p[w], out_s[w] = np.asarray(p, np.float64), s
    for r, t in bufs.items():
        a = t.cpu().numpy()
        o = 0
        for w in local_indices(num_windows, r, world):
            row = a[o:o + SUMMARY_WIDTH]; npar = int(a[o + SUMMARY_WIDTH]); o += SUMMARY_WIDTH + 1
            out_s[w] = row_to_summary(row)
            out_p[w] = a[o:o + npar].copy(); o += npar
    return out_p, out_s


def solve_sharded(windows, solve_fn, max_iters=10, device=None, group=None):
    """Scatter -> every rank solves its windows with `solve_fn` -> gather.  Rank 0 returns (params, summaries) for
    all windows in order; other ranks (None, None)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    local, idx = scatter_windows(windows if rank == 0 else None, device=device, group=group)
    cnt = torch.tensor([len(windows) if rank == 0 else 0], dtype=torch.int64,
                       device=torch.device("cpu") if device is None else torch.device(device))
    dist.broadcast(cnt, 0, group=group)
    if local:
        ps, ss = solve_fn(local, max_iters)
    else:
        ps, ss = [], []
    return gather_results(ps, ss, idx, int(cnt.item()), device=device, group=group)
