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


# ---------------------------------------------------------------------------------------------------------------------
# Device path (round 2): the scattered buffer is solved WHERE NCCL PUT IT.  One buffer per rank, 256-byte aligned
# sections:  header | result region (parameters of every window back to back, then one 48-byte summary per window) |
# per window: camera_index, line_index, fixed_index, observations.  The solver (slslam_lba_solve_batch_device) reads the
# arrays in place, updates the parameters in place and writes the summaries into the result region, so the gather is
# ONE send of one contiguous slice per rank, rank 0 receives all slices into one device buffer and reads it back with
# ONE copy into page-locked memory.  No device -> host -> device round trip, no per-window copies.
# ---------------------------------------------------------------------------------------------------------------------
_ALIGN = 256
SUMMARY_BYTES = 48       # sizeof(slslam_summary): 4 doubles + 4 int32


def _up(x, a=_ALIGN):
    return (x + a - 1) // a * a


class RankLayout:
    """Byte offsets inside one rank's buffer, a pure function of the window shapes [(C, L, N), ...]."""

    def __init__(self, shapes):
        self.shapes = [tuple(int(v) for v in s) for s in shapes]
        n = len(self.shapes)
        self.header = 0
        o = _up(16 + 24 * n)
        self.result_begin = o
        self.param_off = []
        for (Cc, Ll, Nn) in self.shapes:
            self.param_off.append(o)
            o += 8 * ((6 * Cc + 4 * Ll + 1) // 2 * 2)
        self.summary_off = o
        o += SUMMARY_BYTES * n
        self.result_end = o
        o = _up(o)
        self.arrays = []
        for (Cc, Ll, Nn) in self.shapes:
            ci = o; li = ci + 4 * Nn; fi = li + 4 * Nn; ob = _up(fi + 8 * Nn, 16)
            self.arrays.append(dict(camera_index=ci, line_index=li, fixed_index=fi, observations=ob))
            o = _up(ob + 64 * Nn)
        self.total = max(o, _ALIGN)

    @property
    def result_bytes(self):
        return self.result_end - self.result_begin

    def offsets(self):
        return [dict(a, parameters=p) for a, p in zip(self.arrays, self.param_off)]


def pack_rank_buffer(windows, pin=False):
    """The windows of one rank in the RankLayout format (numpy uint8; page-locked when `pin`)."""
    lay = RankLayout([(w.num_cameras, w.num_lines, w.num_observations) for w in windows])
    if pin:
        import torch
        buf = torch.zeros(lay.total, dtype=torch.uint8).pin_memory().numpy()
    else:
        buf = np.zeros(lay.total, np.uint8)
    head = np.array([_MAGIC, len(windows)] + [v for s in lay.shapes for v in s], np.int64)
    buf[:head.nbytes] = head.view(np.uint8)
    for w, a, po in zip(windows, lay.arrays, lay.param_off):
        N = w.num_observations
        buf[a["camera_index"]:a["camera_index"] + 4 * N] = np.ascontiguousarray(w.camera_index, np.int32).view(np.uint8)
        buf[a["line_index"]:a["line_index"] + 4 * N] = np.ascontiguousarray(w.line_index, np.int32).view(np.uint8)
        buf[a["fixed_index"]:a["fixed_index"] + 8 * N] = np.ascontiguousarray(w.fixed_index, np.int32).view(np.uint8)
        buf[a["observations"]:a["observations"] + 64 * N] = np.ascontiguousarray(w.observations, np.float64).view(np.uint8)
        p = np.ascontiguousarray(w.parameters, np.float64)
        buf[po:po + p.nbytes] = p.view(np.uint8)
    return buf, lay


def unpack_results(region: np.ndarray, lay: RankLayout):
    """Result region (host bytes) of one rank -> (parameter arrays, summary dicts)."""
    ps, ss = [], []
    base = lay.result_begin
    for i, (Cc, Ll, Nn) in enumerate(lay.shapes):
        o = lay.param_off[i] - base
        ps.append(region[o:o + 8 * (6 * Cc + 4 * Ll)].view(np.float64).copy())
        so = lay.summary_off - base + SUMMARY_BYTES * i
        d = region[so:so + 32].view(np.float64)
        k = region[so + 32:so + 48].view(np.int32)
        ss.append(dict(initial_cost=float(d[0]), final_cost=float(d[1]), fixed_cost=float(d[2]), gradient_max_norm=float(d[3]),
                       num_successful_steps=int(k[0]), num_unsuccessful_steps=int(k[1]),
                       termination=_TERM[int(k[2])] if 0 <= int(k[2]) < len(_TERM) else "?", iterations=int(k[3])))
    return ps, ss


class DeviceSharder:
    """Scatter -> solve in place -> gather for a fixed set of window shapes, every buffer allocated once.

    rank 0 constructs it with the full window list, the other ranks with None; the shapes travel in one broadcast at
    construction.  `origin`: "host" = rank 0's buffers start in page-locked host memory and are copied to its GPU
    destination by destination, each NCCL send starting as soon as its chunk has landed (rank r solves while later
    chunks are still on their way); "device" = they start in rank 0's HBM (what `value` assumes for its inputs) and the
    scatter is NVLink only.  `solve(recv)` is injected by the caller (the product passes capi.lba_solve_batch_device).
    """

    def __init__(self, windows, device, group=None, pin=True, local_windows=None):
        """`windows`: the full list on rank 0 (None elsewhere).  Alternatively `local_windows` on EVERY rank: the windows
        rank r is going to own (global index i * world + r for its i-th); rank 0 then collects the packed buffers of all
        ranks once, at construction -- a set-up convenience for benchmarks whose ranks generate their windows in parallel."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.device(device)
        cuda = self.dev.type == "cuda"
        if local_windows is not None:
            mine = torch.zeros(1, dtype=torch.int64, device=self.dev)
            mine[0] = len(local_windows)
            counts = [torch.zeros(1, dtype=torch.int64, device=self.dev) for _ in range(self.world)]
            dist.all_gather(counts, mine, group=group)
            counts = [int(c.item()) for c in counts]
            if any(c != counts[0] for c in counts):
                raise ValueError("local_windows: every rank must own the same number of windows (w -> rank w mod world)")
            self.num_windows = sum(counts)
            sh_mine = torch.tensor([[w.num_cameras, w.num_lines, w.num_observations] for w in local_windows], dtype=torch.int64,
                                   device=self.dev).reshape(-1, 3)
            sh_all = [torch.zeros_like(sh_mine) for _ in range(self.world)]
            dist.all_gather(sh_all, sh_mine, group=group)
            per_rank = [t.cpu().numpy() for t in sh_all]
            shapes = np.zeros((self.num_windows, 3), np.int64)
            for r in range(self.world):
                for i, w in enumerate(local_indices(self.num_windows, r, self.world)):
                    shapes[w] = per_rank[r][i]
        else:
            nwin = torch.zeros(1, dtype=torch.int64, device=self.dev)
            if self.rank == 0:
                nwin[0] = len(windows)
            dist.broadcast(nwin, 0, group=group)
            self.num_windows = int(nwin.item())
            shapes = torch.zeros(max(1, self.num_windows), 3, dtype=torch.int64, device=self.dev)
            if self.rank == 0:
                shapes[:self.num_windows] = torch.tensor([[w.num_cameras, w.num_lines, w.num_observations] for w in windows], dtype=torch.int64)
            dist.broadcast(shapes, 0, group=group)
            shapes = shapes.cpu().numpy()[:self.num_windows]
        self.layouts = [RankLayout([shapes[w] for w in local_indices(self.num_windows, r, self.world)]) for r in range(self.world)]
        self.lay = self.layouts[self.rank]
        self.recv = torch.zeros(self.lay.total, dtype=torch.uint8, device=self.dev)        # this rank's windows, solved in place
        if local_windows is not None:
            own = pack_rank_buffer(local_windows, pin=pin and cuda)[0]
            if self.rank == 0:
                self.host_bufs = [own]
                for r in range(1, self.world):
                    t = torch.zeros(self.layouts[r].total, dtype=torch.uint8, device=self.dev)
                    for req in dist.batch_isend_irecv([dist.P2POp(dist.irecv, t, r, group=group)]):
                        req.wait()
                    h = torch.zeros(self.layouts[r].total, dtype=torch.uint8)
                    if pin and cuda:
                        h = h.pin_memory()
                    h.copy_(t)
                    self.host_bufs.append(h.numpy())
            else:
                for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, torch.from_numpy(own).to(self.dev), 0, group=group)]):
                    req.wait()
        elif self.rank == 0:
            self.host_bufs = [pack_rank_buffer([windows[w] for w in local_indices(self.num_windows, r, self.world)], pin=pin and cuda)[0]
                              for r in range(self.world)]
        self.res_off = np.concatenate([[0], np.cumsum([_up(self.layouts[r].result_bytes) for r in range(self.world)])]).astype(np.int64)
        self.shm = None
        if self.rank == 0:
            self.stage = [torch.zeros(self.layouts[r].total, dtype=torch.uint8, device=self.dev) for r in range(self.world)]
            self.res_dev = torch.zeros(int(self.res_off[-1]), dtype=torch.uint8, device=self.dev)
            self.res_host = torch.zeros(int(self.res_off[-1]), dtype=torch.uint8)
            if cuda:
                self.res_host = self.res_host.pin_memory()
        self._pending = []

    def enable_shared_host(self):
        """origin = "host_shared" (one node): the packed windows live in ONE page-locked host segment that every rank's
        process maps (POSIX shared memory, registered with CUDA in each process).  A rank then pulls its own slice over
        its OWN PCIe link and writes its results back the same way: N links instead of rank 0's one, no NCCL on the data
        path (a one-element all-reduce joins the ranks at the end).  This is how a single host producer -- the SLAM front
        end -- would feed N GPUs; the NCCL origins above remain for data that starts on rank 0's GPU or in its private
        memory.  Set-up (untimed): rank 0 copies its packed buffers into the segment."""
        from multiprocessing import resource_tracker, shared_memory
        torch, dist = self.torch, self.dist
        self.seg_off = np.concatenate([[0], np.cumsum([_up(self.layouts[r].total) for r in range(self.world)])]).astype(np.int64)
        self.seg_res = int(self.seg_off[-1])
        size = self.seg_res + int(self.res_off[-1])
        name, err = [None], None
        self._registered = False
        if self.rank == 0:
            try:
                self.shm = shared_memory.SharedMemory(create=True, size=size)
                name[0] = self.shm.name
            except Exception as e:          # e.g. /dev/shm too small in a container
                err = f"shared memory segment of {size} bytes: {e}"
        dist.broadcast_object_list(name, 0, group=self.group)
        try:
            if name[0] is None:
                raise RuntimeError(err or "rank 0 could not create the shared segment")
            if self.rank != 0:
                self.shm = shared_memory.SharedMemory(name=name[0])
                try:                  # only the creator unlinks (Python < 3.13 tracks attachments as if it owned them)
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:
                    pass
            self.seg_np = np.ndarray((size,), np.uint8, buffer=self.shm.buf)
            self.seg = torch.from_numpy(self.seg_np)
            if self.dev.type == "cuda":
                rc = torch.cuda.cudart().cudaHostRegister(self.seg.data_ptr(), size, 0)
                if int(rc) != 0:
                    raise RuntimeError(f"cudaHostRegister of the shared window segment failed: {rc}")
                self._registered = True
        except Exception as e:
            err = str(e)
        # every rank must have the segment, or nobody uses it
        good = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.dev)
        dist.all_reduce(good, op=dist.ReduceOp.MIN, group=self.group)
        if int(good.item()) == 0:
            self.shared_host_error = err or "another rank could not map the shared segment"
            self._drop_segment()
            return False
        if self.rank == 0:
            for r in range(self.world):
                o = int(self.seg_off[r])
                self.seg_np[o:o + self.layouts[r].total] = self.host_bufs[r]
        self._join = torch.zeros(1, dtype=torch.int32, device=self.dev)
        dist.barrier(group=self.group)
        return True

    def _drop_segment(self):
        if getattr(self, "_registered", False):
            self.torch.cuda.cudart().cudaHostUnregister(self.seg.data_ptr())
            self._registered = False
        for a_ in ("seg", "seg_np"):
            if hasattr(self, a_):
                delattr(self, a_)
        shm, self.shm = self.shm, None
        if shm is not None:
            try:
                shm.close()
            except BufferError:
                pass
            if self.rank == 0:
                try:
                    shm.unlink()
                except FileNotFoundError:
                    pass

    def close_shared_host(self):
        if self.shm is None:
            return
        if self.dev.type == "cuda":
            self.torch.cuda.synchronize(self.dev)
        self.dist.barrier(group=self.group)
        self._drop_segment()

    def preload_device(self):
        """origin = "device": put rank 0's packed buffers into its HBM (untimed; the windows then start on the device)."""
        if self.rank == 0:
            for r in range(self.world):
                self.stage[r].copy_(self.torch.from_numpy(self.host_bufs[r]))
            if self.dev.type == "cuda":
                self.torch.cuda.synchronize(self.dev)

    def scatter(self, origin="host"):
        """After this call self.recv holds this rank's buffer (device).  Asynchronous on CUDA: the recv is enqueued."""
        torch, dist = self.torch, self.dist
        if origin == "host_shared":
            o = int(self.seg_off[self.rank])
            self.recv.copy_(self.seg[o:o + self.lay.total], non_blocking=True)     # this rank's own PCIe link
            return
        if self.rank == 0 and origin == "device" and self.dev.type == "cuda":
            # everything is already in HBM: ONE grouped launch sends to all peers side by side (separate sends would
            # queue behind one another on NCCL's stream: 0.54 ms instead of ~0.15 ms for 7 x 6.9 MB, measured)
            ops = [dist.P2POp(dist.isend, self.stage[r], r, group=self.group) for r in range(1, self.world)]
            if ops:
                self._pending += dist.batch_isend_irecv(ops)
            self.recv.copy_(self.stage[0], non_blocking=True)
        elif self.rank == 0:
            order = list(range(1, self.world)) + [0]                      # the other ranks first: they wait for us
            for r in order:
                if origin == "host":
                    src = torch.from_numpy(self.host_bufs[r])
                    (self.recv if r == 0 else self.stage[r]).copy_(src, non_blocking=True)
                elif r == 0:
                    self.recv.copy_(self.stage[0], non_blocking=True)
                if r != 0:
                    # an individual send per destination: it is ordered behind the copy just enqueued on the current
                    # stream, and the next destination's copy overlaps it
                    # (batched-API form even for one op: on NCCL an unbatched send is treated as a collective of the whole
                    # group and serialised with every other operation)
                    if self.dev.type == "cuda":
                        self._pending += dist.batch_isend_irecv([dist.P2POp(dist.isend, self.stage[r], r, group=self.group)])
                    else:
                        dist.send(self.stage[r], r, group=self.group)
        else:
            if self.dev.type == "cuda":
                self._pending += dist.batch_isend_irecv([dist.P2POp(dist.irecv, self.recv, 0, group=self.group)])
            else:
                dist.recv(self.recv, 0, group=self.group)

    def wait_scatter(self):
        """Orders the current stream behind the receive (no host wait on CUDA): call before the solve is enqueued."""
        for req in self._pending:
            req.wait()
        self._pending = []

    def views(self):
        """Host-side helper for CPU tests: numpy views of this rank's windows inside self.recv (CPU tensors only)."""
        buf = self.recv.numpy()
        out = []
        for (Cc, Ll, Nn), off in zip(self.lay.shapes, self.lay.offsets()):
            out.append(Window(Cc, Ll, buf[off["camera_index"]:off["camera_index"] + 4 * Nn].view(np.int32),
                              buf[off["line_index"]:off["line_index"] + 4 * Nn].view(np.int32),
                              buf[off["fixed_index"]:off["fixed_index"] + 8 * Nn].view(np.int32),
                              buf[off["observations"]:off["observations"] + 64 * Nn].view(np.float64),
                              buf[off["parameters"]:off["parameters"] + 8 * (6 * Cc + 4 * Ll)].view(np.float64), None, {}))
        return out

    def store_summary(self, i, s):
        """CPU tests: write window i's summary into the result region the way the device kernel does."""
        buf = self.recv.numpy()
        o = self.lay.summary_off + SUMMARY_BYTES * i
        buf[o:o + 32].view(np.float64)[:] = [s["initial_cost"], s["final_cost"], s.get("fixed_cost", 0.0), s.get("gradient_max_norm", 0.0)]
        buf[o + 32:o + 48].view(np.int32)[:] = [s["num_successful_steps"], s["num_unsuccessful_steps"],
                                                _TERM.index(s["termination"]) if s["termination"] in _TERM else -1, s["iterations"]]

    def gather(self):
        """Rank 0 returns (params, summaries) of all windows in window order; the others (None, None)."""
        if not self.gather_raw():
            return None, None
        return self.unpack_gathered()

    def gather_raw(self):
        """The transfer alone: afterwards rank 0's page-locked self.res_host holds every rank's result region (rank 0
        has waited for the copy; the other ranks have only enqueued their send).  True on rank 0."""
        torch, dist = self.torch, self.dist
        lay = self.lay
        mine = self.recv[lay.result_begin:lay.result_end]
        if self.rank != 0:
            if lay.result_bytes:
                if self.dev.type == "cuda":
                    for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, 0, group=self.group)]):
                        req.wait()
                else:
                    dist.send(mine, 0, group=self.group)
            return False
        ops = []
        for r in range(1, self.world):
            n = self.layouts[r].result_bytes
            if n:
                ops.append(dist.P2POp(dist.irecv, self.res_dev[int(self.res_off[r]):int(self.res_off[r]) + n], r, group=self.group))
        self.res_dev[:lay.result_bytes].copy_(mine, non_blocking=True)
        _exchange(ops, dist)
        self.res_host.copy_(self.res_dev, non_blocking=True)               # ONE device -> host copy for all ranks
        if self.dev.type == "cuda":
            torch.cuda.current_stream(self.dev).synchronize()
        return True

    def gather_raw_shared(self):
        """origin = "host_shared": every rank copies its result region into the shared host segment over its own link; a
        one-element all-reduce, stream-ordered behind the copies, tells rank 0 that all of them have landed."""
        lay = self.lay
        if lay.result_bytes:
            o = self.seg_res + int(self.res_off[self.rank])
            self.seg[o:o + lay.result_bytes].copy_(self.recv[lay.result_begin:lay.result_end], non_blocking=True)
        self.dist.all_reduce(self._join, group=self.group)
        if self.rank != 0:
            return False
        if self.dev.type == "cuda":
            self.torch.cuda.current_stream(self.dev).synchronize()
        return True

    def unpack_gathered(self, shared=False):
        """Rank 0: (params, summaries) of all windows in window order from self.res_host (or the shared segment)."""
        host = self.seg_np[self.seg_res:] if shared else self.res_host.numpy()
        out_p, out_s = [None] * self.num_windows, [None] * self.num_windows
        for r in range(self.world):
            n = self.layouts[r].result_bytes
            ps, ss = unpack_results(host[int(self.res_off[r]):int(self.res_off[r]) + n], self.layouts[r])
            for w, p, s_ in zip(local_indices(self.num_windows, r, self.world), ps, ss):
                out_p[w], out_s[w] = p, s_
        return out_p, out_s
