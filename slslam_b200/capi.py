"""ctypes binding of libslslam_b200.so (the C ABI in include/slslam_b200.h).

Used by tests/, bench.py and __graft_entry__.py.  It only marshals numpy arrays into the C structs: no arithmetic
of the hot path lives here, and there is no fallback — if the library is missing or no sm_100 device is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLSLAM_B200_LIB", os.path.join(_HERE, "libslslam_b200.so"))
TRACE_WIDTH = 8
TERMINATION = {0: "NO_CONVERGENCE", 1: "GRADIENT_TOLERANCE", 2: "FUNCTION_TOLERANCE", 3: "PARAMETER_TOLERANCE",
               4: "NUMERICAL_FAILURE"}

EXPORTS = [
    "slslam_version", "slslam_strerror", "slslam_last_error", "slslam_device_count", "slslam_lba_get_limits", "slslam_measure_fp64_peak",
    "slslam_lba_solve", "slslam_lba_solve_batch", "slslam_lba_solve_batch_device", "slslam_lba_last_timings", "slslam_lba_batch_create", "slslam_lba_batch_solve",
    "slslam_lba_batch_upload_params", "slslam_lba_batch_download", "slslam_lba_batch_info", "slslam_lba_batch_max_active_clusters", "slslam_lba_batch_transfer_bytes", "slslam_lba_batch_phase_cycles", "slslam_lba_batch_plan_cycles", "slslam_lba_batch_destroy",
    "slslam_lba_plan_check", "slslam_lba_launch_shape", "slslam_lba_pipeline_create", "slslam_lba_pipeline_submit", "slslam_lba_pipeline_wait", "slslam_lba_pipeline_destroy",
    "slslam_ransac_score", "slslam_lba_evaluate", "slslam_po_solve", "slslam_po_solve_trace", "slslam_po_evaluate", "slslam_po_last_solve_ms", "slslam_po_last_stats", "slslam_po_get_limits", "slslam_po_plan_check", "slslam_lba_route",
    "slslam_map_create", "slslam_map_destroy", "slslam_map_add_keyframe", "slslam_map_add_landmarks", "slslam_map_set_poses",
    "slslam_map_get_poses", "slslam_map_get_landmarks", "slslam_map_bundle_adjust", "slslam_map_last_timings", "slslam_map_last_window",
    "slslam_geometry_convert",
]

dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)


class LbaDesc(C.Structure):
    _fields_ = [("num_cameras", C.c_int32), ("num_lines", C.c_int32), ("num_observations", C.c_int32),
                ("max_iterations", C.c_int32), ("camera_index", ip), ("line_index", ip), ("fixed_index", ip),
                ("observations", dp), ("robust", C.c_int32), ("huber_delta", C.c_double), ("baseline", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double)]


class PoDesc(C.Structure):
    _fields_ = [("num_poses", C.c_int32), ("num_edges", C.c_int32), ("max_iterations", C.c_int32),
                ("pose_index_1", ip), ("pose_index_2", ip), ("constraints", dp),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_trust_region_radius", C.c_double)]


class Summary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("fixed_cost", C.c_double),
                ("gradient_max_norm", C.c_double), ("num_successful_steps", C.c_int32),
                ("num_unsuccessful_steps", C.c_int32), ("termination_type", C.c_int32), ("iterations", C.c_int32)]


class PoStats(C.Structure):
    _fields_ = [("sparse", C.c_int32), ("free_poses", C.c_int32), ("factor_blocks", C.c_int64),
                ("block_updates", C.c_int64), ("max_column_rows", C.c_int32), ("iterations_enqueued", C.c_int32),
                ("factor_cycles", C.c_int64 * 4)]


class PoPlanInfo(C.Structure):
    _fields_ = [("order", C.c_int32), ("free_poses", C.c_int32), ("stages", C.c_int32), ("widest_stage", C.c_int32),
                ("max_column_rows", C.c_int32), ("reserved", C.c_int32), ("factor_blocks", C.c_int64), ("block_updates", C.c_int64)]


class PoLimits(C.Structure):
    _fields_ = [("max_column_blocks_sparse", C.c_int32), ("max_free_poses_dense", C.c_int32)]


class Limits(C.Structure):
    _fields_ = [("max_cameras", C.c_int32), ("max_free_cameras", C.c_int32),
                ("max_observations_per_line", C.c_int32), ("max_cluster_size", C.c_int32),
                ("max_free_cameras_general", C.c_int32)]


class SlslamError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__(f"slslam error {code}: {detail}")


_LIB = None


def build_library(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a build of the in-tree shared library (cross-compiles without a GPU)."""
    args = ["make", "-C", _HERE, "-s", "libslslam_b200.so"]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SlslamError(-3, f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        L.slslam_version.restype = C.c_int
        L.slslam_strerror.restype = C.c_char_p
        L.slslam_strerror.argtypes = [C.c_int]
        L.slslam_last_error.restype = C.c_char_p
        L.slslam_device_count.restype = C.c_int
        L.slslam_lba_get_limits.argtypes = [C.POINTER(Limits)]
        L.slslam_lba_solve.argtypes = [C.POINTER(LbaDesc), dp, C.POINTER(Summary)]
        L.slslam_lba_solve_batch.argtypes = [C.c_int32, C.POINTER(LbaDesc), C.POINTER(dp), C.POINTER(Summary)]
        L.slslam_lba_solve_batch_device.argtypes = [C.c_int32, C.POINTER(LbaDesc), C.POINTER(dp), C.c_void_p, C.POINTER(Summary),
                                                    C.c_void_p]
        L.slslam_lba_batch_create.argtypes = [C.c_int32, C.POINTER(LbaDesc), C.POINTER(dp), C.c_int32, C.c_int32,
                                              C.POINTER(C.c_void_p)]
        L.slslam_lba_batch_solve.argtypes = [C.c_void_p, C.c_void_p]
        L.slslam_lba_batch_upload_params.argtypes = [C.c_void_p, C.POINTER(dp), C.c_void_p]
        L.slslam_lba_batch_download.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(dp), C.POINTER(Summary), C.POINTER(dp)]
        L.slslam_lba_batch_info.argtypes = [C.c_void_p, ip, ip, ip, ip]
        L.slslam_lba_last_timings.argtypes = [dp]
        L.slslam_lba_last_timings.restype = None
        L.slslam_lba_batch_max_active_clusters.argtypes = [C.c_void_p, ip]
        L.slslam_lba_batch_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.slslam_lba_batch_phase_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.c_int32]
        L.slslam_lba_batch_plan_cycles.argtypes = [C.c_void_p, C.c_int32, ip]
        L.slslam_lba_batch_destroy.argtypes = [C.c_void_p]
        L.slslam_lba_batch_destroy.restype = None
        L.slslam_lba_launch_shape.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, ip, ip]
        L.slslam_lba_plan_check.argtypes = [C.c_int32, C.POINTER(LbaDesc), C.POINTER(dp), C.c_int32, ip]
        L.slslam_lba_pipeline_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.slslam_lba_pipeline_submit.argtypes = [C.c_void_p, C.c_int32, C.POINTER(LbaDesc), C.POINTER(dp), C.POINTER(Summary),
                                                 C.POINTER(C.c_int64)]
        L.slslam_lba_pipeline_wait.argtypes = [C.c_void_p, C.c_int64]
        L.slslam_lba_pipeline_destroy.argtypes = [C.c_void_p]
        L.slslam_lba_pipeline_destroy.restype = None
        L.slslam_ransac_score.argtypes = [C.c_int32, dp, C.c_int32, dp, dp, C.c_double, C.c_double, ip,
                                          C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
        L.slslam_lba_evaluate.argtypes = [C.POINTER(LbaDesc), dp, dp, dp, dp, dp]
        L.slslam_po_solve.argtypes = [C.POINTER(PoDesc), dp, C.POINTER(Summary)]
        L.slslam_po_solve_trace.argtypes = [C.POINTER(PoDesc), dp, C.POINTER(Summary), dp]
        L.slslam_po_evaluate.argtypes = [C.POINTER(PoDesc), dp, dp, dp, dp, dp]
        L.slslam_po_last_solve_ms.restype = C.c_float
        L.slslam_po_last_stats.argtypes = [C.POINTER(PoStats)]
        L.slslam_po_last_stats.restype = None
        L.slslam_po_get_limits.argtypes = [C.POINTER(PoLimits)]
        L.slslam_po_get_limits.restype = None
        L.slslam_lba_route.argtypes = [C.POINTER(LbaDesc)]
        L.slslam_lba_route.restype = C.c_int32
        L.slslam_po_plan_check.argtypes = [C.POINTER(PoDesc), C.c_int32, C.POINTER(PoPlanInfo)]
        L.slslam_po_plan_check.restype = C.c_int32
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        L = lib()
        raise SlslamError(rc, f"{L.slslam_strerror(rc).decode()} [{L.slslam_last_error().decode()}]")


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


class _Keep:
    """A desc plus the numpy arrays it points into."""

    def __init__(self, desc, arrays):
        self.desc, self.arrays = desc, arrays


def lba_desc(w, max_iters=10, robust=True, huber_delta=0.0, baseline=-1.0, lm_opts=None) -> _Keep:
    ci = np.ascontiguousarray(w.camera_index, np.int32)
    li = np.ascontiguousarray(w.line_index, np.int32)
    fi = np.ascontiguousarray(w.fixed_index, np.int32)
    ob = np.ascontiguousarray(w.observations, np.float64)
    o = [0.0, 0.0, 0.0, 0.0] if lm_opts is None else list(lm_opts)
    d = LbaDesc(w.num_cameras, w.num_lines, int(ci.shape[0]), max_iters, _i(ci), _i(li), _i(fi), _d(ob), int(robust),
                huber_delta, baseline, o[0], o[1], o[2], o[3])
    return _Keep(d, (ci, li, fi, ob))


def summary_dict(s: Summary, trace=None):
    d = dict(initial_cost=s.initial_cost, final_cost=s.final_cost, fixed_cost=s.fixed_cost,
             gradient_max_norm=s.gradient_max_norm, num_successful_steps=s.num_successful_steps,
             num_unsuccessful_steps=s.num_unsuccessful_steps, termination=TERMINATION.get(s.termination_type, "?"),
             iterations=s.iterations)
    if trace is not None:
        d["trace"] = trace[:s.iterations].copy()
    return d


def lba_solve(w, params=None, **kw):
    """One window through slslam_lba_solve (host buffers in, host buffers out)."""
    k = lba_desc(w, **kw)
    p = np.ascontiguousarray(w.parameters if params is None else params, np.float64).copy()
    s = Summary()
    _check(lib().slslam_lba_solve(C.byref(k.desc), _d(p), C.byref(s)))
    return p, summary_dict(s)


def measure_fp64_peak(device=-1):
    """(TFLOP/s, SM clock MHz) of a register-only DFMA kernel on `device` (slslam_measure_fp64_peak)."""
    L = lib()
    L.slslam_measure_fp64_peak.argtypes = [C.c_int32, dp, dp]
    t, c = C.c_double(), C.c_double()
    _check(L.slslam_measure_fp64_peak(device, C.cast(C.byref(t), dp), C.cast(C.byref(c), dp)))
    return t.value, c.value


def last_timings():
    """plan | stage + H2D enqueue | launch + solve + D2H | copy-out | total (ms) of the last one-shot solve."""
    t = np.zeros(8)
    lib().slslam_lba_last_timings(_d(t))
    return dict(zip(("plan_ms", "stage_ms", "solve_ms", "copy_out_ms", "total_ms", "dev_h2d_ms", "dev_kernel_ms", "dev_d2h_ms"),
                    t.tolist()))


def lba_solve_batch(windows, **kw):
    keeps = [lba_desc(w, **kw) for w in windows]
    descs = (LbaDesc * len(windows))(*[k.desc for k in keeps])
    ps = [np.ascontiguousarray(w.parameters, np.float64).copy() for w in windows]
    pp = (dp * len(windows))(*[_d(p) for p in ps])
    ss = (Summary * len(windows))()
    _check(lib().slslam_lba_solve_batch(len(windows), descs, pp, ss))
    return ps, [summary_dict(s) for s in ss]


def lba_solve_batch_device(shapes, base_ptr, offsets, max_iters=10, robust=True, summaries_dev_ptr=None,
                           want_host_summaries=True, stream=None, lm_opts=None):
    """slslam_lba_solve_batch_device: windows whose arrays already live in DEVICE memory (for instance in a buffer NCCL
    received).  shapes[i] = (C, L, N); offsets[i] = dict(camera_index=, line_index=, fixed_index=, observations=,
    parameters=) byte offsets from `base_ptr` (a device address).  Parameters are updated in place on the device;
    returns the summaries (host copies) when asked, else None (the call then only enqueues)."""
    n = len(shapes)
    o = [0.0, 0.0, 0.0, 0.0] if lm_opts is None else list(lm_opts)
    descs = (LbaDesc * n)()
    pp = (dp * n)()
    for i, ((Cc, Ll, Nn), off) in enumerate(zip(shapes, offsets)):
        descs[i] = LbaDesc(Cc, Ll, Nn, max_iters, C.cast(base_ptr + off["camera_index"], ip), C.cast(base_ptr + off["line_index"], ip),
                           C.cast(base_ptr + off["fixed_index"], ip), C.cast(base_ptr + off["observations"], dp), int(robust),
                           0.0, -1.0, o[0], o[1], o[2], o[3])
        pp[i] = C.cast(base_ptr + off["parameters"], dp)
    ss = (Summary * n)() if want_host_summaries else None
    _check(lib().slslam_lba_solve_batch_device(n, descs, pp, C.c_void_p(summaries_dev_ptr), ss, C.c_void_p(stream)))
    return [summary_dict(x) for x in ss] if want_host_summaries else None


class LbaBatch:
    """Device-resident batch: plan + upload once, solve many times (slslam_lba_batch_*)."""

    def __init__(self, windows, device=-1, cluster_size=0, **kw):
        self.windows = list(windows)
        self.n = len(self.windows)
        self._keeps = [lba_desc(w, **kw) for w in self.windows]
        self._descs = (LbaDesc * self.n)(*[k.desc for k in self._keeps])
        self._p0 = [np.ascontiguousarray(w.parameters, np.float64) for w in self.windows]
        pp = (dp * self.n)(*[_d(p) for p in self._p0])
        self._h = C.c_void_p()
        _check(lib().slslam_lba_batch_create(self.n, self._descs, pp, device, cluster_size, C.byref(self._h)))
        self.max_iters = [k.desc.max_iterations for k in self._keeps]

    def info(self):
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib().slslam_lba_batch_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        m = C.c_int32()
        _check(lib().slslam_lba_batch_max_active_clusters(self._h, C.byref(m)))
        return dict(cluster_size=a.value, ctas_per_window=a.value, threads_per_cta=b.value, smem_bytes_per_cta=c.value,
                    z_in_smem=d.value, max_active_clusters=m.value, windows_per_wave=m.value)

    def transfer_bytes(self):
        a, b = C.c_int64(), C.c_int64()
        _check(lib().slslam_lba_batch_transfer_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    PHASES = ("init", "linearise", "pairs", "fold", "allreduce", "gradient", "reduced_solve", "trial", "decide", "total",
              "rs_prep", "rs_factor_panel", "rs_trailing", "rs_backsub")

    def phase_cycles(self, window=0, stream=None):
        buf = (C.c_int64 * 14)()
        _check(lib().slslam_lba_batch_phase_cycles(self._h, C.c_void_p(stream), window, buf, 14))
        return dict(zip(self.PHASES, list(buf)))

    PLAN_PHASES = ("counts", "scan", "group_by_line", "partition_tiles", "slot_table", "slot_meta_gather", "pair_lists", "total")

    def plan_cycles(self, window=0):
        """Cumulative SM cycles of the device planner's phases (empty dict when the host planner built this batch)."""
        buf = (C.c_int32 * 8)()
        if lib().slslam_lba_batch_plan_cycles(self._h, window, buf) != 0:
            return {}
        return dict(zip(self.PLAN_PHASES, list(buf)))

    def upload(self, params=None, stream=None):
        ps = self._p0 if params is None else [np.ascontiguousarray(p, np.float64) for p in params]
        pp = (dp * self.n)(*[_d(p) for p in ps])
        _check(lib().slslam_lba_batch_upload_params(self._h, pp, C.c_void_p(stream)))

    def solve(self, stream=None):
        _check(lib().slslam_lba_batch_solve(self._h, C.c_void_p(stream)))

    def download(self, stream=None, trace=False):
        ps = [np.zeros(6 * w.num_cameras + 4 * w.num_lines) for w in self.windows]
        pp = (dp * self.n)(*[_d(p) for p in ps])
        ss = (Summary * self.n)()
        trs = [np.zeros((max(1, m), TRACE_WIDTH)) for m in self.max_iters]
        tp = (dp * self.n)(*[_d(t) for t in trs]) if trace else None
        _check(lib().slslam_lba_batch_download(self._h, C.c_void_p(stream), pp, ss, tp))
        return ps, [summary_dict(s, trs[i] if trace else None) for i, s in enumerate(ss)]

    def close(self):
        if self._h:
            lib().slslam_lba_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lba_launch_shape(num_windows, max_observations, max_lines, resident_ctas=148, smem_bytes_per_cta=232448):
    """(CTAs per window, windows per wave) the planner chooses; pure host arithmetic (slslam_lba_launch_shape)."""
    a, b = C.c_int32(), C.c_int32()
    _check(lib().slslam_lba_launch_shape(num_windows, max_observations, max_lines, resident_ctas, smem_bytes_per_cta,
                                         C.byref(a), C.byref(b)))
    return a.value, b.value


def lba_plan_check(windows, cluster_size=0, **kw):
    """Device planner against host planner (slslam_lba_plan_check): returns (code, [window, field, index])."""
    pb = PreparedBatch(windows, **kw)
    pp = (dp * pb.n)(*[_d(p) for p in pb.p0])
    detail = (C.c_int32 * 3)()
    rc = lib().slslam_lba_plan_check(pb.n, pb.descs, pp, cluster_size, detail)
    if rc < 0:
        _check(rc)
    return rc, list(detail)


class PreparedBatch:
    """Descs of a list of windows marshalled once (the numpy arrays stay referenced), for repeated submission."""

    def __init__(self, windows, pin=False, **kw):
        self.windows = list(windows)
        self.n = len(self.windows)
        if pin:
            # observations in page-locked host memory: the library then DMAs them straight from the caller's array
            import copy
            import torch
            pinned = []
            for w in self.windows:
                w = copy.copy(w)
                w.observations = torch.from_numpy(np.ascontiguousarray(w.observations, np.float64)).pin_memory().numpy()
                pinned.append(w)
            self.windows = pinned
        self.keeps = [lba_desc(w, **kw) for w in self.windows]
        self.descs = (LbaDesc * self.n)(*[k.desc for k in self.keeps])
        self.p0 = [np.ascontiguousarray(w.parameters, np.float64) for w in self.windows]


class LbaPipeline:
    """slslam_lba_pipeline_*: submit() returns a ticket at once, wait(ticket) returns (params, summaries)."""

    ASYNC_HOST = 2

    def __init__(self, device=-1, depth=2, flags=0):
        self._h = C.c_void_p()
        self._inflight = {}
        _check(lib().slslam_lba_pipeline_create(device, depth, flags, C.byref(self._h)))

    def submit(self, batch, **kw):
        pb = batch if isinstance(batch, PreparedBatch) else PreparedBatch(batch, **kw)
        ps = [p.copy() for p in pb.p0]
        pp = (dp * pb.n)(*[_d(p) for p in ps])
        ss = (Summary * pb.n)()
        t = C.c_int64()
        _check(lib().slslam_lba_pipeline_submit(self._h, pb.n, pb.descs, pp, ss, C.byref(t)))
        self._inflight[t.value] = (pb, ps, pp, ss)
        return t.value

    def wait(self, ticket):
        _check(lib().slslam_lba_pipeline_wait(self._h, ticket))
        pb, ps, pp, ss = self._inflight.pop(ticket)
        return ps, [summary_dict(s) for s in ss]

    def close(self):
        if self._h:
            lib().slslam_lba_pipeline_destroy(self._h)
            self._h = C.c_void_p()
            self._inflight.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lba_evaluate(w, params=None, robust=True):
    k = lba_desc(w, robust=robust)
    p = np.ascontiguousarray(w.parameters if params is None else params, np.float64)
    N = w.num_observations
    r, Jc, Jl = np.zeros((N, 4)), np.zeros((N, 4, 6)), np.zeros((N, 4, 4))
    cost = C.c_double()
    _check(lib().slslam_lba_evaluate(C.byref(k.desc), _d(p), _d(r), _d(Jc), _d(Jl), C.cast(C.byref(cost), dp)))
    return r, Jc, Jl, cost.value


def po_desc(g, max_iters=10, lm_opts=None) -> _Keep:
    a = np.ascontiguousarray(g.pose_index_1, np.int32)
    b = np.ascontiguousarray(g.pose_index_2, np.int32)
    c = np.ascontiguousarray(g.constraints, np.float64)
    o = [0.0, 0.0, 0.0, 0.0] if lm_opts is None else list(lm_opts)
    return _Keep(PoDesc(g.num_poses, int(a.shape[0]), max_iters, _i(a), _i(b), _d(c), o[0], o[1], o[2], o[3]), (a, b, c))


def po_solve(g, params=None, max_iters=10, lm_opts=None):
    k = po_desc(g, max_iters, lm_opts)
    p = np.ascontiguousarray(g.parameters if params is None else params, np.float64).copy()
    s = Summary()
    tr = np.zeros((max(1, max_iters), TRACE_WIDTH))
    _check(lib().slslam_po_solve_trace(C.byref(k.desc), _d(p), C.byref(s), _d(tr)))
    return p, summary_dict(s, tr)


def po_last_stats():
    """How the last po_solve of this thread factored the normal equations (slslam_po_last_stats)."""
    st = PoStats()
    lib().slslam_po_last_stats(C.byref(st))
    return {k: (list(getattr(st, k)) if k == "factor_cycles" else getattr(st, k)) for k, _ in PoStats._fields_}


ROUTES = {0: "tiled", 1: "motion_only", 2: "general"}


def lba_route(w, **kw):
    """slslam_lba_route: the kernel a window would go to, decided on the host (no device needed)."""
    k = lba_desc(w, **kw)
    rc = lib().slslam_lba_route(C.byref(k.desc))
    _check(rc if rc < 0 else 0)
    return ROUTES[rc]


def po_plan_check(g, force_columns=False):
    """slslam_po_plan_check: the symbolic plan of a pose graph, built and verified on the host (no device needed)."""
    k = po_desc(g)
    info = PoPlanInfo()
    _check(lib().slslam_po_plan_check(C.byref(k.desc), 1 if force_columns else 0, C.byref(info)))
    return {f: getattr(info, f) for f, _ in PoPlanInfo._fields_ if f != "reserved"}


def po_limits():
    lim = PoLimits()
    lib().slslam_po_get_limits(C.byref(lim))
    return {k: getattr(lim, k) for k, _ in PoLimits._fields_}


def po_evaluate(g, params=None):
    k = po_desc(g)
    p = np.ascontiguousarray(g.parameters if params is None else params, np.float64)
    E = g.num_edges
    r, J1, J2 = np.zeros((E, 6)), np.zeros((E, 6, 6)), np.zeros((E, 6, 6))
    cost = C.c_double()
    _check(lib().slslam_po_evaluate(C.byref(k.desc), _d(p), _d(r), _d(J1), _d(J2), C.cast(C.byref(cost), dp)))
    return r, J1, J2, cost.value


def ransac_score(poses, lines, obs, baseline=0.12, thr=5.0 / 406.05, want_inliers=True, want_errors=True):
    """slslam_ransac_score: poses [H][12] (R row-major, t), lines [K][6], obs [K][8] -> (scores, inlier, errors)."""
    poses = np.ascontiguousarray(poses, np.float64); lines = np.ascontiguousarray(lines, np.float64)
    obs = np.ascontiguousarray(obs, np.float64)
    H, K = poses.shape[0], lines.shape[0]
    scores = np.zeros(H, np.int32)
    inl = np.zeros((H, K), np.uint8) if want_inliers else None
    err = np.zeros((H, K), np.float32) if want_errors else None
    _check(lib().slslam_ransac_score(H, _d(poses), K, _d(lines), _d(obs), baseline, thr, _i(scores),
                                     inl.ctypes.data_as(C.POINTER(C.c_uint8)) if want_inliers else None,
                                     err.ctypes.data_as(C.POINTER(C.c_float)) if want_errors else None))
    return scores, inl, err


class MapTimings(C.Structure):
    _fields_ = [("assemble_ms", C.c_double), ("solve_and_writeback_ms", C.c_double), ("total_ms", C.c_double),
                ("h2d_bytes", C.c_int64), ("candidates", C.c_int32)]


def geometry_convert(mode, arr):
    """slslam_geometry_convert: 0 av[n,6] -> orth[n,4]; 1 orth -> av; 2 R[n,9] (row-major) -> w[n,3]; 3 w -> R."""
    L = lib()
    L.slslam_geometry_convert.argtypes = [C.c_int32, C.c_int32, dp, dp]
    a = np.ascontiguousarray(arr, np.float64)
    n = a.shape[0]
    out = np.zeros((n, (4, 6, 3, 9)[mode]))
    _check(L.slslam_geometry_convert(mode, n, _d(a), _d(out)))
    return out


class DeviceMap:
    """slslam_map_*: keyframe poses, landmark lines and observations resident on the device; per keyframe only the new
    observations go up, the window is assembled, solved and written back where it lives."""

    def __init__(self, max_keyframes, max_landmarks, max_observations, device=-1):
        L = lib()
        L.slslam_map_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.slslam_map_destroy.argtypes = [C.c_void_p]; L.slslam_map_destroy.restype = None
        L.slslam_map_add_keyframe.argtypes = [C.c_void_p, C.c_int32, dp, C.c_int32, ip, dp]
        L.slslam_map_add_landmarks.argtypes = [C.c_void_p, C.c_int32, ip, ip, dp]
        L.slslam_map_set_poses.argtypes = [C.c_void_p, C.c_int32, ip, dp]
        L.slslam_map_get_poses.argtypes = [C.c_void_p, C.c_int32, ip, dp]
        L.slslam_map_get_landmarks.argtypes = [C.c_void_p, C.c_int32, ip, dp]
        L.slslam_map_bundle_adjust.argtypes = [C.c_void_p, C.c_int32, ip, ip, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Summary), ip]
        L.slslam_map_last_timings.argtypes = [C.c_void_p, C.POINTER(MapTimings)]; L.slslam_map_last_timings.restype = None
        L.slslam_map_last_window.argtypes = [C.c_void_p, ip, ip, ip, dp, dp, ip, ip]
        self._h = C.c_void_p()
        _check(L.slslam_map_create(device, max_keyframes, max_landmarks, max_observations, C.byref(self._h)))
        self.sizes = (0, 0, 0)

    def add_keyframe(self, kf_id, T12, lm_ids, obs8):
        T = np.ascontiguousarray(T12, np.float64).ravel()
        li = np.ascontiguousarray(lm_ids, np.int32)
        ob = np.ascontiguousarray(obs8, np.float64).ravel()
        _check(lib().slslam_map_add_keyframe(self._h, kf_id, _d(T), int(li.shape[0]), _i(li), _d(ob)))

    def add_landmarks(self, lm_ids, init_kf_ids, lines_av6):
        li = np.ascontiguousarray(lm_ids, np.int32); ki = np.ascontiguousarray(init_kf_ids, np.int32)
        av = np.ascontiguousarray(lines_av6, np.float64).ravel()
        _check(lib().slslam_map_add_landmarks(self._h, int(li.shape[0]), _i(li), _i(ki), _d(av)))

    def set_poses(self, kf_ids, T12):
        ki = np.ascontiguousarray(kf_ids, np.int32); T = np.ascontiguousarray(T12, np.float64).ravel()
        _check(lib().slslam_map_set_poses(self._h, int(ki.shape[0]), _i(ki), _d(T)))

    def get_poses(self, kf_ids):
        ki = np.ascontiguousarray(kf_ids, np.int32)
        out = np.zeros((ki.shape[0], 12))
        _check(lib().slslam_map_get_poses(self._h, int(ki.shape[0]), _i(ki), _d(out)))
        return out

    def get_landmarks(self, lm_ids):
        li = np.ascontiguousarray(lm_ids, np.int32)
        out = np.zeros((li.shape[0], 6))
        _check(lib().slslam_map_get_landmarks(self._h, int(li.shape[0]), _i(li), _d(out)))
        return out

    def bundle_adjust(self, ba_kf_ids, ba_order, window_size, max_iters=10, robust=True):
        ki = np.ascontiguousarray(ba_kf_ids, np.int32); oi = np.ascontiguousarray(ba_order, np.int32)
        s = Summary()
        sz = (C.c_int32 * 3)()
        _check(lib().slslam_map_bundle_adjust(self._h, int(ki.shape[0]), _i(ki), _i(oi), window_size, max_iters, int(robust), C.byref(s), sz))
        self.sizes = tuple(sz)
        return summary_dict(s)

    def last_timings(self):
        t = MapTimings()
        lib().slslam_map_last_timings(self._h, C.byref(t))
        return {k: getattr(t, k) for k, _ in MapTimings._fields_}

    def last_window(self):
        """The window the last bundle_adjust assembled: dict of arrays (parameters as assembled, before the solve)."""
        Cc, Ll, Nn = self.sizes
        ci, li, fi = np.zeros(Nn, np.int32), np.zeros(Nn, np.int32), np.zeros(2 * Nn, np.int32)
        ob, pr = np.zeros(8 * Nn), np.zeros(6 * Cc + 4 * Ll)
        llm, ckf = np.zeros(Ll, np.int32), np.zeros(Cc, np.int32)
        _check(lib().slslam_map_last_window(self._h, _i(ci), _i(li), _i(fi), _d(ob), _d(pr), _i(llm), _i(ckf)))
        return dict(num_cameras=Cc, num_lines=Ll, camera_index=ci, line_index=li, fixed_index=fi, observations=ob,
                    parameters=pr, line_landmark=llm, camera_keyframe=ckf)

    def close(self):
        if self._h:
            lib().slslam_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
