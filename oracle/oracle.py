"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and --impl reference).
Nothing under slslam_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TERMINATION = {0: "NO_CONVERGENCE", 1: "GRADIENT_TOLERANCE", 2: "FUNCTION_TOLERANCE", 3: "PARAMETER_TOLERANCE",
               4: "NUMERICAL_FAILURE"}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("slslam_oracle.cpp", "jet.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        try:
            _LIB = C.CDLL(so)
        except OSError:
            _LIB = C.CDLL(build(force=True))
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L = _LIB
        L.oracle_trace_width.restype = C.c_int
        L.oracle_lba_residual_jacobian.argtypes = [dp, dp, dp, C.c_double, dp, dp, dp]
        L.oracle_lba_residual_jacobian.restype = None
        L.oracle_po_residual_jacobian.argtypes = [dp, dp, dp, dp, dp, dp]
        L.oracle_po_residual_jacobian.restype = None
        L.oracle_lba_cost.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, dp, C.c_int, C.c_double, C.c_double, dp]
        L.oracle_lba_cost.restype = C.c_double
        L.oracle_lba_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip, dp, C.c_int, C.c_double,
                                       C.c_double, C.c_int, dp, dp, dp, dp]
        L.oracle_lba_solve.restype = C.c_int
        L.oracle_po_cost.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp]
        L.oracle_po_cost.restype = C.c_double
        L.oracle_po_solve.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp, dp, dp]
        L.oracle_po_solve.restype = C.c_int
        L.oracle_po_solve2.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, C.c_int, dp, dp, dp, dp]
        L.oracle_po_solve2.restype = C.c_int
        L.oracle_ransac_score.argtypes = [C.c_int, dp, C.c_int, dp, dp, C.c_double, C.c_double, ip,
                                          C.POINTER(C.c_ubyte), C.POINTER(C.c_float)]
        L.oracle_ransac_score.restype = None
    return _LIB


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def lba_residual_jacobian(cam, line, ob, baseline=0.12):
    cam, line, ob = _f64(cam), _f64(line), _f64(ob)
    r, Jc, Jl = np.zeros(4), np.zeros((4, 6)), np.zeros((4, 4))
    lib().oracle_lba_residual_jacobian(_d(cam), _d(line), _d(ob), baseline, _d(r), _d(Jc), _d(Jl))
    return r, Jc, Jl


def po_residual_jacobian(p1, p2, c):
    p1, p2, c = _f64(p1), _f64(p2), _f64(c)
    r, J1, J2 = np.zeros(6), np.zeros((6, 6)), np.zeros((6, 6))
    lib().oracle_po_residual_jacobian(_d(p1), _d(p2), _d(c), _d(r), _d(J1), _d(J2))
    return r, J1, J2


def _summary(s, trace, iters):
    return dict(initial_cost=float(s[0]), final_cost=float(s[1]), num_successful_steps=int(s[2]),
                num_unsuccessful_steps=int(s[3]), termination=TERMINATION[int(s[4])], iterations=int(s[5]),
                fixed_cost=float(s[6]), gradient_max_norm=float(s[7]), trace=trace[:int(s[5])].copy())


def lba_cost(w, params=None, robust=True, huber_delta=1.0 / 406.05, baseline=0.12):
    p = _f64(w.parameters if params is None else params)
    ci, li, ob = _i32(w.camera_index), _i32(w.line_index), _f64(w.observations)
    return lib().oracle_lba_cost(w.num_cameras, w.num_lines, w.num_observations, _i(ci), _i(li), _d(ob),
                                 int(robust), huber_delta, baseline, _d(p))


def lba_solve(w, max_iters=10, robust=True, huber_delta=1.0 / 406.05, baseline=0.12, solver=1, lm_opts=None,
              params=None):
    """Returns (parameters, summary).  solver 0 = full normal equations (reference), 1 = Schur (same step)."""
    p = _f64(w.parameters if params is None else params).copy()
    ci, li, fi, ob = _i32(w.camera_index), _i32(w.line_index), _i32(w.fixed_index), _f64(w.observations)
    tw = lib().oracle_trace_width()
    trace = np.zeros((max(max_iters, 1), tw))
    s = np.zeros(8)
    o = None if lm_opts is None else _f64(lm_opts)
    lib().oracle_lba_solve(w.num_cameras, w.num_lines, w.num_observations, max_iters, _i(ci), _i(li), _i(fi), _d(ob),
                           int(robust), huber_delta, baseline, solver, None if o is None else _d(o), _d(p), _d(s),
                           _d(trace))
    return p, _summary(s, trace, max_iters)


def po_cost(g, params=None):
    p = _f64(g.parameters if params is None else params)
    a, b, c = _i32(g.pose_index_1), _i32(g.pose_index_2), _f64(g.constraints)
    return lib().oracle_po_cost(g.num_poses, g.num_edges, _i(a), _i(b), _d(c), _d(p))


def po_solve(g, max_iters=10, lm_opts=None, params=None, solver=1, want_stats=False):
    """solver 1 = sparse Cholesky of the normal equations (the reference's SPARSE_NORMAL_CHOLESKY, po_problem.cpp:68),
    0 = dense Cholesky (cross-check; same step)."""
    p = _f64(g.parameters if params is None else params).copy()
    a, b, c = _i32(g.pose_index_1), _i32(g.pose_index_2), _f64(g.constraints)
    tw = lib().oracle_trace_width()
    trace = np.zeros((max(max_iters, 1), tw))
    s = np.zeros(8)
    o = None if lm_opts is None else _f64(lm_opts)
    stats = np.zeros(2)
    lib().oracle_po_solve2(g.num_poses, g.num_edges, max_iters, _i(a), _i(b), _d(c), None if o is None else _d(o),
                           int(solver), _d(p), _d(s), _d(trace), _d(stats))
    out = _summary(s, trace, max_iters)
    if want_stats:
        out["factor_flops"], out["factor_nnz"] = float(stats[0]), float(stats[1])
    return p, out


def ransac_score(poses, lines, obs, baseline=0.12, thr=5.0 / 406.05):
    """poses [H][12] (R row-major, t), lines [K][6] (closest point, direction), obs [K][8] -> (scores [H], inlier [H][K], errors [H][K])."""
    poses = np.ascontiguousarray(poses, np.float64); lines = np.ascontiguousarray(lines, np.float64)
    obs = np.ascontiguousarray(obs, np.float64)
    H, K = poses.shape[0], lines.shape[0]
    scores = np.zeros(H, np.int32); inl = np.zeros((H, K), np.uint8); err = np.zeros((H, K), np.float32)
    lib().oracle_ransac_score(H, _d(poses), K, _d(lines), _d(obs), baseline, thr,
                              scores.ctypes.data_as(C.POINTER(C.c_int)), inl.ctypes.data_as(C.POINTER(C.c_ubyte)),
                              err.ctypes.data_as(C.POINTER(C.c_float)))
    return scores, inl, err
