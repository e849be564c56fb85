"""A second, independently written restatement of the trust-region loop the oracle implements (SURVEY.md Appendix A3,
Ceres 1.7.0 semantics): dense numpy normal equations, Jacobians from torch.autograd on the matrix / Pluecker form of
the residual (tests/golden/make_golden.py) instead of dual numbers on AngleAxisRotatePoint.  The reference itself cannot
run here (Ceres absent), so this does not pin the oracle to Ceres; it pins the oracle's IMPLEMENTATION of the stated
semantics: cost sequence, radius schedule, accept / reject decisions, termination and result must agree."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from slslam_b200 import synth  # noqa: E402

A_HUBER = 1.0 / 406.05


def _evaluate(w, x, robust, want_jac, free_col):
    """cost, fixed cost, corrected residual vector, corrected Jacobian over the free columns (unscaled)."""
    import torch
    from make_golden import lba_residual
    C = w.num_cameras
    N = w.num_observations
    ncol = max(free_col.values(), default=-1) + 1 if free_col else 0
    r_all = np.zeros(4 * N)
    J = np.zeros((4 * N, len(free_col))) if want_jac else None
    cost = fixed = 0.0
    ob_all = w.observations.reshape(-1, 8)
    for i in range(N):
        c, l = int(w.camera_index[i]), int(w.line_index[i])
        cam = torch.tensor(x[6 * c:6 * c + 6]); line = torch.tensor(x[6 * C + 4 * l:6 * C + 4 * l + 4]); ob = torch.tensor(ob_all[i])
        r = lba_residual(cam, line, ob).detach().numpy()
        s = float(r @ r)
        if robust and s > A_HUBER ** 2:
            rho, rho1 = 2 * A_HUBER * np.sqrt(s) - A_HUBER ** 2, A_HUBER / np.sqrt(s)
        else:
            rho, rho1 = s, 1.0
        cam_free, line_free = ("c", c, 0) in free_col, ("l", l, 0) in free_col
        if cam_free or line_free:
            cost += 0.5 * rho
        else:
            fixed += 0.5 * rho
            continue
        sw = np.sqrt(rho1)
        r_all[4 * i:4 * i + 4] = sw * r
        if want_jac:
            Jc, Jl = torch.autograd.functional.jacobian(lambda a, b: lba_residual(a, b, ob), (cam, line))
            if cam_free:
                for k in range(6):
                    J[4 * i:4 * i + 4, free_col[("c", c, k)]] = sw * Jc[:, k].numpy()
            if line_free:
                for k in range(4):
                    J[4 * i:4 * i + 4, free_col[("l", l, k)]] = sw * Jl[:, k].numpy()
    return cost, fixed, r_all, J


def lm_numpy(w, max_iters, robust=True):
    C, L = w.num_cameras, w.num_lines
    fx = w.fixed_index.reshape(-1, 2)
    cam_used, cam_const = np.zeros(C, bool), np.zeros(C, bool)
    line_used, line_const = np.zeros(L, bool), np.zeros(L, bool)
    for i in range(w.num_observations):                   # constants are sticky per block (lba_problem.cpp:88-91)
        c, l = w.camera_index[i], w.line_index[i]
        cam_used[c] = line_used[l] = True
        cam_const[c] |= bool(fx[i, 0]); line_const[l] |= bool(fx[i, 1])
    free_col, col_index = {}, []
    for c in range(C):
        if cam_used[c] and not cam_const[c]:
            for k in range(6):
                free_col[("c", c, k)] = len(col_index); col_index.append(6 * c + k)
    for l in range(L):
        if line_used[l] and not line_const[l]:
            for k in range(4):
                free_col[("l", l, k)] = len(col_index); col_index.append(6 * C + 4 * l + k)
    col_index = np.array(col_index)
    x = w.parameters.astype(np.float64).copy()
    cost, fixed, r, J = _evaluate(w, x, robust, True, free_col)
    initial = cost + fixed
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))          # Jacobi scaling, once, at x0
    radius, shrink = 1e4, 2.0
    trace, succ, unsucc, invalid, term = [], 0, 0, 0, "NO_CONVERGENCE"
    gtol_abs = None
    for it in range(max_iters):
        g_unscaled = J.T @ r
        gmax = np.abs(g_unscaled).max()
        if gtol_abs is None:
            gtol_abs = 1e-10 * max(gmax, np.finfo(float).eps)
        if gmax <= gtol_abs:
            term = "GRADIENT_TOLERANCE"; break
        Js = J * scale
        H, g = Js.T @ Js, Js.T @ r
        diag = np.clip(np.diag(H), 1e-6, 1e32) / radius
        rec = [cost, 0.0, 0.0, radius, 0.0, 0.0]
        try:
            y = np.linalg.solve(H + np.diag(diag), g)
            ok = np.all(np.isfinite(y))
        except np.linalg.LinAlgError:
            ok = False
        model = 0.5 * float(y @ (g + diag * y)) if ok else 0.0   # = -(m.(r + m/2)), m = J step, for an exact solve
        rec[2] = model
        if not ok or model < 0.0:
            unsucc += 1; rec[5] = -1.0; trace.append(rec); invalid += 1
            if invalid >= 5:
                term = "NUMERICAL_FAILURE"; break
            radius *= 0.5
            continue
        invalid = 0
        delta = -y * scale
        xt = x.copy(); xt[col_index] += delta
        new_cost, _, _, _ = _evaluate(w, xt, robust, False, free_col)
        step_norm, x_norm = np.linalg.norm(delta), np.linalg.norm(x[col_index])
        rec[1], rec[4] = new_cost, step_norm
        if step_norm <= 1e-8 * (x_norm + 1e-8):
            term = "PARAMETER_TOLERANCE"; trace.append(rec); break
        change = cost - new_cost
        if abs(change) < 1e-6 * cost:
            term = "FUNCTION_TOLERANCE"; trace.append(rec); break
        q = change / model
        if q > 1e-3:
            succ += 1; rec[5] = 1.0
            x = xt
            cost, fixed, r, J = _evaluate(w, x, robust, True, free_col)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * q - 1.0) ** 3)); shrink = 2.0
        else:
            unsucc += 1
            radius /= shrink; shrink *= 2.0
        trace.append(rec)
    return x, dict(initial_cost=initial, final_cost=cost + fixed, iterations=len(trace), termination=term,
                   num_successful_steps=succ, num_unsuccessful_steps=unsucc, trace=np.array(trace))


@pytest.mark.parametrize("case", ["anchored_far", "gauge_free", "not_robust", "motion_only", "converges"])
def test_oracle_lm_loop_matches_independent_restatement(case):
    from oracle import oracle
    kw, robust, iters = dict(sigma_px=1.0, start="far"), True, 6
    if case == "gauge_free":
        kw["anchored"] = False
    if case == "not_robust":
        robust = False
    if case == "motion_only":
        w = synth.motion_only_window(4, num_lines=25, sigma_px=1.0)
    elif case == "converges":
        w, iters = synth.make_window(5, 3, 14, 40, sigma_px=0.3, start="near"), 25      # reaches the function tolerance
    else:
        w = synth.make_window(1, 4, 18, 64, **kw)
    xo, so = oracle.lba_solve(w, max_iters=iters, robust=robust, solver=0)      # full normal equations, as the reference
    xn, sn = lm_numpy(w, iters, robust)
    assert abs(so["initial_cost"] - sn["initial_cost"]) <= 1e-12 * sn["initial_cost"]
    assert so["iterations"] == sn["iterations"] and so["termination"] == sn["termination"]
    assert so["num_successful_steps"] == sn["num_successful_steps"]
    to, tn = so["trace"], sn["trace"]
    for k in range(sn["iterations"]):
        assert to[k, 5] == tn[k, 5], (k, to[k], tn[k])                           # accepted / rejected / invalid
        tol = 1e-9 * 10 ** min(k, 4)        # rounding differences compound through ill-conditioned steps (as in test_lba_gpu.py)
        assert abs(to[k, 0] - tn[k, 0]) <= tol * tn[k, 0], (k, to[k], tn[k])     # cost at the linearisation point
        assert abs(to[k, 3] - tn[k, 3]) <= 1e-7 * tn[k, 3], (k, to[k], tn[k])    # trust-region radius
        assert abs(to[k, 2] - tn[k, 2]) <= 1e-6 * abs(tn[k, 2]) + 1e-18          # model cost change
    assert abs(so["final_cost"] - sn["final_cost"]) <= 1e-6 * sn["final_cost"]          # SURVEY.md §8c tolerance
    assert np.abs(xo - xn).max() < 1e-5
    if case == "converges":
        assert sn["termination"] != "NO_CONVERGENCE"
