"""The reference-facing C++ boundary: include/lba_problem.h, include/po_problem.h and the ceres shim, driven through
the reference's own construct -> setters -> build -> set_options -> ceres::Solve sequence (reference
src/slam.cpp:924-944, 1283-1293) by slslam_b200/host/slslam_host_demo.cpp."""
import os
import subprocess

import numpy as np
import pytest

from slslam_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "slslam_b200", "host")
DEMO = os.path.join(HOST, "slslam_host_demo")


@pytest.fixture(scope="module")
def demo():
    if not os.path.exists(capi.LIB_PATH):
        capi.build_library()
    subprocess.check_call(["make", "-s", "-C", HOST])
    return DEMO


def _write_lba(path, w, max_iters, robust=True):
    with open(path, "wb") as f:
        np.array([w.num_cameras, w.num_lines, w.num_observations, max_iters, int(robust)], np.int32).tofile(f)
        np.ascontiguousarray(w.camera_index, np.int32).tofile(f)
        np.ascontiguousarray(w.line_index, np.int32).tofile(f)
        np.ascontiguousarray(w.fixed_index, np.int32).tofile(f)
        np.ascontiguousarray(w.observations, np.float64).tofile(f)
        np.ascontiguousarray(w.parameters, np.float64).tofile(f)


def _write_po(path, g, max_iters):
    with open(path, "wb") as f:
        np.array([g.num_poses, g.num_edges, max_iters], np.int32).tofile(f)
        np.ascontiguousarray(g.pose_index_1, np.int32).tofile(f)
        np.ascontiguousarray(g.pose_index_2, np.int32).tofile(f)
        np.ascontiguousarray(g.constraints, np.float64).tofile(f)
        np.ascontiguousarray(g.parameters, np.float64).tofile(f)


def _run(demo, kind, inp, out):
    p = subprocess.run([demo, kind, inp, out], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    d = np.fromfile(out, np.float64)
    head = dict(error_code=int(d[0]), initial_cost=d[1], final_cost=d[2], successful=int(d[3]), unsuccessful=int(d[4]),
                termination=int(d[5]))
    return head, d[6:], p.stdout


def test_headers_compile_as_cxx98():
    """The reference is C++03-style code; the drop-in headers must not force a newer standard on slam.cpp."""
    src = '#include "lba_problem.h"\n#include "po_problem.h"\nint main() { ceres::Problem p; ceres::Solver::Options o; ' \
          'ceres::Solver::Summary s; ceres::Solve(o, &p, &s); double R[9], w[3] = {0.1, 0.2, 0.3}; ' \
          'ceres::AngleAxisToRotationMatrix(w, R); ceres::RotationMatrixToAngleAxis(R, w); return s.num_successful_steps; }\n'
    subprocess.run(["g++", "-std=gnu++98", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"],
                   input=src, text=True, check=True)


def test_rotation_shim_round_trip(tmp_path):
    """ceres/rotation.h as gc.cpp uses it (reference src/gc.cpp:24-48): column-major, log(exp(w)) = w, near-pi branch."""
    src = r'''
#include <cstdio>
#include "ceres/rotation.h"
int main() {
  const double ws[5][3] = {{0.1, -0.2, 0.3}, {0, 0, 0}, {1e-9, 0, 0}, {3.1, 0.2, -0.1}, {0, 3.14159265358979, 0}};
  for (int i = 0; i < 5; ++i) {
    double R[9], w[3];
    ceres::AngleAxisToRotationMatrix(ws[i], R);
    ceres::RotationMatrixToAngleAxis(R, w);
    printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", w[0], w[1], w[2], R[1], R[3], R[2], R[6]);
  }
  return 0;
}'''
    exe = str(tmp_path / "rot")
    subprocess.run(["g++", "-O1", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-", "-o", exe], input=src, text=True, check=True)
    out = subprocess.check_output([exe], text=True).strip().splitlines()
    ws = [[0.1, -0.2, 0.3], [0, 0, 0], [1e-9, 0, 0], [3.1, 0.2, -0.1], [0, 3.14159265358979, 0]]
    for line, w in zip(out, ws):
        v = np.array([float(x) for x in line.split()])
        R = synth.rodrigues(np.array(w, float))
        assert np.abs(v[:3] - np.array(w)).max() < 1e-7 if np.linalg.norm(w) > 3 else np.abs(v[:3] - np.array(w)).max() < 1e-12
        # column-major: R[1] = R_10, R[3] = R_01, R[2] = R_20, R[6] = R_02
        assert np.abs(v[3:] - np.array([R[1, 0], R[0, 1], R[2, 0], R[0, 2]])).max() < 1e-12


def test_no_gpu_leaves_parameters_untouched(demo, tmp_path):
    if capi.lib().slslam_device_count() > 0:
        pytest.skip("a GPU is present")
    w = synth.make_window(0, 3, 12, 40)
    inp, out = str(tmp_path / "w.bin"), str(tmp_path / "o.bin")
    _write_lba(inp, w, 5)
    head, params, stdout = _run(demo, "lba", inp, out)
    assert head["error_code"] == -3 and "failed" in stdout
    assert np.array_equal(params, w.parameters)
    # a failed solve contributes ZEROS to the statistics the reference accumulates without looking (src/slam.cpp:949-952),
    # never the -1 "did not run" markers, and says so on stderr although LBAProblem::set_options forces SILENT
    assert head["initial_cost"] == 0.0 and head["final_cost"] == 0.0 and head["successful"] == 0 and head["unsuccessful"] == 0
    p = subprocess.run([demo, "lba", inp, out], capture_output=True, text=True, timeout=120)
    assert "solve failed" in p.stderr


@pytest.mark.gpu
def test_lba_through_cpp_boundary(demo, gpu, tmp_path):
    from oracle import oracle
    for seed, robust in ((2, True), (3, False)):
        w = synth.window_S(seed, sigma_px=0.5)
        inp, out = str(tmp_path / "w.bin"), str(tmp_path / "o.bin")
        _write_lba(inp, w, 10, robust)
        head, params, _ = _run(demo, "lba", inp, out)
        po, so = oracle.lba_solve(w, max_iters=10, solver=1, robust=robust)
        assert head["error_code"] == 0
        assert abs(head["final_cost"] - so["final_cost"]) < 1e-6 * so["final_cost"]
        assert abs(head["initial_cost"] - so["initial_cost"]) < 1e-11 * so["initial_cost"]
        assert head["successful"] + head["unsuccessful"] == so["num_successful_steps"] + so["num_unsuccessful_steps"]
        assert np.abs(params[:6 * w.num_cameras] - po[:6 * w.num_cameras]).max() < 1e-6
        # identical bits to the ctypes route: both are the same C-ABI call
        pc, sc = gpu.lba_solve(w, max_iters=10, robust=robust)
        assert np.array_equal(pc, params)


@pytest.mark.gpu
def test_motion_only_ba_through_cpp_boundary(demo, gpu, tmp_path):
    """SLAM::motion_only_ba's problem (reference src/slam.cpp:578-675) through the same LBAProblem call sequence: the
    library routes it to the motion-only kernel; only camera 0 may change."""
    from oracle import oracle
    w = synth.motion_only_window(21, num_lines=80)
    inp, out = str(tmp_path / "m.bin"), str(tmp_path / "o.bin")
    _write_lba(inp, w, 10)
    head, params, _ = _run(demo, "lba", inp, out)
    po, so = oracle.lba_solve(w, max_iters=10, solver=1)
    assert head["error_code"] == 0
    assert abs(head["final_cost"] - so["final_cost"]) < 1e-6 * so["final_cost"]
    assert np.abs(params[:6] - po[:6]).max() < 1e-8
    assert np.array_equal(params[6:], w.parameters[6:])


@pytest.mark.gpu
def test_po_through_cpp_boundary(demo, gpu, tmp_path):
    from oracle import oracle
    g = synth.make_pose_graph(1, num_poses=40, neighbours=2, num_loops=3)
    inp, out = str(tmp_path / "g.bin"), str(tmp_path / "o.bin")
    _write_po(inp, g, 10)
    head, params, _ = _run(demo, "po", inp, out)
    po, so = oracle.po_solve(g, max_iters=10)
    assert head["error_code"] == 0
    assert abs(head["final_cost"] - so["final_cost"]) < 1e-6 * so["final_cost"]
    assert np.abs(params - po).max() < 1e-6


@pytest.mark.gpu
def test_wide_window_through_cpp_boundary(demo, gpu, tmp_path):
    """A --ba_window_size 20 window (20 free + 20 constant keyframes = 40 camera blocks, beyond the tiled kernel) through
    LBAProblem -> ceres::Solve: routed to the general kernel, no silent no-op (reference src/slam.cpp:1376-1382)."""
    from oracle import oracle
    w = synth.make_window(31, 20, 150, 1900, num_fixed_cameras=20, sigma_px=0.5)
    assert w.num_cameras == 40
    inp, out = str(tmp_path / "w.bin"), str(tmp_path / "o.bin")
    _write_lba(inp, w, 10)
    head, params, _ = _run(demo, "lba", inp, out)
    po, so = oracle.lba_solve(w, max_iters=10, solver=1)
    assert head["error_code"] == 0
    assert abs(head["final_cost"] - so["final_cost"]) < 1e-6 * so["final_cost"]
    assert head["successful"] == so["num_successful_steps"] and head["unsuccessful"] == so["num_unsuccessful_steps"]
    assert np.abs(params[:6 * 40] - po[:6 * 40]).max() < 1e-6
