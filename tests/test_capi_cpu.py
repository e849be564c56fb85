"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/slslam_b200.h
declares, and fails loudly (no fallback) when there is no GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from slslam_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        capi.build_library()
    return capi.lib()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "slslam_b200.h")).read()
    declared = set(re.findall(r"\b(slslam_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (slslam_[a-z_0-9]+)", out))
    missing = declared - exported
    assert not missing, f"declared in the header but not exported: {missing}"
    assert set(capi.EXPORTS) <= exported


def test_only_sm100a_code_in_library():
    out = subprocess.check_output(["cuobjdump", "-lelf", capi.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header(lib):
    import ctypes as C
    assert C.sizeof(capi.Summary) == 48
    assert C.sizeof(capi.LbaDesc) == 16 + 4 * 8 + 8 + 6 * 8
    lim = capi.Limits()
    lib.slslam_lba_get_limits(C.byref(lim))
    assert lim.max_cameras >= 20 and lim.max_free_cameras >= 10 and lim.max_cluster_size >= 16


def test_no_cpu_fallback_without_gpu(lib):
    """Without a device every compute entry point must return SLSLAM_ERR_CUDA and leave the parameters alone."""
    if lib.slslam_device_count() > 0:
        pytest.skip("a GPU is present")
    w = synth.make_window(0, 3, 12, 40)
    before = w.parameters.copy()
    with pytest.raises(capi.SlslamError) as e:
        capi.lba_solve(w)
    assert e.value.code == -3
    assert np.array_equal(w.parameters, before)
    with pytest.raises(capi.SlslamError):
        capi.lba_evaluate(w)


def test_validation_precedes_device(lib):
    w = synth.make_window(0, 3, 12, 40)
    w.line_index = w.line_index.copy(); w.line_index[0] = 10 ** 6
    with pytest.raises(capi.SlslamError) as e:
        capi.lba_solve(w)
    assert e.value.code == -1
    assert lib.slslam_strerror(-2).decode().startswith("problem exceeds")


def test_new_entry_points_fail_loudly_without_gpu(lib):
    """Pipeline, planner check and RANSAC scoring: argument errors first, then SLSLAM_ERR_CUDA -- never a CPU result."""
    import ctypes as C
    if lib.slslam_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.slslam_lba_pipeline_create(-1, 0, 0, C.byref(h)) == -1            # depth out of range
    assert lib.slslam_lba_pipeline_create(-1, 2, 0, C.byref(h)) == -3 and not h.value
    poses, lines, obs = np.zeros((2, 12)), np.zeros((3, 6)), np.zeros((3, 8))
    scores = np.full(2, 7, np.int32)
    with pytest.raises(capi.SlslamError) as e:
        capi.ransac_score(poses, lines, obs)
    assert e.value.code == -3
    assert lib.slslam_ransac_score(2, None, 3, capi._d(lines), capi._d(obs), 0.12, 0.01, capi._i(scores), None, None) == -1
    assert (scores == 7).all()
    w = synth.make_window(0, 3, 12, 40)
    with pytest.raises(capi.SlslamError) as e:
        capi.lba_plan_check([w])
    assert e.value.code == -3
