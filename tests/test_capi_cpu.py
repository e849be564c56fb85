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


def test_launch_shape_policy(lib):
    """Group size and waves are chosen together (no GPU needed): one launch when the batch fits, equally full waves when
    it does not, never fewer than three tiles per CTA, never more CTAs than the per-line state needs shared memory for."""
    assert capi.lba_launch_shape(8, 10000, 2000) == (18, 8)          # the bench shape: 8 x 18 = 144 of 148 SMs
    assert capi.lba_launch_shape(1, 10000, 2000) == (48, 1)          # a single window: capped at 48 CTAs
    assert capi.lba_launch_shape(16, 10000, 2000) == (9, 16)
    assert capi.lba_launch_shape(64, 10000, 2000) == (9, 16)         # 4 x 16, not 18 + 18 + 18 + 10
    assert capi.lba_launch_shape(32, 10000, 2000) == (9, 16)
    cs, per = capi.lba_launch_shape(1, 1000, 200)                    # a small window keeps >= 3 tiles per CTA
    assert 1 <= cs <= 12 and per == 1
    cs, per = capi.lba_launch_shape(300, 1000, 200)                  # more windows than SMs: one CTA each, in waves
    assert cs == 1 and per * ((300 + per - 1) // per) >= 300 and per <= 148
    cs, per = capi.lba_launch_shape(148, 10000, 2000)                # 2 k lines do not fit one CTA's shared memory
    assert cs >= 4 and cs * per <= 148
    assert lib.slslam_lba_launch_shape(0, 1, 1, 148, 232448, None, None) == -1
