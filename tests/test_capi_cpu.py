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
