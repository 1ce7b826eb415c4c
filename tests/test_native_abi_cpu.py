"""The C-ABI library: builds for sm_100a without a GPU, loads, exports every symbol include/biovil_b200.h declares,
and fails loudly (no CPU fallback) when there is no device.  No compute call is made here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "biovil_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    declared = _declared_symbols()
    assert len(declared) >= 12
    assert sorted(N.EXPORTED_SYMBOLS) == declared
    for sym in declared:
        assert getattr(native_lib, sym) is not None
    assert b"sm_100a" in native_lib.bv_version()


def test_struct_layouts_match_header(native_lib):
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    assert ctypes.sizeof(N.BvConv) == 40
    assert ctypes.sizeof(N.BvWeights) == 40 * (3 + 4 * 16 + 1) + 16 + 40
    assert ctypes.sizeof(N.BvOutputs) == 80
    assert ctypes.sizeof(N.BvLaunchInfo) == 88


def test_workspace_and_shape_helpers(native_lib):
    assert native_lib.bv_patch_grid(480) == 15 and native_lib.bv_patch_grid(512) == 16
    w1 = native_lib.bv_workspace_bytes(1, 1, 480, 480)
    w512 = native_lib.bv_workspace_bytes(512, 1, 480, 480)
    assert 0 < w1 < w512 < 16 * 2 ** 30
    assert native_lib.bv_workspace_bytes(512, 3, 480, 480) > w512          # 3-channel float stem gathers K=192
    assert native_lib.bv_workspace_bytes(1, 1, 470, 480) == 0               # not a multiple of 32
    assert native_lib.bv_workspace_bytes(0, 1, 480, 480) == 0
    assert native_lib.bv_workspace_bytes(1, 2, 480, 480) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_device_fails_loudly(native_lib):
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    handle = ctypes.c_void_p()
    w = N.BvWeights()
    rc = native_lib.bv_create(ctypes.byref(handle), ctypes.byref(w), 0)
    assert rc == N.BV_ERR_NO_DEVICE
    assert b"no CPU fallback" in native_lib.bv_last_error() or b"CUDA" in native_lib.bv_last_error()
    with pytest.raises(N.NativeError):
        N.check(rc)


def test_sass_contains_blackwell_instructions(native_lib):
    """tcgen05.mma / tcgen05.ld / TMA (tiled + im2col) must be in the binary (B200_PROFILING.md mnemonics)."""
    import shutil
    import subprocess
    from incremental_multimodal_medical_learning_ii_b200 import _native as N
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(N.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG.2D", "UTMALDG.4D.IM2COL", "UTMASTG.2D"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass          # no legacy mma.sync path
