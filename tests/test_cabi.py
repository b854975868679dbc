"""CPU tests of the drop-in boundary: libsrb200.so builds, loads, and exports exactly the entry
points include/srb200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "srb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srb_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared_functions()
    assert "srb_eval" in names and "srb_create" in names and len(names) >= 25


def test_library_exports_every_declared_symbol():
    import srb200
    lib = srb200.load_library()
    for name in _declared_functions():
        assert hasattr(lib, name), "libsrb200.so does not export %s" % name
    # the Python binding covers the whole header, nothing more
    assert sorted(srb200.engine.SIGNATURES) == _declared_functions()
    assert lib.srb_version().startswith(b"srb200")


def test_library_is_sm100a_and_has_no_oracle_dependency():
    import subprocess
    import srb200
    path = srb200.engine.library_path()
    out = subprocess.run(["cuobjdump", "--list-elf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "sr_oracle" not in ldd and "sr_ref" not in ldd


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product fails loudly instead of computing on the CPU."""
    import srb200
    if srb200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(srb200.SrbError) as ei:
        srb200.Engine((4, 1, 8, 8), 2)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "super-resolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "sr_oracle" not in text and "sr_ref" not in text and "oracle/" not in text, f
                assert "/root/reference" not in text, f


def test_invalid_arguments_are_rejected_before_touching_cuda():
    import srb200
    lib = srb200.load_library()
    ctx = ctypes.c_void_p()
    desc = srb200.engine.ModelDesc(4, 4, 1, 0, 2, 0, None, None)     # 0 observations
    st = lib.srb_create(ctypes.byref(desc), 0, ctypes.byref(ctx))
    assert st == 1 and b"0 observations" in lib.srb_last_error(ctx)
    lib.srb_destroy(ctx)
    psf = np.ones((2, 2))
    desc = srb200.engine.ModelDesc(4, 4, 1, 1, 2, 2, psf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None)
    st = lib.srb_create(ctypes.byref(desc), 0, ctypes.byref(ctx))
    assert st == 1
    lib.srb_destroy(ctx)


def test_multi_device_entry_points_fail_loudly_without_a_device_or_context():
    """The multi-GPU forms behind the same boundary: a NULL context is SRB_ERR_INVALID, and without a CUDA device
    srb_multi_create fails with SRB_ERR_CUDA (there is no CPU path to fall back to)."""
    import srb200
    lib = srb200.load_library()
    x = np.zeros(4)
    xp = x.ctypes.data_as(ctypes.c_void_p)
    assert lib.srb_multi_eval(None, xp, None, None) == 1
    assert lib.srb_multi_cg_minimize(None, xp, None, None) == 1
    assert lib.srb_multi_lbfgs_minimize(None, xp, None, None) == 1
    assert lib.srb_multi_solve_irls(None, xp, None, 20, 1e-5, None) == 1
    if srb200.device_count() > 0:
        return
    with pytest.raises(srb200.SrbError) as ei:
        srb200.MultiEngine((4, 1, 8, 8), 2, n_gpus=2, partition=srb200.PARTITION_ROWS)
    assert "no CPU fallback" in str(ei.value)
