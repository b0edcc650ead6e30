"""The shipped C-ABI library loads and exports every symbol include/mvdecon.h declares (no compute without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mvdecon.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"MVD_API\s+[\w\s\*]+?\b(\w+)\s*\(", src)))


def test_header_symbols_are_bound_by_the_python_binding():
    import mvrecon_b200 as m
    assert set(_declared_symbols()) == set(m.SYMBOLS)


def test_product_library_exports_every_declared_symbol():
    import ctypes
    import mvrecon_b200 as m
    if not os.path.exists(m.LIBRARY_PATH):
        pytest.fail("libmvdecon.so has not been built (run `make` or __graft_entry__.build())")
    dll = ctypes.CDLL(m.LIBRARY_PATH)
    for name in _declared_symbols():
        assert hasattr(dll, name), name
    assert dll.mvd_version() >= 200
    assert dll.mvd_reference_threads() == max(4, os.cpu_count() or 1)


def test_no_cpu_fallback_without_device():
    """Without a usable device every compute entry point must fail loudly (never silently fall back)."""
    import numpy as np
    import mvrecon_b200 as m
    lib = m.lib()
    if lib.getNumDevicesCUDA() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(m.MvdError):
        lib.convolve(np.zeros((8, 8, 8), np.float32), np.ones((3, 3, 3), np.float32))
    with pytest.raises(m.MvdError):
        m.DeconViews([m.DeconView(np.ones((8, 8, 8), np.float32), np.ones((8, 8, 8), np.float32), np.ones((3, 3, 3), np.float32))])


def test_package_never_references_the_oracle_or_the_emulator():
    pkg = os.path.join(ROOT, "multiview-reconstruction_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py",)):
                text = open(os.path.join(dirpath, fn)).read()
                assert "mvdecon_oracle" not in text and "hostemu" not in text, fn


def test_legacy_device_query_symbols_do_not_crash():
    import mvrecon_b200 as m
    lib = m.lib()
    n = lib.getNumDevicesCUDA()
    assert n >= -1
    if n > 0:
        assert isinstance(lib.getNameDeviceCUDA(0), str) and lib.dll.getMemDeviceCUDA(0) > 0
        assert lib.dll.getCUDAcomputeCapabilityMajorVersion(0) >= 9
