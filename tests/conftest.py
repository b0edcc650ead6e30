import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import mvrecon_b200 as m
        return m.lib().getNumDevicesCUDA() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import mvdecon_oracle
    return mvdecon_oracle


@pytest.fixture(scope="session")
def product_lib():
    """The shipped CUDA library through its C ABI (GPU tests)."""
    import mvrecon_b200 as m
    return m.lib()


@pytest.fixture(scope="session")
def hostemu_lib():
    """TEST-ONLY build of the same kernel bodies for the CPU (index-math validation without a GPU).
    Never used by the product package; see multiview-reconstruction_b200/csrc/backend.h."""
    import mvrecon_b200 as m
    path = os.path.join(ROOT, "tests", "host", "libmvdecon_hostemu.so")
    r = subprocess.run(["make", "-j8", "hostemu"], cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(path):
        pytest.fail("could not build the host emulator:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return m.Lib(path)


@pytest.fixture(scope="session")
def small_dataset(oracle):
    return oracle.make_synthetic((33, 36, 40), 3, seed=1, psf_size_xyz=(7, 5, 7), psf_sigma_xyz=(1.2, 1.0, 2.0), bead_density=512)
