import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without an NVIDIA device node the gpu-marked tests are skipped (a plain `pytest` on a CPU-only box); on a box
    that HAS a GPU nothing is skipped: a missing libidp_contact.so or a failing idp_create must fail loudly there."""
    import glob
    if glob.glob("/dev/nvidia[0-9]*"):
        return
    skip = pytest.mark.skip(reason="no NVIDIA device on this machine (gpu-marked tests run on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    from oracle.binding import Oracle
    return Oracle("parity")


@pytest.fixture(scope="session")
def lib_built():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "idp_b200", "csrc"), "-j4"], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "idp_b200", "libidp_contact.so")


@pytest.fixture(scope="session")
def gpu_ctx(lib_built):
    from idp_b200 import ContactContext
    ctx = ContactContext(0)
    yield ctx
    ctx.close()


def lexsorted(a):
    a = np.asarray(a)
    if len(a) == 0:
        return a
    return a[np.lexsort(a.T[::-1])]


def make_cases():
    """Small meshes the oracle finishes in seconds: (name, mesh, direction, dhat list)."""
    from idp_b200 import meshgen
    cases = []
    m, d = meshgen.nested_icospheres(nu=12, gap=4e-2, jitter=1e-3, seed=11)
    cases.append(("icospheres12", m, d, [1e-2, 5e-2, 9e-2]))
    m, d = meshgen.sheet_stack(n_sheets=3, nx=24, ny=20, h=3e-2, A=1.1e-2, jitter=1e-4, seed=5, dir_sigma=8e-3, dir_seed=6)
    cases.append(("sheets3", m, d, [1.5e-2, 3e-2]))
    m = meshgen.random_soup(150, seed=3, scale=1.0, tri_size=0.12)
    rng = np.random.default_rng(9)
    cases.append(("soup150", m, rng.normal(0, 0.05, m.X.shape), [2e-2, 8e-2]))
    return cases
