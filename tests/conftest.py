import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def asph():
    import asph_b200
    return asph_b200


def _ensure_oracle():
    libs = [os.path.join(ROOT, "oracle", f"liboracle_{k}.so") for k in ("f32", "f64")]
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("oracle_capi.cpp", "sph_oracle.hpp")] + \
           [os.path.join(ROOT, "include", "asph.h")]
    stale = any(not os.path.exists(l) or os.path.getmtime(l) < max(os.path.getmtime(s) for s in srcs) for l in libs)
    if stale:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j2"], stdout=subprocess.DEVNULL)
    return libs


@pytest.fixture(scope="session")
def oracle32(asph):
    """CPU oracle, fp32 build (tests only)."""
    return asph.load_library(_ensure_oracle()[0])


@pytest.fixture(scope="session")
def oracle64(asph):
    return asph.load_library(_ensure_oracle()[1])


@pytest.fixture(scope="session")
def cuda_lib(asph):
    """The product library; GPU tests call the kernels through this C ABI only."""
    return asph.load_library()


@pytest.fixture(scope="session")
def default_params(asph):
    return asph.SimulationParams.from_yaml(os.path.join(ROOT, "configs", "default-config.yaml"))


@pytest.fixture(scope="session")
def split_patterns(asph):
    return asph.load_split_patterns_from_file()
