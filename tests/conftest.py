import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "isolated(timeout): run the test body in a child pytest process (kernels that have never run on hardware)")


@pytest.hookimpl(tryfirst=True)
def pytest_pyfunc_call(pyfuncitem):
    """Tests marked `isolated` exercise kernels that have had no hardware run yet.  Their bodies run in a child pytest
    process (own CUDA context, own process group, hard time limit), so a hung kernel or a sticky CUDA error there costs one
    test, not the rest of the suite.  The child's verdict is the test's verdict."""
    mark = pyfuncitem.get_closest_marker("isolated")
    if mark is None or os.environ.get("ASPH_TEST_CHILD"):
        return None
    limit = float(mark.kwargs.get("timeout", mark.args[0] if mark.args else 75))
    node = f"{pyfuncitem.path}::{pyfuncitem.name}"
    cmd = [sys.executable, "-m", "pytest", node, "-q", "-x", "--runxfail", "--tb=short", "-p", "no:cacheprovider"]
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=ROOT, start_new_session=True,
                         env=dict(os.environ, ASPH_TEST_CHILD="1"))
    try:
        out, _ = p.communicate(timeout=limit)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, 9)
        except OSError:
            pass
        out, _ = p.communicate()
        pytest.fail(f"isolated test killed after {limit:.0f} s\n{(out or '')[-1500:]}", pytrace=False)
    if p.returncode == 5 or " skipped" in (out or "") and " passed" not in (out or "") and p.returncode == 0:
        pytest.skip((out or "").strip().splitlines()[-1] if out else "skipped in the child")
    if p.returncode != 0:
        pytest.fail(f"isolated test failed in its child process (rc {p.returncode})\n{(out or '')[-3000:]}", pytrace=False)
    return True


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
