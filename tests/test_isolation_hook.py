"""The `isolated` marker of tests/conftest.py (bodies of tests whose kernels never ran on hardware execute in a child pytest
process): pass, failure, time-out and skip in the child become the verdict of the parent test."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = '''
import time
import pytest

@pytest.mark.isolated(timeout=20)
def test_passes():
    assert True

@pytest.mark.isolated(timeout=20)
def test_fails():
    assert False, "boom in the child"

@pytest.mark.isolated(timeout=2)
def test_hangs():
    time.sleep(300)

@pytest.mark.isolated(timeout=20)
def test_skips():
    pytest.skip("needs 2 GPUs")

@pytest.mark.isolated(timeout=20)
@pytest.mark.xfail(strict=False, reason="pending")
@pytest.mark.parametrize("mode", ["a", "b"])
def test_pending(mode):
    assert mode == "a"
'''


def test_isolated_marker_outcomes(tmp_path):
    f = tmp_path / "test_cases.py"
    f.write_text(CASES)
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests") + os.pathsep + os.environ.get("PYTHONPATH", ""),
               PYTEST_ADDOPTS="-p conftest")   # tests/conftest.py as a plugin, in the parent and (inherited) in the children
    env.pop("ASPH_TEST_CHILD", None)
    out = subprocess.run([sys.executable, "-m", "pytest", str(f), "-q", "-rA", "-p", "no:cacheprovider"], capture_output=True, text=True,
                         timeout=300, env=env, cwd=str(tmp_path))
    text = out.stdout
    assert "PASSED test_cases.py::test_passes" in text, text[-3000:]
    assert "FAILED test_cases.py::test_fails" in text and "boom in the child" in text
    assert "FAILED test_cases.py::test_hangs" in text and "killed after 2 s" in text
    assert "SKIPPED" in text
    assert "XPASS test_cases.py::test_pending[a]" in text and "XFAIL test_cases.py::test_pending[b]" in text
    assert "2 failed, 1 passed, 1 skipped, 1 xfailed, 1 xpassed" in text
